"""Video Swin Transformer backbone on the clover_b200 kernels.

Same class name, constructor, parameter / buffer names and return structure as the reference
(mmaction/models/backbones/swin_transformer_3d.py), so configs select it unchanged and reference
checkpoints load with strict=True.  Activations stay channels-last ([tokens, C]) end to end; the
(B, C, D, H, W) layout of the reference is only produced as a view at the module boundary.
Sub-modules (nn.Linear / nn.LayerNorm / nn.Conv3d) are parameter containers: their own forward is
never called, all arithmetic goes through the C ABI.
"""
import numpy as np
import torch
import torch.nn as nn

from . import functional as Fn
from . import ops, tables
from .tables import get_window_size


def trunc_normal_(t, mean=0.0, std=1.0):
    return nn.init.trunc_normal_(t, mean=mean, std=std, a=-2.0, b=2.0)


_DEV_TABLES = {}


def _device_tables(dims, window, shift, window_cfg, device):
    """(rel_code int32 [N], code_off, region int32 [nWin,N] | None) on `device`, cached."""
    key = (dims, window, shift, tuple(window_cfg), str(device))
    ent = _DEV_TABLES.get(key)
    if ent is None:
        N = window[0] * window[1] * window[2]
        code, off = tables.rel_code(N, tuple(window_cfg))
        region = None
        if any(s > 0 for s in shift):
            region = torch.from_numpy(tables.region_ids(*dims, window, shift)).to(device)
        ent = (torch.from_numpy(code).to(device), off, region)
        _DEV_TABLES[key] = ent
    return ent


_W7_SPECS = {}


def _w7_spec(dims, window, shift, window_cfg, device):
    """ops.W7Spec for (wd, 7, 7) windows of a (Wd, 7, 7) configuration (specialised attention kernels), else None.
    `dims` is the padded frame extent the shift-mask regions are computed on."""
    if tuple(window[1:]) != (7, 7) or tuple(window_cfg[1:]) != (7, 7) or window[0] % 2 or not 2 <= window[0] <= 8:
        return None
    key = (dims, tuple(window), tuple(shift), tuple(window_cfg), str(device))
    spec = _W7_SPECS.get(key)
    if spec is None:
        q_ext = k_ext = None
        if any(s > 0 for s in shift):
            q, k = tables.w7_ext_tables(tables.region_ids(*dims, tuple(window), tuple(shift)))
            q_ext = torch.from_numpy(q).to(device=device, dtype=torch.bfloat16).contiguous()
            k_ext = torch.from_numpy(k).to(device=device, dtype=torch.bfloat16).contiguous()
        spec = ops.W7Spec(window[0], window_cfg[0], q_ext, k_ext)
        _W7_SPECS[key] = spec
    return spec


class Mlp(nn.Module):
    """reference :250-268 (parameter container)."""

    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, drop=0.0):
        super().__init__()
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.act = act_layer()
        self.fc2 = nn.Linear(hidden_features, out_features)
        self.drop = nn.Dropout(drop)


class WindowAttention3D(nn.Module):
    """reference :318-400.  ``forward(x, mask)`` keeps the reference contract for windows (B_, N, C);
    ``mask`` must come from :func:`compute_mask` (it carries the region ids the kernel needs)."""

    def __init__(self, dim, window_size, num_heads, qkv_bias=False, qk_scale=None, attn_drop=0.0, proj_drop=0.0):
        super().__init__()
        self.dim, self.window_size, self.num_heads = dim, tuple(window_size), num_heads
        head_dim = dim // num_heads
        if head_dim not in (32, 64):
            raise ValueError(f"clover_b200 attention kernels support head_dim 32 or 64, got {head_dim}")
        if qk_scale is not None and abs(qk_scale - head_dim ** -0.5) > 1e-12:
            self.scale = qk_scale
        else:
            self.scale = head_dim ** -0.5
        n_bias = (2 * window_size[0] - 1) * (2 * window_size[1] - 1) * (2 * window_size[2] - 1)
        self.relative_position_bias_table = nn.Parameter(torch.zeros(n_bias, num_heads))
        self.register_buffer("relative_position_index",
                             torch.from_numpy(tables.relative_position_index(tuple(window_size)).copy()))
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Dropout(proj_drop)
        trunc_normal_(self.relative_position_bias_table, std=0.02)

    def forward(self, x, mask=None):
        B_, N, C = x.shape
        region = None
        if mask is not None:
            region = getattr(mask, "regions", None)
            if region is None:
                raise TypeError("WindowAttention3D: pass the mask returned by clover_b200.swin.compute_mask")
        code, off = tables.rel_code(N, self.window_size)
        code = torch.from_numpy(code).to(x.device)
        x2 = x.reshape(B_ * N, C)
        x16 = x2 if x2.dtype == torch.bfloat16 else Fn.CastFn.apply(x2.contiguous(), torch.bfloat16)
        y = Fn.WindowAttnFn.apply(x16, B_, N, self.num_heads, code, off, region, self.scale,
                                  self.qkv.weight, self.qkv.bias, self.relative_position_bias_table,
                                  self.proj.weight, self.proj.bias)
        return y.view(B_, N, C)


class ShiftMask(torch.Tensor):
    """(nW, N, N) 0 / -100 tensor that also carries the O(N) region ids the kernels consume."""
    regions = None


def compute_mask(D, H, W, window_size, shift_size, device, dtype=torch.float32):
    """reference :548-562 (values identical); the returned tensor has a ``.regions`` int32 (nW, N)."""
    rid = tables.region_ids(D, H, W, tuple(window_size), tuple(shift_size))
    m = torch.from_numpy(tables.attn_mask_from_regions(rid)).to(device=device, dtype=dtype).as_subclass(ShiftMask)
    m.regions = torch.from_numpy(rid).to(device)
    return m


class DropPath(nn.Module):
    """timm.models.layers.DropPath (stochastic depth per sample), the class swin_transformer_3d.py:6,443 instantiates.
    Inside SwinTransformerBlock3D the factor keep / (1 - p) rides in the proj / fc2 GEMM epilogues; this forward is the
    stand-alone form for callers that use the module directly."""

    def __init__(self, drop_prob=0.0, scale_by_keep=True):
        super().__init__()
        self.drop_prob, self.scale_by_keep = float(drop_prob), scale_by_keep

    def forward(self, x):
        if self.drop_prob == 0.0 or not self.training:
            return x
        from . import rng
        s = rng.drop_path_scales(x.shape[0], self.drop_prob, x.device)
        if not self.scale_by_keep:
            s = s * (1.0 - self.drop_prob)
        return x * s.view((-1,) + (1,) * (x.dim() - 1)).to(x.dtype)

    def extra_repr(self):
        return f"drop_prob={self.drop_prob:0.3f}"


class SwinTransformerBlock3D(nn.Module):
    """reference :403-505."""

    def __init__(self, dim, num_heads, window_size=(2, 7, 7), shift_size=(0, 0, 0), mlp_ratio=4.0, qkv_bias=True,
                 qk_scale=None, drop=0.0, attn_drop=0.0, drop_path=0.0, act_layer=nn.GELU, norm_layer=nn.LayerNorm,
                 use_checkpoint=False):
        super().__init__()
        self.dim, self.num_heads = dim, num_heads
        self.window_size, self.shift_size = tuple(window_size), tuple(shift_size)
        self.mlp_ratio, self.use_checkpoint = mlp_ratio, use_checkpoint
        for s, w in zip(self.shift_size, self.window_size):
            assert 0 <= s < w, "shift_size must in 0-window_size"
        self.norm1 = norm_layer(dim)
        self.attn = WindowAttention3D(dim, window_size=self.window_size, num_heads=num_heads, qkv_bias=qkv_bias,
                                      qk_scale=qk_scale, attn_drop=attn_drop, proj_drop=drop)
        self.drop_path_rate = float(drop_path)
        self.drop_path = DropPath(drop_path) if drop_path > 0.0 else nn.Identity()       # :443
        self.norm2 = norm_layer(dim)
        self.mlp = Mlp(in_features=dim, hidden_features=int(dim * mlp_ratio), act_layer=act_layer, drop=drop)
        if drop > 0 or attn_drop > 0:
            raise NotImplementedError("clover_b200: drop / attn_drop > 0 are not supported (0 in every Clover config)")

    def forward_tokens(self, x, B, D, H, W, prev_dp=None, return_dp=False):
        """x fp32 [B*D*H*W, C] channels-last tokens -> same shape.  prev_dp: the MLP-branch DropPath factors of the block
        that produced x (its gradient copy is pre-scaled in this block's backward); return_dp: also return this block's."""
        dp = None
        if self.training and self.drop_path_rate > 0:      # two independent per-sample draws: attention branch, MLP branch
            from . import rng
            dp = (rng.drop_path_scales(B, self.drop_path_rate, x.device),      # a tuple: the same tensor OBJECTS tag the
                  rng.drop_path_scales(B, self.drop_path_rate, x.device))      # pre-scaled gradient copies in backward
        window, shift = get_window_size((D, H, W), self.window_size, self.shift_size)
        wg = ops.Window(B, D, H, W, window, shift)
        code, off, region = _device_tables((wg.Dp, wg.Hp, wg.Wp), window, shift, self.window_size, x.device)
        wg.w7 = _w7_spec((wg.Dp, wg.Hp, wg.Wp), window, shift, self.window_size, x.device)
        a, m = self.attn, self.mlp
        y = Fn.SwinBlockFn.apply(x, wg, self.num_heads, code, off, region, dp, prev_dp,
                                 self.norm1.weight, self.norm1.bias, a.qkv.weight, a.qkv.bias,
                                 a.relative_position_bias_table, a.proj.weight, a.proj.bias,
                                 self.norm2.weight, self.norm2.bias, m.fc1.weight, m.fc1.bias, m.fc2.weight, m.fc2.bias)
        return (y, dp[1] if dp is not None else None) if return_dp else y

    def forward(self, x, mask_matrix=None):
        """reference contract: x (B, D, H, W, C) -> same."""
        B, D, H, W, C = x.shape
        y = self.forward_tokens(x.reshape(-1, C).float().contiguous(), B, D, H, W)
        return y.view(B, D, H, W, C)


class PatchMerging(nn.Module):
    """reference :508-544."""

    def __init__(self, dim, norm_layer=nn.LayerNorm):
        super().__init__()
        self.dim = dim
        self.reduction = nn.Linear(4 * dim, 2 * dim, bias=False)
        self.norm = norm_layer(4 * dim)

    def forward_tokens(self, x, B, D, H, W, prev_dp=None):
        y = Fn.PatchMergeFn.apply(x, (B, D, H, W), self.norm.weight, self.norm.bias, self.reduction.weight, prev_dp)
        return y, (H + 1) // 2, (W + 1) // 2

    def forward(self, x):
        B, D, H, W, C = x.shape
        y, H2, W2 = self.forward_tokens(x.reshape(-1, C).float().contiguous(), B, D, H, W)
        return y.view(B, D, H2, W2, 2 * C)


class BasicLayer(nn.Module):
    """reference :565-646."""

    def __init__(self, dim, depth, num_heads, window_size=(1, 7, 7), mlp_ratio=4.0, qkv_bias=False, qk_scale=None,
                 drop=0.0, attn_drop=0.0, drop_path=0.0, norm_layer=nn.LayerNorm, downsample=None, use_checkpoint=False):
        super().__init__()
        self.window_size = tuple(window_size)
        self.shift_size = tuple(i // 2 for i in window_size)
        self.depth, self.use_checkpoint = depth, use_checkpoint
        self.blocks = nn.ModuleList([
            SwinTransformerBlock3D(dim=dim, num_heads=num_heads, window_size=window_size,
                                   shift_size=(0, 0, 0) if (i % 2 == 0) else self.shift_size, mlp_ratio=mlp_ratio,
                                   qkv_bias=qkv_bias, qk_scale=qk_scale, drop=drop, attn_drop=attn_drop,
                                   drop_path=drop_path[i] if isinstance(drop_path, list) else drop_path,
                                   norm_layer=norm_layer, use_checkpoint=use_checkpoint)
            for i in range(depth)])
        self.downsample = downsample(dim=dim, norm_layer=norm_layer) if downsample is not None else None

    def forward_tokens(self, x, B, D, H, W):
        prev = None                       # DropPath factors of the previous block's MLP branch (same layer only)
        for blk in self.blocks:
            x, prev = blk.forward_tokens(x, B, D, H, W, prev_dp=prev, return_dp=True)
        if self.downsample is not None:
            x, H, W = self.downsample.forward_tokens(x, B, D, H, W, prev_dp=prev)
        return x, H, W

    def forward(self, x):
        """reference contract: (B, C, D, H, W) -> (B, C', D, H', W')."""
        B, C, D, H, W = x.shape
        t = x.permute(0, 2, 3, 4, 1).reshape(-1, C).float().contiguous()
        y, H2, W2 = self.forward_tokens(t, B, D, H, W)
        return y.view(B, D, H2, W2, -1).permute(0, 4, 1, 2, 3)


class PatchEmbed3D(nn.Module):
    """reference :649-688."""

    def __init__(self, patch_size=(2, 4, 4), in_chans=3, embed_dim=96, norm_layer=None, stride=(2, 4, 4)):
        super().__init__()
        self.patch_size, self.in_chans, self.embed_dim = tuple(patch_size), in_chans, embed_dim
        if tuple(stride) != tuple(patch_size):
            raise NotImplementedError("clover_b200: PatchEmbed3D needs stride == patch_size (true for every Clover config)")
        self.proj = nn.Conv3d(in_chans, embed_dim, kernel_size=patch_size, stride=stride)
        self.norm = norm_layer(embed_dim) if norm_layer is not None else None
        self.input_norm = None            # (mean, 1/std) fp32 [Cin]: raw uint8 frames are normalised inside the patch gather

    def set_input_normalization(self, mean, std):
        """Accept raw uint8 clips and normalise them as (v - mean[c]) / std[c] inside the patch-gather kernel -- the
        reference's GPUNormalize module hook (utils/module_hooks.py:35-87) without its extra pass over the clip.
        mean / std: per-channel sequences (img_norm_cfg of configs/_base_/datasets_local/*.py); None switches it off."""
        if mean is None:
            self.input_norm = None
            return
        m = torch.as_tensor(mean, dtype=torch.float32).reshape(-1)
        s = torch.as_tensor(std, dtype=torch.float32).reshape(-1)
        if m.numel() != self.in_chans or s.numel() != self.in_chans:
            raise ValueError(f"set_input_normalization: need {self.in_chans} means and stds")
        self.input_norm = (m, 1.0 / s)

    def pair_supported(self, x, mask):
        """Whether forward_tokens(..., pair=True) can share the patch gather / projection between a masked and a clean pass."""
        if mask is None or self.norm is None or not ops.lnr_supported(self.embed_dim):
            return False
        Hp, Wp = -(-x.shape[-2] // self.patch_size[1]), -(-x.shape[-1] // self.patch_size[2])
        return Hp % mask.shape[-2] == 0 and Wp % mask.shape[-1] == 0

    def forward_tokens(self, x, mask=None, mask_token=None, pair=False):
        """x fp32 (or uint8 after set_input_normalization) (B, Cin, F, H, W) -> (tokens fp32 [B*D*Hp*Wp, C], (B, D, Hp, Wp)).
        pair=True: tokens of [masked pass ; clean pass] of the same clips, (2B, D, Hp, Wp)."""
        if mask is not None and self.norm is None:
            raise NotImplementedError("clover_b200: mask-token blend requires patch_norm=True")
        B, _, Fr, H, W = x.shape
        pd, ph, pw = self.patch_size
        D, Hp, Wp = -(-Fr // pd), -(-H // ph), -(-W // pw)
        nw, nb = (self.norm.weight, self.norm.bias) if self.norm is not None else (None, None)
        norm = None
        if x.dtype == torch.uint8:
            if self.input_norm is None:
                raise TypeError("PatchEmbed3D: uint8 clips need set_input_normalization(mean, std) (GPUNormalize)")
            if self.input_norm[0].device != x.device:
                self.input_norm = tuple(t.to(x.device) for t in self.input_norm)
            norm = self.input_norm
        else:
            x = x.float()
        y = Fn.PatchEmbedFn.apply(x, self.proj.weight, self.proj.bias, nw, nb, mask, mask_token, self.patch_size, norm, pair)
        return y, ((2 * B if pair else B), D, Hp, Wp)

    def forward(self, x):
        y, (B, D, Hp, Wp) = self.forward_tokens(x)
        return y.view(B, D, Hp, Wp, self.embed_dim).permute(0, 4, 1, 2, 3)


class SwinTransformer3D(nn.Module):
    """reference :18-247; registered under the same name (clover_b200.registry)."""

    def __init__(self, pretrained=None, pretrained2d=True, patch_size=(2, 4, 4), stride=(2, 4, 4), in_chans=3,
                 embed_dim=96, depths=[2, 2, 6, 2], num_heads=[3, 6, 12, 24], window_size=(8, 7, 7), mlp_ratio=4.0,
                 qkv_bias=True, qk_scale=None, drop_rate=0.0, attn_drop_rate=0.0, drop_path_rate=0.1,
                 norm_layer=nn.LayerNorm, patch_norm=True, frozen_stages=-1, use_checkpoint=False, mask_token=False):
        super().__init__()
        self.pretrained, self.pretrained2d = pretrained, pretrained2d
        self.num_layers, self.embed_dim = len(depths), embed_dim
        self.patch_norm, self.frozen_stages = patch_norm, frozen_stages
        self.window_size, self.patch_size = tuple(window_size), tuple(patch_size)
        self.fp16_enabled = False
        self.patch_embed = PatchEmbed3D(patch_size=patch_size, in_chans=in_chans, embed_dim=embed_dim, stride=stride,
                                        norm_layer=norm_layer if patch_norm else None)
        if drop_rate > 0:
            raise NotImplementedError("clover_b200: drop_rate > 0 is not supported (0 in every Clover config)")
        self.pos_drop = nn.Dropout(p=drop_rate)
        dpr = [x.item() for x in torch.linspace(0, drop_path_rate, sum(depths), device="cpu")]
        self.layers = nn.ModuleList()
        for i in range(self.num_layers):
            self.layers.append(BasicLayer(
                dim=int(embed_dim * 2 ** i), depth=depths[i], num_heads=num_heads[i], window_size=window_size,
                mlp_ratio=mlp_ratio, qkv_bias=qkv_bias, qk_scale=qk_scale, drop=drop_rate, attn_drop=attn_drop_rate,
                drop_path=dpr[sum(depths[:i]):sum(depths[:i + 1])], norm_layer=norm_layer,
                downsample=PatchMerging if i < self.num_layers - 1 else None, use_checkpoint=use_checkpoint))
        self.num_features = int(embed_dim * 2 ** (self.num_layers - 1))
        self.norm = norm_layer(self.num_features)
        if mask_token:
            self.mask_token = nn.Parameter(torch.zeros(1, self.embed_dim, 1, 1, 1))
            trunc_normal_(self.mask_token, mean=0.0, std=0.02)
        self._freeze_stages()

    def set_input_normalization(self, mean, std):
        """Raw uint8 clips in, normalisation folded into the patch gather (see PatchEmbed3D.set_input_normalization)."""
        self.patch_embed.set_input_normalization(mean, std)

    def _freeze_stages(self):
        if self.frozen_stages >= 0:
            self.patch_embed.eval()
            for p in self.patch_embed.parameters():
                p.requires_grad = False
        if self.frozen_stages >= 1:
            self.pos_drop.eval()
            for i in range(self.frozen_stages):
                m = self.layers[i]
                m.eval()
                for p in m.parameters():
                    p.requires_grad = False

    # ---- weights ---------------------------------------------------------------------------------
    def init_weights(self, pretrained=None):
        def _init(m):
            if isinstance(m, nn.Linear):
                trunc_normal_(m.weight, std=0.02)
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)
            elif isinstance(m, nn.LayerNorm):
                nn.init.constant_(m.bias, 0)
                nn.init.constant_(m.weight, 1.0)
        if pretrained:
            self.pretrained = pretrained
        if isinstance(self.pretrained, str):
            self.apply(_init)
            if self.pretrained2d:
                self.inflate_weights()
            else:
                ck = torch.load(self.pretrained, map_location="cpu")
                self.load_state_dict(ck.get("state_dict", ck), strict=False)
        elif self.pretrained is None:
            self.apply(_init)
        else:
            raise TypeError("pretrained must be a str or None")

    def inflate_weights(self, logger=None):
        """2-D Swin checkpoint -> 3-D (reference :130-181): repeat the patch-embed kernel over time / pd,
        bicubic-resize and tile the relative-position bias tables."""
        ck = torch.load(self.pretrained, map_location="cpu")
        sd = ck["state_dict"]
        for k in [k for k in sd if "relative_position_index" in k or "attn_mask" in k]:
            del sd[k]
        pd = self.patch_size[0]
        sd["patch_embed.proj.weight"] = sd["patch_embed.proj.weight"].unsqueeze(2).repeat(1, 1, pd, 1, 1) / pd
        own = self.state_dict()
        wd, wh, ww = self.window_size
        for k in [k for k in sd if "relative_position_bias_table" in k]:
            t = sd[k]
            L1, nH1 = t.shape
            nH2 = own[k].shape[1]
            L2 = (2 * wh - 1) * (2 * ww - 1)
            if nH1 == nH2:
                if L1 != L2:
                    S1 = int(L1 ** 0.5)
                    t = torch.nn.functional.interpolate(t.permute(1, 0).view(1, nH1, S1, S1), size=(2 * wh - 1, 2 * ww - 1),
                                                        mode="bicubic").view(nH2, L2).permute(1, 0)
                sd[k] = t.repeat(2 * wd - 1, 1)
        self.load_state_dict(sd, strict=False)

    # ---- forward ---------------------------------------------------------------------------------
    def forward_tokens(self, x, mask=None, pair=False):
        """x (B, 3, F, H, W) -> (tokens fp32 [B*D*h*w, C_out] channels-last, (B, D, h, w)).  pair=True (with a mask): the masked
        and the clean pass of the same clips in one doubled batch [masked ; clean], sharing the patch embedding."""
        tok = getattr(self, "mask_token", None) if mask is not None else None
        if mask is not None and tok is None:
            raise ValueError("SwinTransformer3D: mask given but the backbone was built with mask_token=False")
        t, (B, D, H, W) = self.patch_embed.forward_tokens(x, mask, tok, pair=pair)
        if self.training:                  # all DropPath factors of this pass in one draw (two per block, in block order)
            from . import rng
            ps = [blk.drop_path_rate for layer in self.layers for blk in layer.blocks for _ in range(2) if blk.drop_path_rate > 0]
            rng.predraw_drop_path(B, ps, t.device)
        for layer in self.layers:
            t, H, W = layer.forward_tokens(t, B, D, H, W)
        t = Fn.layer_norm(t, self.norm.weight, self.norm.bias, 1e-5, out_fp32=True)
        return t, (B, D, H, W)

    def forward(self, x, mask=None):
        """reference :217-242: returns (B, C, D, h, w) [and w (B,1,D,56,56) when mask is given]."""
        t, (B, D, H, W) = self.forward_tokens(x, mask)
        out = t.view(B, D, H, W, self.num_features).permute(0, 4, 1, 2, 3)
        if mask is not None:
            return out, self.mask_weight(mask, x.shape)
        return out

    def mask_weight(self, mask, x_shape):
        """w of reference :227-229 (nearest up-sampling of the 7x7 mask, broadcast over time)."""
        B, _, Fr, H, W = x_shape
        pd, ph, pw = self.patch_size
        D, Hp, Wp = -(-Fr // pd), -(-H // ph), -(-W // pw)
        mh, mw = mask.shape[-2:]
        yy = torch.arange(Hp, device=mask.device) // (Hp // mh)
        xx = torch.arange(Wp, device=mask.device) // (Wp // mw)
        w2d = mask[:, 0][:, yy][:, :, xx].to(torch.float32)
        return w2d[:, None, None].expand(B, 1, D, Hp, Wp)

    def train(self, mode=True):
        super().train(mode)
        self._freeze_stages()
        return self


class GPUNormalize:
    """Drop-in for the reference's module hook of the same name (mmaction/utils/module_hooks.py:35-87;
    ``module_hooks=[dict(type='GPUNormalize', hooked_module='backbone', hook_pos='forward_pre', input_format='NCTHW',
    mean=[...], std=[...])]``).  The reference hook converts the uint8 clip to float and normalises it in a separate pass;
    this one hands mean / std to the backbone's patch gather and lets the uint8 clip through untouched."""

    def __init__(self, input_format, mean, std):
        if input_format != "NCTHW":
            raise ValueError(f"clover_b200.GPUNormalize supports input_format='NCTHW' (video clips), got {input_format}")
        self.input_format, self.mean, self.std = input_format, list(mean), list(std)

    def hook_func(self):
        def normalize_hook(module, inputs):
            x = inputs[0]
            assert x.dtype == torch.uint8, f"The previous augmentation should use uint8 data type, but get {x.dtype}"
            if not hasattr(module, "set_input_normalization"):
                raise TypeError("GPUNormalize: the hooked module must be a clover_b200 SwinTransformer3D / PatchEmbed3D")
            module.set_input_normalization(self.mean, self.std)
            return inputs
        return normalize_hook
