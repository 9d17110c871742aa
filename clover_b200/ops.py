"""Tensor-level wrappers over the C ABI (one Python function per kernel family).

PyTorch is used for device memory and streams only: every function checks shapes / dtypes /
contiguity, takes ``tensor.data_ptr()`` and calls into libclover_b200.so on the current CUDA stream.
CUDA tensors are mandatory -- there is no CPU path.
"""
import ctypes as C
import os

import torch

from . import _lib
from ._lib import AttnDesc, AttnW7Desc, GemmEpilogue, LnBwd, LnDesc, LnrBwd, LnrDesc, RowsAffine, WindowGeom

BF16 = torch.bfloat16
F32 = torch.float32


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("clover_b200 ops need CUDA tensors (no CPU fallback exists)")


def _is_bf16(t):
    if t.dtype == BF16:
        return 1
    if t.dtype == F32:
        return 0
    raise TypeError(f"expected bf16 or fp32 tensor, got {t.dtype}")


def _rowmajor2d(t, name):
    if t.dim() != 2 or t.stride(1) != 1:
        raise ValueError(f"{name}: expected a 2-D tensor with unit inner stride, got shape {tuple(t.shape)} strides {t.stride()}")
    return t.stride(0) if t.shape[0] > 1 else max(t.stride(0), t.shape[1])


# ------------------------------------------------------------------------------------------------
# optional per-launch CUDA-event timing (bench.py's roofline leg); off by default
# ------------------------------------------------------------------------------------------------
_PROF = None
PROFILE_SHAPES = os.environ.get("CLOVER_B200_PROFILE_SHAPES", "0") == "1"       # bench / tools: family names carry the shapes
# Kernel selection is fixed (no environment switches): (., 7, 7) windows with tokens in (d, h, w) order run on the
# specialised kernels of attention_w7.cu, other head_dim-32 windows on the generic tcgen05 kernel of attention_tc.cu.
# Tests flip these module attributes to compare the kernels against each other.
USE_TC_ATTENTION = True
USE_W7_ATTENTION = True
USE_TC64_ATTENTION = True      # head_dim-64 BERT / fusion attention on tcgen05 (attention_h64.cu); False: the mma.sync kernels


def set_tunable(name, value):
    """Experiment knobs of the library for tools / tests (clv_set_tunable); -1 restores the default."""
    _lib.check(_lib.load().clv_set_tunable(name.encode(), int(value)), "clv_set_tunable")


def profile_begin():
    global _PROF
    _PROF = []


def profile_end():
    """Returns [(family, algorithmic_flops, algorithmic_bytes, milliseconds)] after synchronising."""
    return profile_resolve(profile_detach())


def profile_detach():
    """Stop recording WITHOUT synchronising; hand the raw event list to profile_resolve() later (bench.py times only the
    first step of its timed region this way, so the per-launch events do not tax the other steps)."""
    global _PROF
    rec, _PROF = _PROF or [], None
    return rec


def profile_resolve(rec):
    torch.cuda.synchronize()
    return [(fam, fl, by, e0.elapsed_time(e1)) for fam, fl, by, e0, e1 in rec]


def _prof_open():
    if _PROF is None:
        return None
    e0 = torch.cuda.Event(enable_timing=True)
    e0.record()
    return e0


def _prof_close(e0, family, flops, nbytes):
    if e0 is not None:
        e1 = torch.cuda.Event(enable_timing=True)
        e1.record()
        _PROF.append((family, flops, nbytes, e0, e1))


def _tb(*ts):
    """bytes of the given tensors (None entries skipped)"""
    return float(sum(t.numel() * t.element_size() for t in ts if t is not None))


def _profiled(family, tag=None, nbytes=None):
    """nbytes(*args, **kwargs) -> ALGORITHMIC bytes of the launch (inputs read once + outputs written once): the numerator of
    the HBM roofline fraction bench.py reports for the memory-bound families."""
    def deco(fn):
        def wrapped(*a, **k):
            if _PROF is None:
                return fn(*a, **k)
            e0 = _prof_open()
            out = fn(*a, **k)
            name = family
            if PROFILE_SHAPES and tag is not None:
                name = family + " " + tag(*a, **k)
            _prof_close(e0, name, 0.0, nbytes(*a, **k) if nbytes is not None else 0.0)
            return out
        wrapped.__name__ = fn.__name__
        wrapped.__doc__ = fn.__doc__
        return wrapped
    return deco


class Window:
    """Clamped window geometry of one Swin block call (get_window_size, swin_transformer_3d.py:302-315)."""

    def __init__(self, B, D, H, W, window, shift):
        self.B, self.D, self.H, self.W = B, D, H, W
        self.window, self.shift = tuple(window), tuple(shift)
        self.c = WindowGeom(B, D, H, W, *self.window, *self.shift)
        self.Dp = -(-D // window[0]) * window[0]
        self.Hp = -(-H // window[1]) * window[1]
        self.Wp = -(-W // window[2]) * window[2]
        self.N = window[0] * window[1] * window[2]
        self.nwin = (self.Dp // window[0]) * (self.Hp // window[1]) * (self.Wp // window[2])
        self.rows = B * self.nwin * self.N            # window-order rows (incl. zero padding)
        self.tokens = B * D * H * W
        self.padded = (self.Dp, self.Hp, self.Wp) != (D, H, W)
        self.shifted = any(s > 0 for s in self.shift)
        self.w7 = None                                # ops.W7Spec when the specialised (., 7, 7) attention kernels apply

    def ref(self):
        return C.pointer(self.c)

    def row_map(self, device):
        """int32 [D*H*W] device tensor: window-order row (within a clip) of every spatial token (unpadded frames)."""
        if self.padded:
            raise ValueError("row_map: the frame is padded to window multiples; use the closed-form kernels")
        key = (self.D, self.H, self.W, self.window, self.shift, str(device))
        t = _ROW_MAPS.get(key)
        if t is None:
            from . import tables
            t = torch.from_numpy(tables.window_row_map(self.D, self.H, self.W, self.window, self.shift)).to(device)
            _ROW_MAPS[key] = t
        return t


_ROW_MAPS = {}


# ------------------------------------------------------------------------------------------------
def gemm(a, b, out, *, a_t=False, b_t=False, bias=None, residual=None, act=None, out_pre=None, gelu_pre=None,
         scale_cols=0, scale=1.0, window=None, k_splits=1, accumulate=False, row_scale=None, row_scale_rows=0, rowsum=None):
    """out = epilogue(A @ B^T).  a: [M,K] (or [K,M] if a_t), b: [N,K] (or [K,N] if b_t), bf16.
    See clv_gemm_bf16 in include/clover_b200.h."""
    _need_cuda(a, b, out)
    if a.dtype != BF16 or b.dtype != BF16:
        raise TypeError("gemm operands must be bf16")
    lda, ldb = _rowmajor2d(a, "A"), _rowmajor2d(b, "B")
    M, K = (a.shape[1], a.shape[0]) if a_t else (a.shape[0], a.shape[1])
    N, Kb = (b.shape[1], b.shape[0]) if b_t else (b.shape[0], b.shape[1])
    if K != Kb:
        raise ValueError(f"gemm: K mismatch {K} vs {Kb}")
    ldo = _rowmajor2d(out, "out")
    if out.shape[1] != N or (window is None and out.shape[0] != M):
        raise ValueError(f"gemm: out shape {tuple(out.shape)} does not match M={M} N={N}")
    e = GemmEpilogue()
    e.bias = _ptr(bias)
    if bias is not None and (bias.dtype != F32 or bias.numel() != N):
        raise ValueError("gemm: bias must be fp32 [N]")
    if residual is not None:
        e.residual = _ptr(residual)
        e.residual_is_bf16 = _is_bf16(residual)
        e.ld_residual = _rowmajor2d(residual, "residual")
    e.out, e.out_is_bf16, e.ld_out = _ptr(out), _is_bf16(out), ldo
    if out_pre is not None:
        e.out_pre, e.ld_pre = _ptr(out_pre), _rowmajor2d(out_pre, "out_pre")
    if gelu_pre is not None:
        e.gelu_pre, e.ld_gelu_pre = _ptr(gelu_pre), _rowmajor2d(gelu_pre, "gelu_pre")
    e.act = 1 if act == "gelu" else 0
    e.scale_cols, e.scale = int(scale_cols), float(scale)
    if window is not None:
        e.window = window.ref()
    e.k_splits, e.accumulate = int(k_splits), int(bool(accumulate))
    if row_scale is not None:
        if row_scale.dtype != F32 or not row_scale.is_contiguous() or row_scale_rows <= 0 or \
                row_scale.numel() * row_scale_rows < M:
            raise ValueError("gemm: row_scale must be contiguous fp32 with numel * row_scale_rows >= M")
        _need_cuda(row_scale)
        e.row_scale, e.row_scale_rows = _ptr(row_scale), int(row_scale_rows)
    if rowsum is not None:
        if rowsum.dtype != F32 or not rowsum.is_contiguous() or rowsum.numel() != M:
            raise ValueError("gemm: rowsum must be contiguous fp32 [M]")
        _need_cuda(rowsum)
        e.rowsum = _ptr(rowsum)
    lib = _lib.load()
    ev = _prof_open()
    _lib.check(lib.clv_gemm_bf16(_ptr(a), lda, int(a_t), _ptr(b), ldb, int(b_t), M, N, K, C.byref(e), _stream()), "clv_gemm_bf16")
    _prof_close(ev, "gemm" if not PROFILE_SHAPES else f"gemm M={M} N={N} K={K} at={int(a_t)} bt={int(b_t)} ks={int(k_splits)} "
                f"ep={'b' if bias is not None else ''}{'g' if act else ''}{'r' if residual is not None else ''}"
                f"{'w' if window is not None else ''}{'p' if gelu_pre is not None else ''}{'o16' if out.dtype == BF16 else 'o32'}",
                2.0 * M * N * K, 2.0 * (M * K + N * K) + out.element_size() * M * N)
    return out


def wgrad_splits(M, N, K):
    """Split-K factor for a weight-gradient GEMM (few output tiles, very long K): one wave of CTAs (148 SMs, one resident CTA
    each) over the 128 x 256 tiles the kernel takes when N fills them, 128 x 128 otherwise -- two waves were 18 % slower on
    the N = 128 stage-1 gradients (tools/gemm_probe.py).  Splitting costs a clear of the output and fp32 reductions in L2
    instead of plain stores: on the BERT gradients (K = 4096 caption tokens, >= 54 tiles) one pass is faster than two halves,
    so every split keeps at least 4096 of K (2048 when the tiles fill a quarter of the SMs or less)."""
    bn = 256 if N % 256 == 0 else 128
    tiles = -(-M // 128) * -(-N // bn)
    want = max(1, 148 // tiles)
    kmin = 2048 if tiles <= 36 else 4096          # a quarter wave or less: two halves of K = 4096 still beat one pass
    return max(1, min(want, K // kmin))


# ------------------------------------------------------------------------------------------------
def _ln_desc(x, gamma, beta, eps, rows, Cn, mean, rstd, *, window=None, merge=None, add0=None, add1=None, add2=None,
             group=None, blend=None, row_index=None):
    d = LnDesc()
    d.x, d.x_is_bf16 = _ptr(x), _is_bf16(x)
    d.ld_x = x.stride(-2) if x.dim() >= 2 else Cn
    d.gamma, d.beta, d.eps = _ptr(gamma), _ptr(beta), float(eps)
    d.mean, d.rstd = _ptr(mean), _ptr(rstd)
    d.rows, d.C = int(rows), int(Cn)
    if window is not None:
        d.window = window.ref()
    if merge is not None:
        d.merge_B, d.merge_D, d.merge_H, d.merge_W, d.merge_C = merge
    if add0 is not None:
        d.add0 = _ptr(add0)
    if add1 is not None:
        d.add1, d.div1, d.mod1 = _ptr(add1[0]), int(add1[1]), int(add1[2])
    if add2 is not None:
        d.add2, d.div2, d.mod2 = _ptr(add2[0]), int(add2[1]), int(add2[2])
    if group is not None:
        d.group_rows, d.group_stride, d.row_offset = group
    if blend is not None:
        mask, token, (bD, bH, bW) = blend
        d.blend_mask, d.blend_token = _ptr(mask), _ptr(token)
        d.blend_D, d.blend_H, d.blend_W = bD, bH, bW
        d.blend_mh, d.blend_mw = mask.shape[-2], mask.shape[-1]
    if row_index is not None:
        d.row_index = _ptr(row_index)
    return d


def _ln_tag(x, gamma, *a, **k):
    kinds = [n for n in ("window", "merge", "blend", "row_index", "add1", "group", "dres", "dx_copy", "dx_bf16", "row_map") if k.get(n) is not None]
    return f"generic rows={k.get('rows') or x.shape[0]} C={gamma.numel()} x={str(x.dtype)[-4:]} {'+'.join(kinds)}"


def _ln_fwd_bytes(x, gamma, beta, eps, y, *, rows=None, **kw):
    r = y.shape[0] if rows is None else rows
    return float(r * gamma.numel() * (x.element_size() + y.element_size()))


def _ln_bwd_bytes(x, gamma, beta, eps, mean, rstd, dy, *, rows, dx=None, dres=None, dx_copy=None, **kw):
    n = rows * gamma.numel()
    return float(n * (x.element_size() + dy.element_size() + (4 if dx is not None else 0) + (4 if dres is not None else 0)
                      + (dx_copy.element_size() if dx_copy is not None else 0)))


@_profiled("ln_fwd", _ln_tag, _ln_fwd_bytes)
def layernorm_fwd(x, gamma, beta, eps, y, *, rows=None, mean=None, rstd=None, **kw):
    """y = LN(gather(x) + adds) (* blend).  x: [..., C] fp32/bf16 2-D; y: 2-D bf16/fp32."""
    _need_cuda(x, gamma, beta, y)
    Cn = gamma.numel()
    rows = y.shape[0] if rows is None else rows
    d = _ln_desc(x, gamma, beta, eps, rows, Cn, mean, rstd, **kw)
    _lib.check(_lib.load().clv_layernorm_fwd(C.byref(d), _ptr(y), _is_bf16(y), y.stride(0), _stream()), "clv_layernorm_fwd")
    return y


@_profiled("ln_bwd", _ln_tag, _ln_bwd_bytes)
def layernorm_bwd(x, gamma, beta, eps, mean, rstd, dy, *, rows, dx=None, dres=None, dx_copy=None, copy_window=None,
                  dgamma=None, dbeta=None, dtoken=None, dx_dense=False, dxsum=None, copy_scale=None, copy_scale_rows=0, **kw):
    _need_cuda(x, gamma, dy)
    Cn = gamma.numel()
    d = _ln_desc(x, gamma, beta, eps, rows, Cn, mean, rstd, **kw)
    b = LnBwd()
    b.dy, b.dy_is_bf16, b.ld_dy = _ptr(dy), _is_bf16(dy), dy.stride(0)
    if dx is not None:
        if dx.dtype != F32:
            raise TypeError("layernorm_bwd: dx must be fp32")
        b.dx, b.ld_dx = _ptr(dx), dx.stride(0)
    if dres is not None:
        b.dres, b.ld_dres = _ptr(dres), dres.stride(0)
    if dx_copy is not None:
        b.dx_copy, b.dx_copy_is_bf16, b.ld_copy = _ptr(dx_copy), _is_bf16(dx_copy), dx_copy.stride(0)
    if copy_window is not None:
        b.copy_window = copy_window.ref()
    b.dgamma, b.dbeta, b.dtoken = _ptr(dgamma), _ptr(dbeta), _ptr(dtoken)
    b.dx_dense = int(dx_dense)
    if dxsum is not None:
        if dxsum.dtype != F32 or not dxsum.is_contiguous():
            raise TypeError("layernorm_bwd: dxsum must be contiguous fp32")
        _need_cuda(dxsum)
        b.dxsum = _ptr(dxsum)
    if copy_scale is not None:
        if copy_scale.dtype != F32 or not copy_scale.is_contiguous() or copy_scale_rows <= 0:
            raise ValueError("layernorm_bwd: copy_scale must be contiguous fp32 with copy_scale_rows > 0")
        _need_cuda(copy_scale)
        b.copy_scale, b.copy_scale_rows = _ptr(copy_scale), int(copy_scale_rows)
    _lib.check(_lib.load().clv_layernorm_bwd(C.byref(d), C.byref(b), _stream()), "clv_layernorm_bwd")


def ln_merge_fast(Cn):
    """Source widths with a dedicated PatchMerging LayerNorm kernel (bf16 copy / column sums of dx supported)."""
    return int(Cn) in (128, 256, 512)


def lnr_supported(Cn):
    return bool(_lib.load().clv_lnr_supported(int(Cn)))


def _lnr_desc(x, gamma, beta, eps, mean, rstd, row_map):
    if x.dim() != 2 or not x.is_contiguous():
        raise ValueError("lnr: x must be a contiguous 2-D tensor")
    d = LnrDesc()
    d.x, d.x_is_bf16 = _ptr(x), _is_bf16(x)
    d.gamma, d.beta, d.eps = _ptr(gamma), _ptr(beta), float(eps)
    d.mean, d.rstd = _ptr(mean), _ptr(rstd)
    d.rows, d.C = x.shape[0], x.shape[1]
    if row_map is not None:
        if row_map.dtype != torch.int32 or not row_map.is_contiguous():
            raise TypeError("lnr: row_map must be contiguous int32")
        d.row_map, d.map_period = _ptr(row_map), row_map.numel()
    return d


def _lnr_tag(x, gamma, *a, **k):
    kinds = [n for n in ("dres", "dx", "dx_bf16", "row_map", "blend") if k.get(n) is not None]
    return f"rows={x.shape[0]} C={x.shape[1]} x={str(x.dtype)[-4:]} {'+'.join(kinds)}"


def _lnr_blend(d, x, blend):
    if blend is None:
        return
    w, token = blend                        # fp32 [rows] weights, fp32 [C] token
    if w.dtype != F32 or token.dtype != F32 or w.numel() != x.shape[0] or token.numel() != x.shape[1] or \
            not w.is_contiguous() or not token.is_contiguous():
        raise ValueError("lnr: blend = (fp32 [rows] weights, fp32 [C] token), both contiguous")
    _need_cuda(w, token)
    d.row_blend, d.blend_token = _ptr(w), _ptr(token)


@_profiled("ln_fwd", _lnr_tag, lambda x, gamma, beta, eps, y, **k: _tb(x, y))
def lnr_fwd(x, gamma, beta, eps, y, *, mean=None, rstd=None, row_map=None, y_mapped=False, blend=None):
    """y[m(s)] = LN(x[s]) on dense rows (clv_lnr_fwd); m = identity unless y_mapped.  blend = (w [rows], token [C]):
    y = LN(x) (1 - w) + token w (mask-token blend after the patch-embed LN)."""
    _need_cuda(x, gamma, beta, y)
    if y.shape != x.shape or not y.is_contiguous():
        raise ValueError("lnr_fwd: y must be contiguous with x's shape")
    d = _lnr_desc(x, gamma, beta, eps, mean, rstd, row_map)
    _lnr_blend(d, x, blend)
    _lib.check(_lib.load().clv_lnr_fwd(C.byref(d), _ptr(y), _is_bf16(y), int(y_mapped), _stream()), "clv_lnr_fwd")
    return y


@_profiled("ln_bwd", _lnr_tag, lambda x, gamma, beta, eps, mean, rstd, dy, *, dx=None, dres=None, dx_bf16=None, **k:
           _tb(x, dy, dx, dres if dres is not dx else None, dx_bf16))
def lnr_bwd(x, gamma, beta, eps, mean, rstd, dy, *, dx=None, dres=None, dx_bf16=None, row_map=None, dy_mapped=False,
            dx_bf16_mapped=False, dgamma=None, dbeta=None, dxsum=None, copy_scale=None, copy_scale_rows=0, blend=None,
            dtoken=None):
    _need_cuda(x, gamma, dy)
    for t, n in ((dy, "dy"), (dx, "dx"), (dres, "dres"), (dx_bf16, "dx_bf16")):
        if t is not None and (t.shape != x.shape or not t.is_contiguous()):
            raise ValueError(f"lnr_bwd: {n} must be contiguous with x's shape")
    if (dx is not None and dx.dtype != F32) or (dres is not None and dres.dtype != F32) or (dx_bf16 is not None and dx_bf16.dtype != BF16):
        raise TypeError("lnr_bwd: dx / dres fp32 and dx_bf16 bf16 required")
    d = _lnr_desc(x, gamma, beta, eps, mean, rstd, row_map)
    _lnr_blend(d, x, blend)
    b = LnrBwd()
    if blend is not None:
        if dtoken is None or dtoken.dtype != F32 or not dtoken.is_contiguous() or dtoken.numel() != x.shape[1]:
            raise ValueError("lnr_bwd: the blend needs a contiguous fp32 [C] dtoken accumulator")
        _need_cuda(dtoken)
        b.dtoken = _ptr(dtoken)
    b.dy, b.dy_is_bf16, b.dy_mapped = _ptr(dy), _is_bf16(dy), int(dy_mapped)
    b.dres, b.dx, b.dx_bf16, b.dx_bf16_mapped = _ptr(dres), _ptr(dx), _ptr(dx_bf16), int(dx_bf16_mapped)
    b.dgamma, b.dbeta, b.dxsum = _ptr(dgamma), _ptr(dbeta), _ptr(dxsum)
    if copy_scale is not None:
        if copy_scale.dtype != F32 or not copy_scale.is_contiguous() or copy_scale_rows <= 0 or \
                copy_scale.numel() * copy_scale_rows < x.shape[0]:
            raise ValueError("lnr_bwd: copy_scale must be contiguous fp32 with numel * copy_scale_rows >= rows")
        _need_cuda(copy_scale)
        b.copy_scale, b.copy_scale_rows = _ptr(copy_scale), int(copy_scale_rows)
    _lib.check(_lib.load().clv_lnr_bwd(C.byref(d), C.byref(b), _stream()), "clv_lnr_bwd")


# ------------------------------------------------------------------------------------------------
def _attn_desc(batch, seq, heads, hd, bias_table=None, rel_code=None, code_off=0, region=None, key_mask=None, drop=None):
    d = AttnDesc()
    d.batch, d.seq, d.heads, d.head_dim = batch, seq, heads, hd
    if bias_table is not None:
        d.bias_table, d.table_len = _ptr(bias_table), bias_table.shape[0]
        d.rel_code, d.code_off = _ptr(rel_code), int(code_off)
    if region is not None:
        d.region, d.nwin = _ptr(region), region.shape[0]
    if key_mask is not None:
        d.key_mask = _ptr(key_mask)
    if drop is not None:                      # (p, seed, offset) of the attention-probability dropout stream
        d.drop_p, d.drop_seed, d.drop_offset = float(drop[0]), int(drop[1]), int(drop[2])
    return d


class W7Spec:
    """Static description of a (wd, 7, 7) window-attention call for the specialised kernels: temporal extent of the
    clamped window, the configured temporal window (bias-table size) and the bf16 K-extension tables that encode the
    shift mask (None for unshifted blocks)."""

    def __init__(self, wd, cfg_wd, q_ext=None, k_ext=None):
        self.wd, self.cfg_wd, self.q_ext, self.k_ext = int(wd), int(cfg_wd), q_ext, k_ext
        self.nwin = q_ext.shape[0] if q_ext is not None else 0

    def desc(self, batch, heads, bias_table):
        d = AttnW7Desc()
        d.batch, d.heads, d.wd = batch, heads, self.wd
        d.bias_table, d.table_len, d.cfg_wd = _ptr(bias_table), bias_table.shape[0], self.cfg_wd
        d.q_ext, d.k_ext, d.nwin = _ptr(self.q_ext), _ptr(self.k_ext), self.nwin
        return d

    def fwd_ok(self):
        return USE_W7_ATTENTION and self.wd % 2 == 0 and 2 <= self.wd <= 8

    def bwd_ok(self):
        return USE_W7_ATTENTION and self.wd in (2, 4, 8)


def attention_fwd(qkv, batch, seq, heads, hd, out, lse, w7=None, **bias):
    _need_cuda(qkv, out, lse)
    if qkv.dtype != BF16 or not qkv.is_contiguous() or qkv.shape != (batch * seq, 3 * heads * hd):
        raise ValueError(f"attention_fwd: qkv must be contiguous bf16 [{batch * seq}, {3 * heads * hd}], got {tuple(qkv.shape)} {qkv.dtype}")
    d = _attn_desc(batch, seq, heads, hd, **bias)
    ev = _prof_open()
    lib = _lib.load()
    if w7 is not None and hd == 32 and w7.fwd_ok() and seq == 49 * w7.wd and bias.get("bias_table") is not None:
        if bias["bias_table"].dtype != F32 or not bias["bias_table"].is_contiguous():
            raise TypeError("attention_fwd: bias_table must be contiguous fp32")
        dw = w7.desc(batch, heads, bias["bias_table"])
        ws = torch.empty(lib.clv_attention_w7_fwd_workspace_bytes(C.byref(dw)), dtype=torch.uint8, device=qkv.device)
        _lib.check(lib.clv_attention_w7_fwd(C.byref(dw), _ptr(qkv), _ptr(out), _ptr(lse), _ptr(ws), _stream()),
                   "clv_attention_w7_fwd")
    elif hd == 32 and bias.get("key_mask") is None and bias.get("drop") is None and 33 <= seq <= 416 and USE_TC_ATTENTION:
        _lib.check(lib.clv_attention_fwd_tc(C.byref(d), _ptr(qkv), _ptr(out), _ptr(lse), _stream()), "clv_attention_fwd_tc")
    elif hd == 64 and USE_TC64_ATTENTION and bias.get("bias_table") is None and bias.get("region") is None and \
            lib.clv_attention_tc64_supported(seq):
        _lib.check(lib.clv_attention_fwd_tc64(C.byref(d), _ptr(qkv), _ptr(out), _ptr(lse), _stream()), "clv_attention_fwd_tc64")
    else:
        _lib.check(lib.clv_attention_fwd(C.byref(d), _ptr(qkv), _ptr(out), _ptr(lse), _stream()), "clv_attention_fwd")
    _prof_close(ev, ("attn_fwd_hd%d" % hd) + (f" b={batch} n={seq} h={heads}" if PROFILE_SHAPES else ""), 4.0 * batch * heads * seq * seq * hd, 2.0 * batch * seq * heads * hd * 4)
    return out


def attention_bwd(qkv, out, dout, lse, batch, seq, heads, hd, dqkv, q_scale, dbias_table=None, w7=None, **bias):
    _need_cuda(qkv, out, dout, lse, dqkv)
    for t, n in ((qkv, "qkv"), (out, "out"), (dout, "dout"), (dqkv, "dqkv")):
        if t.dtype != BF16 or not t.is_contiguous():
            raise ValueError(f"attention_bwd: {n} must be contiguous bf16")
    d = _attn_desc(batch, seq, heads, hd, **bias)
    lib = _lib.load()
    ev = _prof_open()
    if w7 is not None and hd == 32 and w7.bwd_ok() and seq == 49 * w7.wd and bias.get("bias_table") is not None:
        dw = w7.desc(batch, heads, bias["bias_table"])
        nbytes = lib.clv_attention_w7_bwd_workspace_bytes(C.byref(dw), int(dbias_table is not None))
        ws = torch.empty(nbytes, dtype=torch.uint8, device=qkv.device)
        _lib.check(lib.clv_attention_w7_bwd(C.byref(dw), _ptr(qkv), _ptr(out), _ptr(dout), _ptr(lse), _ptr(dqkv),
                                            float(q_scale), _ptr(dbias_table), _ptr(ws), _stream()), "clv_attention_w7_bwd")
    elif hd == 32 and bias.get("key_mask") is None and bias.get("drop") is None and 33 <= seq <= 224 and USE_TC_ATTENTION:
        nbytes = lib.clv_attention_bwd_tc_workspace_bytes(C.byref(d), int(dbias_table is not None))
        ws = torch.empty(nbytes, dtype=torch.uint8, device=qkv.device)
        _lib.check(lib.clv_attention_bwd_tc(C.byref(d), _ptr(qkv), _ptr(out), _ptr(dout), _ptr(lse), _ptr(dqkv),
                                            float(q_scale), _ptr(dbias_table), _ptr(ws), _stream()), "clv_attention_bwd_tc")
    elif hd == 64 and USE_TC64_ATTENTION and bias.get("bias_table") is None and bias.get("region") is None and \
            dbias_table is None and lib.clv_attention_tc64_supported(seq):
        ws = torch.empty(batch * heads * seq, dtype=F32, device=qkv.device)
        _lib.check(lib.clv_attention_bwd_tc64(C.byref(d), _ptr(qkv), _ptr(out), _ptr(dout), _ptr(lse), _ptr(dqkv),
                                              float(q_scale), _ptr(ws), _stream()), "clv_attention_bwd_tc64")
    else:
        ws = torch.empty(batch * heads * seq, dtype=F32, device=qkv.device)
        _lib.check(lib.clv_attention_bwd(C.byref(d), _ptr(qkv), _ptr(out), _ptr(dout), _ptr(lse), _ptr(dqkv),
                                         float(q_scale), _ptr(dbias_table), _ptr(ws), _stream()), "clv_attention_bwd")
    _prof_close(ev, ("attn_bwd_hd%d" % hd) + (f" b={batch} n={seq} h={heads}" if PROFILE_SHAPES else ""), 8.0 * batch * heads * seq * seq * hd, 2.0 * batch * seq * heads * hd * 8)
    return dqkv


def attention_probs_mean(qkv, batch, seq, heads, hd, **bias):
    """fp32 [batch, seq, seq]: head-mean of the attention probabilities (evaluation; finetune.py:192)."""
    _need_cuda(qkv)
    if qkv.dtype != BF16 or not qkv.is_contiguous() or qkv.shape != (batch * seq, 3 * heads * hd):
        raise ValueError("attention_probs_mean: qkv must be contiguous bf16 [batch*seq, 3*heads*hd]")
    d = _attn_desc(batch, seq, heads, hd, **bias)
    out = torch.empty(batch, seq, seq, dtype=F32, device=qkv.device)
    _lib.check(_lib.load().clv_attention_probs_mean(C.byref(d), _ptr(qkv), _ptr(out), _stream()), "clv_attention_probs_mean")
    return out


# ------------------------------------------------------------------------------------------------
@_profiled("dropout", None, lambda x, y, p, seed, offset, residual=None: _tb(x, y, residual))
def dropout(x, y, p, seed, offset, residual=None):
    """y = residual + x * keep / (1 - p) with keep drawn from stream (seed, offset + i); see clv_dropout."""
    _need_cuda(x, y, residual)
    n = x.numel()
    if y.numel() != n or not x.is_contiguous() or not y.is_contiguous() or (residual is not None and (
            residual.numel() != n or not residual.is_contiguous())):
        raise ValueError("dropout: x, y (and residual) must be contiguous with the same number of elements")
    _lib.check(_lib.load().clv_dropout(_ptr(x), _is_bf16(x), _ptr(residual), _is_bf16(residual) if residual is not None else 0,
                                       _ptr(y), _is_bf16(y), n, float(p), int(seed), int(offset), _stream()), "clv_dropout")
    return y


def keep_mask(n, p, seed, offset, device):
    """uint8 [n]: the keep decisions of stream elements offset .. offset + n - 1."""
    out = torch.empty(n, dtype=torch.uint8, device=device)
    _need_cuda(out)
    _lib.check(_lib.load().clv_keep_mask(_ptr(out), n, float(p), int(seed), int(offset), _stream()), "clv_keep_mask")
    return out


@_profiled("rows_scale", None, lambda x, y, scale, rows_per_group: _tb(x, y))
def rows_scale(x, y, scale, rows_per_group):
    """y[r, :] = x[r, :] * scale[r // rows_per_group]  (DropPath factor on a gradient)."""
    _need_cuda(x, y, scale)
    rows, Cn = x.shape
    if y.shape != x.shape or not x.is_contiguous() or not y.is_contiguous() or scale.dtype != F32 or \
            scale.numel() * rows_per_group < rows:
        raise ValueError("rows_scale: bad arguments")
    _lib.check(_lib.load().clv_rows_scale(_ptr(x), _is_bf16(x), _ptr(y), _is_bf16(y), rows, Cn, _ptr(scale), int(rows_per_group),
                                          _stream()), "clv_rows_scale")
    return y


# ------------------------------------------------------------------------------------------------
@_profiled("cast", lambda src, dst, *a, **k: f"n={src.numel()} {str(src.dtype)[-4:]}->{str(dst.dtype)[-4:]}",
           lambda src, dst, *a, **k: _tb(src, dst))
def cast(src, dst, scale=1.0):
    _need_cuda(src, dst)
    if not (src.is_contiguous() and dst.is_contiguous()) or src.numel() != dst.numel():
        raise ValueError("cast: contiguous tensors of equal size required")
    n = src.numel()
    if n % 4:
        raise ValueError("cast: numel must be a multiple of 4")
    _lib.check(_lib.load().clv_cast(_ptr(src), _is_bf16(src), _ptr(dst), _is_bf16(dst), n, float(scale), _stream()), "clv_cast")
    return dst


def to_bf16(src):
    return cast(src, torch.empty(src.shape, dtype=BF16, device=src.device))


@_profiled("gelu", None, lambda x, y, dy=None: _tb(x, y, dy))
def gelu(x, y, dy=None):
    """y = GELU(x), or y = dy * GELU'(x) when dy is given (contiguous tensors of equal size)."""
    _need_cuda(x, y, dy)
    n = x.numel()
    if n % 4 or y.numel() != n or not (x.is_contiguous() and y.is_contiguous()) or (dy is not None and not dy.is_contiguous()):
        raise ValueError("gelu: contiguous tensors with numel % 4 == 0 required")
    _lib.check(_lib.load().clv_gelu(_ptr(x), _is_bf16(x), _ptr(dy), _is_bf16(dy) if dy is not None else 0, _ptr(y),
                                   _is_bf16(y), n, _stream()), "clv_gelu")
    return y


@_profiled("gelu")
def tanh(x, y, dy=None):
    """y = tanh(x), or y = dy * (1 - x^2) when dy is given and x is the saved tanh output."""
    _need_cuda(x, y, dy)
    n = x.numel()
    if n % 4 or y.numel() != n or not (x.is_contiguous() and y.is_contiguous()) or (dy is not None and not dy.is_contiguous()):
        raise ValueError("tanh: contiguous tensors with numel % 4 == 0 required")
    _lib.check(_lib.load().clv_tanh(_ptr(x), _is_bf16(x), _ptr(dy), _is_bf16(dy) if dy is not None else 0, _ptr(y),
                                   _is_bf16(y), n, _stream()), "clv_tanh")
    return y


@_profiled("patchify", None, lambda x, patch, norm=None: _tb(x) + 2.0 * x.numel())
def patchify(x, patch, norm=None):
    """x fp32 (B,Cin,F,H,W) -> bf16 [B*D*Hp*Wp, Cin*pd*ph*pw] and (D,Hp,Wp).
    x uint8 + norm=(mean, inv_std) fp32 [Cin] device tensors: raw frames normalised in the load (GPUNormalize)."""
    _need_cuda(x)
    if x.dtype not in (F32, torch.uint8) or not x.is_contiguous():
        raise ValueError("patchify: contiguous fp32 (or uint8 + norm) input required")
    B, Cin, Fr, H, W = x.shape
    pd, ph, pw = patch
    D, Hp, Wp = -(-Fr // pd), -(-H // ph), -(-W // pw)
    out = torch.empty(B * D * Hp * Wp, Cin * pd * ph * pw, dtype=BF16, device=x.device)
    if x.dtype == torch.uint8:
        if norm is None or norm[0].numel() != Cin or norm[1].numel() != Cin or norm[0].dtype != F32 or norm[1].dtype != F32:
            raise ValueError("patchify: uint8 frames need norm=(mean, inv_std) fp32 [Cin] (set_input_normalization)")
        _need_cuda(*norm)
        _lib.check(_lib.load().clv_patchify_u8(_ptr(x), _ptr(norm[0]), _ptr(norm[1]), _ptr(out), B, Cin, Fr, H, W, pd, ph, pw,
                                               _stream()), "clv_patchify_u8")
    else:
        _lib.check(_lib.load().clv_patchify(_ptr(x), _ptr(out), B, Cin, Fr, H, W, pd, ph, pw, _stream()), "clv_patchify")
    return out, (D, Hp, Wp)


@_profiled("colsum", lambda x, out, *a, **k: f"rows={k.get('rows') or x.shape[0]} C={x.shape[-1]} {str(x.dtype)[-4:]} mod={k.get('mod', 1)}")
def grouped_colsum(x, out, div=1, mod=1, scale=1.0, accumulate=False, rows=None):
    _need_cuda(x, out)
    if out.dtype != F32 or not out.is_contiguous():
        raise ValueError("grouped_colsum: out must be contiguous fp32")
    Cn = x.shape[-1]
    rows = x.shape[0] if rows is None else rows
    _lib.check(_lib.load().clv_grouped_colsum(_ptr(x), _is_bf16(x), x.stride(0), rows, Cn, div, mod, float(scale), _ptr(out),
                                             int(accumulate), _stream()), "clv_grouped_colsum")
    return out


@_profiled("rows_affine")
def rows_affine(y, rows, Cn, x=None, in_group=None, out_group=None, add0=None, bvec=None, bdiv=1, bscale=1.0):
    _need_cuda(y)
    d = RowsAffine()
    if x is not None:
        d.x, d.x_is_bf16, d.ld_x = _ptr(x), _is_bf16(x), x.stride(0)
    if in_group is not None:
        d.in_group_rows, d.in_group_stride, d.in_offset = in_group
    d.y, d.y_is_bf16, d.ld_y = _ptr(y), _is_bf16(y), y.stride(0)
    if out_group is not None:
        d.out_group_rows, d.out_group_stride, d.out_offset = out_group
    d.add0 = _ptr(add0)
    if bvec is not None:
        d.bvec, d.bdiv, d.bscale = _ptr(bvec), int(bdiv), float(bscale)
    d.rows, d.C = int(rows), int(Cn)
    _lib.check(_lib.load().clv_rows_affine(C.byref(d), _stream()), "clv_rows_affine")
    return y


@_profiled("scatter_add")
def scatter_add_rows(src, index, dst):
    _need_cuda(src, index, dst)
    if src.dtype != F32 or dst.dtype != F32 or index.dtype != torch.int64:
        raise TypeError("scatter_add_rows: fp32 src/dst and int64 index required")
    _lib.check(_lib.load().clv_scatter_add_rows(_ptr(src), _ptr(index), _ptr(dst), src.shape[0], src.shape[1], _stream()),
               "clv_scatter_add_rows")
    return dst


# ------------------------------------------------------------------------------------------------
@_profiled("nce_fwd")
def nce_rank_fwd(embs, temperature, margin, use_rank, eps=1e-8):
    """embs: list of (nblk+1) contiguous fp32 [Bg, D] (query side first).  Returns (losses[2], workspace)."""
    _need_cuda(*embs)
    nblk = len(embs) - 1
    Bg, D = embs[0].shape
    for e in embs:
        if e.dtype != F32 or not e.is_contiguous() or e.shape != (Bg, D):
            raise ValueError("nce_rank_fwd: contiguous fp32 [Bg, D] embeddings required")
    lib = _lib.load()
    ws = torch.empty(lib.clv_nce_workspace_floats(nblk, Bg, D), dtype=F32, device=embs[0].device)
    out = torch.empty(2, dtype=F32, device=embs[0].device)
    arr = (C.c_void_p * (nblk + 1))(*[e.data_ptr() for e in embs])
    _lib.check(lib.clv_nce_rank_fwd(arr, nblk, Bg, D, float(temperature), float(margin), int(use_rank), float(eps),
                                    _ptr(ws), _ptr(out), _stream()), "clv_nce_rank_fwd")
    return out, ws


@_profiled("nce_bwd")
def nce_rank_bwd(ws, nblk, Bg, D, temperature, use_rank, g_nce, g_rank):
    grads = [torch.empty(Bg, D, dtype=F32, device=ws.device) for _ in range(nblk + 1)]
    arr = (C.c_void_p * (nblk + 1))(*[g.data_ptr() for g in grads])
    _lib.check(_lib.load().clv_nce_rank_bwd(nblk, Bg, D, float(temperature), int(use_rank), _ptr(ws), _ptr(g_nce),
                                           _ptr(g_rank), arr, _stream()), "clv_nce_rank_bwd")
    return grads


@_profiled("focal_fwd")
def softmax_focal_fwd(logits, target, V, gamma, ignore_index=-100):
    """logits fp32 [rows, >=V]; returns (loss[1], stats, sums)."""
    _need_cuda(logits, target)
    if logits.dtype != F32 or target.dtype != torch.int64:
        raise TypeError("softmax_focal_fwd: fp32 logits and int64 targets required")
    rows = logits.shape[0]
    stats = torch.empty(rows, 3, dtype=F32, device=logits.device)
    sums = torch.empty(2, dtype=F32, device=logits.device)
    loss = torch.empty(1, dtype=F32, device=logits.device)
    _lib.check(_lib.load().clv_softmax_focal_fwd(_ptr(logits), logits.stride(0), rows, V, _ptr(target), ignore_index,
                                                float(gamma), _ptr(stats), _ptr(sums), _ptr(loss), _stream()),
               "clv_softmax_focal_fwd")
    return loss, stats, sums


@_profiled("focal_bwd")
def softmax_focal_bwd(logits, target, V, gamma, stats, sums, g_loss, dlogits):
    Vpad = dlogits.shape[1]
    _lib.check(_lib.load().clv_softmax_focal_bwd(_ptr(logits), logits.stride(0), logits.shape[0], V, Vpad, _ptr(target),
                                                float(gamma), _ptr(stats), _ptr(sums), _ptr(g_loss), _ptr(dlogits),
                                                _is_bf16(dlogits), dlogits.stride(0), _stream()), "clv_softmax_focal_bwd")
    return dlogits
