"""Closed-form integer tables of the Video Swin backbone (host side, numpy, cached).

The kernels never read an (N, N) index or mask tensor: the relative-position bias index
factorises as ``idx(i, j) = code[i] - code[j] + off`` and the shift mask as
``region[i] != region[j]``, so only O(N) integers per window are shipped to the device.
The full (N, N) forms are kept for state-dict parity (``relative_position_index`` buffer) and for
the bit-exact tests against the reference's tables (swin_transformer_3d.py:302-315, 345-359, 548-562).
"""
from functools import lru_cache

import numpy as np


def get_window_size(x_size, window_size, shift_size=None):
    """Window clamped to the input extent; shift zeroed on clamped axes (swin_transformer_3d.py:302-315)."""
    use_w = tuple(x if x <= w else w for x, w in zip(x_size, window_size))
    if shift_size is None:
        return use_w
    use_s = tuple(0 if x <= w else s for x, w, s in zip(x_size, window_size, shift_size))
    return use_w, use_s


@lru_cache(maxsize=None)
def rel_code(n_tokens, window_cfg):
    """code[n] = d*(2Wh-1)(2Ww-1) + h*(2Ww-1) + w with (d,h,w) the decomposition of token n over the
    CONFIGURED window (the reference slices relative_position_index[:N,:N], :382), and the offset
    that makes ``code[i] - code[j] + off`` equal relative_position_index[i, j]."""
    Wd, Wh, Ww = window_cfg
    n = np.arange(n_tokens, dtype=np.int64)
    d, h, w = n // (Wh * Ww), (n // Ww) % Wh, n % Ww
    sh, sw = (2 * Wh - 1) * (2 * Ww - 1), (2 * Ww - 1)
    code = d * sh + h * sw + w
    off = (Wd - 1) * sh + (Wh - 1) * sw + (Ww - 1)
    return code.astype(np.int32), int(off)


@lru_cache(maxsize=None)
def relative_position_index(window_cfg):
    """int64 (N, N) buffer with the reference's values (swin_transformer_3d.py:345-359)."""
    Wd, Wh, Ww = window_cfg
    code, off = rel_code(Wd * Wh * Ww, tuple(window_cfg))
    c = code.astype(np.int64)
    return c[:, None] - c[None, :] + off


def _axis_region(L, w, s):
    i = np.arange(L)
    if s == 0:
        return np.zeros(L, dtype=np.int32)
    return ((i >= L - w).astype(np.int32) + (i >= L - s).astype(np.int32))


@lru_cache(maxsize=None)
def region_ids(D, H, W, window, shift):
    """int32 (nWin, N): shift-mask region of every window token in the rolled frame
    (compute_mask, swin_transformer_3d.py:548-562).  (D,H,W) is the padded extent."""
    wd, wh, ww = window
    nD, nH, nW = D // wd, H // wh, W // ww
    g = np.arange(nD * nH * nW)[:, None]
    n = np.arange(wd * wh * ww)[None, :]
    d = (g // (nH * nW)) * wd + n // (wh * ww)
    h = ((g // nW) % nH) * wh + (n // ww) % wh
    w = (g % nW) * ww + n % ww
    rid = _axis_region(D, wd, shift[0])[d] * 9 + _axis_region(H, wh, shift[1])[h] * 3 + _axis_region(W, ww, shift[2])[w]
    return np.ascontiguousarray(rid.astype(np.int32))


def attn_mask_from_regions(rid):
    """(nWin, N, N) float32 0 / -100 mask, the reference's materialised form (tests only)."""
    return np.where(rid[:, None, :] != rid[:, :, None], np.float32(-100.0), np.float32(0.0))


def window_gather_index(B, D, H, W, window, shift):
    """int64 (B*nWin, N) flat source rows of fused roll(-shift)+window_partition over the padded
    frame (tests / documentation; the kernels evaluate the same map in registers)."""
    wd, wh, ww = window
    nD, nH, nW = D // wd, H // wh, W // ww
    g = np.arange(nD * nH * nW, dtype=np.int64)[:, None]
    n = np.arange(wd * wh * ww, dtype=np.int64)[None, :]
    d = ((g // (nH * nW)) * wd + n // (wh * ww) + shift[0]) % D
    h = (((g // nW) % nH) * wh + (n // ww) % wh + shift[1]) % H
    w = ((g % nW) * ww + n % ww + shift[2]) % W
    src = (d * H + h) * W + w
    return (np.arange(B, dtype=np.int64)[:, None, None] * (D * H * W) + src[None]).reshape(B * src.shape[0], -1)


@lru_cache(maxsize=None)
def window_row_map(D, H, W, window, shift):
    """int32 (D*H*W,): window-order row (within one clip) of every spatial token s of an UNPADDED frame --
    the inverse permutation of :func:`window_gather_index`, i.e. where roll(-shift) + window_partition
    (swin_transformer_3d.py:456-466) puts token s; window_reverse + roll(+shift) (:471-479) reads it back."""
    src = window_gather_index(1, D, H, W, tuple(window), tuple(shift)).reshape(-1)
    inv = np.empty(D * H * W, dtype=np.int32)
    inv[src] = np.arange(D * H * W, dtype=np.int32)
    return inv


def w7_ext_tables(rid):
    """K-extension rows of the 7x7-window attention kernels (include/clover_b200.h, clv_attn_w7_desc_t) from the
    shift-mask region ids ``rid`` (nWin, N) of :func:`region_ids` (id = 9*rd + 3*rh + rw).  Returns float32
    (q_ext, k_ext), each (nWin, N, 16); every value is exactly representable in bf16.  <q_ext[i], k_ext[j]> is
    100 * (#axes on which the regions of i and j agree) - 300: 0 where compute_mask (swin_transformer_3d.py:548-562)
    gives 0, and -100, -200 or -300 where it gives -100."""
    rid = np.asarray(rid)
    nW, N = rid.shape
    axes = np.stack([rid // 9, (rid // 3) % 3, rid % 3], -1)                      # (nW, N, 3)
    onehot = (axes[..., None] == np.arange(3)).astype(np.float32).reshape(nW, N, 9) * 10.0
    q = np.zeros((nW, N, 16), np.float32)
    k = np.zeros((nW, N, 16), np.float32)
    q[..., 4:13] = onehot
    k[..., 4:13] = onehot
    q[..., 13], q[..., 14] = 1.0, 1.0
    k[..., 0], k[..., 1] = 1.0, 1.0
    k[..., 13], k[..., 14] = -256.0, -44.0
    return q, k
