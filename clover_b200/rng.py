"""Host side of the counter-based random streams used by the stochastic regularisers (dropout, DropPath).

The kernels draw element ``e`` of a stream as ``keep = hash32(seed, e) >= p * 2**32`` (clover_b200/csrc/common.cuh,
rand_u32; restated in numpy by tests/rng_ref.py).  Nothing random lives on the device: this module hands every
dropout site a disjoint ``[offset, offset + n)`` range of one 64-bit stream per process, so the backward pass (and
the attention kernels) regenerate the identical mask from ``(seed, offset)``.

``manual_seed(s)`` pins the stream (tests); otherwise the seed comes from ``torch.initial_seed()`` mixed with the
rank, so data-parallel ranks draw different masks as they do in the reference (independent torch generators).
``LOG`` (a list, or None) records every site in call order so that tests can hand the same masks to the oracle.
"""
import torch

_MASK64 = (1 << 64) - 1
_state = {"seed": None, "offset": 0, "pinned": False, "key": None}
LOG = None


def manual_seed(seed):
    """Pin the stream to `seed` (offset 0) until reseed() is called."""
    _state.update(seed=int(seed) & _MASK64, offset=0, pinned=True, key=None)


def reseed():
    """Drop a pinned / derived seed: the next draw re-derives it from torch.initial_seed() and the rank."""
    _state.update(seed=None, offset=0, pinned=False, key=None)


def get_state():
    """(seed, offset, pinned) -- put it into a checkpoint next to torch's RNG state to make a resume reproducible."""
    return _seed(), _state["offset"], _state["pinned"]


def set_state(state):
    seed, offset, pinned = state
    _state.update(seed=int(seed) & _MASK64, offset=int(offset), pinned=bool(pinned), key=_key() if not pinned else None)


def _key():
    rank = 0
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        rank = torch.distributed.get_rank()
    return int(torch.initial_seed()), rank


def _seed():
    """Unless pinned, the seed follows torch.initial_seed() and the data-parallel rank: a later torch.manual_seed() (the
    reference's set_random_seed, tools/train.py) or init_process_group() re-derives it and restarts the stream, so ranks
    never share masks because the first draw happened before the process group existed."""
    if _state["pinned"]:
        return _state["seed"]
    key = _key()
    if _state["seed"] is None or _state["key"] != key:
        _state.update(seed=(key[0] ^ (key[1] * 0x9E3779B97F4A7C15)) & _MASK64, offset=0, key=key)
    return _state["seed"]


def next_stream(n, kind="dropout", shape=None, p=0.0):
    """Reserve n consecutive stream elements; returns (seed, offset)."""
    seed, off = _seed(), _state["offset"]
    _state["offset"] = (off + int(n) + 3) // 4 * 4
    if LOG is not None:
        LOG.append(dict(kind=kind, shape=tuple(shape) if shape is not None else (int(n),), p=float(p), seed=seed, offset=off))
    return seed, off


_PREDRAWN = []          # FIFO of (batch, p, tensor) filled by predraw_drop_path and consumed by drop_path_scales


def predraw_drop_path(batch, ps, device):
    """Draws the DropPath factors of a whole backbone pass (one entry of `ps` per later drop_path_scales call, in call order)
    with ONE uniform draw instead of two tiny kernels per residual branch (48 branches in Swin-B).  Same distribution as
    bernoulli_(keep) / keep; the following drop_path_scales(batch, p, device) calls return rows of the batch draw."""
    _PREDRAWN.clear()
    ps = [float(p) for p in ps]
    if not ps:
        return
    keep = 1.0 - torch.tensor(ps, dtype=torch.float32).clamp_(0.0, 1.0)
    keep_d = keep.to(device)[:, None]
    u = torch.rand(len(ps), batch, dtype=torch.float32, device=device)
    s = (u < keep_d).to(torch.float32) / keep_d.clamp_min(1e-30)
    for p, row in zip(ps, s.unbind(0)):
        _PREDRAWN.append((batch, p, row))


def drop_path_scales(batch, p, device):
    """timm DropPath factors (drop.py: x.new_empty(B,1,..).bernoulli_(keep) / keep): fp32 [batch] on `device`, drawn
    from torch's generator of that device (torch.manual_seed controls it, as in the reference)."""
    keep = 1.0 - float(p)
    s = None
    if _PREDRAWN:
        b0, p0, row = _PREDRAWN.pop(0)
        if b0 == batch and p0 == float(p) and row.device == torch.device(device):
            s = row
        else:
            _PREDRAWN.clear()               # out of step with the pre-drawn sequence: fall back to individual draws
    if s is None:
        s = torch.empty(batch, dtype=torch.float32, device=device).bernoulli_(keep)
        if keep > 0.0:
            s.div_(keep)
    if LOG is not None:
        LOG.append(dict(kind="drop_path", shape=(batch,), p=float(p), scale=s.detach().cpu().clone()))
    return s
