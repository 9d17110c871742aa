"""numpy restatement of the counter-based random stream of the dropout kernels (clover_b200/csrc/common.cuh rand_u32 /
drop_threshold) and a replay helper that feeds the product's recorded draws (clover_b200.rng.LOG) to the oracle."""
import numpy as np
import torch

_M = np.uint64(0xFFFFFFFFFFFFFFFF)


def rand_u32(seed, idx):
    """idx: uint64 array of stream positions -> uint32 array (high word of splitmix64(idx + seed * phi))."""
    with np.errstate(over="ignore"):
        z = idx.astype(np.uint64) + np.uint64(seed) * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return (z >> np.uint64(32)).astype(np.uint32)


def threshold(p):
    t = float(np.float32(p)) * 4294967296.0
    return np.uint32(0 if t <= 0 else min(int(t), 4294967295))


def keep_mask(n, p, seed, offset):
    idx = np.arange(n, dtype=np.uint64) + np.uint64(offset)
    return rand_u32(seed, idx) >= threshold(p)


class Replay:
    """Callable (kind, x) -> x * keep / (1 - p) that consumes the product's logged dropout sites in call order."""

    def __init__(self, log):
        self.sites = [e for e in log if e["kind"] != "drop_path"]
        self.paths = [e["scale"] for e in log if e["kind"] == "drop_path"]
        self.pos = 0

    def __call__(self, kind, x):
        e = self.sites[self.pos]
        self.pos += 1
        assert e["kind"] == kind, (e["kind"], kind)
        assert int(np.prod(e["shape"])) == x.numel(), (kind, e["shape"], tuple(x.shape))
        keep = torch.from_numpy(keep_mask(x.numel(), e["p"], e["seed"], e["offset"])).view(x.shape)
        return x * keep.to(x.dtype) / (1.0 - float(np.float32(e["p"])))

    def drop_path_pairs(self, rates):
        """One (attention-branch, MLP-branch) factor pair per Swin block in network order; blocks whose drop_path rate
        is 0 hold an nn.Identity in the reference and draw nothing -> None."""
        it = iter(self.paths)
        out = [None if r == 0 else (next(it), next(it)) for r in rates]
        assert next(it, None) is None, "unconsumed DropPath draws"
        return out

    def done(self):
        return self.pos == len(self.sites)
