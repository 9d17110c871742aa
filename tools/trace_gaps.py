"""Reads a torch.profiler chrome trace written by `bench.py --trace DIR` and prints, per CUDA stream, the busy time, the
idle gaps on the compute stream and the NCCL kernels (start relative to the step, duration) -- the evidence for where the
multi-GPU step loses time against the single-GPU one."""
import gzip
import json
import sys
from collections import defaultdict


def main(path, top=25):
    ev = json.load(gzip.open(path))["traceEvents"]
    ks = [e for e in ev if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and "dur" in e]
    ks.sort(key=lambda e: e["ts"])
    t0 = ks[0]["ts"]
    t1 = max(e["ts"] + e["dur"] for e in ks)
    by = defaultdict(list)
    for e in ks:
        by[e["args"].get("stream", e.get("tid"))].append(e)
    print(f"span {(t1 - t0) / 1e3:.2f} ms, {len(ks)} device activities, {len(by)} streams")
    main_stream = max(by, key=lambda s: sum(e["dur"] for e in by[s]))
    for s, l in sorted(by.items(), key=lambda kv: -sum(e["dur"] for e in kv[1])):
        print(f"stream {s}: {len(l)} activities, busy {sum(e['dur'] for e in l) / 1e3:.2f} ms" + ("  <- compute" if s == main_stream else ""))
    l = by[main_stream]
    gaps = []
    for a, b in zip(l, l[1:]):
        g = b["ts"] - (a["ts"] + a["dur"])
        if g > 0:
            gaps.append((g, a, b))
    print(f"compute-stream idle: {sum(g for g, _, _ in gaps) / 1e3:.2f} ms in {len(gaps)} gaps; the {top} largest:")
    for g, a, b in sorted(gaps, key=lambda x: -x[0])[:top]:
        print(f"  {g / 1e3:7.3f} ms at t={(a['ts'] + a['dur'] - t0) / 1e3:8.2f}  after {a['name'][:60]}  before {b['name'][:60]}")
    print("NCCL / other-stream activities:")
    for s, l2 in by.items():
        if s == main_stream:
            continue
        for e in l2:
            if e["dur"] > 50:
                print(f"  stream {s} t={(e['ts'] - t0) / 1e3:8.2f} dur {e['dur'] / 1e3:7.3f} ms  {e['name'][:70]}")
    tot = defaultdict(float)
    for e in l:
        tot[e["name"][:70]] += e["dur"]
    print("compute-stream top kernels by total time:")
    for n, d in sorted(tot.items(), key=lambda kv: -kv[1])[:15]:
        print(f"  {d / 1e3:8.3f} ms  {n}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)
