"""Prints the same-box A/B comparison written by tools/ab_bench.sh (step time and the per-shape family table)."""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")


def line(tag, r):
    d = json.loads([l for l in open(os.path.join(G, f"ab_{tag}_{r}.json")) if l.startswith("{")][-1])
    return d


def fam(tag):
    out = {}
    for l in open(os.path.join(G, f"ab_{tag}_family_times.txt")).read().splitlines()[1:]:
        p = [x.strip() for x in l.split("|")]
        out[p[0]] = (int(p[1]), float(p[2]))
    return out


for r in (1, 2):
    b, n = line("base", r), line("new", r)
    print(f"round {r}: base {b['ms_per_step']:.2f} ms ({b['value']:.1f} clips/s, {b['clocks']['sm_mhz']} MHz)   "
          f"new {n['ms_per_step']:.2f} ms ({n['value']:.1f} clips/s, {n['clocks']['sm_mhz']} MHz)   "
          f"ratio {b['ms_per_step'] / n['ms_per_step']:.4f}")
fb, fn = fam("base"), fam("new")
rows = sorted(set(fb) | set(fn), key=lambda k: -abs(fb.get(k, (0, 0))[1] - fn.get(k, (0, 0))[1]))
print("family | launches | base ms | new ms | delta")
tb = tn = 0.0
for k in rows:
    tb += fb.get(k, (0, 0))[1]; tn += fn.get(k, (0, 0))[1]
for k in rows[:28]:
    print(f"{k[:88]:88s} | {fn.get(k, fb.get(k))[0]:3d} | {fb.get(k, (0, 0))[1]:7.3f} | {fn.get(k, (0, 0))[1]:7.3f} | {fn.get(k, (0, 0))[1] - fb.get(k, (0, 0))[1]:+7.3f}")
print(f"sum of families: base {tb:.2f} ms  new {tn:.2f} ms")
