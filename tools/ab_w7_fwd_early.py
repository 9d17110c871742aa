"""A/B of the forward issue order of the (wd,7,7) window attention:  python tools/ab_w7_fwd_early.py [shapes]"""
import runpy
import sys

sys.path.insert(0, ".")
from clover_b200 import ops

shapes = sys.argv[1] if len(sys.argv) > 1 else "s3,s1,s2,c2"
for mode in (0, 1, 0, 1):
    ops.set_tunable("w7_fwd_early", mode)
    print(f"# w7_fwd_early = {mode}", flush=True)
    sys.argv = ["tools/attn_microbench.py", "--shapes", shapes, "--iters", "10", "--which", "fwd"]
    runpy.run_path("tools/attn_microbench.py", run_name="__main__")
