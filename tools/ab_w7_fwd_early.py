"""A/B of the forward issue orders of the (wd,7,7) window attention:  python tools/ab_w7_fwd_early.py [shapes]
(w7_fwd_early: next tile's score MMAs before the O read-out; w7_fwd_pvsplit: P V products of the first key bodies while the rest is
still being exponentiated); also prints the per-phase cycle sums of the softmax warp."""
import runpy
import sys

sys.path.insert(0, ".")
from clover_b200 import ops

shapes = sys.argv[1] if len(sys.argv) > 1 else "s3,s1,s2,c2"
for early, split in ((0, 0), (1, 1), (0, 0), (1, 1)):
    ops.set_tunable("w7_fwd_early", early)
    ops.set_tunable("w7_fwd_pvsplit", split)
    print(f"# w7_fwd_early = {early} w7_fwd_pvsplit = {split}", flush=True)
    sys.argv = ["tools/attn_microbench.py", "--shapes", shapes, "--iters", "10", "--which", "fwd,bwd", "--phases", "1"]
    runpy.run_path("tools/attn_microbench.py", run_name="__main__")
