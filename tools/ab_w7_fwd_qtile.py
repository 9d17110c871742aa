"""A/B of the query-tile height of the (wd,7,7) window-attention forward (98 rows = two slabs vs 128 rows = every TMEM lane):
python tools/ab_w7_fwd_qtile.py [shapes]"""
import runpy
import sys

sys.path.insert(0, ".")
from clover_b200 import ops

shapes = sys.argv[1] if len(sys.argv) > 1 else "s3,s1,s2,s4,c2"
clips = sys.argv[2] if len(sys.argv) > 2 else "0"
for q in (0, 1, -1, 0, 1, -1):
    ops.set_tunable("w7_fwd_qtile", q)
    print(f"# w7_fwd_qtile = {q}", flush=True)
    sys.argv = ["tools/attn_microbench.py", "--shapes", shapes, "--iters", "20", "--which", "fwd", "--phases", "1", "--clips", clips]
    runpy.run_path("tools/attn_microbench.py", run_name="__main__")
