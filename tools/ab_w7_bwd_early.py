import sys, runpy
sys.path.insert(0, ".")
from clover_b200 import ops
for mode in (1, 0):
    ops.set_tunable("w7_bwd_early", mode)
    print(f"# w7_bwd_early = {mode}", flush=True)
    sys.argv = ["tools/attn_microbench.py", "--shapes", "s3,s1,c2", "--iters", "10", "--which", "bwd"]
    runpy.run_path("tools/attn_microbench.py", run_name="__main__")
