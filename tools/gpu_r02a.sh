#!/bin/bash
# Round-2 capture A (one B200): gpu tests with achieved-error recording, the default bench line, ncu --set full captures of the
# HBM-bound kernels (LayerNorm, optimizer, losses).  Everything lands in gpurun_out/.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -60 > gpurun_out/r02a_pytest.txt
tail -5 gpurun_out/r02a_pytest.txt
python bench.py --steps 10 --warmup 3 > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err
tail -c 600 gpurun_out/r02a_bench.err
B="python bench.py --steps 1 --warmup 1 --kernels-only"
ncu --set full --clock-control none --import-source on -k regex:lnr_bwd_kernel -s 115 -c 10 -o gpurun_out/r02a_lnr_bwd $B > gpurun_out/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:lnr_fwd_kernel -s 100 -c 8 -o gpurun_out/r02a_lnr_fwd $B > gpurun_out/ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"adamw_kernel|grad_sqnorm_kernel" -s 2 -c 2 -o gpurun_out/r02a_adamw $B > gpurun_out/ncu3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"nce_|focal_|cos_norm" -s 14 -c 14 -o gpurun_out/r02a_loss $B > gpurun_out/ncu4.log 2>&1
ls -la gpurun_out/
