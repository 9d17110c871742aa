#!/bin/bash
# Round-2 capture I (one B200): fine-tune bench lines (c4 / c5), the ncu launch list of the default bench command (time + DRAM
# bytes per launch), ncu --set full captures of the kernels that are new or changed this round.  Everything lands in gpurun_out/.
set -x
mkdir -p gpurun_out
python bench.py --workload c4 --steps 10 --warmup 3 > gpurun_out/r02i_bench_c4.json 2> gpurun_out/r02i_bench_c4.err
python bench.py --workload c5 --steps 10 --warmup 3 > gpurun_out/r02i_bench_c5.json 2> gpurun_out/r02i_bench_c5.err
tail -c 300 gpurun_out/r02i_bench_c5.err
B="python bench.py --steps 1 --warmup 1 --kernels-only"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r02i_launches.csv $B > gpurun_out/ncu_ll.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn64_ -s 30 -c 6 -o gpurun_out/r02i_attn64 $B > gpurun_out/ncu1.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"attn_w7_bwd2_kernel|attn_w7_fwd2_kernel|attn_w7_dbias" -c 6 -o gpurun_out/r02i_attn_c2 python tools/attn_microbench.py --shapes c2 --iters 1 > gpurun_out/ncu2.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"gemm_bf16_kernel<0, 0, 256, false, true>" -s 40 -c 3 -o gpurun_out/r02i_gemm_fc1_ts $B > gpurun_out/ncu3.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"attn_w7_bwd2_kernel|attn_w7_fwd_kernel" -s 20 -c 4 -o gpurun_out/r02i_attn_s3 $B > gpurun_out/ncu4.log 2>&1
ls -la gpurun_out/ | grep r02i
