#!/bin/bash
# Round-2 final capture (one B200): default bench line (with the CPU and GPU-eager baselines), fine-tune lines, ncu launch list of
# the default command (time + DRAM bytes per launch), per-family CUDA-event table.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02z_bench.json 2> gpurun_out/r02z_bench.err
CLOVER_B200_PROFILE_SHAPES=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline > gpurun_out/r02z_bench_shapes.json 2>/dev/null
cp gpurun_out/family_times.txt gpurun_out/r02z_family_times.txt
timeout 600 python bench.py --workload c4 --steps 10 --warmup 3 > gpurun_out/r02z_bench_c4.json 2>/dev/null
timeout 600 python bench.py --workload c5 --steps 10 --warmup 3 > gpurun_out/r02z_bench_c5.json 2>/dev/null
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r02z_launches.csv python bench.py --steps 2 --warmup 2 --kernels-only > gpurun_out/ncu_ll.log 2>&1
for f in r02z_bench r02z_bench_c4 r02z_bench_c5; do python -c "
import json; d=json.loads([l for l in open('gpurun_out/$f.json') if l.startswith('{')][0]); print('$f', d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'], d['roofline']['frac'], d.get('cpu_baseline'), (d.get('gpu_eager_baseline') or {}).get('bf16_autocast'))"; done
