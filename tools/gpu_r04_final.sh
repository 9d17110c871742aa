#!/bin/bash
# Final capture of round 2 (one B200): GPU tests, smoke(), default bench line (with the CPU and GPU-eager baselines), per-family
# CUDA-event table, fine-tune lines.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/r04z_pytest_gpu.txt 2>&1; tail -2 gpurun_out/r04z_pytest_gpu.txt
cp gpurun_out/parity_errors.json gpurun_out/r04z_parity_errors.json 2>/dev/null
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r04z_bench.json 2> gpurun_out/r04z_bench.err
CLOVER_B200_PROFILE_SHAPES=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline > gpurun_out/r04z_bench_shapes.json 2>/dev/null
cp gpurun_out/family_times.txt gpurun_out/r04z_family_times.txt
timeout 300 python bench.py --workload c4 --steps 10 --warmup 3 > gpurun_out/r04z_bench_c4.json 2>/dev/null
timeout 300 python bench.py --workload c5 --steps 10 --warmup 3 > gpurun_out/r04z_bench_c5.json 2>/dev/null
for f in r04z_bench r04z_bench_shapes r04z_bench_c4 r04z_bench_c5; do python -c "
import json; d=json.loads([l for l in open('gpurun_out/$f.json') if l.startswith('{')][0]); print('$f', round(d['value'],1), round(d['ms_per_step'],2), round(d['e2e']['value'],1), d['clocks'], round(d['roofline']['frac'],3), d.get('cpu_baseline'), (d.get('gpu_eager_baseline') or {}).get('bf16_autocast'), d['roofline'].get('window_attn_c2_fwd_bwd'))"; done
