#!/bin/bash
# Round-2 capture L (8 B200s of one box): c3 / c4 / c5 bench lines at N = 8 (bf16 gradient buckets, gradient sinks).
mkdir -p gpurun_out
for wl in c3 c4 c5; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 10 --warmup 3 --workload $wl > gpurun_out/r02l_bench_8gpu_$wl.json 2> gpurun_out/r02l_bench_8gpu_$wl.err
  grep '^{' gpurun_out/r02l_bench_8gpu_$wl.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$wl', d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'], d.get('rank_consistency'))"
  tail -c 400 gpurun_out/r02l_bench_8gpu_$wl.err | grep -v OMP
done
