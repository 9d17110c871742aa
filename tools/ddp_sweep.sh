#!/bin/bash
# DDP knob sweep at N = 2 (developer): bucket size and static_graph, kernels-only timing
run() { echo "== $*"; python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 4 --kernels-only "$@" 2>&1 | grep '^{' | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'])"; }
run
run --bucket-mb 50
run --bucket-mb 200
run --static-graph
run --bucket-mb 25
run
python bench.py --steps 10 --warmup 3 --kernels-only | grep '^{'
