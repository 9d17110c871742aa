run() { echo "== $*"; env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --kernels-only 2>&1 | grep '^{' | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'])"; }
run X=1
run NCCL_MAX_CTAS=2
run NCCL_MAX_CTAS=4
run NCCL_MAX_CTAS=8
run NCCL_MAX_CTAS=4 NCCL_PROTO=Simple
run X=1
python bench.py --steps 10 --warmup 3 --kernels-only | grep '^{'
