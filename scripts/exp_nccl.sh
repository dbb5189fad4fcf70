run() { echo "== $*"; env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 30 --warmup 5 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value'],1), round(d['ms_per_step'],3))"; }
run G2_DUMMY=1
run NCCL_MAX_NCHANNELS=2
run NCCL_MAX_NCHANNELS=4
run NCCL_MAX_NCHANNELS=8
run NCCL_MAX_NCHANNELS=4 NCCL_ALGO=Ring
