#!/bin/bash
# 8-GPU measurements of one round (gpurun --gpus 8): c2 (headline) with and without all-reduce overlap, c3, c4 and c5 (BASELINE configs),
# and the multi-GPU gradient-equivalence tests.
set -u
mkdir -p gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 30 --warmup 3 "$@" 2>gpurun_out/n8.err | tail -1; }
run > gpurun_out/r02_bench_c2_n8.json; cut -c1-160 gpurun_out/r02_bench_c2_n8.json
run --no-overlap > gpurun_out/r02_bench_c2_n8_nooverlap.json; cut -c1-160 gpurun_out/r02_bench_c2_n8_nooverlap.json
for W in c3 c4 c5; do run --workload $W > gpurun_out/r02_bench_${W}_n8.json; cut -c1-160 gpurun_out/r02_bench_${W}_n8.json; done
tail -3 gpurun_out/n8.err
python -m pytest tests/test_multigpu_gpu.py -q -m gpu 2>&1 | tail -3
