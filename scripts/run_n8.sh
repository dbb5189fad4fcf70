#!/bin/bash
# 8-GPU measurements of one round (gpurun --gpus 8): c2 (headline) with and without all-reduce overlap, c4 and c5 (BASELINE configs[3,4]).
set -u
mkdir -p gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 20 --warmup 3 "$@" 2>gpurun_out/n8.err | tail -1; }
run > gpurun_out/r02_bench_c2_n8.json; cut -c1-160 gpurun_out/r02_bench_c2_n8.json
run --no-overlap > gpurun_out/r02_bench_c2_n8_nooverlap.json; cut -c1-160 gpurun_out/r02_bench_c2_n8_nooverlap.json
run --workload c4 > gpurun_out/r02_bench_c4_n8.json; cut -c1-160 gpurun_out/r02_bench_c4_n8.json
run --workload c5 > gpurun_out/r02_bench_c5_n8.json; cut -c1-160 gpurun_out/r02_bench_c5_n8.json
tail -3 gpurun_out/n8.err
