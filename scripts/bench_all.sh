#!/bin/bash
# One-GPU bench lines of every BASELINE workload (c2 with the CPU baseline; c3-c5 per-GPU shards of the multi-GPU configs).
python bench.py --steps 50 --warmup 5 > gpurun_out/r02_bench_c2_n1.json 2> gpurun_out/r02_bench_c2_n1.err
for W in c3 c4 c5; do python bench.py --workload $W --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_${W}_n1.json 2> gpurun_out/r02_bench_${W}_n1.err; done
