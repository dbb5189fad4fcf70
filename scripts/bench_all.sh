#!/bin/bash
# One-GPU bench lines of every BASELINE workload (each with the CPU baseline = the reference's own modules on the host cores).
python bench.py --steps 50 --warmup 5 > gpurun_out/r02_bench_c2_n1.json 2> gpurun_out/r02_bench_c2_n1.err
for W in c3 c4 c5; do python bench.py --workload $W --steps 20 --warmup 3 > gpurun_out/r02_bench_${W}_n1.json 2> gpurun_out/r02_bench_${W}_n1.err; done
