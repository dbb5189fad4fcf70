#!/bin/bash
# Round-2 profile artifacts on one B200 (outputs under gpurun_out/; summaries are copied to profiles/ by hand).
W=${1:-c2}
ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/r02_launches_$W.csv \
    python bench.py --workload $W --steps 2 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/r02_launches_$W.log 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
    -k regex:'conv_halo|conv_tc|wgrad' --csv --log-file gpurun_out/r02_traffic_$W.csv \
    python scripts/profile_step.py gpurun_out/r02_step_profile_$W.json --workload $W > gpurun_out/r02_step_profile_$W.txt 2>&1
if [ "$W" = c2 ]; then
  for S in c2_bdec_fwd70 c2_att64_fwd c2_att64_dgrad; do
    ncu --set full --import-source on --clock-control none -k regex:conv_halo -c 1 -f -o gpurun_out/r02_ncu_halo_$S \
        python scripts/conv_bench.py --only $S --reps 1 > gpurun_out/ncu_$S.log 2>&1
  done
  ncu --set full --import-source on --clock-control none -k regex:wgrad_halo -c 1 -f -o gpurun_out/r02_ncu_wgrad_c2_att64 \
      python scripts/conv_bench.py --wgrad --only c2_att64_fwd --reps 1 > gpurun_out/ncu_wgrad.log 2>&1
fi
