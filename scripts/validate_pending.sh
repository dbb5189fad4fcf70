#!/bin/bash
# HISTORICAL (round 2, first GPU call): ran everything round 1 had left un-run on a B200 -- the then-pending tests, the A/B
# micro-benchmarks of the default-off kernels and a whole-step A/B of every switch.  Its output is profiles/r02_pending_validation.txt.
# The tests it ran are now part of `pytest -m gpu` and the kernels it compared are default routes or deleted,
# so what is left here is the A/B of the switches that still exist.
mkdir -p gpurun_out
S=c2_bdec_fwd70,c2_bdec_dgrad,c2_att64_fwd,c2_att64_dgrad,c2_att32_fwd,c2_att_up64_fwd,c3_unet64,c3_dec_up64,c5_unet128,c5_bdec_fwd
for p in 0 1 2; do echo "-- G2_HALO_PERSISTENT=$p"; G2_HALO_PERSISTENT=$p timeout 200 python scripts/conv_bench.py --only $S 2>&1 | grep halo; done
for w in 0 1; do echo "-- G2_WGRAD_HALO=$w"; G2_WGRAD_HALO=$w timeout 200 python scripts/conv_bench.py --wgrad --only $S 2>&1 | grep wgrad; done
for cfg in "G2_DUMMY=1" "G2_FUSED_LATENT=0" "G2_HALO_PERSISTENT=0" "G2_WGRAD_HALO=0" "G2_NORM_DIRECT=0" "G2_GRAD_STREAM=0" "G2_SIDE_STREAMS=0"; do
  echo "-- $cfg"
  env $cfg python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'])"
done
