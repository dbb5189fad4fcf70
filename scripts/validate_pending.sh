#!/bin/bash
# First GPU call of the next round: validate everything that was written after the round-1 GPU budget was spent.
#   gpurun --timeout 1800 -- 'bash scripts/validate_pending.sh'   (about 20 GPU-minutes: two test passes incl. child-process suites, three A/B benches)
# Writes gpurun_out/pending_*.txt.  Nothing here changes defaults; flip them in the source once the numbers are in.
set -u
mkdir -p gpurun_out
echo "== 1. regular GPU gate" | tee gpurun_out/pending_summary.txt
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee -a gpurun_out/pending_summary.txt
echo "== 2. pending tests (fused Adam, LSTM first step, fused latent kernels, eval forward, VAE, variants, MN-major probe)" | tee -a gpurun_out/pending_summary.txt
G2_RUN_PENDING=1 timeout 900 python -m pytest tests/test_pending_next_round.py -q -rf 2>&1 | tail -40 | tee gpurun_out/pending_tests.txt | tail -15 | tee -a gpurun_out/pending_summary.txt
echo "== 3. persistent halo kernel: A/B micro-benchmark (same process, env read once -> two runs)" | tee -a gpurun_out/pending_summary.txt
S=c2_bdec_fwd70,c2_bdec_dgrad,c2_att64_fwd,c2_att64_dgrad,c2_att32_fwd,c2_att_up64_fwd,c3_unet64,c3_dec_up64,c5_unet128,c5_bdec_fwd
for p in 0 1; do
  echo "-- G2_HALO_PERSISTENT=$p" | tee -a gpurun_out/pending_conv_bench.txt
  G2_HALO_PERSISTENT=$p timeout 200 python scripts/conv_bench.py --only $S 2>&1 | tee -a gpurun_out/pending_conv_bench.txt
done
tail -24 gpurun_out/pending_conv_bench.txt >> gpurun_out/pending_summary.txt
echo "== 3b. halo weight-gradient kernel: A/B micro-benchmark" | tee -a gpurun_out/pending_summary.txt
for p in 0 1; do
  echo "-- G2_WGRAD_HALO=$p" | tee -a gpurun_out/pending_wgrad_bench.txt
  G2_WGRAD_HALO=$p timeout 200 python scripts/conv_bench.py --wgrad --only $S 2>&1 | grep wgrad | tee -a gpurun_out/pending_wgrad_bench.txt
done
tail -24 gpurun_out/pending_wgrad_bench.txt >> gpurun_out/pending_summary.txt
echo "== 4. whole step: fused latent kernels off / on, persistent off / on, halo wgrad off / on, skinny GEMM off / on" | tee -a gpurun_out/pending_summary.txt
for cfg in "G2_FUSED_LATENT=0 G2_HALO_PERSISTENT=0" "G2_FUSED_LATENT=1 G2_HALO_PERSISTENT=0" "G2_FUSED_LATENT=1 G2_HALO_PERSISTENT=1" "G2_FUSED_LATENT=0 G2_WGRAD_HALO=1" "G2_FUSED_LATENT=0 G2_SKINNY_GEMM=1" "G2_FUSED_LATENT=1 G2_SKINNY_GEMM=1" "G2_FUSED_LATENT=1 G2_SKINNY_GEMM=1 G2_NORM_DIRECT=1" "G2_FUSED_LATENT=1 G2_HALO_PERSISTENT=1 G2_WGRAD_HALO=1"; do
  echo "-- $cfg" | tee -a gpurun_out/pending_summary.txt
  env $cfg timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['config']['last_elbo'])" | tee -a gpurun_out/pending_summary.txt
done
# Optional follow-ups (each a separate, short GPU call; ncu replays every kernel ~40x, so keep -c small):
#   ncu --set full --clock-control none --import-source on -k regex:conv_halo_persistent -c 2 -o gpurun_out/r02_halo_persistent \
#       env G2_HALO_PERSISTENT=1 python scripts/conv_bench.py --only c2_bdec_fwd70 --reps 1
#   ncu --set full --clock-control none --import-source on -k regex:wgrad_halo -c 2 -o gpurun_out/r02_wgrad_halo \
#       env G2_WGRAD_HALO=1 python scripts/conv_bench.py --wgrad --only c2_att64_fwd --reps 1
#   python scripts/ncu_summary.py gpurun_out/r02_*.ncu-rep > profiles/r02_ncu_<kernel>_summary.txt
