S=c2_bdec_fwd70,c2_att64_fwd,c2_att64_dgrad,c2_att32_fwd,c3_unet64,c5_bdec_fwd
echo "== staged"; python scripts/conv_bench.py --only $S 2>&1 | grep halo
echo "== direct"; G2_HALO_EPI=2 python scripts/conv_bench.py --only $S 2>&1 | grep halo
