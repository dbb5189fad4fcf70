S=c2_bdec_fwd70,c2_bdec_dgrad,c2_att64_fwd,c2_att64_dgrad,c2_att32_fwd,c2_att16_fwd,c2_att_up64_fwd,c3_unet64,c3_unet32,c3_dec_up64,c5_unet128,c5_bdec_fwd
echo "== default"; python scripts/conv_bench.py --only $S 2>&1 | grep halo
echo "== persistent"; G2_HALO_PERSISTENT=1 python scripts/conv_bench.py --only $S 2>&1 | grep halo

echo "== persistent tests"; G2_HALO_PERSISTENT=1 python -m pytest tests/test_halo_gpu.py tests/test_tc_gpu.py -q -x 2>&1 | tail -3
echo "== wgrad (halo layout, two issuers)"; python scripts/conv_bench.py --wgrad --only c2_bdec_fwd70,c2_bdec_dgrad,c2_att64_fwd,c2_att64_dgrad,c2_att32_fwd,c3_unet64,c5_unet128,c5_bdec_fwd 2>&1 | grep wgrad
echo "== wgrad tests"; python -m pytest tests/test_tc_gpu.py tests/test_halo_gpu.py -q -x -k "wgrad" 2>&1 | tail -2
