S=c3_unet64,c3_unet32,c5_unet128,c5_bdec_fwd,c2_bdec_fwd70,c3_dec_up64
echo "== 3xTF32, persistent kernel, streamed weights"; G2_HALO_X3_RESIDENT=0 python scripts/conv_bench.py --x3 --only $S 2>&1 | grep halo
echo "== 3xTF32, persistent kernel, streamed weights, 16 KB ring"; G2_HALO_X3_RESIDENT=0 G2_HALO_RING_KB=16 python scripts/conv_bench.py --x3 --only $S 2>&1 | grep halo
