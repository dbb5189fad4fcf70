S=c2_bdec_fwd70,c2_att64_fwd,c2_att64_dgrad,c3_unet64
echo "== default"; python scripts/conv_bench.py --only $S 2>&1 | grep halo
echo "== persistent"; G2_HALO_PERSISTENT=1 python scripts/conv_bench.py --only $S 2>&1 | grep halo
echo "== persistent kouter"; G2_HALO_PERSISTENT=1 G2_HALO_KOUTER=1 python scripts/conv_bench.py --only $S 2>&1 | grep halo
echo "== persistent 2 CTAs/SM"; G2_HALO_PERSISTENT=1 G2_HALO_PERSISTENT_CTAS=296 G2_HALO_PERSISTENT_SMEM_KB=113 G2_HALO_PERSISTENT_COLS=128 python scripts/conv_bench.py --only $S 2>&1 | grep halo
echo "== persistent 2 CTAs/SM kouter"; G2_HALO_PERSISTENT=1 G2_HALO_KOUTER=1 G2_HALO_PERSISTENT_CTAS=296 G2_HALO_PERSISTENT_SMEM_KB=113 G2_HALO_PERSISTENT_COLS=128 python scripts/conv_bench.py --only $S 2>&1 | grep halo
echo "== persistent 3 CTAs/SM"; G2_HALO_PERSISTENT=1 G2_HALO_PERSISTENT_CTAS=444 G2_HALO_PERSISTENT_SMEM_KB=75 G2_HALO_PERSISTENT_COLS=64 python scripts/conv_bench.py --only $S 2>&1 | grep halo
