S=c2_att64_fwd,c2_att64_dgrad,c2_att32_fwd,c2_att16_fwd,c2_att_up64_fwd,c3_unet64,c3_dec_up64
for kb in 32 64 96; do echo "== persistent everywhere, weight ring $kb KB"; G2_HALO_PERSISTENT=1 G2_HALO_RING_KB=$kb python scripts/conv_bench.py --only $S 2>&1 | grep halo; done
