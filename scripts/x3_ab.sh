#!/bin/bash
# A/B of the two 3xTF32 routes of the precise layers on a B200: in-kernel split (igemm_halo.cu) vs pre-split [hi|hi|lo] copy.
mkdir -p gpurun_out
{
echo "== in-kernel (G2_X3_INKERNEL=1), cross-check of every layer against the pre-split route"
G2_X3_DEBUG=1 python scripts/parity_report.py genesisv2 --modes tf32x3 2>&1 | grep -v Warning
echo "== pre-split (G2_X3_INKERNEL=0)"
G2_X3_INKERNEL=0 python scripts/parity_report.py genesisv2 monet --modes tf32x3 2>&1 | grep -v Warning
echo "== in-kernel monet"
python scripts/parity_report.py monet --modes tf32x3 2>&1 | grep -v Warning
} > gpurun_out/r02_x3_ab.txt 2>&1
tail -5 gpurun_out/r02_x3_ab.txt
