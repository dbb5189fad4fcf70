G2_X3_DEBUG=1 python scripts/parity_report.py genesisv2 --modes tf32x3 2>&1 | grep X3DEBUG > gpurun_out/r02_x3_dbg.txt
