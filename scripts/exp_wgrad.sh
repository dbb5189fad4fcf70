echo "== wgrad (all shapes)"; python scripts/conv_bench.py --wgrad 2>&1 | grep wgrad
python -m pytest tests/test_tc_gpu.py tests/test_halo_gpu.py tests/test_ops_gpu.py -q -x 2>&1 | tail -2
