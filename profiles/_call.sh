timeout 300 python profiles/plain_probe.py 1024 2>&1 | tail -3
timeout 900 python -m pytest tests/test_modules_gpu.py tests/test_gemm_gpu.py tests/test_adapted_mlp_gpu.py -x -q -m gpu 2>&1 | tail -3
echo ALLDONE_MARK37
