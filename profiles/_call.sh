timeout 120 python profiles/panel_tc32_probe.py 2>&1 | grep " us " | head -2
timeout 400 python -m pytest tests/test_panel_gpu.py tests/test_adapted_mlp_gpu.py -x -q -m gpu 2>&1 | tail -3
echo ALLDONE_MARK58
