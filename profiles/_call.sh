timeout 900 python -m pytest tests/test_hypernet_gpu.py tests/test_graphs_gpu.py -x -q -m gpu 2>&1 | tail -4
timeout 300 python profiles/hyper_stage_probe.py 2>&1 | tail -6
echo ALLDONE_MARK48
