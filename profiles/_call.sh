timeout 900 python -m pytest tests/test_gemm_gpu.py tests/test_hypernet_gpu.py tests/test_graphs_gpu.py tests/test_adapted_mlp_gpu.py -x -q -m gpu 2>&1 | tail -4
timeout 900 python bench.py --steps 20 --warmup 5 --no-llm --no-e2e --no-sweep --no-cpu-baseline --no-gpu-eager --no-kernel-breakdown > gpurun_out/r2_bench15.json 2> gpurun_out/r2_bench15.err; tail -c 400 gpurun_out/r2_bench15.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench15.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'])
oc=d['other_configs']
print({k:v for k,v in oc['hypernet_microstep_B4_K128'].items() if k.startswith('ms')})
PY
echo ALLDONE_MARK42
