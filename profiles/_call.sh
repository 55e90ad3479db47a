timeout 900 python -m pytest tests/test_hypernet_gpu.py tests/test_graphs_gpu.py -x -q -m gpu 2>&1 | tail -4
timeout 300 python profiles/hyper_microstep_probe.py 2>&1 | tail -2 || exit 1
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_hyper_microstep_launches.csv python profiles/hyper_microstep_probe.py > gpurun_out/hyper_ncu.log 2>&1
tail -2 gpurun_out/hyper_ncu.log
timeout 900 python bench.py --steps 20 --warmup 5 --no-llm --no-e2e --no-sweep --no-cpu-baseline --no-gpu-eager --no-kernel-breakdown > gpurun_out/r2_bench14.json 2> gpurun_out/r2_bench14.err; tail -c 400 gpurun_out/r2_bench14.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench14.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'])
oc=d['other_configs']
print({k:v for k,v in oc['hypernet_microstep_B4_K128'].items() if k.startswith('ms')})
print(oc['hypernet_forward_K128'])
PY
echo ALLDONE_MARK41
