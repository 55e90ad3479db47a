set -x
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -8
timeout 600 python bench.py --steps 20 --warmup 5 --no-sweep --no-llm > gpurun_out/r2_bench9_pdl.json 2> gpurun_out/r2_bench9.err; tail -c 600 gpurun_out/r2_bench9.err
timeout 600 python bench.py --steps 20 --warmup 5 --no-sweep --no-llm --no-pdl > gpurun_out/r2_bench9_nopdl.json 2>> gpurun_out/r2_bench9.err
timeout 600 python bench.py --steps 200 --warmup 10 --no-sweep --no-llm --no-extras --no-e2e --no-cpu-baseline > gpurun_out/r2_bench9_pdl200.json 2>> gpurun_out/r2_bench9.err
python - <<'PY'
import json
for f in ['pdl','nopdl','pdl200']:
    try:
        d=json.loads(open(f'gpurun_out/r2_bench9_{f}.json').read().strip().splitlines()[-1])
    except Exception as e:
        print(f,'FAILED',e); continue
    print(f, d['value'], d['ms_per_step'], d.get('clocks'))
    oc=d.get('other_configs',{}).get('hypernet_microstep_B4_K128',{})
    print('   ', {k:v for k,v in oc.items() if k.startswith('ms_')})
    print('   e2e', d.get('e2e'))
PY
echo ALLDONE_MARK19
