set -x
timeout 900 python -m pytest tests/test_hypernet_gpu.py tests/test_modules_gpu.py tests/test_graphs_gpu.py tests/test_gemm_gpu.py -x -q -m gpu 2>&1 | tail -15
timeout 600 python bench.py --steps 20 --warmup 5 --no-sweep > gpurun_out/r2_bench8.json 2> gpurun_out/r2_bench8.err; tail -c 600 gpurun_out/r2_bench8.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench8.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'])
def walk(o,p=''):
    if isinstance(o,dict):
        for k,v in o.items(): walk(v,p+'/'+k)
    elif isinstance(o,(int,float,str)) and ('hyper' in p or 'micro' in p): print(p,o)
walk(d)
PY
echo ALLDONE_MARK18
