timeout 900 python bench.py --steps 20 --warmup 5 --no-llm --no-e2e --no-cpu-baseline --no-gpu-eager --no-kernel-breakdown > gpurun_out/r2_bench11_sweep.json 2> gpurun_out/r2_bench11.err; tail -c 400 gpurun_out/r2_bench11.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench11_sweep.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'])
sw=d['other_configs']['configs4_dim_sweep']
for r in sw['batch_sweep_D768_r32']: print(r)
for r in sw['points']: print(r['D'],r['r'],round(r['ms_per_step'],4),round(r['samples_per_s']/1e6,2),round(r['frac_of_bf16_burst_peak'],3))
PY
echo ALLDONE_MARK33
