s=$(date +%s)
timeout 1200 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_final_n1_s20.json 2> gpurun_out/r2_final_n1_s20.err; echo "rc=$? wall=$(( $(date +%s) - s ))s"; tail -c 300 gpurun_out/r2_final_n1_s20.err
s=$(date +%s)
timeout 1200 python bench.py > gpurun_out/r2_final_n1_default.json 2> gpurun_out/r2_final_n1_default.err; echo "rc=$? wall=$(( $(date +%s) - s ))s"
s=$(date +%s)
timeout 1200 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_final_ref.json 2> gpurun_out/r2_final_ref.err; echo "rc=$? wall=$(( $(date +%s) - s ))s"; tail -c 600 gpurun_out/r2_final_ref.json
python - <<'PY'
import json
for f in ['r2_final_n1_s20','r2_final_n1_default']:
    d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
    print(f, d['value'], d['ms_per_step'], d['steps'], d.get('clocks'), 'launches', d.get('gpu_launches'))
    print('  roofline', json.dumps(d.get('roofline'))[:400])
    print('  sustained', json.dumps(d.get('sustained'))[:300])
    print('  e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e'].get('h2d_gbs'))
    print('  cpu', json.dumps(d.get('cpu_baseline'))[:300])
    print('  h1', json.dumps(d.get('as_written_h1'))[:300])
    print('  eager', json.dumps(d.get('gpu_eager_baseline'))[:500])
    oc=d['other_configs']
    for k in ('hypernet_microstep_B4_K128','v4_microstep_with_llama1b','train_projector_B1024_fwd_bwd','fewshot_merged_projector_B256_fwd_bwd','fewshot_generate_4_adapters_and_merge','fewshot_mean_adapter_16_sets','hypernet_forward_K128','optimizer_step_hypernet'):
        print('  ',k, json.dumps({a:b for a,b in oc.get(k,{}).items() if a!='what'})[:500])
PY
echo ALLDONE_MARK43
