TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 $TR --nproc-per-node 8 --master-port 29513 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2g_n8.json 2> gpurun_out/r2g_n8.err
echo "N=8 rc=$?"; tail -n 1 gpurun_out/r2g_n8.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('   ', round(d['value']/1e6,2), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']/1e6,2), 'exposed', d['comm']['ms_exposed_per_step']); oc=d['other_configs']; print(json.dumps({k:v for k,v in oc['train_projector_B1024_global_dp'].items() if k!='what'})); print(json.dumps({k:v for k,v in oc['hypernet_path_dp_ga'].items() if k!='what'}))" 2>&1 | tail -4
echo ALLDONE_MARK55
