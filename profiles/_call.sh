TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for rep in 1 2; do
  timeout 600 $TR --nproc-per-node 4 --master-port 29512 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/r2m_n4_$rep.json 2> gpurun_out/r2m_n4_$rep.err
  echo "N=4 rep $rep rc=$?"; tail -n 1 gpurun_out/r2m_n4_$rep.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('   ', round(d['value']/1e6,2), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']/1e6,2), list(d['other_configs'].keys()))" 2>&1 | tail -1
done
echo ALLDONE_MARK54
