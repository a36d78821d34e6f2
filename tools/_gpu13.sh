N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
timeout 600 $TR tools/boot_multi.py 1e8 2048 2>&1 | grep "^{" | tail -1
timeout 900 $TR bench.py --gpus $N --configs stats,lsm --steps 3 > gpurun_out/bench_r2_${N}gpu_b.json 2> gpurun_out/bench_r2_${N}gpu_b.err
grep "^{" gpurun_out/bench_r2_${N}gpu_b.json | python -c "
import json,sys
j=json.loads(sys.stdin.read())
print('value %.4e ms %.1f e2e_ms %.1f'%(j['value'], j['ms_per_step'], j['e2e']['ms_per_step']))
for k,v in j['configs'].items(): print(k, json.dumps(v)[:1500])
"
