set -x
export PICGOLF_PEER_TIMEOUT_S=30
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2_61_bench_8gpu.json 2> gpurun_out/r2_61_bench_8gpu.err; echo bench8 rc=$?; tail -5 gpurun_out/r2_61_bench_8gpu.err
timeout 300 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu --no-warm > gpurun_out/r2_61_bench_1gpu.json 2>/dev/null
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_61_bench_8gpu.json').read().strip().splitlines()[-1])
o=json.loads(open('gpurun_out/r2_61_bench_1gpu.json').read().strip().splitlines()[-1])
print('8gpu ms/step', d['ms_per_step'], 'value', d['value'], 'launches', d['gpu_launches'], d['config']['parallelism'], d['roofline']['stage_ms_per_step'], d['clocks'])
print('1gpu ms/step', o['ms_per_step'], 'eff', o['ms_per_step']/d['ms_per_step'])
print('e2e', d['e2e']['ms_per_step'], d['e2e']['value']); print('warm', {k:(v['ms_per_step'], v['resorts']) for k,v in d['warm_regime'].items() if isinstance(v,dict)})
for k in ('ngp','2d3v'):
    print(k, '8gpu', d['other_workloads'][k]['ms_per_step'], '1gpu', o['other_workloads'][k]['ms_per_step'], 'eff', o['other_workloads'][k]['ms_per_step']/d['other_workloads'][k]['ms_per_step'])
PY
