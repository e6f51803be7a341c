export PICGOLF_PEER_TIMEOUT_S=30
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 --steps 20 --warmup 5 --no-others --no-e2e --no-cpu > gpurun_out/r2_66_bench_2gpu.json 2> gpurun_out/r2_66_bench_2gpu.err; echo bench rc=$?; tail -3 gpurun_out/r2_66_bench_2gpu.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_66_bench_2gpu.json').read().strip().splitlines()[-1])
print('ms/step', d['ms_per_step'], 'value', d['value'], 'sorts in timed region', d['sorts_in_timed_region'])
print('warm', {k:(v['ms_per_step'], v['resorts'], v['resorts_fused_into_the_passes']) for k,v in d['warm_regime'].items() if isinstance(v,dict)})
PY
