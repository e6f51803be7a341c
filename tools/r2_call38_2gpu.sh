set -x
export PICGOLF_PEER_TIMEOUT_S=30
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/multigpu_check.py > gpurun_out/r2_38_multigpu_check_2gpu.txt 2>&1; echo check rc=$?; grep -E "MULTIGPU_CHECK|ok=False|bit-identical=False|Error" gpurun_out/r2_38_multigpu_check_2gpu.txt | head
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --no-others --no-e2e > gpurun_out/r2_38_bench_2gpu.json 2> gpurun_out/r2_38_bench_2gpu.err; echo bench rc=$?; tail -3 gpurun_out/r2_38_bench_2gpu.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_38_bench_2gpu.json').read().strip().splitlines()[-1])
print('ms/step', d['ms_per_step'], 'value', d['value'])
print('warm', {k:(v['ms_per_step'], v['resorts'], v['resorts_fused_into_the_passes']) for k,v in d['warm_regime'].items() if isinstance(v,dict)})
PY
