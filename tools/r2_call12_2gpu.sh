set -x
export PICGOLF_PEER_TIMEOUT_S=20
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/multigpu_check.py > gpurun_out/r2_12_multigpu_check.txt 2>&1; echo check rc=$?; tail -25 gpurun_out/r2_12_multigpu_check.txt
timeout 600 python -m pytest tests/test_multigpu.py -m gpu -x -q > gpurun_out/r2_12_test_multigpu.txt 2>&1; tail -5 gpurun_out/r2_12_test_multigpu.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_12_bench_2gpu.json 2> gpurun_out/r2_12_bench_2gpu.err; echo bench rc=$?; tail -5 gpurun_out/r2_12_bench_2gpu.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_12_bench_2gpu.json').read().strip().splitlines()[-1])
print('ms/step', d['ms_per_step'], 'value', d['value'], 'launches', d['gpu_launches'], d['config']['parallelism'], d['roofline']['stage_ms_per_step'])
print('e2e', json.dumps(d['e2e'])[:600])
print('others', json.dumps(d['other_workloads'])[:1500])
PY
timeout 300 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu --no-warm --no-others > gpurun_out/r2_12_bench_1gpu.json 2>/dev/null
python -c "
import json
d=json.loads(open('gpurun_out/r2_12_bench_1gpu.json').read().strip().splitlines()[-1]); print('1gpu ms/step', d['ms_per_step'])"
