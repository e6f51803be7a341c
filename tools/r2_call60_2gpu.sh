export PICGOLF_PEER_TIMEOUT_S=30
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29516 tools/multigpu_check.py > gpurun_out/r2_60_multigpu_check_2gpu.txt 2>&1; echo check rc=$?; grep -E "MULTIGPU_CHECK|ok=False|bit-identical=False|Error" gpurun_out/r2_60_multigpu_check_2gpu.txt | head
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 40 --warmup 5 --no-others --no-e2e --no-cpu --no-warm > gpurun_out/r2_60_bench_2gpu.json 2> gpurun_out/r2_60_bench_2gpu.err; echo bench rc=$?
timeout 300 python bench.py --steps 40 --warmup 5 --no-e2e --no-cpu --no-warm --no-others > gpurun_out/r2_60_bench_1gpu.json 2>/dev/null
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_60_bench_2gpu.json').read().strip().splitlines()[-1])
o=json.loads(open('gpurun_out/r2_60_bench_1gpu.json').read().strip().splitlines()[-1])
print('2gpu ms/step', d['ms_per_step'], '1gpu', o['ms_per_step'], 'eff', o['ms_per_step']/d['ms_per_step'], d['roofline']['stage_ms_per_step'])
PY
