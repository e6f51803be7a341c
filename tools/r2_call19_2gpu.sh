set -x
export PICGOLF_PEER_TIMEOUT_S=30
(nproc; lscpu | grep -E "Model name|Socket|NUMA|Thread|Core"; nvidia-smi topo -m; for d in /sys/bus/pci/devices/*; do if [ -f $d/class ] && grep -q "^0x0302" $d/class; then echo $d $(cat $d/numa_node) $(cat $d/local_cpulist); fi; done) > gpurun_out/r2_19_box_probe.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --no-warm --no-others > gpurun_out/r2_19_bench_2gpu.json 2> gpurun_out/r2_19_bench_2gpu.err; echo bench rc=$?; tail -3 gpurun_out/r2_19_bench_2gpu.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_19_bench_2gpu.json').read().strip().splitlines()[-1])
print('ms/step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], d['e2e']['host_affinity'], d['e2e']['serial_set_step_get']['ms_per_step'])
PY
head -40 gpurun_out/r2_19_box_probe.txt
