bash tools/build_variants.sh solveprof "-DPG_SOLVE_PROF" > /dev/null 2>&1
PICGOLF_LIB=particleincellcodegolf.jl_b200/lib/variants/libpicgolf_solveprof.so python tools/solve_prof.py 2>&1 | tail -6 | tee gpurun_out/r2_37_solve_prof.txt
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/r2_37_gpu_tests.txt 2>&1; tail -4 gpurun_out/r2_37_gpu_tests.txt
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/r2_37_bench.json 2> gpurun_out/r2_37_bench.err; tail -3 gpurun_out/r2_37_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_37_bench.json').read().strip().splitlines()[-1])
print('ms/step', d['ms_per_step'], 'launches', d['gpu_launches'], 'roof', d['roofline']['frac'], d['roofline']['launch_ms'], d['hbm_roofline_frac_step'], d['roofline']['stage_ms_per_step'])
print('warm', {k:(v['ms_per_step'], v['resorts'], v['resorts_fused_into_the_passes']) for k,v in d['warm_regime'].items() if isinstance(v,dict)})
print('others', {k:(v['ms_per_step'], v['hbm_roofline_frac_step']) for k,v in d['other_workloads'].items()})
PY
