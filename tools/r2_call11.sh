set -x
timeout 900 python -m pytest tests -m gpu -x -q -k "2d3v or 2d_stage or streamed" > gpurun_out/r2_11_tests.txt 2>&1; tail -6 gpurun_out/r2_11_tests.txt
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests -m gpu -x -q -k "variants_match and stream48_512" > gpurun_out/r2_11_memcheck.txt 2>&1; echo memcheck rc=$?; tail -4 gpurun_out/r2_11_memcheck.txt
for v in stream stream44_512 stream82_512 stream28_512 stream11_512 stream44_768 stream24_1024 ring42; do
  PICGOLF_2D_KERNEL=$v timeout 300 python bench.py --workload 2d3v --no-e2e --no-cpu --steps 20 --warmup 5 > gpurun_out/r2_11_2d3v_$v.json 2> gpurun_out/r2_11_2d3v_$v.err
  python -c "
import json
d=json.loads(open('gpurun_out/r2_11_2d3v_$v.json').read().strip().splitlines()[-1]); print('$v', round(d['ms_per_step'],3), round(d['roofline']['launch_ms'],3), round(d['roofline']['frac'],3), round(d['hbm_roofline_frac_step'],3), d['sorts_in_timed_region'], d['roofline']['stage_ms_per_step'])"
done
timeout 400 ncu --set full --clock-control none --import-source on -k regex:particles_2d3v_stream -s 4 -c 1 -f -o gpurun_out/r2_11_stream python bench.py --workload 2d3v --no-e2e --no-cpu --steps 3 --warmup 3 > gpurun_out/r2_11_ncu.log 2>&1
tail -2 gpurun_out/r2_11_ncu.log | cut -c1-300
