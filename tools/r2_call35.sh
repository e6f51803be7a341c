timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/r2_35_gpu_tests.txt 2>&1; tail -4 gpurun_out/r2_35_gpu_tests.txt
grep -h "solve1d" PARITY.json | head -12
bash tools/build_variants.sh solveprof "-DPG_SOLVE_PROF" > /dev/null 2>&1
PICGOLF_LIB=particleincellcodegolf.jl_b200/lib/variants/libpicgolf_solveprof.so python tools/solve_prof.py 2>&1 | tail -5
