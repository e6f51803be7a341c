bash tools/build_variants.sh solveprof "-DPG_SOLVE_PROF" > /dev/null 2>&1
PICGOLF_LIB=particleincellcodegolf.jl_b200/lib/variants/libpicgolf_solveprof.so python tools/solve_prof.py 2>&1 | tail -5
timeout 600 python -m pytest tests -m gpu -q -x -k "solve1d or c1_ngp or c2_ or growth" 2>&1 | tail -3
