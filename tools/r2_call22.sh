set -x
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests -m gpu -x -q -k "variants_match and stream-" > gpurun_out/r2_22_racecheck_2d.txt 2>&1; echo rc=$?; tail -6 gpurun_out/r2_22_racecheck_2d.txt
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_zz_esfield_gpu.py -m gpu -x -q -k "test_loop_matches_oracle and 3" > gpurun_out/r2_22_racecheck_es.txt 2>&1; echo rc=$?; tail -6 gpurun_out/r2_22_racecheck_es.txt
timeout 900 compute-sanitizer --tool synccheck --error-exitcode 9 python -m pytest tests -m gpu -x -q -k "variants_match and stream-" > gpurun_out/r2_22_synccheck_2d.txt 2>&1; echo rc=$?; tail -4 gpurun_out/r2_22_synccheck_2d.txt
