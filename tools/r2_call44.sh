timeout 600 python -m pytest tests -m gpu -q -x -k "ngp or leapfrog or explicit_gaussian or mod" 2>&1 | tail -2
bash tools/r2_call43.sh
