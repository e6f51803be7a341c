for part in poly solve tma; do
  timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_fused.py $part > gpurun_out/r2_55_racecheck_$part.txt 2>&1; echo "racecheck $part rc=$?"; tail -2 gpurun_out/r2_55_racecheck_$part.txt
done
PICGOLF_LOOP=0 timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_fused.py poly > gpurun_out/r2_55_racecheck_poly_noloop.txt 2>&1; echo "racecheck poly (fixed schedule) rc=$?"; tail -2 gpurun_out/r2_55_racecheck_poly_noloop.txt
