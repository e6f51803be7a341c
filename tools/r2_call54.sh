for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python tools/sanitize_fused.py > gpurun_out/r2_54_${tool}.txt 2>&1; echo "$tool rc=$?"; tail -3 gpurun_out/r2_54_${tool}.txt
done
