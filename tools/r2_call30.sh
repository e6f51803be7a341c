timeout 1200 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "fused or poly or sorted or config4 or c2 or checkpoint or streamed" > gpurun_out/r2_30_tests.txt 2>&1; tail -5 gpurun_out/r2_30_tests.txt
timeout 900 python tools/fused_sort_timing.py 28 > gpurun_out/r2_30_fused_sort_timing.txt 2>&1; head -20 gpurun_out/r2_30_fused_sort_timing.txt
export PICGOLF_LOOP=0
for cfg in "0.0 1"; do set -- $cfg
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_30_launches_vth$1_se$2.csv python tools/fused_sort_launches.py $1 $2 > gpurun_out/r2_30_ncu_$1_$2.log 2>&1
echo "== vth $1 sort_every $2"
python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2_30_launches_vth$1_se$2.csv')) if len(r)>5 and r[0].isdigit()]
rows=[r for r in rows if int(r[-1].replace(',',''))>2500]
for r in rows[-12:]:
    print(r[4][:60].ljust(60), r[-1])
PY
done
