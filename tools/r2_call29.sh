export PICGOLF_LOOP=0
for cfg in "0.0 1" "0.0 16" "0.3 1"; do set -- $cfg
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_29_launches_vth$1_se$2.csv python tools/fused_sort_launches.py $1 $2 > gpurun_out/r2_29_ncu_$1_$2.log 2>&1
echo "== vth $1 sort_every $2"
python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2_29_launches_vth$1_se$2.csv')) if len(r)>5 and r[0].isdigit()]
rows=[r for r in rows if int(r[-1].replace(',',''))>8000]
for r in rows[-22:]:
    print(r[4][:60].ljust(60), r[-1])
PY
done
