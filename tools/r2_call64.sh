timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2_64_launches_gauss_fp.csv python bench.py --no-e2e --no-cpu --no-warm --no-others --steps 20 --warmup 5 > gpurun_out/r2_64_ncu_gauss.log 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/r2_64_launches_gauss_fp.csv')) if len(r)>5 and r[0].isdigit()]
tot=collections.Counter(); cnt=collections.Counter()
for r in rows[-400:]:
    n=r[4].split('(')[0]; tot[n]+=int(r[-1].replace(',','')); cnt[n]+=1
T=sum(tot.values())
for n,v in tot.most_common(10): print(f"{n[:50]:50s} {cnt[n]:4d} launches {v/1e6:9.3f} ms  share {v/T:.3f}")
PY
