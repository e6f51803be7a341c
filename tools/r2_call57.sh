for v in "" limbuncond; do
  if [ -n "$v" ]; then export PICGOLF_LIB=particleincellcodegolf.jl_b200/lib/variants/libpicgolf_$v.so; else unset PICGOLF_LIB; fi
  echo "== variant '$v'"
  timeout 600 python tools/f_rows_timing.py --schemes simpson_gauss,simpson_area,1d2v 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    try: d=json.loads(l)
    except: continue
    print(d['scheme'], round(d['ms_per_step'],3), 'ms/step', round(d['particle_steps_per_s']/1e9,2), 'G/s')"
done
