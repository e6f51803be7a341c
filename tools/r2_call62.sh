for v in "" unroll2 unroll4 stages3 stages6 ""; do
  if [ -n "$v" ]; then export PICGOLF_LIB=particleincellcodegolf.jl_b200/lib/variants/libpicgolf_$v.so; else unset PICGOLF_LIB; fi
  timeout 300 python bench.py --steps 40 --warmup 8 --no-cpu --no-e2e --no-warm --no-others 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('variant [$v] gauss ms/step', round(d['ms_per_step'],4), 'pass', round(d['roofline']['launch_ms'],4))"
done
