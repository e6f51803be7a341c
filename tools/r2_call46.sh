for v in "" mod1fast "" mod1fast; do
  if [ -n "$v" ]; then export PICGOLF_LIB=particleincellcodegolf.jl_b200/lib/variants/libpicgolf_$v.so; else unset PICGOLF_LIB; fi
  timeout 600 python bench.py --steps 40 --warmup 8 --no-cpu --no-e2e --no-warm --no-others 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('variant [$v] gauss ms/step', d['ms_per_step'], d['roofline']['launch_ms'])"
done
