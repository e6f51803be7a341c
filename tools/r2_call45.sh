for v in "" head mod1fast "" head mod1fast; do
  if [ -n "$v" ]; then export PICGOLF_LIB=particleincellcodegolf.jl_b200/lib/variants/libpicgolf_$v.so; else unset PICGOLF_LIB; fi
  timeout 600 python bench.py --workload ngp --steps 40 --warmup 5 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('variant [$v] ngp ms/step', d['ms_per_step'], 'particles stage/step', d['roofline']['stage_ms_per_step']['particles'], d['gpu_launches'])"
done
