set -x
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fp_pass_poly -s 10 -c 8 -f -o gpurun_out/r2_20_poly python bench.py --no-e2e --no-cpu --no-warm --no-others --steps 4 --warmup 3 > gpurun_out/r2_20_ncu_poly.log 2>&1
tail -1 gpurun_out/r2_20_ncu_poly.log | cut -c1-200
ls -la gpurun_out/r2_20_poly.ncu-rep
