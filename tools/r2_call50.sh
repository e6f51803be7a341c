timeout 900 ncu --set full --clock-control none --import-source on -k regex:fp_pass_poly -s 10 -c 3 -f -o gpurun_out/r2_50_poly python bench.py --no-e2e --no-cpu --no-warm --no-others --steps 4 --warmup 3 > gpurun_out/r2_50_ncu_poly.log 2>&1
tail -1 gpurun_out/r2_50_ncu_poly.log | cut -c1-200
