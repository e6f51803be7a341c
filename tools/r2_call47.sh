timeout 900 ncu --set full --clock-control none --import-source on -k regex:fp_pass_poly -s 10 -c 3 -f -o gpurun_out/r2_47_poly python bench.py --no-e2e --no-cpu --no-warm --no-others --steps 4 --warmup 3 > gpurun_out/r2_47_ncu_poly.log 2>&1
tail -1 gpurun_out/r2_47_ncu_poly.log | cut -c1-200
export PICGOLF_LOOP=0
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fp_pass_poly -s 9 -c 3 -f -o gpurun_out/r2_47_poly_fused python tools/fused_sort_launches.py 0.3 1 > gpurun_out/r2_47_ncu_poly_fused.log 2>&1
tail -1 gpurun_out/r2_47_ncu_poly_fused.log | cut -c1-200
ls -la gpurun_out/r2_47_*.ncu-rep
