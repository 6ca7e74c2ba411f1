out=gpurun_out/r03u; mkdir -p $out
timeout 600 ncu --set full --clock-control none -k regex:"rk_build|diag_cells|cell_moments|pair_prefix|site_count|site_enum|conf_scan|ncrow" -c 24 -o /tmp/small -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-fp64-peak > $out/ncu_small.log 2>&1; tail -1 $out/ncu_small.log | cut -c1-200
python scripts/ncu_summary.py full /tmp/small.ncu-rep > $out/small_kernels_summary.md; wc -l $out/small_kernels_summary.md
