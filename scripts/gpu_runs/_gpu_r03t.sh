out=gpurun_out/r03t; mkdir -p $out
L=$PWD/b-spline-two-e_b200/lib
for v in libbs2e_gpu.so libbs2e_gpu_stwb.so libbs2e_gpu_stcg.so; do BS2E_LIB=$L/$v timeout 300 python scripts/fill_ab.py cfg4 0,4,6 5 > $out/ab_$v.json 2>$out/ab_$v.err; python -c "
import json; d=json.load(open('$out/ab_$v.json')); print('$v', d['sum_median_ms'], d['frac_hbm'], [round(b['frac_hbm_min'],3) for b in d['blocks']])"; done
