out=gpurun_out/r03y; mkdir -p $out
L=$PWD/b-spline-two-e_b200/lib
BS2E_LIB=$L/libbs2e_gpu_ntx512.so timeout 300 python -X faulthandler -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "both_site or small_factor or cfg3" > $out/pytest.log 2>&1; echo rc=$?; tail -2 $out/pytest.log
for v in libbs2e_gpu.so libbs2e_gpu_ntx512.so; do
BS2E_LIB=$L/$v BS2E_ONLY_BLOCKS=0,6 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"site_mma" -c 8 --csv --log-file $out/t_$v.csv python scripts/sharded_run.py cfg4 > /dev/null 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(l for l in open('$out/t_$v.csv') if l.startswith('"'))]
h=rows[0]; kn,mv=h.index('Kernel Name'),h.index('Metric Value')
print('$v', [(r[kn][24:32], round(float(r[mv].replace(',',''))/1e6,3)) for r in rows[1:]])
PY
BS2E_LIB=$L/$v timeout 200 python scripts/fill_ab.py cfg4 0,2,6 5 > $out/ab_$v.json 2>/dev/null; python -c "
import json; d=json.load(open('$out/ab_$v.json')); print('$v', d['sum_median_ms'], [round(b['frac_hbm_min'],3) for b in d['blocks']])"
done
