out=gpurun_out/r03k; mkdir -p $out
L=$PWD/b-spline-two-e_b200/lib
timeout 600 python -X faulthandler -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "both_site or small_factor or cfg2 or cfg3" > $out/pytest.log 2>&1; tail -2 $out/pytest.log
for v in libbs2e_gpu.so libbs2e_gpu_nopf.so libbs2e_gpu_skiprow.so; do BS2E_LIB=$L/$v timeout 300 python scripts/fill_ab.py cfg4 all 5 > $out/ab_cfg4_$v.json 2>$out/ab_$v.err; python -c "
import json; d=json.load(open('$out/ab_cfg4_$v.json')); print('$v cfg4', d['sum_median_ms'], d['frac_hbm'], [round(b['frac_hbm_min'],3) for b in d['blocks']])"; done
for v in libbs2e_gpu.so libbs2e_gpu_skiprow.so; do BS2E_LIB=$L/$v timeout 300 python scripts/fill_ab.py cfg3 all 7 > $out/ab_cfg3_$v.json 2>>$out/ab_$v.err; python -c "
import json; d=json.load(open('$out/ab_cfg3_$v.json')); print('$v cfg3', d['sum_median_ms'], d['frac_hbm'])"; done
