out=gpurun_out/r03w; mkdir -p $out
timeout 300 python -X faulthandler -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "both_site or small_factor" > $out/pytest_a.log 2>&1; echo "rc=$?"; tail -2 $out/pytest_a.log
timeout 400 python -X faulthandler -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "cfg2 or cfg3 or cfg4_scale or site_partition or cfg5" > $out/pytest_b.log 2>&1; echo "rc=$?"; tail -2 $out/pytest_b.log
timeout 200 python scripts/fill_ab.py cfg4 all 5 > $out/ab_cfg4.json 2>$out/ab.err; python -c "
import json; d=json.load(open('$out/ab_cfg4.json')); print('cfg4', d['sum_median_ms'], d['frac_hbm'], [round(b['frac_hbm_min'],3) for b in d['blocks']])"
timeout 200 python scripts/fill_ab.py cfg3 all 7 > $out/ab_cfg3.json 2>>$out/ab.err; python -c "
import json; d=json.load(open('$out/ab_cfg3.json')); print('cfg3', d['sum_median_ms'], d['frac_hbm'], [round(b['frac_hbm_min'],3) for b in d['blocks']])"
