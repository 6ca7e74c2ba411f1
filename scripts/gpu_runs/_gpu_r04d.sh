out=gpurun_out/r04d; mkdir -p $out
timeout 200 python scripts/fill_ab.py cfg4 all 5 > $out/ab_fork.json 2>/dev/null; python -c "
import json; d=json.load(open('$out/ab_fork.json')); print('fork  ', d['sum_median_ms'], [round(b['frac_hbm_min'],3) for b in d['blocks']])"
BS2E_NOFORK=1 timeout 200 python scripts/fill_ab.py cfg4 all 5 > $out/ab_nofork.json 2>/dev/null; python -c "
import json; d=json.load(open('$out/ab_nofork.json')); print('nofork', d['sum_median_ms'], [round(b['frac_hbm_min'],3) for b in d['blocks']])"
