show() { python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', d['fill'], 'sum %.2f ms frac %.3f' % (d['sum_median_ms'], d['frac_hbm']), [(b['block'], round(b['min_ms'],2), round(b['median_ms'],2)) for b in d['blocks']])"; }
for f in mma fma; do BS2E_FILL=$f python scripts/fill_ab.py cfg4 6 7 | show fork; BS2E_NOFORK=1 BS2E_FILL=$f python scripts/fill_ab.py cfg4 6 7 | show nofork; done
