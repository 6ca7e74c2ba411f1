show() { python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', d['fill'], 'sum %.2f ms frac %.3f' % (d['sum_median_ms'], d['frac_hbm']), [(b['block'], round(b['median_ms'],2)) for b in d['blocks']])"; }
for wl in cfg4 cfg3 cfg1; do
BS2E_FILL=mma python scripts/fill_ab.py $wl all 5 | show "$wl fork"
BS2E_NOFORK=1 BS2E_FILL=mma python scripts/fill_ab.py $wl all 5 | show "$wl nofork"
BS2E_FILL=fma python scripts/fill_ab.py $wl all 5 | show "$wl fork"
done
