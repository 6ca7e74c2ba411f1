out=gpurun_out/r02x; mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "rk_row_slices or site_partition or both_site" 2>&1 | tail -3
L=$PWD/b-spline-two-e_b200/lib
for v in libbs2e_gpu.so libbs2e_gpu_full.so; do BS2E_LIB=$L/$v BS2E_FILL=mma timeout 300 python scripts/fill_ab.py cfg4 all 5 > $out/ab_cfg4_$v.json 2>$out/ab_$v.err; python -c "
import json; d=json.load(open('$out/ab_cfg4_$v.json')); print('$v cfg4', d['sum_median_ms'], d['frac_hbm'], [round(b['frac_hbm_min'],3) for b in d['blocks']])"; done
