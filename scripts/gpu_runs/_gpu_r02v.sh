out=gpurun_out/r02v; mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "both_site or small or cfg2" 2>&1 | tail -3
for f in mma fma; do BS2E_FILL=$f timeout 300 python scripts/fill_ab.py cfg4 all 5 > $out/ab_cfg4_$f.json 2>$out/ab_$f.err; python -c "
import json; d=json.load(open('$out/ab_cfg4_$f.json')); print('$f cfg4', d['sum_median_ms'], d['frac_hbm'], [round(b['frac_hbm_min'],3) for b in d['blocks']])"; done
for f in mma fma; do BS2E_FILL=$f timeout 300 python scripts/fill_ab.py cfg3 all 7 > $out/ab_cfg3_$f.json 2>>$out/ab_$f.err; python -c "
import json; d=json.load(open('$out/ab_cfg3_$f.json')); print('$f cfg3', d['sum_median_ms'], d['frac_hbm'], [round(b['frac_hbm_min'],3) for b in d['blocks']])"; done
BS2E_FILL=mma BS2E_ONLY_BLOCKS=6 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"site_mma" -c 2 -o $out/cfg4_mma -f python scripts/sharded_run.py cfg4 > $out/ncu_cfg4_mma.log 2>&1
BS2E_FILL=fma BS2E_ONLY_BLOCKS=6 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"site_fill" -c 2 -o $out/cfg4_fma -f python scripts/sharded_run.py cfg4 > $out/ncu_cfg4_fma.log 2>&1
tail -2 $out/ncu_cfg4_fma.log
