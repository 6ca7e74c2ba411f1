out=gpurun_out/r03d; mkdir -p $out
timeout 900 python -X faulthandler -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "stage_A or stage_B or cfg1_Rk or rk_row_slices or cfg3 or cfg4_stage_AB" > $out/pytest.log 2>&1; tail -3 $out/pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-fp64-peak > $out/bench_n1.json 2> $out/bench_n1.err; python -c "
import json
txt=open('$out/bench_n1.json').read().strip().splitlines()
d=json.loads([l for l in txt if l.startswith('{')][-1]); print('N=1', d['value'], d['stage_ms_per_step'], d['roofline']['frac'], d['roofline']['rk_build'])"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"rk_build" -c 1 -o $out/rk_build -f python scripts/sharded_run.py cfg4 > $out/ncu_rk.log 2>&1; tail -1 $out/ncu_rk.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"dip_fill" -s 4 -c 4 -o $out/dip_fill -f python scripts/dipole_probe.py cfg3 v > $out/ncu_dip.log 2>&1; tail -2 $out/ncu_dip.log
for kb in 12 16; do BS2E_SITE_CHUNK_KB=$kb timeout 300 python scripts/fill_ab.py cfg4 4,6 5 > $out/ab_chunk$kb.json 2>/dev/null; python -c "
import json; d=json.load(open('$out/ab_chunk$kb.json')); print('chunk $kb', d['sum_median_ms'], d['frac_hbm'])"; done
timeout 300 python scripts/fill_ab.py cfg4 4,6 5 > $out/ab_chunk8.json 2>/dev/null; python -c "
import json; d=json.load(open('$out/ab_chunk8.json')); print('chunk 8', d['sum_median_ms'], d['frac_hbm'])"
