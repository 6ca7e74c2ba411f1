out=gpurun_out/r03c; mkdir -p $out
timeout 900 python -X faulthandler -m pytest tests/test_dipole.py -m gpu -x -q > $out/pytest_dip.log 2>&1; tail -3 $out/pytest_dip.log
BS2E_BENCH_TRACE=1 timeout 600 python bench.py --steps 6 --warmup 3 --no-e2e --no-cpu-baseline --no-fp64-peak > $out/bench_n1.json 2> $out/bench_n1.err; python -c "
import json
txt=open('$out/bench_n1.json').read().strip().splitlines()
d=json.loads([l for l in txt if l.startswith('{')][-1]); print('N=1', d['value'], d['stage_ms_per_step'], d['roofline']['frac'])"
grep trace $out/bench_n1.err | cut -c1-400 | head -3
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"dip_fill" -s 20 -c 3 -o $out/dip_fill -f python scripts/dipole_probe.py cfg3 v > $out/ncu_dip.log 2>&1; tail -3 $out/ncu_dip.log
