out=gpurun_out/r03f; mkdir -p $out
timeout 900 python -X faulthandler -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "stage_B or cfg1_Rk or rk_row_slices or cfg3" > $out/pytest.log 2>&1; tail -2 $out/pytest.log
for t in a b; do
timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-fp64-peak > $out/bench_n1_$t.json 2> $out/bench_n1_$t.err; python -c "
import json
txt=open('$out/bench_n1_$t.json').read().strip().splitlines()
d=json.loads([l for l in txt if l.startswith('{')][-1]); print('N=1', d['value'], d['stage_ms_per_step'], d['roofline']['frac'], d['roofline']['rk_build']['frac'])"
done
