out=gpurun_out/r03h; mkdir -p $out
timeout 1500 python -X faulthandler -m pytest tests -m gpu -x -q > $out/pytest_gpu.log 2>&1; tail -4 $out/pytest_gpu.log
for t in a b; do
timeout 600 python bench.py --steps 8 --warmup 3 --no-e2e --no-cpu-baseline --no-fp64-peak > $out/bench_n1_$t.json 2> $out/bench_n1_$t.err; python -c "
import json
txt=open('$out/bench_n1_$t.json').read().strip().splitlines()
d=json.loads([l for l in txt if l.startswith('{')][-1]); print('N=1', d['value'], d['stage_ms_per_step'], d['roofline']['frac'])"
done
