out=gpurun_out/r03e; mkdir -p $out
for t in a b c; do
BS2E_BENCH_TRACE=1 timeout 600 python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu-baseline --no-fp64-peak > $out/bench_n1_$t.json 2> $out/bench_n1_$t.err; python -c "
import json
txt=open('$out/bench_n1_$t.json').read().strip().splitlines()
d=json.loads([l for l in txt if l.startswith('{')][-1]); print('N=1', d['value'], d['stage_ms_per_step']['C_blocks_of_every_step'])"
grep "trace rank 0 step 0" $out/bench_n1_$t.err | cut -c1-600
done
