out=gpurun_out/r03x; mkdir -p $out
BS2E_TRACE=1 timeout 300 python bench.py --steps 8 --warmup 3 --no-e2e --no-cpu-baseline --no-fp64-peak > $out/bench_a.json 2> $out/bench_a.err
grep -c "out_take: no cached" $out/bench_a.err; grep "out_take: no cached" $out/bench_a.err | tail -12
python -c "
import json
d=json.loads([l for l in open('$out/bench_a.json') if l.startswith('{')][-1]); print('default', d['value'], d['stage_ms_per_step']['C_blocks_of_every_step'])"
BS2E_OUT_CACHE_FRAC=0.7 timeout 300 python bench.py --steps 8 --warmup 3 --no-e2e --no-cpu-baseline --no-fp64-peak > $out/bench_b.json 2> $out/bench_b.err
python -c "
import json
d=json.loads([l for l in open('$out/bench_b.json') if l.startswith('{')][-1]); print('frac0.7', d['value'], d['stage_ms_per_step']['C_blocks_of_every_step'])"
