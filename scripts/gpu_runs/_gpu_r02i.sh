out=gpurun_out/r02i; mkdir -p $out
timeout 600 python -m pytest tests -m gpu -x -q > $out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 $out/pytest_gpu.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-fp64-peak > $out/bench_cfg4.json 2> $out/bench_cfg4.err; echo "bench rc=$?"; tail -5 $out/bench_cfg4.err; python -c "
import json; d=json.load(open('$out/bench_cfg4.json')); print(d['value'], d['stage_ms_per_step'], d['roofline']['frac'], d['roofline']['per_block_frac'])"
BS2E_FILL=fma timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-fp64-peak > $out/bench_cfg4_fma.json 2> $out/bench_cfg4_fma.err; python -c "
import json; d=json.load(open('$out/bench_cfg4_fma.json')); print('fma', d['value'], d['stage_ms_per_step'], d['roofline']['frac'])"
timeout 300 python bench.py --workload cfg3 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-fp64-peak > $out/bench_cfg3.json 2> $out/bench_cfg3.err; python -c "
import json; d=json.load(open('$out/bench_cfg3.json')); print('cfg3', d['value'], d['stage_ms_per_step'], d['roofline']['frac'], d['roofline']['per_block_frac'])"
