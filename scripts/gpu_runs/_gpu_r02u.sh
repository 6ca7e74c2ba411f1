out=gpurun_out/r02u; mkdir -p $out
timeout 900 python -m pytest tests -m gpu -x -q -k "not cfg5 and not cfg4" 2>&1 | tail -3
timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-fp64-peak > $out/bench_n1.json 2> $out/bench_n1.err; python -c "
import json
txt=open('$out/bench_n1.json').read().strip().splitlines()
d=json.loads([l for l in txt if l.startswith('{')][-1]); print('N=1', d['value'], d['stage_ms_per_step'], d['roofline']['frac'])"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 5 --warmup 3 --no-fp64-peak --no-e2e > $out/bench_n2.json 2> $out/bench_n2.err; python -c "
import json
txt=open('$out/bench_n2.json').read().strip().splitlines()
d=json.loads([l for l in txt if l.startswith('{')][-1]); print('N=2', d['value'], d['stage_ms_per_step'], d['roofline']['frac'])"
