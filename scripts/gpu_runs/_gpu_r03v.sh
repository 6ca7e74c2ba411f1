out=gpurun_out/r03v; mkdir -p $out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 20 --warmup 3 > $out/bench_n2.json 2> $out/bench_n2.err; echo rc=$?; python -c "
import json
txt=open('$out/bench_n2.json').read().strip().splitlines()
d=json.loads([l for l in txt if l.startswith('{')][-1]); print('N=2', d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e'].get('pageable',{}).get('value'), d['roofline']['frac'], d['roofline']['traffic'])"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --impl reference --gpus 2 --steps 20 --warmup 3 > $out/ref_n2.json 2> $out/ref_n2.err; echo rc=$?; tail -c 200 $out/ref_n2.json
