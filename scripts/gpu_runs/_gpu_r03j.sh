out=gpurun_out/r03j; mkdir -p $out
run() { n=$1; tag=$2; shift 2; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $n "$@" > $out/$tag.json 2> $out/$tag.err; python -c "
import json
txt=open('$out/$tag.json').read().strip().splitlines()
d=json.loads([l for l in txt if l.startswith('{')][-1]); print('$tag', d.get('value'), d.get('stage_ms_per_step'), (d.get('e2e') or {}).get('value'))"; }
run 8 cfg4_n8 --steps 10 --warmup 3 --no-fp64-peak
run 4 cfg4_n4 --steps 10 --warmup 3 --no-fp64-peak --no-e2e
run 8 cfg5_n8 --workload cfg5 --steps 3 --warmup 3 --no-fp64-peak --no-e2e
run 8 ref_n8 --impl reference --steps 2 --warmup 1
