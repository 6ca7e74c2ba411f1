out=gpurun_out/r03i; mkdir -p $out
for tag in a b c; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 10 --warmup 3 --no-fp64-peak --no-e2e > $out/bench_n2_$tag.json 2> $out/bench_n2_$tag.err; python -c "
import json
txt=open('$out/bench_n2_$tag.json').read().strip().splitlines()
d=json.loads([l for l in txt if l.startswith('{')][-1]); print('N=2 $tag', d['value'], d['stage_ms_per_step']['A_cells'], d['stage_ms_per_step']['B_rk'], d['per_rank_ms_per_step']['C_blocks_ms_of_every_step'])"
done
