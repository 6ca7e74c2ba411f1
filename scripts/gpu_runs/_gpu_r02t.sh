out=gpurun_out/r02t; mkdir -p $out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 5 --warmup 3 --no-fp64-peak > $out/bench_n2.json 2> $out/bench_n2.err; echo "rc=$?"; tail -3 $out/bench_n2.err; python -c "
import json; d=json.load(open('$out/bench_n2.json')); print(d['value'], d['stage_ms_per_step'], d['roofline']['frac'], d['e2e'])"
