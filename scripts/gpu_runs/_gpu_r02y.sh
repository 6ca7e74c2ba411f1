out=gpurun_out/r02y; mkdir -p $out
timeout 600 python -X faulthandler -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "rk_row_slices or site_partition or both_site" > $out/pytest.log 2>&1; tail -3 $out/pytest.log
for extra in "" "--whole-rk --rebalance 0"; do
tag=$(echo $extra | tr -d ' -'); 
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 5 --warmup 3 --no-fp64-peak --no-e2e $extra > $out/bench_n2_$tag.json 2> $out/bench_n2_$tag.err; python -c "
import json
txt=open('$out/bench_n2_$tag.json').read().strip().splitlines()
d=json.loads([l for l in txt if l.startswith('{')][-1]); print('N=2 $extra', d['value'], d['stage_ms_per_step'], d['per_rank_ms_per_step'], d['partition'])"
done
