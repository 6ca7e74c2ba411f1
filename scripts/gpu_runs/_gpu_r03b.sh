out=gpurun_out/r03b; mkdir -p $out
timeout 900 python -X faulthandler -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "stage_A or stage_B or cfg1_Rk or rk_row_slices or cfg3 or cfg4_stage_AB or stage_C_blocks" > $out/pytest.log 2>&1; tail -4 $out/pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-fp64-peak > $out/bench_n1.json 2> $out/bench_n1.err; python -c "
import json
txt=open('$out/bench_n1.json').read().strip().splitlines()
d=json.loads([l for l in txt if l.startswith('{')][-1]); print('N=1', d['value'], d['stage_ms_per_step'], d['roofline']['frac'], d['roofline']['rk_build'])"
timeout 600 python bench.py --workload cfg3 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-fp64-peak > $out/bench_cfg3.json 2> $out/bench_cfg3.err; python -c "
import json
txt=open('$out/bench_cfg3.json').read().strip().splitlines()
d=json.loads([l for l in txt if l.startswith('{')][-1]); print('cfg3', d['value'], d['stage_ms_per_step'], d['roofline']['frac'], d['roofline']['rk_build'])"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"rk_build" -c 1 -o $out/rk_build -f python scripts/sharded_run.py cfg4 > $out/ncu_rk.log 2>&1; tail -1 $out/ncu_rk.log
