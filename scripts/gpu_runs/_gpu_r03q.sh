out=gpurun_out/r03q; mkdir -p $out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > $out/gpu.txt 2>&1; nproc >> $out/gpu.txt
timeout 300 python -X faulthandler -m pytest tests/test_dipole.py -m gpu -x -q > $out/pytest_dip.log 2>&1; tail -2 $out/pytest_dip.log
timeout 200 python scripts/dipole_probe.py cfg3 v 2>&1 | tail -1
timeout 900 python bench.py > $out/bench.json 2> $out/bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference > $out/bench_reference.json 2>> $out/bench.err; echo "ref rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-fp64-peak > $out/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:"site_mma" -c 18 -o /tmp/fill_all -f python scripts/sharded_run.py cfg4 > $out/ncu_fill_all.log 2>&1; tail -1 $out/ncu_fill_all.log
python scripts/ncu_summary.py traffic /tmp/fill_all.ncu-rep cfg4 9 > $out/traffic_cfg4.json; python scripts/ncu_summary.py full /tmp/fill_all.ncu-rep > $out/fill_all_summary.md
timeout 200 ncu --set full --clock-control none -k regex:"dip_fill" -s 4 -c 4 -o /tmp/dip -f python scripts/dipole_probe.py cfg3 v > /dev/null 2>&1; python scripts/ncu_summary.py full /tmp/dip.ncu-rep > $out/dip_fill_summary.md
ls -la $out
