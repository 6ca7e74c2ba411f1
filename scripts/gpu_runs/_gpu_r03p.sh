out=gpurun_out/r03p; mkdir -p $out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > $out/gpu.txt 2>&1; nproc >> $out/gpu.txt
timeout 1500 python -X faulthandler -m pytest tests -m gpu -x -q > $out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $out/pytest_gpu.log; tail -3 $out/pytest_gpu.log
timeout 900 python bench.py > $out/bench.json 2> $out/bench.err; echo "bench rc=$?"; tail -c 400 $out/bench.json
timeout 600 python bench.py --impl reference > $out/bench_reference.json 2>> $out/bench.err; echo "ref rc=$?"; tail -c 300 $out/bench_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-fp64-peak > $out/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:"site_mma" -c 18 -o $out/fill_all -f python scripts/sharded_run.py cfg4 > $out/ncu_fill_all.log 2>&1; tail -1 $out/ncu_fill_all.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
