#!/bin/bash
# One GPU-box pass: parity tests, bench line, ncu launch list, ncu full capture of one
# whole step (stage A, B and C kernels).  Usage (from the repo root, through gpurun):
#   gpurun --timeout 1500 -- 'bash scripts/gpu_capture.sh r01m'
tag=${1:-rXX}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > $out/gpu.txt 2>&1
lscpu | head -20 > $out/cpu.txt; nproc >> $out/cpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > $out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $out/pytest_gpu.log
tail -5 $out/pytest_gpu.log
timeout 600 python bench.py > $out/bench.json 2> $out/bench.err; echo "bench rc=$?"
cat $out/bench.json
timeout 600 python bench.py --impl reference > $out/bench_reference.json 2>> $out/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv \
    --log-file $out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-fp64-peak > $out/ncu_launch.log 2>&1
# full captures are written to /tmp and only their summaries are kept: gpurun brings back at most 64 MiB
# (a) stages A, B and the plan kernels of the set-up pass and the first blocks
timeout 900 ncu --set full --clock-control none \
    -k regex:"rk_build|diag_cells|cell_moments|pair_prefix|site_count|site_enum|conf_scan|ncrow" -c 24 -o /tmp/small -f \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-fp64-peak > $out/ncu_small.log 2>&1
python scripts/ncu_summary.py full /tmp/small.ncu-rep > $out/small_kernels_summary.md
# (b) the fill launches of every block (sharded_run fills every block twice: 36 launches at cfg4)
timeout 1500 ncu --set full --clock-control none -k regex:"site_mma|site_fill" -c 36 -o /tmp/fill_all -f \
    python scripts/sharded_run.py cfg4 > $out/ncu_fill_all.log 2>&1
python scripts/ncu_summary.py full /tmp/fill_all.ncu-rep > $out/fill_all_summary.md
ls -la $out
