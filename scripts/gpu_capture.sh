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
timeout 300 python bench.py --impl reference --steps 1 --warmup 0 > $out/bench_reference.json 2>> $out/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file $out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > $out/ncu_launch.log 2>&1
# setup (4 + 5 counts + 10 fills) and one warm-up step (19) are skipped, the timed step is captured
timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:"site_fill|block_fill|rk_build|diag_cells|site_count|block_count|cell_moments|pair_prefix" -s 38 -c 19 -o $out/full \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > $out/ncu_full.log 2>&1
ls -la $out
