out=gpurun_out/r02a; mkdir -p $out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > $out/gpu.txt 2>&1
(lscpu | head -25; nproc; free -g; numactl -H 2>/dev/null | head) > $out/cpu.txt 2>&1
python scripts/fp64_peak.py > $out/fp64_peak.json 2> $out/fp64.err
cat $out/fp64_peak.json
timeout 300 python scripts/sharded_run.py cfg4 > $out/cfg4_base.json 2> $out/cfg4_base.err; cat $out/cfg4_base.json | cut -c1-400
BS2E_ONLY_BLOCKS=6 timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"site_fill|rk_build|diag_cells" -c 4 -o $out/cfg4_full -f python scripts/sharded_run.py cfg4 > $out/ncu_cfg4.log 2>&1
tail -3 $out/ncu_cfg4.log
ls -la $out
