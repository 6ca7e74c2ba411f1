out=gpurun_out/r02c; mkdir -p $out
BS2E_ONLY_BLOCKS=6 timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"site_fill" -c 2 -o $out/cfg4_fill -f python scripts/sharded_run.py cfg4 > $out/ncu_cfg4.log 2>&1
tail -3 $out/ncu_cfg4.log
BS2E_ONLY_BLOCKS=6 python scripts/sharded_run.py cfg4 | cut -c1-300
