out=gpurun_out/r02j; mkdir -p $out
BS2E_ONLY_BLOCKS=6 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"site_mma" -c 2 -o $out/cfg4_mma -f python scripts/sharded_run.py cfg4 > $out/ncu_cfg4.log 2>&1
tail -3 $out/ncu_cfg4.log
