out=gpurun_out/r03l; mkdir -p $out
BS2E_ONLY_BLOCKS=6 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"site_mma" -c 2 -o $out/cfg4_mma -f python scripts/sharded_run.py cfg4 > $out/ncu_cfg4_mma.log 2>&1; tail -1 $out/ncu_cfg4_mma.log
BS2E_ONLY_BLOCKS=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"site_mma" -c 2 -o $out/cfg4_mma_L0 -f python scripts/sharded_run.py cfg4 > $out/ncu_cfg4_mma_L0.log 2>&1; tail -1 $out/ncu_cfg4_mma_L0.log
