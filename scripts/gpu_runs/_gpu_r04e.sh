out=gpurun_out/r04e; mkdir -p $out
BS2E_ONLY_BLOCKS=6 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"site_mma" -c 1 -o /tmp/x -f python scripts/sharded_run.py cfg4 > /dev/null 2>&1
ncu -i /tmp/x.ncu-rep --page source --csv > $out/src_x.csv 2>/dev/null
ncu -i /tmp/x.ncu-rep --page raw --csv > $out/raw_x.csv 2>/dev/null
ls -la $out
