out=gpurun_out/r02h; mkdir -p $out
timeout 900 python bench.py --steps 5 --warmup 3 --cpu-budget 16 > $out/bench_cfg4.json 2> $out/bench_cfg4.err; echo "bench rc=$?"; tail -5 $out/bench_cfg4.err; cut -c1-3000 $out/bench_cfg4.json
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > $out/ref_cfg4.json 2> $out/ref_cfg4.err; echo "ref rc=$?"; tail -3 $out/ref_cfg4.err; cut -c1-1500 $out/ref_cfg4.json
