out=gpurun_out/r02b; mkdir -p $out
timeout 900 python -m pytest tests -m gpu -x -q > $out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $out/pytest_gpu.log
tail -15 $out/pytest_gpu.log
timeout 300 python scripts/sharded_run.py cfg4 > $out/cfg4.json 2> $out/cfg4.err; cut -c1-600 $out/cfg4.json; tail -3 $out/cfg4.err
timeout 300 python scripts/sharded_run.py cfg3 > $out/cfg3.json 2> $out/cfg3.err; cut -c1-600 $out/cfg3.json; tail -3 $out/cfg3.err
