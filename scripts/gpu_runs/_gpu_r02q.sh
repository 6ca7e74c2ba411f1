out=gpurun_out/r02q; mkdir -p $out
BS2E_PARITY_REPORT=$out/parity.json timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 > $out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -25 $out/pytest_gpu.log
