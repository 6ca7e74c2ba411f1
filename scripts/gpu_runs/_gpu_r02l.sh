L=b-spline-two-e_b200/lib
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
run() { echo "== $1"; env $3 BS2E_LIB=$PWD/$L/$2 BS2E_ONLY_BLOCKS=6 python scripts/sharded_run.py cfg4 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['stage_C_ms'], d['elements_per_s'], d['checksum_xor_rank0'])"; }
run mma libbs2e_gpu.so X=1
run mma_c24 libbs2e_gpu.so BS2E_SITE_CHUNK_KB=24
run minb1 libbs2e_gpu_minb1.so X=1
run minb1_c24 libbs2e_gpu_minb1.so BS2E_SITE_CHUNK_KB=24
run fma libbs2e_gpu.so BS2E_FILL=fma
