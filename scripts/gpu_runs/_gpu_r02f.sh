L=b-spline-two-e_b200/lib
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
run() { echo "== $1"; env $3 BS2E_LIB=$PWD/$L/$2 BS2E_ONLY_BLOCKS=6 python scripts/sharded_run.py cfg4 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['stage_C_ms'], d['elements_per_s'], d['checksum_xor_rank0'], d.get('site_phase_cycles'))"; }
run base libbs2e_gpu.so X=1
run base_c16 libbs2e_gpu.so BS2E_SITE_CHUNK_KB=16
run dregs160 libbs2e_gpu_dregs160.so X=1
run rc1nt256 libbs2e_gpu_rc1nt256.so X=1
run rc1nt128 libbs2e_gpu_rc1nt128.so X=1
run xrc2 libbs2e_gpu_xrc2.so X=1
run nothing libbs2e_gpu_nothing.so X=1
run time libbs2e_gpu_time.so X=1
echo "== cfg3 base"; python scripts/sharded_run.py cfg3 | cut -c1-330
echo "== cfg4 base"; python scripts/sharded_run.py cfg4 | cut -c1-330
