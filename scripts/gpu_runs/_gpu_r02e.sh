L=b-spline-two-e_b200/lib
run() { echo "== $1"; env $3 BS2E_LIB=$PWD/$L/$2 BS2E_ONLY_BLOCKS=6 python scripts/sharded_run.py cfg4 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['stage_C_ms'], d['elements_per_s'], d['checksum_xor_rank0'], d.get('site_phase_cycles'))"; }
run base_c8 libbs2e_gpu_rb1.so BS2E_SITE_CHUNK_KB=8
run base_c4 libbs2e_gpu_rb1.so BS2E_SITE_CHUNK_KB=4
run base_c2 libbs2e_gpu_rb1.so BS2E_SITE_CHUNK_KB=2
run nostore_c8 libbs2e_gpu_nostore.so BS2E_SITE_CHUNK_KB=8
run nodot_c8 libbs2e_gpu_nodot.so BS2E_SITE_CHUNK_KB=8
run nothing_c8 libbs2e_gpu_nothing.so BS2E_SITE_CHUNK_KB=8
run time_c8 libbs2e_gpu_time.so BS2E_SITE_CHUNK_KB=8
