out=gpurun_out/r02d; mkdir -p $out
./scripts/microbench/dfma_lat > $out/dfma.txt 2>&1; cat $out/dfma.txt
L=b-spline-two-e_b200/lib
run() { # name, lib, extra env
  echo "== $1"; env $3 BS2E_LIB=$PWD/$L/$2 BS2E_ONLY_BLOCKS=6 python scripts/sharded_run.py cfg4 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['stage_C_ms'], d['elements_per_s'], d['checksum_xor_rank0'], d.get('site_phase_cycles'))"
}
run base libbs2e_gpu.so X=1
run base_c16 libbs2e_gpu.so BS2E_SITE_CHUNK_KB=16
run base_c8 libbs2e_gpu.so BS2E_SITE_CHUNK_KB=8
run rb1 libbs2e_gpu_rb1.so X=1
run rb4 libbs2e_gpu_rb4.so X=1
run rb4_c16 libbs2e_gpu_rb4.so BS2E_SITE_CHUNK_KB=16
run d768_c16 libbs2e_gpu_d768.so BS2E_SITE_CHUNK_KB=16
run d768_c12 libbs2e_gpu_d768.so BS2E_SITE_CHUNK_KB=12
run rb4d768_c16 libbs2e_gpu_rb4d768.so BS2E_SITE_CHUNK_KB=16
run time libbs2e_gpu_time.so X=1
echo "== cfg3 base"; python scripts/sharded_run.py cfg3 | cut -c1-330
echo "== cfg3 rb4"; BS2E_LIB=$PWD/$L/libbs2e_gpu_rb4.so python scripts/sharded_run.py cfg3 | cut -c1-330
echo "== cfg3 time"; BS2E_LIB=$PWD/$L/libbs2e_gpu_time.so python scripts/sharded_run.py cfg3 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['stage_C_ms'], d.get('site_phase_cycles'))"
