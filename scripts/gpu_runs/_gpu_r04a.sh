out=gpurun_out/r04a; mkdir -p $out
timeout 1500 python -X faulthandler -m pytest tests -m gpu -x -q > $out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $out/pytest_gpu.log; tail -3 $out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-fp64-peak > $out/bench.json 2> $out/bench.err; python -c "
import json
d=json.loads([l for l in open('$out/bench.json') if l.startswith('{')][-1]); print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['traffic'], d['gpu_launches'])"
