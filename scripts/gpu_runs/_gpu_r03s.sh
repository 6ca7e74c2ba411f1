out=gpurun_out/r03s; mkdir -p $out
L=$PWD/b-spline-two-e_b200/lib
for v in libbs2e_gpu.so libbs2e_gpu_decoupled.so; do
BS2E_LIB=$L/$v BS2E_ONLY_BLOCKS=2,6 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"site_mma" -c 8 --csv --log-file $out/t_$v.csv python scripts/sharded_run.py cfg4 > /dev/null 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(l for l in open('$out/t_$v.csv') if l.startswith('"'))]
h=rows[0]; kn,mv=h.index('Kernel Name'),h.index('Metric Value')
print('$v', [(r[kn][24:32], round(float(r[mv].replace(',',''))/1e6,3)) for r in rows[1:]])
PY
done
