out=gpurun_out/r04c; mkdir -p $out
L=$PWD/b-spline-two-e_b200/lib
for v in libbs2e_gpu.so libbs2e_gpu_rows16.so; do
BS2E_LIB=$L/$v BS2E_ONLY_BLOCKS=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"rk_build" -c 1 --csv --log-file $out/t_$v.csv python scripts/sharded_run.py cfg4 > /dev/null 2>&1
python -c "
import csv
rows=[r for r in csv.reader(l for l in open('$out/t_$v.csv') if l.startswith('\"'))]
print('$v', rows[1][-1])"
done
