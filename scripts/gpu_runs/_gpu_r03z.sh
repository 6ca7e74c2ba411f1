out=gpurun_out/r03z; mkdir -p $out
timeout 300 python -X faulthandler -m pytest tests/test_dipole.py -m gpu -x -q > $out/pytest_dip.log 2>&1; echo rc=$?; tail -2 $out/pytest_dip.log
timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_write.sum,smsp__inst_executed.sum --clock-control none -k regex:"dip_fill" -s 4 -c 4 --csv --log-file $out/dip.csv python scripts/dipole_probe.py cfg3 v > /dev/null 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(l for l in open('$out/dip.csv') if l.startswith('"'))]
h=rows[0]; mn,mv,mu=h.index('Metric Name'),h.index('Metric Value'),h.index('Metric Unit')
for r in rows[1:]: print(r[mn], r[mv], r[mu])
PY
