out=gpurun_out/r03r; mkdir -p $out
BS2E_ONLY_BLOCKS=5,6,7,8 timeout 900 ncu --set full --clock-control none -k regex:"site_mma" -c 16 -o /tmp/fill_rest -f python scripts/sharded_run.py cfg4 > $out/ncu_fill_rest.log 2>&1; tail -1 $out/ncu_fill_rest.log | cut -c1-100
python scripts/ncu_summary.py full /tmp/fill_rest.ncu-rep > $out/fill_rest_summary.md
python - <<'PY' > $out/fill_rest_bytes.json
import csv, io, json, subprocess
txt = subprocess.run(["ncu", "-i", "/tmp/fill_rest.ncu-rep", "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt))); hdr, units, data = rows[0], rows[1], rows[2:]
U = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
out = []
for r in data:
    f = lambda m: float(r[hdr.index(m)].replace(",", "")) * U.get(units[hdr.index(m)], 1.0)
    out.append({"kernel": r[hdr.index("Kernel Name")][:40], "grid": r[hdr.index("launch__grid_size")], "read": f("dram__bytes_read.sum"), "write": f("dram__bytes_write.sum"), "ms": float(r[hdr.index("gpu__time_duration.sum")].replace(",", ""))})
print(json.dumps(out))
PY
