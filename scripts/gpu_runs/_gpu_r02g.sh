out=gpurun_out/r02g; mkdir -p $out
./scripts/microbench/dmma_probe > $out/dmma.txt 2>&1; cat $out/dmma.txt
python - <<'PY' > $out/plan_timing.txt 2>&1
import sys, time, os
sys.path.insert(0, "b-spline-two-e_b200"); sys.path.insert(0, ".")
import bs2e
for wl in ("cfg3", "cfg4"):
    setup = bs2e.BasisSetup(device=0, **bs2e.CONFIGS[wl])
    t0 = time.perf_counter(); S, H_vec, syms = setup.host_inputs(); t1 = time.perf_counter()
    ctx = setup.open(); ctx.slater_cells(); ctx.rk_build(); ctx.set_one_particle(H_vec, S); ctx.sync()
    print(wl, "host_inputs %.2f s" % (t1 - t0), "n_sym", len(syms))
    cfgs = [ctx.configs_upload(s) for s in syms]
    for rep in range(3):
        tt = []
        for s, c in zip(syms, cfgs):
            a = time.perf_counter(); b = ctx.block_plan(s, False, cfg=c); ctx.sync(); tt.append((time.perf_counter() - a) * 1e3); b.free()
        print(wl, "rep", rep, "plan_dev ms per block:", " ".join("%.2f" % t for t in tt), "sum %.2f" % sum(tt))
    for rep in range(2):
        tt = []
        for s in syms:
            a = time.perf_counter(); b = ctx.block_plan(s, False); ctx.sync(); tt.append((time.perf_counter() - a) * 1e3); b.free()
        print(wl, "rep", rep, "plan(host configs) ms per block:", " ".join("%.2f" % t for t in tt), "sum %.2f" % sum(tt))
    ctx.close()
PY
cat $out/plan_timing.txt
