"""host timeline of the per-block sequence plan -> assemble -> free (profiling aid)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "b-spline-two-e_b200"))
import torch, bs2e
from bs2e.sharding import exchange_cost, site_partition
wl = sys.argv[1] if len(sys.argv) > 1 else "cfg4"
nshare = int(sys.argv[2]) if len(sys.argv) > 2 else 1
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
setup = bs2e.BasisSetup(device=0, **bs2e.CONFIGS[wl])
S, H_vec, syms = setup.host_inputs()
ctx = setup.open()
stream = torch.cuda.Stream(); ctx.set_stream(stream.cuda_stream)
ctx.slater_cells(); ctx.rk_build(); ctx.set_one_particle(H_vec, S)
cfgs = [ctx.configs_upload(s) for s in syms]
ranges = []
for s, c in zip(syms, cfgs):
    if nshare == 1:
        ranges.append([(1, s.n_config)])
    else:
        tmp = ctx.block_plan(s, False, cfg=c); cH, cS = tmp.row_counts(); tmp.free()
        ranges.append(site_partition(s.conf_n, cH + cS, nshare, setup.k, exchange_cost(setup.p['max_k']))[which])
ctx.sync()
for rep in range(4):
    t = []
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    w0 = time.perf_counter()
    e0.record(stream)
    for s, c, r in zip(syms, cfgs, ranges):
        a = time.perf_counter(); blk = ctx.block_plan(s, False, ranges=r, cfg=c)
        b = time.perf_counter(); blk.assemble()
        cc = time.perf_counter(); blk.free()
        d = time.perf_counter(); t.append((b - a, cc - b, d - cc))
    e1.record(stream)
    w1 = time.perf_counter()
    torch.cuda.synchronize()
    w2 = time.perf_counter()
    print("rep", rep, "device %.2f ms, host loop %.2f ms, drain %.2f ms" % (e0.elapsed_time(e1), (w1 - w0) * 1e3, (w2 - w1) * 1e3))
    print("   plan ms:", " ".join("%.2f" % (x[0] * 1e3) for x in t))
    print("   assemble ms:", " ".join("%.2f" % (x[1] * 1e3) for x in t), " free ms:", " ".join("%.2f" % (x[2] * 1e3) for x in t))
