#!/usr/bin/env python
"""bench.py -- two-electron matrix-element hot path of basis_setup on B200.

One "step" = one pass of the hot path over one BASELINE workload:
    stage A  setup_Slater_integrals   (cell integrals)
    stage B  compute_R_k_map          (R^k tensor)
    stage C  count_nnz + construct_block_tensor for every symmetry block
Metric: 2e matrix elements/s = sum_sym (nnz_H+nnz_S)
divided by the time of stage C (count pass + CSR build + fill); the R^k
integrals/s of stages A+B is reported beside it ("rk_integrals_per_s").

  value : configuration lists, one-particle matrices and R^k resident in HBM; device time
          (CUDA events on the library's stream) of stage C = for every symmetry block the
          PLAN (group structure, radial sites, count pass, scan -- built on the device, the
          host reads the totals), the allocation of the CSR arrays and the fill; one block
          at a time (the CSR of cfg4 is 159 GB), arrays recycled through the memory pool
  roofline : the fill launches of every block bracketed by CUDA events inside the timed
          steps; algorithmic bytes = 24 B per stored element
  e2e   : the same metric through the C ABI with HOST buffers
          (bs2e_set_one_particle / bs2e_block_count / bs2e_block_fill; H2D and D2H inside
          the timed region) -- into pinned arrays, and once into pageable arrays as the
          Fortran caller allocates them
  --impl reference : the CPU oracle port of the reference path on the host cores

N>1 (torchrun, one rank per GPU): strong scaling.  Every rank assembles its
share of the rows of every symmetry block -- rows are dealt by the first radial
index of their configuration so that radial sites stay whole, balanced on the
stored entries and then on measured times (bs2e.sharding: site_units,
refine_bounds; untimed set-up) -- and builds the rows of the R^k tensor its
radial sites read (bs2e_rk_rows: a slice by first spline index; cheaper than an
all-gather of the whole tensor); no data-path collective.
value = elements of all ranks / max-over-ranks stage-C time.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "b-spline-two-e_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC = "2e matrix elements/s (R^k integrals/s alongside) on basis_setup two-electron path"
UNIT = "matrix elements/s"
DEFAULT_WORKLOAD = "cfg4"   # BASELINE.json configs[3]: the largest configuration that runs on ONE GPU (block by block)
# config.workload of both arms (sizes follow from the namelist; checked against the generated basis)
WORKLOADS = {
    "cfg1": "cfg1: k=8 n_b=96 k_GL=14 max_k=4 max_l_1p=3 max_L=2 n_sym=9 sum_n_config=129979",
    "cfg2": "cfg2: k=7 n_b=105 k_GL=13 max_k=6 max_l_1p=3 max_L=2 n_sym=3 sum_n_config=91264",
    "cfg3": "cfg3: k=8 n_b=206 k_GL=14 max_k=12 max_l_1p=6 max_L=4 n_sym=5 sum_n_config=501166",
    "cfg4": "cfg4: k=8 n_b=307 k_GL=18 max_k=20 max_l_1p=10 max_L=8 n_sym=9 sum_n_config=2689060",
    "cfg5": "cfg5: k=8 n_b=606 k_GL=23 max_k=30 max_l_1p=15 max_L=12 n_sym=13 sum_n_config=13493894",
}


def env_int(name, default):
    return int(os.environ.get(name, default))


# ---------------------------------------------------------------------------
# clocks sampler (B200_PROFILING.md: sample DURING the timed region)
# ---------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-i", str(self.idx), "-lms", "25"], stdout=subprocess.PIPE, text=True)
            self.thr = threading.Thread(target=self._read, daemon=True)
            self.thr.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def window(self, t0, t1):
        self.t0, self.t1 = t0, t1

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        t0, t1 = getattr(self, "t0", None), getattr(self, "t1", None)
        inside = [ln for ts, ln in self.lines if t0 is None or (t0 <= ts <= t1 + 0.03)]
        for ln in inside:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                  "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(max(mx)) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def measured_peak_hbm():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic_per_launch(workload):
    """dram bytes per block_fill launch from the committed ncu capture, if one exists"""
    path = os.path.join(ROOT, "profiles", "ncu_fill_traffic.json")
    try:
        d = json.load(open(path))
        return d.get(workload, {}).get("dram_bytes_per_launch")
    except Exception:
        return None


# ---------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    import bs2e
    from bs2e.sharding import exchange_cost, ranges_of_bounds, refine_bounds, rk_rows_needed, site_units, unit_bounds

    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus N>1 must be launched with torch.distributed.run (one rank per GPU)")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce(x, op):
        if world == 1:
            return float(x)
        t = torch.tensor([float(x)], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=op)
        return float(t.item())

    max_over_ranks = lambda x: reduce(x, dist.ReduceOp.MAX)
    sum_over_ranks = lambda x: reduce(x, dist.ReduceOp.SUM)

    params = bs2e.CONFIGS[args.workload]
    setup = bs2e.BasisSetup(device=local, **params)
    S, H_vec, syms = setup.host_inputs()
    if describe(args.workload, setup, syms) != WORKLOADS[args.workload]:
        raise SystemExit(f"workload table out of date: {describe(args.workload, setup, syms)}")
    ctx = setup.open()
    stream = torch.cuda.Stream()
    ctx.set_stream(stream.cuda_stream)
    full = setup.p["full"]
    K1 = setup.p["max_k"] + 1
    n_rk = ctx.P * ctx.P * K1

    # ---- untimed setup: the inputs of the path made resident in HBM ----
    ctx.slater_cells(); ctx.rk_build(); ctx.set_one_particle(H_vec, S)
    cfgs = [ctx.configs_upload(s) for s in syms]           # term%configs of every symmetry
    ranges, units, bounds = [], [], []
    for s, c in zip(syms, cfgs):
        if world == 1:
            ranges.append([(1, s.n_config)])
        else:   # rows dealt by their first radial index (radial sites stay whole), balanced on the stored
                # entries; the partition depends on the basis only and is computed once
            tmp = ctx.block_plan(s, full, cfg=c)
            cH, cS = tmp.row_counts()
            tmp.free()
            present, unit_w = site_units(s.conf_n, cH + cS, setup.k, exchange_cost(setup.p['max_k']))
            units.append((present, unit_w))
            bounds.append(unit_bounds(present, unit_w, world))
            ranges.append(ranges_of_bounds(s.conf_n, bounds[-1])[rank])
            if not ranges[-1]:
                raise SystemExit(f"rank {rank}: empty share of block L={s.l} (more GPUs than radial indices)")
    ctx.sync()
    rebalance_log = []
    if world > 1:
        # measured rebalancing (untimed set-up, like the partition itself): every rank times stage C of its
        # share of every block, the times are gathered and the n1 axis is cut again into intervals of equal
        # measured cost (bs2e.sharding.refine_bounds); every rank computes the same cuts from the same numbers
        for it in range(args.rebalance + 1):
            mine = []
            for s, c, r in zip(syms, cfgs, ranges):
                best = None
                for rep in range(2):
                    with torch.cuda.stream(stream):
                        a = torch.cuda.Event(enable_timing=True); a.record(stream)
                        blk = ctx.block_plan(s, full, ranges=r, cfg=c)
                        blk.assemble()
                        blk.free()
                        b = torch.cuda.Event(enable_timing=True); b.record(stream)
                    ctx.sync()
                    t = a.elapsed_time(b)
                    best = t if best is None else min(best, t)
                mine.append(best)
            t_all = torch.zeros(world, len(syms), device="cuda", dtype=torch.float64)
            dist.all_gather_into_tensor(t_all, torch.tensor(mine, device="cuda", dtype=torch.float64))
            t_all = t_all.cpu().numpy()
            rebalance_log.append({"max_over_mean": [float(t_all[:, q].max() / t_all[:, q].mean()) for q in range(len(syms))],
                                  "sum_of_max_ms": float(t_all.max(axis=0).sum())})
            if it == args.rebalance:
                break
            for q, s in enumerate(syms):
                bounds[q] = refine_bounds(units[q][0], units[q][1], bounds[q], t_all[:, q])
                ranges[q] = ranges_of_bounds(s.conf_n, bounds[q])[rank]
                if not ranges[q]:
                    raise SystemExit(f"rank {rank}: empty share of block L={s.l} after rebalancing")
    rk_rows = None
    if world > 1 and not args.whole_rk:
        # stages A and B per rank: only the rows of the R^k tensor this rank's radial sites read (bs2e_rk_rows)
        need = [rk_rows_needed(s.conf_n, bounds[q][rank], setup.k) for q, s in enumerate(syms)]
        a_lo, a_hi = min(n[0] for n in need), max(n[1] for n in need)
        ctx.rk_rows(a_lo, a_hi)
        ctx.slater_cells(); ctx.rk_build(); ctx.sync()
        g = torch.zeros(world, 2, device="cuda", dtype=torch.float64)
        dist.all_gather_into_tensor(g, torch.tensor([a_lo, a_hi], device="cuda", dtype=torch.float64))
        rk_rows = [[int(a), int(b)] for a, b in g.cpu().numpy()]
    ctx.sync()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")   # > 126 MB L2

    def ev():
        e = torch.cuda.Event(enable_timing=True)
        e.record(stream)
        return e

    elems_of_step = [0]
    host_ms = []   # host time of every block_plan / assemble call (BS2E_BENCH_TRACE)

    def one_step(rec):
        """stage A, stage B, then stage C block by block: plan (device-built, the host reads the totals),
        allocation of the CSR arrays from the pool, fill, release"""
        with torch.cuda.stream(stream):
            flush.zero_()                                   # evict L2 between timed iterations
            e0 = ev()
            ctx.slater_cells()
            ea = ev()
            ctx.rk_build()
            e1 = ev()
            per_block, el = [], 0
            for s, c, r in zip(syms, cfgs, ranges):
                h0 = time.perf_counter()
                blk = ctx.block_plan(s, full, ranges=r, cfg=c)
                h1 = time.perf_counter()
                f0 = ev()
                blk.assemble()                              # the two concurrent launches of the site kernel
                f1 = ev()
                host_ms.append((1e3 * (h1 - h0), 1e3 * (time.perf_counter() - h1)))
                n = blk.nnz_H + blk.nnz_S
                per_block.append((f0, f1, n, blk.nnz_H, blk.nnz_S, blk.nrows))
                blk.free()                                  # stream-ordered: the pool hands the pages to the next block
                el += n
            e2 = ev()
        elems_of_step[0] = el
        if rec is not None:
            rec.append((e0, ea, e1, e2, per_block))

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        # let nvidia-smi reach its sampling loop: its start-up (NVML initialisation) stalls the GPU for 0.1-0.5 s
        # on some boxes and must not land in a timed step (seen as one 126 / 549 ms first step, gpurun_out/r03b,d)
        t_wait = time.perf_counter()
        while len(sampler.lines) < 3 and time.perf_counter() - t_wait < 10.0 and sampler.proc is not None:
            time.sleep(0.05)
    for _ in range(args.warmup):
        one_step(None)
    barrier()
    launches0 = bs2e.launch_count()
    rec = []
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        one_step(rec)
    barrier()
    wall = time.perf_counter() - t0
    sampler.window(t0, t0 + wall)
    launches = bs2e.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None

    if os.environ.get("BS2E_BENCH_TRACE"):
        nb_ = len(syms)
        for k_, r in enumerate(rec):
            fm = [round(q[0].elapsed_time(q[1]), 2) for q in r[4]]
            hm = host_ms[-(len(rec) - k_) * nb_:][:nb_]
            print(f"[trace rank {rank} step {k_}] stage C {r[2].elapsed_time(r[3]):.2f} ms; fill per block {fm}; "
                  f"host plan/assemble ms {[(round(a, 2), round(b, 2)) for a, b in hm]}", file=sys.stderr)
    my_elems = elems_of_step[0]
    total_elems = sum_over_ranks(my_elems)
    tA = sum(r[0].elapsed_time(r[1]) for r in rec)
    tB = sum(r[1].elapsed_time(r[2]) for r in rec)
    tC = sum(r[2].elapsed_time(r[3]) for r in rec)
    fill_ms = [q[0].elapsed_time(q[1]) for r in rec for q in r[4]]
    per_rank = None
    if world > 1:   # stage times and fill times of every rank (the value uses the max)
        g = torch.zeros(world, 4, device="cuda", dtype=torch.float64)
        dist.all_gather_into_tensor(g, torch.tensor([tA, tB, tC, sum(fill_ms)], device="cuda", dtype=torch.float64))
        g = (g / args.steps).cpu().numpy()
        per_rank = {"A_cells_ms": g[:, 0].tolist(), "B_rk_ms": g[:, 1].tolist(), "C_blocks_ms": g[:, 2].tolist(),
                    "C_fill_only_ms": g[:, 3].tolist()}
        gs = torch.zeros(world, args.steps, device="cuda", dtype=torch.float64)
        dist.all_gather_into_tensor(gs, torch.tensor([r[2].elapsed_time(r[3]) for r in rec], device="cuda",
                                                     dtype=torch.float64))
        per_rank["C_blocks_ms_of_every_step"] = gs.cpu().numpy().round(3).tolist()
    tA, tB, tC = max_over_ranks(tA), max_over_ranks(tB), max_over_ranks(tC)
    wall = max_over_ranks(wall)
    K = args.steps
    value = total_elems * K / (tC * 1e-3)
    rk_per_s = n_rk * K / ((tA + tB) * 1e-3)

    # ---- roofline of the dominant kernel (site_fill_kernel), this rank ----
    fill_total_ms = sum(fill_ms)
    n_fill = len(fill_ms)
    alg_bytes_per_launch = 24.0 * my_elems / len(syms)         # 16 B data + 8 B index per element
    avg_fill_ms = fill_total_ms / n_fill
    peak, peak_src = measured_peak_hbm()
    achieved = alg_bytes_per_launch / (avg_fill_ms * 1e-3) / 1e9
    fill_mode = os.environ.get("BS2E_FILL")
    fill_kernel = ("site_mma_kernel" if (fill_mode == "mma" or (fill_mode is None and K1 > 7)) else
                   "block_fill_kernel" if fill_mode == "row" else "site_fill_kernel")
    roofline = {"kernel": fill_kernel, "bound": "hbm", "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak,
                # ncu DRAM bytes per launch of the whole blocks (captured on one GPU); a share's launches at N>1 were not captured
                "traffic": ncu_traffic_per_launch(args.workload) if world == 1 else None,
                "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes_per_launch,
                "avg_launch_ms": avg_fill_ms,
                "timed": "CUDA events around the fill of every block inside the timed steps (a launch = the two "
                         "concurrent launches of the site kernel -- sites with / without exchange windows -- of one symmetry "
                         "block; the stream is idle when they start)",
                "share_of_stage_C": (fill_total_ms / K) / max(tC / K, 1e-9),
                # the same bytes over the whole timed stage C (plans, count passes, scans, allocation included)
                "achieved_over_timed_stage_C": 24.0 * my_elems * K / (tC * 1e-3) / 1e9,
                "frac_over_timed_stage_C": 24.0 * my_elems * K / (tC * 1e-3) / 1e9 / peak,
                "per_block_frac": [24.0 * q[2] / (q[0].elapsed_time(q[1]) * 1e-3) / 1e9 / peak for q in rec[-1][4]],
                "rk_build": {"achieved": 8.0 * n_rk * K / (tB * 1e-3) / 1e9, "unit": "GB/s",
                             "frac": 8.0 * n_rk * K / (tB * 1e-3) / 1e9 / peak, "bound": "hbm"}}

    fp64 = fp64_peak(torch) if rank == 0 and not args.no_fp64_peak else None

    # ---- e2e: through the C ABI with host buffers ----
    e2e = None
    if not args.no_e2e:
        sizes = [(q[5], q[3], q[4]) for q in rec[-1][4]]        # (rows, nnz_H, nnz_S) of this rank's share of every block
        e2e = run_e2e(args, bs2e, ctx, setup, syms, ranges, sizes, S, H_vec, world, barrier,
                      max_over_ranks, total_elems, n_rk)

    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_base = cpu_baseline(args.workload, budget_s=args.cpu_budget)

    if rank == 0:
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K,
            "warmup": args.warmup, "ms_per_step": wall * 1e3 / K, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOADS[args.workload],
                       "elements_per_step": total_elems, "rk_integrals_per_step": n_rk,
                       "stage_C": "per symmetry block: device-built plan + count pass + scan (host reads the totals), "
                                  "CSR arrays from the memory pool, fill, release; one block at a time",
                       "l2": "256 MiB buffer written between timed iterations; R^k and CSR output exceed L2",
                       "parallelism": ("rows of every symmetry block dealt by first radial index (radial sites stay whole), every GPU builds "
                                       "the R^k rows its sites read, no collective") if world > 1 else "single GPU"},
            "stage_ms_per_step": {"A_cells": tA / K, "B_rk": tB / K, "C_blocks": tC / K,
                                  "C_fill_only": fill_total_ms / K,
                                  "C_blocks_of_every_step": [round(r[2].elapsed_time(r[3]), 3) for r in rec]},
            "per_rank_ms_per_step": per_rank,
            "partition": ({"unit": "first radial index n1 (radial sites stay whole)",
                           "rebalance_steps": args.rebalance, "measured": rebalance_log,
                           "rk_rows_per_rank": rk_rows, "n_b": int(ctx.n_b)} if world > 1 else None),
            "rk_integrals_per_s": rk_per_s,
            "whole_step_elements_per_s": total_elems * K / ((tA + tB + tC) * 1e-3),
            "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "roofline": roofline,
            "fp64_peak": fp64, "cpu_baseline": cpu_base,
        }
        print(json.dumps(out))
    for c in cfgs:
        c.free()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def fp64_peak(torch, seconds=2.0):
    """cuBLAS DGEMM 8192^3 on this box (BASELINE.md section 2 asks for the FP64 denominator): best single
    call and back-to-back sustained rate"""
    n = 8192
    a = torch.randn(n, n, device="cuda", dtype=torch.float64)
    b = torch.randn(n, n, device="cuda", dtype=torch.float64)
    c = torch.empty_like(a)
    torch.matmul(a, b, out=c)
    torch.cuda.synchronize()
    best = 0.0
    for _ in range(3):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(a, b, out=c); e1.record(); torch.cuda.synchronize()
        best = max(best, 2.0 * n ** 3 / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    reps = max(3, int(seconds / (2.0 * n ** 3 / (best * 1e12))))
    e0.record()
    for _ in range(reps):
        torch.matmul(a, b, out=c)
    e1.record(); torch.cuda.synchronize()
    return {"what": "cuBLAS DGEMM 8192^3 (torch.matmul float64)", "burst_tflops": best,
            "sustained_tflops": 2.0 * n ** 3 * reps / (e0.elapsed_time(e1) * 1e-3) / 1e12, "sustained_reps": reps,
            "dmma_issue_peak_tflops": 148 * 0.25 * 512 * 1.965e9 / 1e12,
            "note": "mma.sync.m8n8k4.f64 issues at 0.25 per cycle per SM (scripts/microbench/dmma_probe.cu)"}


def describe(name, setup, syms):
    p = setup.p
    return (f"{name}: k={p['k']} n_b={setup.n_b} k_GL={p['k_GL']} max_k={p['max_k']} max_l_1p={p['max_l_1p']} "
            f"max_L={p['max_L']} n_sym={len(syms)} sum_n_config={sum(s.n_config for s in syms)}")


class PinnedArrays:
    """CSR output arrays in pinned host memory (bs2e_host_alloc)."""

    def __init__(self, bs2e, nrows, nnzH, nnzS):
        self.bs2e = bs2e
        self.ptrs = []
        self.arrs = (self._mk(nrows + 1, np.int64), self._mk(max(nnzH, 1), np.int64),
                     self._mk(2 * max(nnzH, 1), np.float64), self._mk(nrows + 1, np.int64),
                     self._mk(max(nnzS, 1), np.int64), self._mk(2 * max(nnzS, 1), np.float64))

    def _mk(self, n, dtype):
        p = ctypes.c_void_p()
        nbytes = int(n) * np.dtype(dtype).itemsize
        rc = self.bs2e.lib().bs2e_host_alloc(nbytes, ctypes.byref(p))
        if rc != 0:
            raise RuntimeError(self.bs2e.lib().bs2e_last_error().decode())
        self.ptrs.append(p)
        buf = (ctypes.c_char * nbytes).from_address(p.value)
        return np.frombuffer(buf, dtype=dtype, count=int(n))

    def free(self):
        self.arrs = None
        for p in self.ptrs:
            self.bs2e.lib().bs2e_host_free(p)
        self.ptrs = []


def run_e2e(args, bs2e, ctx, setup, syms, ranges, sizes, S, H_vec, world, barrier,
            max_over_ranks, total_elems, n_rk):
    """Same step through the reference-facing calls with host buffers.  One set of destination arrays sized
    for the largest block is reused for every block (the consumer -- block_diag_CS%store -- takes a block
    before the next one is built; the CSR of cfg4 is 159 GB, more than the host can pin)."""
    full = setup.p["full"]
    nr = max(q[0] for q in sizes)
    mH = max(q[1] for q in sizes)
    mS = max(q[2] for q in sizes)
    steps = max(1, min(args.steps, args.e2e_steps))
    counted = {"h2d": 0, "d2h": 0}

    def step(arrs, count_bytes):
        t0 = time.perf_counter()
        ctx.slater_cells()
        ctx.rk_build()
        ctx.sync()
        t1 = time.perf_counter()
        ctx.set_one_particle(H_vec, S)                      # H2D: one-particle matrices
        if count_bytes:
            counted["h2d"] += sum(h.nbytes for h in H_vec) + S.nbytes
        for s, r, (n, nH, nS) in zip(syms, ranges, sizes):
            out = tuple(a[:m] for a, m in zip(arrs, (n + 1, max(nH, 1), 2 * max(nH, 1),
                                                     n + 1, max(nS, 1), 2 * max(nS, 1))))
            if world == 1:
                nnz = ctx.block_count(s, full)              # H2D configs + device plan + count pass
                ctx.block_fill(s, full, nnz, out=out)       # fill + D2H into the caller's arrays
            else:
                blk = ctx.block_plan(s, full, ranges=r)
                blk.assemble()
                blk.download(out=out)
                blk.free()
            if count_bytes:
                counted["h2d"] += s.conf_n.nbytes + s.conf_l.nbytes
                counted["d2h"] += 2 * 8 * (n + 1) + 24 * (nH + nS)
        ctx.sync()
        t2 = time.perf_counter()
        return t1 - t0, t2 - t1

    pin = PinnedArrays(bs2e, nr, mH, mS)
    step(pin.arrs, False)                                   # warm-up (page-locks, caches)
    barrier()
    tAB = tC = 0.0
    for _ in range(steps):
        a, c = step(pin.arrs, True)
        tAB += a
        tC += c
    barrier()
    tC = max_over_ranks(tC)
    tAB = max_over_ranks(tAB)
    pin.free()
    out = {"value": total_elems * steps / tC, "unit": UNIT, "steps": steps,
           "h2d_bytes_per_step": counted["h2d"] // steps, "d2h_bytes_per_step": counted["d2h"] // steps,
           "ms_per_step_stage_C": tC * 1e3 / steps,
           "d2h_gb_per_s": counted["d2h"] / steps / (tC / steps) / 1e9,
           "rk_integrals_per_s": n_rk * steps / tAB,
           "destination": "pinned host arrays (bs2e_host_alloc), sized for the largest block and reused",
           "api": "bs2e_set_one_particle + bs2e_block_count + bs2e_block_fill"
                  if world == 1 else "bs2e_block_plan_ranges + assemble + download"}
    if not args.no_pageable:
        # the Fortran caller's arrays (H_sp%init, sparse_array_tools.f90:557-569) are ordinary pageable
        # memory: the library stages such downloads through its pinned ring (csrc/download.cu)
        pag = (np.zeros(nr + 1, np.int64), np.zeros(max(mH, 1), np.int64), np.zeros(2 * max(mH, 1)),
               np.zeros(nr + 1, np.int64), np.zeros(max(mS, 1), np.int64), np.zeros(2 * max(mS, 1)))
        for a in pag:
            a.fill(0)                                       # init_CS zero-initialises: the pages exist
        barrier()
        _, c = step(pag, False)
        barrier()
        c = max_over_ranks(c)
        out["pageable"] = {"value": total_elems / c, "unit": UNIT, "steps": 1, "ms_per_step_stage_C": c * 1e3,
                           "destination": "pageable numpy arrays, as the reference caller allocates them"}
        del pag
    return out


# ---------------------------------------------------------------------------
# CPU baseline / reference arm: the oracle port on the host cores
# ---------------------------------------------------------------------------
class CpuArm:
    """The oracle (a C port of the reference path, kind="port") on the host cores: -O3 -march=native build,
    thread count set explicitly (a multi-rank launcher exports OMP_NUM_THREADS=1).  The tensors stage C reads are
    prepared once (untimed, tabulated evaluation on all cores); stage A is timed reference-faithfully on a sample
    of its outermost index, stage B in full (serial like the reference), stage C on row samples."""

    def __init__(self, workload, budget_s=20.0):
        import bs2e
        from oracle import bs2e_oracle as O
        self.O, self.workload = O, workload
        O.use_native_build()
        self.threads = int(os.environ.get("BS2E_CPU_THREADS", os.cpu_count() or 1))
        O.lib().orc_set_threads(self.threads)
        p = self.p = O.basis_params(**bs2e.CONFIGS[workload])
        run = self.run = O.OracleRun(**p)
        K1 = p["max_k"] + 1
        P = run.bs.num_pairs()
        run.slater(tabulate=1, par_mode=1)                      # untimed preparation
        t0 = time.perf_counter()
        run.rk_map()                                             # stage B, serial like the reference
        self.tB = time.perf_counter() - t0
        nnz6 = run.s6.nnz
        est_full = nnz6 * K1 / 40e3 / min(self.threads, K1)      # ~40k values/s/thread
        self.jp_step = max(1, int(np.ceil(est_full / (0.25 * budget_s))))
        t0 = time.perf_counter()
        _, done = O.time_Slater_diag_sample(run.bs, p["max_k"], p["k_GL"], self.jp_step)
        self.tA_s = time.perf_counter() - t0
        self.tA = self.tA_s * (nnz6 * K1) / max(done, 1)
        self.rk_per_s = P * P * K1 / (self.tA + self.tB)
        run.one_particle()
        self.syms = run.basis()
        # calibration of the row cost (count_nnz + construct_block_tensor scan the whole configuration list per row)
        big = max(self.syms, key=lambda q: q.n_config)
        t0 = time.perf_counter()
        self._rows(big, [(big.n_config // 2, big.n_config // 2 + 3)])
        self.row_s_per_config = (time.perf_counter() - t0) / 4 / big.n_config

    def _rows(self, s, chunks):
        O, p, el = self.O, self.p, 0
        for lo, hi in chunks:
            cap = O.count_nnz(self.run.bs.k, s, p["max_k"], p["full"], rows=(lo, hi))   # count_nnz scan
            _, _, em = self.run.block(s, rows=(lo, hi), nnz=cap)                          # construct_block_tensor
            el += em[0] + em[1]
        return el

    def stage_C_sample(self, budget_s):
        from concurrent.futures import ThreadPoolExecutor
        syms, chunk = self.syms, 8
        est = sum(self.row_s_per_config * s.n_config * s.n_config for s in syms) / self.threads
        frac = min(1.0, budget_s / max(est, 1e-9))
        tasks = []
        for s in syms:
            n = s.n_config
            nch = max(1, int(round(frac * n / chunk)))
            starts = np.linspace(1, max(1, n - chunk + 1), nch).astype(int)
            tasks += [(s, [(int(a), int(min(n, a + chunk - 1)))]) for a in starts]
        t0 = time.perf_counter()
        with ThreadPoolExecutor(max_workers=self.threads) as ex:
            elems = sum(ex.map(lambda t: self._rows(*t), tasks))
        dt = time.perf_counter() - t0
        if frac < 1.0:                                           # keep the next sample on its budget
            self.row_s_per_config *= min(4.0, max(0.25, dt / budget_s))
        return elems, dt, frac, chunk

    def result(self, elems, dt, frac, chunk):
        return {"value": elems / dt, "unit": UNIT, "cores": self.threads, "kind": "port",
                "rk_integrals_per_s": self.rk_per_s,
                "sample": (f"{self.workload}: stage C on {frac * 100:.3f}% of the rows of every symmetry block "
                           f"(evenly spaced {chunk}-row chunks, count_nnz + construct_block_tensor, "
                           f"chunks dealt to {self.threads} threads, {dt:.1f} s); stage A on every "
                           f"{self.jp_step}-th outer index ({self.tA_s:.1f} s, extrapolated to {self.tA:.0f} s), "
                           f"stage B full ({self.tB:.1f} s). The port memoises the 3j/6j factors and indexes R^k "
                           "directly, so it is faster than the Fortran; built -O3 -march=native on this box."),
                "omp_threads_set": self.threads, "cpu_count": os.cpu_count()}


def cpu_baseline(workload, budget_s=20.0):
    arm = CpuArm(workload, budget_s)
    return arm.result(*arm.stage_C_sample(0.5 * budget_s))


def run_reference(args):
    rank = env_int("RANK", 0)
    if rank != 0:
        return
    arm = CpuArm(args.workload, args.cpu_budget)            # preparation once, not per step
    per_step = max(0.5, min(0.5 * args.cpu_budget, 150.0 / max(1, args.steps + args.warmup)))
    for _ in range(args.warmup):
        arm.stage_C_sample(per_step)
    vals, last = [], None
    t0 = time.perf_counter()
    for _ in range(args.steps):
        last = arm.stage_C_sample(per_step)
        vals.append(last[0] / last[1])
    wall = time.perf_counter() - t0
    v = float(np.mean(vals))
    cb = arm.result(*last)
    cb["value"] = v
    out = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": wall * 1e3 / max(1, args.steps),
           "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
           "data": "synthetic", "config": {"workload": WORKLOADS[args.workload]},
           "rk_integrals_per_s": arm.rk_per_s, "cpu_baseline": cb,
           "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0,
           "note": "reference cannot be built here (no Fortran compiler); the C oracle port is timed instead; "
                   "a step = one bounded row sample of stage C (the value), stage A/B timed once"}
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD)
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--cpu-budget", type=float, default=20.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer leg (profiling runs)")
    ap.add_argument("--no-pageable", action="store_true", help="skip the pageable-destination pass of the e2e leg")
    ap.add_argument("--no-fp64-peak", action="store_true", help="skip the cuBLAS DGEMM measurement")
    ap.add_argument("--whole-rk", action="store_true", help="N>1: every rank builds the whole R^k tensor (A/B)")
    ap.add_argument("--rebalance", type=int, default=3,
                    help="N>1: steps of measured rebalancing of the row partition in the untimed set-up")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
