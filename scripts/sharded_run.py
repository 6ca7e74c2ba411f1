#!/usr/bin/env python
"""A whole BASELINE configuration sharded over the GPUs of one box, one symmetry block
at a time, output kept on the device and reduced to checksums (the CSR of cfg4/cfg5 is
0.5-3 TB, far more than the host can take).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
        --master-port 29520 scripts/sharded_run.py cfg5

Every rank builds the whole R^k tensor, then for every symmetry block: count pass on all
rows (weights of the partition), rows dealt by first radial index (bs2e.sharding.
site_partition), plan + count + fill of the rank's share (device-timed), checksum, free.
Rank 0 prints one JSON line: stored elements, elements/s over all ranks (max-over-ranks
stage-C time), per-block figures and the XOR of the fragment checksums.
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "b-spline-two-e_b200")):
    sys.path.insert(0, p)


def main():
    import torch
    import torch.distributed as dist
    import bs2e
    from bs2e.sharding import exchange_cost, site_partition

    workload = sys.argv[1] if len(sys.argv) > 1 else "cfg4"
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    # profiling aid: BS2E_FAKE_SHARD="r/n" runs the share of rank r of n on a single GPU
    fake = os.environ.get("BS2E_FAKE_SHARD")
    torch.cuda.set_device(local)
    if fake:
        fr, fn = (int(v) for v in fake.split("/"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def allred(x, op):
        if world == 1:
            return float(x)
        t = torch.tensor([float(x)], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=op)
        return float(t.item())

    setup = bs2e.BasisSetup(device=local, **bs2e.CONFIGS[workload])
    S, H_vec, syms = setup.host_inputs()
    ctx = setup.open()
    stream = torch.cuda.Stream()
    ctx.set_stream(stream.cuda_stream)
    full = setup.p["full"]
    ev = lambda: torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0 = ev(); e0.record(stream)
        ctx.slater_cells()
        ctx.rk_build()
        e1 = ev(); e1.record(stream)
    ctx.set_one_particle(H_vec, S)
    ctx.sync()
    t_ab = e0.elapsed_time(e1)
    n_rk = ctx.P * ctx.P * (setup.p["max_k"] + 1)

    per_block, tot_ms, my_elems, xor = [], 0.0, 0, 0
    t_wall = time.perf_counter()
    only = os.environ.get("BS2E_ONLY_BLOCKS")
    if only:
        syms = [syms[int(q)] for q in only.split(",")]
    for s in syms:
        whole = ctx.block_plan(s, full)                     # counts of all rows: partition weights
        cH, cS = whole.row_counts()
        nnz_all = whole.nnz_H + whole.nnz_S
        whole.free()
        xc = exchange_cost(setup.p['max_k'])
        mine = (site_partition(s.conf_n, cH + cS, fn, setup.k, xc)[fr] if fake
                else site_partition(s.conf_n, cH + cS, world, setup.k, xc)[rank])
        ms, el = 0.0, 0
        if mine:
            blk = ctx.block_plan(s, full, ranges=mine)
            blk.assemble()                                  # allocation + first fill, untimed
            ctx.sync()
            with torch.cuda.stream(stream):
                a = ev(); a.record(stream)
                ctx.blocks_run([blk], recount=True)         # count + scan + fill, device-timed
                b = ev(); b.record(stream)
            ctx.sync()
            ms = a.elapsed_time(b)
            el = blk.nnz_H + blk.nnz_S
            cs = blk.checksum()
            xor ^= cs[0] ^ cs[1]
            blk.free()
        ms_max = allred(ms, dist.ReduceOp.MAX if world > 1 else None)
        el_sum = allred(el, dist.ReduceOp.SUM if world > 1 else None)
        assert fake or int(el_sum) == nnz_all, (el_sum, nnz_all)    # the shares tile the block
        per_block.append({"L": s.l, "pi": s.pi, "n_config": s.n_config, "elements": int(el_sum),
                          "ms": ms_max, "csr_gb": 24e-9 * el_sum})
        tot_ms += ms_max
        my_elems += el
    wall = time.perf_counter() - t_wall
    total = sum(b["elements"] for b in per_block)
    phases = bs2e.site_phase_cycles().tolist()
    if rank == 0:
        print(json.dumps({"workload": workload, "n_gpus": world, "elements": total, "csr_tb": 24e-12 * total,
                          "stage_C_ms": tot_ms, "elements_per_s": total / (tot_ms * 1e-3),
                          "stage_AB_ms": t_ab, "rk_integrals": n_rk, "rk_integrals_per_s": n_rk / (t_ab * 1e-3),
                          "wall_s_incl_planning": wall, "checksum_xor_rank0": xor,
                          "site_phase_cycles": phases if any(phases) else None, "blocks": per_block}))
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
