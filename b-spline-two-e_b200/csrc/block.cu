// block.cu -- stage C: assembly of one (L, parity) symmetry block of the
// coupled two-electron Hamiltonian H and overlap S as CSR matrices.
//
// Stands in for count_nnz (src/mat_els/hamiltonian.f90:348-416) and
// construct_block_tensor (:106-283) with the element formulas of
// src/mat_els/mat_els.f90:552-571,608-633,664-715.  The reference scans all
// n_config^2 configuration pairs twice and evaluates 3j/6j symbols for every
// pair; here
//   * the angular factors are tabulated once per pair of (l1,l2) blocks on
//     the host (exact arithmetic, wigner.cpp) and uploaded,
//   * the band partners of a row are GENERATED from the block structure of
//     the configuration list (no pair scan): block_count_kernel counts them
//     in closed form, an exclusive scan builds index_ptr, and
//     block_fill_kernel (one warp per row) writes indices and values in
//     ascending column order with coalesced 8/16-byte stores.
// Bound: HBM (24 B written per stored element + R^k gathers).
#include <cub/device/device_scan.cuh>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <set>
#include <unordered_map>

#include <cuda.h>

#include "ctx.h"
#include "plan.h"
#include "site_core.h"

namespace bs2e {

// ---------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------
__global__ void block_count_kernel(Geom g, Plan pl, long long row_lo, long long nrows,
                                   long long* __restrict__ cntH, long long* __restrict__ cntS)
{
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx > nrows) return;
    long long h = 0, s = 0;
    if (idx < nrows) row_count(g, pl, (int)(row_lo + idx), &h, &s);
    cntH[idx] = h;  // slot nrows holds 0 so that the scan yields the total
    cntS[idx] = s;
}

__global__ void ptr_one_based_kernel(long long n, long long* __restrict__ a, long long* __restrict__ b)
{
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n) return;
    a[idx] += 1;
    b[idx] += 1;
}

constexpr int kFillWarps = 8;

__global__ void __launch_bounds__(kFillWarps * 32)
block_fill_kernel(Geom g, Plan pl, OneBody ob, const double* __restrict__ R, long long row_lo,
                  long long nrows, const long long* __restrict__ Hptr,
                  const long long* __restrict__ Sptr, long long* __restrict__ Hidx,
                  double2* __restrict__ Hdat, long long* __restrict__ Sidx,
                  double2* __restrict__ Sdat)
{
    const long long wrow = (long long)blockIdx.x * kFillWarps + (threadIdx.x >> 5);
    if (wrow >= nrows) return;
    const int lane = threadIdx.x & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    const RowInfo r = row_info(pl, (int)(row_lo + wrow));
    long long hpos = Hptr[wrow] - 1, spos = Sptr[wrow] - 1;

    for_each_chunk(g, pl, r, [&](int bj, int nc, const Segment& s, const Coupling& c, int base, int hi) {
        const int nd = base + lane;
        const bool act = nd <= hi;
        bool storeS = false;
        Element e;
        if (act) {
            const bool sup = nd >= s.dlo && nd <= s.dhi;
            const bool sup_ex = nd >= s.xlo && nd <= s.xhi;
            e = element_value(g, pl, ob, R, r, c, bj, nc, nd, sup, sup_ex);
            storeS = e.storeS;
            Hidx[hpos + lane] = (long long)s.jbase + nd;
            Hdat[hpos + lane] = make_double2(e.H.re, e.H.im);
        }
        hpos += imin(32, hi - base + 1);
        const unsigned m = __ballot_sync(0xffffffffu, storeS);
        if (storeS) {
            const long long pos = spos + __popc(m & lt_mask);
            Sidx[pos] = (long long)s.jbase + nd;
            Sdat[pos] = make_double2(e.S.re, e.S.im);
        }
        spos += __popc(m);
    });
}

// ---------------------------------------------------------------------------
// site-centric fill: one CTA per radial site (n_a,n_b), see site_core.h.
//   phase 1  per column block bj and n_c slot: clipped windows T[bj][q];
//            R^k values of both windows -> shared memory, Rv[k][slot]
//   phase 2  per (bj, storage mode): prefix over the n_c slots of the number of
//            stored entries, hp[bj][mode][q] (mode: D, X, D+X, diagonal pair)
//   phase 3  per row of the site: offsets of its column blocks inside the row
//   phase 4  warps grab (row, column block) pairs from a shared counter and
//            walk the OUTPUT positions of the pair 32 at a time (all lanes
//            busy): n_c slot by binary search in hp, n_d by interval
//            arithmetic, sum_k ang_k R^k from shared memory, coalesced stores
// Nothing is computed per pair except the copy of its 2*K1 angular factors.
// ---------------------------------------------------------------------------
struct SiteList {
    const unsigned* key;  // [nsites]  n_a << 16 | n_b
    const int* ptr;       // [nsites+1] into rows
    const int* rows;      // 1-based configuration (row) indices, ascending per site
    int nsites;
};

struct alignas(16) RowCache {  // per row of the site: what a pair needs to know about its row
    long long hbase, sbase;    // 0-based position of the first H / S entry of the row
    int bi, la, lb, pad;
};

struct SiteSmem {   // element counts of the dynamic shared memory carve-up
    int nsmax;      // slots (stride of Rv over k) = entries of list DX at most
    int ncmax;      // n_c slots
    int cap;        // (row, column block) pairs per group
    size_t bytes;
};

__host__ __device__ inline size_t site_smem_bytes(const Geom& g, int nblk, int nw, int cap)
{
    const size_t nsmax = (size_t)site_max_slots(g), ncmax = (size_t)site_max_nc(g);
    size_t b = 0;
    b += sizeof(double) * 2 * (size_t)site_win_doubles(g);             // Rv: D and X windows
    b += sizeof(double) * (size_t)nw * 2 * g.K1;                       // wang
    b += sizeof(SiteEntry) * (size_t)nblk * ncmax;                     // T
    b += sizeof(RowCache) * (size_t)nblk;                              // rcache
    b += sizeof(Cand) * 2 * nsmax;                                     // listD + listX (half each), listDX
    b += sizeof(int2) * (size_t)cap;                                   // plist
    b += sizeof(int) * (ncmax + 1);                                    // cprefix
    b += sizeof(unsigned short) * (size_t)nblk * kModes * (ncmax + 1);  // hp
    b += sizeof(unsigned short) * (size_t)nblk * (ncmax + 1);           // sp
    return b + 16;
}

__device__ __forceinline__ int warp_incl_scan(int v, int lane)
{
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += u;
    }
    return v;
}

constexpr int kSiteWarps = 8;
constexpr int kSiteMinBlocks = 4;
constexpr size_t kSiteSmemLimit = 200 * 1024;

template <int NW>
__global__ void __launch_bounds__(NW * 32, kSiteMinBlocks)
site_fill_kernel(const __grid_constant__ CUtensorMap tmapR, Geom g, Plan pl, OneBody ob, SiteList sl, SiteSmem lay,
                 long long row_lo, const long long* __restrict__ Hptr,
                 const long long* __restrict__ Sptr, long long* __restrict__ Hidx,
                 double2* __restrict__ Hdat, long long* __restrict__ Sidx,
                 double2* __restrict__ Sdat)
{
    extern __shared__ __align__(128) unsigned char smraw[];
    const int K1 = g.K1, nsmax = lay.nsmax, ncmax = lay.ncmax, cap = lay.cap;
    const int nblk = pl.nblk;
    double* Rv = reinterpret_cast<double*>(smraw);
    double* wang_all = Rv + 2 * (size_t)site_win_doubles(g);
    SiteEntry* T = reinterpret_cast<SiteEntry*>(wang_all + (size_t)NW * 2 * K1);
    RowCache* rcache = reinterpret_cast<RowCache*>(T + (size_t)nblk * ncmax);  // rows of the group
    Cand* listD = reinterpret_cast<Cand*>(rcache + nblk);
    Cand* listX = listD + nsmax / 2;
    Cand* listDX = listX + nsmax / 2;
    int2* plist = reinterpret_cast<int2*>(listDX + nsmax);                     // coupled pairs of the group
    int* cprefix = reinterpret_cast<int*>(plist + cap);
    unsigned short* hp = reinterpret_cast<unsigned short*>(cprefix + ncmax + 1);
    unsigned short* sp = hp + (size_t)nblk * kModes * (ncmax + 1);
    __shared__ int s_next, s_npairs;
    __shared__ __align__(8) unsigned long long s_bar;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int sidx = blockIdx.x;
    const unsigned key = sl.key[sidx];
    const Site s = make_site(g, (int)(key >> 16), (int)(key & 0xffffu));
    const int* srows = sl.rows + sl.ptr[sidx];
    const int nr = sl.ptr[sidx + 1] - sl.ptr[sidx];
    const int nnc = s.nnc;

    // ---- phase 0: both R^k windows of the site, [K1][2w+1][cpad] boxes of the tensor
    //      R[k][p1][p2], dropped into shared memory by two TMA tile loads ----
    const unsigned bar = (unsigned)__cvta_generic_to_shared(&s_bar);
    // if n_a - w exceeds every n_2 of the basis, all exchange windows of the site are clipped away
    const bool wantX = s.dXlo <= pl.max_nd;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        const unsigned box_bytes = (unsigned)(K1 * s.kst * sizeof(double));
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar),
                     "r"(wantX ? 2 * box_bytes : box_bytes)
                     : "memory");
        const unsigned dstD = (unsigned)__cvta_generic_to_shared(Rv);
        asm volatile(
            "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
            " [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dstD),
            "l"(&tmapR), "r"(site_colD(g, s)), "r"(site_rowD(g, s)), "r"(0), "r"(bar)
            : "memory");
        if (wantX) {
            const unsigned dstX = (unsigned)__cvta_generic_to_shared(Rv + s.xoff);
            asm volatile(
                "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
                " [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dstX),
                "l"(&tmapR), "r"(site_colX(g, s)), "r"(site_rowX(g, s)), "r"(0), "r"(bar)
                : "memory");
        }
    }
    // ---- phase 1a: clipped windows per column block, candidate lists D and X,
    //      slot prefix of list DX ----
    if (warp == NW - 1 && wantX) {
        int run = 0;
        for (int q0 = 0; q0 < nnc; q0 += 32) {
            const int q = q0 + lane;
            const int c = q < nnc ? site_cand_DX_count(s, q) : 0;
            const int inc = warp_incl_scan(c, lane);
            if (q < nnc) cprefix[q] = run + inc - c;
            run += __shfl_sync(0xffffffffu, inc, 31);
        }
        if (lane == 0) cprefix[nnc] = run;
    }
    for (int q = lane; q < nnc; q += 32)
        for (int bj = warp; bj < nblk; bj += NW) T[bj * ncmax + q] = site_entry(g, pl, s, bj, q);
    for (int t = tid; t < s.nD; t += NW * 32) listD[t] = site_cand_D(s, t);
    if (wantX)
        for (int t = tid; t < s.nX; t += NW * 32) listX[t] = site_cand_X(s, t);
    __syncthreads();
    // ---- phase 1b: list DX (CSR order over both windows) ----
    for (int q = warp; wantX && q < nnc; q += NW) {
        const int n = cprefix[q + 1] - cprefix[q], base = cprefix[q];
        for (int idx = lane; idx < n; idx += 32) listDX[base + idx] = site_cand_DX(s, q, idx);
    }
    // ---- phase 2: prefix of stored entries over the n_c slots, per (bj, mode);
    //      one thread per (bj, mode), serial over the slots ----
    for (int task = tid; task < nblk * kModes; task += NW * 32) {
        const int bj = task / kModes, mode = task - bj * kModes;
        if (!wantX && (mode == kModeX || mode == kModeDX)) continue;  // never read, see tot_of()
        const bool useD = mode_useD(mode), useX = mode_useX(mode);
        const bool diag = mode == kModeDiag;
        const bool samex = diag && pl.blk[bj].l1 == pl.blk[bj].l2;
        unsigned short* hpq = hp + task * (ncmax + 1);
        unsigned short* spq = sp + bj * (ncmax + 1);
        int run = 0, srun = 0;
        for (int q = 0; q < nnc; ++q) {
            SiteEntry e = T[bj * ncmax + q];
            if (diag && !pl.full) e = entry_cut(e, s, site_nc(s, q));
            hpq[q] = (unsigned short)run;
            run += entry_count(e, useD, useX);
            if (diag) {
                spq[q] = (unsigned short)srun;
                srun += entry_count(e, true, samex);
            }
        }
        hpq[nnc] = (unsigned short)run;
    }

    double* wang = wang_all + warp * 2 * K1;
    const int G = cap / nblk;  // rows per group (host guarantees >= 1)
    double* const Hd = reinterpret_cast<double*>(Hdat);
    double* const Sd = reinterpret_cast<double*>(Sdat);
    // first live candidate of each list when n_c < n_a is cut away (diagonal pair, j >= i)
    const int cutD = (s.na - s.cDlo) * s.dw;
    const int cutX = imin(s.nX, imax(0, s.na - s.cXlo) * s.xw);

    for (int g0 = 0; g0 < nr; g0 += G) {
        const int gr = imin(G, nr - g0);
        __syncthreads();  // phase 2 / previous group finished
        if (tid == 0) { s_next = 0; s_npairs = 0; }
        __syncthreads();
        // ---- phase 3: coupled pairs of each row and their offsets inside the row ----
        for (int ri = warp; ri < gr; ri += NW) {
            const int rowi = srows[g0 + ri];
            const RowInfo r = row_info(pl, rowi);
            if (lane == 0) {
                const long long wrow = (long long)rowi - row_lo;
                rcache[ri] = RowCache{Hptr[wrow] - 1, Sptr[wrow] - 1, r.bi, r.la, r.lb, 0};
            }
            int run = 0;
            for (int b0 = 0; b0 < nblk; b0 += 32) {
                const int bj = b0 + lane;
                int c = 0, mode = -1;
                if (bj < nblk) {
                    mode = pair_mode(pl, r, bj);
                    if (mode >= 0) {
                        const unsigned short* hb = hp + (bj * kModes) * (ncmax + 1) + nnc;
                        mode = effective_mode(mode, hb[kModeD * (ncmax + 1)], wantX ? hb[kModeX * (ncmax + 1)] : 0);
                        c = (wantX || mode != kModeX) ? hb[mode * (ncmax + 1)] : 0;
                    }
                }
                const int inc = warp_incl_scan(c, lane);
                const unsigned live = __ballot_sync(0xffffffffu, c > 0);
                int base = 0;
                if (lane == 0 && live) base = atomicAdd(&s_npairs, __popc(live));
                base = __shfl_sync(0xffffffffu, base, 0);
                if (c > 0) {
                    const int slot = base + __popc(live & ((1u << lane) - 1u));
                    plist[slot] = make_int2(ri | (bj << 12) | (mode << 24), run + inc - c);
                }
                run += __shfl_sync(0xffffffffu, inc, 31);
            }
        }
        __syncthreads();
        const int npairs = s_npairs;
        if (g0 == 0) {  // the staged windows must have landed
            unsigned done = 0;
            while (!done)
                asm volatile(
                    "{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }"
                    : "=r"(done)
                    : "r"(bar)
                    : "memory");
        }
        // ---- phase 4: fill ----
        for (;;) {
            int p = 0;
            if (lane == 0) p = atomicAdd(&s_next, 1);
            p = __shfl_sync(0xffffffffu, p, 0);
            if (p >= npairs) break;
            const int2 pe = plist[p];
            const int ri = pe.x & 0xfff, bj = (pe.x >> 12) & 0xfff, mode = pe.x >> 24;
            const RowCache rc = rcache[ri];
            RowInfo r;
            r.i = 0;
            r.bi = rc.bi;
            r.na = s.na;
            r.nb = s.nb;
            r.la = rc.la;
            r.lb = rc.lb;
            const int cpl = r.bi * nblk + bj;
            const unsigned fl = pl.flags[cpl];
            PairCtx pc;
            pc.pk = pair_k(pl.krange[cpl]);
            __syncwarp();
            for (int i = lane; i < pc.pk.nkd; i += 32) wang[i] = pl.angD[(size_t)cpl * K1 + pc.pk.dlo + 2 * i];
            for (int i = lane; i < pc.pk.nkx; i += 32) wang[K1 + i] = pl.angX[(size_t)cpl * K1 + pc.pk.xlo + 2 * i];
            __syncwarp();
            pc.Tb = T + bj * ncmax;
            pc.hpq = hp + (bj * kModes + mode) * (ncmax + 1);
            pc.spq = sp + bj * (ncmax + 1);
            pc.Rv = Rv;
            pc.wa_d = wang;
            pc.wa_x = wang + K1;
            pc.kst = s.kst;
            pc.bj = bj;
            pc.diag = mode == kModeDiag;
            pc.dirany = (fl & kDirAny) != 0;
            pc.exany = (fl & kExAny) != 0;
            pc.samex = pc.diag && r.la == r.lb;
            pc.cut = pc.diag && !pl.full;
            pc.hbase = rc.hbase + pe.y;
            pc.sbase = rc.sbase;
            const unsigned short* hb = hp + (bj * kModes) * (ncmax + 1) + nnc;
            const int win = pair_window(mode, hb[kModeD * (ncmax + 1)], wantX ? hb[kModeX * (ncmax + 1)] : 0);
            if (win == kModeD) {
                for (int t = (pc.cut ? cutD : 0) + lane; t < s.nD; t += 32)
                    site_item<kModeD>(g, pl, ob, s, r, pc, listD[t], Hidx, Hd, Sidx, Sd);
            } else if (win == kModeX) {
                for (int t = (pc.cut ? cutX : 0) + lane; t < s.nX; t += 32)
                    site_item<kModeX>(g, pl, ob, s, r, pc, listX[t], Hidx, Hd, Sidx, Sd);
            } else {
                const int ncand = cprefix[nnc];
                for (int t = (pc.cut ? cprefix[union_pos(s, s.na)] : 0) + lane; t < ncand; t += 32)
                    site_item<kModeDX>(g, pl, ob, s, r, pc, listDX[t], Hidx, Hd, Sidx, Sd);
            }
        }
    }
}

__global__ void checksum_kernel(long long n, const long long* __restrict__ idx,
                                const double* __restrict__ dat, unsigned long long* out)
{
    unsigned long long acc = 0;
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < n;
         q += (long long)gridDim.x * blockDim.x) {
        const unsigned long long w = (unsigned long long)(q + 1);
        acc += w * (unsigned long long)idx[q];
        acc += w * (unsigned long long)__double_as_longlong(dat[2 * q]);
        acc += (w << 1) * (unsigned long long)__double_as_longlong(dat[2 * q + 1]);
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(out, acc);
}

// count pass + exclusive scan -> 1-based index_ptr of the planned rows
void block_count_scan(bs2e_block* b, bool read_totals)
{
    bs2e_ctx* c = b->ctx;
    cudaStream_t st = c->stream;
    const long long nrows = b->row_hi - b->row_lo + 1;
    block_count_kernel<<<(unsigned)((nrows + 1 + 127) / 128), 128, 0, st>>>(
        c->dg, b->dplan, b->row_lo, nrows, b->d_cntH, b->d_cntS);
    BS2E_LAUNCHED();
    size_t tmp = b->scan_tmp_bytes;
    BS2E_CUDA(cub::DeviceScan::ExclusiveSum(b->d_scan_tmp, tmp, b->d_cntH, b->d_Hptr, nrows + 1, st));
    g_launches.fetch_add(2);  // DeviceScanInitKernel + DeviceScanKernel
    BS2E_CUDA(cub::DeviceScan::ExclusiveSum(b->d_scan_tmp, tmp, b->d_cntS, b->d_Sptr, nrows + 1, st));
    g_launches.fetch_add(2);
    ptr_one_based_kernel<<<(unsigned)((nrows + 1 + 255) / 256), 256, 0, st>>>(nrows + 1, b->d_Hptr,
                                                                             b->d_Sptr);
    BS2E_LAUNCHED();
    if (read_totals) {
        long long lastH = 0, lastS = 0;
        BS2E_CUDA(cudaMemcpyAsync(&lastH, b->d_Hptr + nrows, sizeof(long long),
                                  cudaMemcpyDeviceToHost, st));
        BS2E_CUDA(cudaMemcpyAsync(&lastS, b->d_Sptr + nrows, sizeof(long long),
                                  cudaMemcpyDeviceToHost, st));
        BS2E_CUDA(cudaStreamSynchronize(st));
        b->nnzH = lastH - 1;
        b->nnzS = lastS - 1;
    }
}

// ---------------------------------------------------------------------------
// plan: derive the block structure from the configuration list, count, scan
// ---------------------------------------------------------------------------
bs2e_block* block_plan(bs2e_ctx* c, int L, long long n_config, const int64_t* conf_n,
                       const int64_t* conf_l, int full, long long row_lo, long long row_hi)
{
    HostPlan hp;
    try {
        hp = build_host_plan(c->hg, L, n_config, conf_n, conf_l, full, row_lo, row_hi);
    } catch (const std::invalid_argument& e) {
        throw Error(e.what());
    }
    const int nblk = hp.nblk;
    bs2e_block* b = new bs2e_block();
    b->ctx = c;
    b->L = L;
    b->full = full ? 1 : 0;
    b->n_config = n_config;
    b->row_lo = row_lo;
    b->row_hi = row_hi;
    b->lmax = hp.lmax;
    try {
        cudaStream_t st = c->stream;
        b->d_blk = dev_upload(hp.blocks, st);
        b->d_ncrow = dev_upload(hp.ncrow, st);
        b->d_flags = dev_upload(hp.flags, st);
        b->d_krange = dev_upload(hp.krange, st);
        b->d_angD = dev_upload(hp.angD, st);
        b->d_angX = dev_upload(hp.angX, st);
        b->d_row_n1 = dev_upload(hp.row_n1, st);
        b->d_row_n2 = dev_upload(hp.row_n2, st);
        b->d_row_blk = dev_upload(hp.row_blk, st);
        b->nsites = (int)hp.site_key.size();
        if (b->nsites > 0) {
            b->d_site_key = dev_upload(hp.site_key, st);
            b->d_site_ptr = dev_upload(hp.site_ptr, st);
            b->d_site_rows = dev_upload(hp.site_rows, st);
        }
        Plan& pl = b->dplan;
        pl.nblk = nblk;
        pl.n_config = (int)n_config;
        pl.full = b->full;
        pl.L = L;
        pl.blk = b->d_blk;
        pl.ncrow = b->d_ncrow;
        pl.flags = b->d_flags;
        pl.krange = b->d_krange;
        pl.angD = b->d_angD;
        pl.angX = b->d_angX;
        pl.row_n1 = b->d_row_n1;
        pl.row_n2 = b->d_row_n2;
        pl.row_blk = b->d_row_blk;
        pl.max_nd = hp.max_nd;

        const long long nrows = row_hi - row_lo + 1;
        b->d_cntH = dev_alloc<long long>(nrows + 1);
        b->d_cntS = dev_alloc<long long>(nrows + 1);
        b->d_Hptr = dev_alloc<long long>(nrows + 1);
        b->d_Sptr = dev_alloc<long long>(nrows + 1);
        size_t tmp = 0;
        BS2E_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp, b->d_cntH, b->d_Hptr, nrows + 1, st));
        b->scan_tmp_bytes = tmp;
        BS2E_CUDA(cudaMalloc(&b->d_scan_tmp, tmp ? tmp : 1));
        block_count_scan(b, true);
    } catch (...) {
        block_free(b);
        throw;
    }
    return b;
}

void block_assemble(bs2e_block* b)
{
    bs2e_ctx* c = b->ctx;
    if (!c->have_R) throw Error("block_assemble: call bs2e_rk_build first");
    if (!c->have_1p) throw Error("block_assemble: call bs2e_set_one_particle first");
    if (b->lmax > c->lmax_1p) throw Error("block_assemble: configuration l exceeds max_l_1p of H_vec");
    if (!b->d_Hidx) {
        b->d_Hidx = dev_alloc<long long>(b->nnzH);
        b->d_Hdat = dev_alloc<double>(2 * (size_t)b->nnzH);
        b->d_Sidx = dev_alloc<long long>(b->nnzS);
        b->d_Sdat = dev_alloc<double>(2 * (size_t)b->nnzS);
    }
    const long long nrows = b->row_hi - b->row_lo + 1;
    // site kernel unless its shared-memory windows do not fit (large k_s * max_k)
    // or BS2E_FILL=row asks for the row kernel (kept for A/B measurements)
    constexpr int NW = kSiteWarps;
    const Geom& g = c->dg;
    const char* mode = getenv("BS2E_FILL");
    bool use_site = b->nsites > 0 && c->have_tmap && !(mode && strcmp(mode, "row") == 0) && site_max_nc(g) <= 255;
    SiteSmem lay{};
    if (use_site) {
        lay.nsmax = site_max_slots(g);
        lay.ncmax = site_max_nc(g);
        const int cap_want = std::max(b->dplan.nblk, std::min(2048, b->dplan.nblk * b->dplan.nblk));
        lay.cap = cap_want;
        lay.bytes = site_smem_bytes(g, b->dplan.nblk, NW, lay.cap);
        if (lay.bytes > kSiteSmemLimit) {  // shrink the pair group before giving up
            lay.cap = b->dplan.nblk;
            lay.bytes = site_smem_bytes(g, b->dplan.nblk, NW, lay.cap);
        }
        if (lay.bytes > kSiteSmemLimit || 2 * site_win_doubles(g) + site_kst(g) > 65535 || b->dplan.nblk > 4095)
            use_site = false;
    }
    if (use_site) {
        BS2E_CUDA(cudaFuncSetAttribute(site_fill_kernel<NW>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)lay.bytes));
        const SiteList sl{b->d_site_key, b->d_site_ptr, b->d_site_rows, b->nsites};
        site_fill_kernel<NW><<<(unsigned)b->nsites, NW * 32, lay.bytes, c->stream>>>(
            c->tmapR, c->dg, b->dplan, c->one_body(), sl, lay, b->row_lo, b->d_Hptr, b->d_Sptr, b->d_Hidx,
            reinterpret_cast<double2*>(b->d_Hdat), b->d_Sidx, reinterpret_cast<double2*>(b->d_Sdat));
    } else {
        block_fill_kernel<<<(unsigned)((nrows + kFillWarps - 1) / kFillWarps), kFillWarps * 32, 0,
                            c->stream>>>(c->dg, b->dplan, c->one_body(), c->d_R, b->row_lo, nrows,
                                         b->d_Hptr, b->d_Sptr, b->d_Hidx,
                                         reinterpret_cast<double2*>(b->d_Hdat), b->d_Sidx,
                                         reinterpret_cast<double2*>(b->d_Sdat));
    }
    BS2E_LAUNCHED();
    b->assembled = true;
}

void block_download(bs2e_block* b, int64_t* H_ptr, int64_t* H_idx, double* H_dat, int64_t* S_ptr,
                    int64_t* S_idx, double* S_dat)
{
    if (!b->assembled) throw Error("block_download: call bs2e_block_assemble first");
    cudaStream_t st = b->ctx->stream;
    const long long nrows = b->row_hi - b->row_lo + 1;
    auto d2h = [&](void* dst, const void* src, size_t bytes) {
        if (dst && bytes) BS2E_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, st));
    };
    d2h(H_ptr, b->d_Hptr, sizeof(long long) * (nrows + 1));
    d2h(S_ptr, b->d_Sptr, sizeof(long long) * (nrows + 1));
    d2h(H_idx, b->d_Hidx, sizeof(long long) * b->nnzH);
    d2h(S_idx, b->d_Sidx, sizeof(long long) * b->nnzS);
    d2h(H_dat, b->d_Hdat, sizeof(double) * 2 * b->nnzH);
    d2h(S_dat, b->d_Sdat, sizeof(double) * 2 * b->nnzS);
    BS2E_CUDA(cudaStreamSynchronize(st));
}

void block_row_counts(bs2e_block* b, int64_t* cH, int64_t* cS)
{
    cudaStream_t st = b->ctx->stream;
    const long long nrows = b->row_hi - b->row_lo + 1;
    if (cH) BS2E_CUDA(cudaMemcpyAsync(cH, b->d_cntH, sizeof(long long) * nrows, cudaMemcpyDeviceToHost, st));
    if (cS) BS2E_CUDA(cudaMemcpyAsync(cS, b->d_cntS, sizeof(long long) * nrows, cudaMemcpyDeviceToHost, st));
    BS2E_CUDA(cudaStreamSynchronize(st));
}

void block_checksum(bs2e_block* b, uint64_t* sH, uint64_t* sS)
{
    if (!b->assembled) throw Error("block_checksum: call bs2e_block_assemble first");
    cudaStream_t st = b->ctx->stream;
    unsigned long long* d = dev_alloc<unsigned long long>(2);
    unsigned long long h[2] = {0, 0};
    try {
        BS2E_CUDA(cudaMemsetAsync(d, 0, 2 * sizeof(unsigned long long), st));
        if (b->nnzH > 0) {
            checksum_kernel<<<148 * 8, 256, 0, st>>>(b->nnzH, b->d_Hidx, b->d_Hdat, d);
            BS2E_LAUNCHED();
        }
        if (b->nnzS > 0) {
            checksum_kernel<<<148 * 8, 256, 0, st>>>(b->nnzS, b->d_Sidx, b->d_Sdat, d + 1);
            BS2E_LAUNCHED();
        }
        BS2E_CUDA(cudaMemcpyAsync(h, d, sizeof(h), cudaMemcpyDeviceToHost, st));
        BS2E_CUDA(cudaStreamSynchronize(st));
    } catch (...) {
        cudaFree(d);
        throw;
    }
    cudaFree(d);
    if (sH) *sH = h[0];
    if (sS) *sS = h[1];
}

void block_free(bs2e_block* b)
{
    if (!b) return;
    cudaFree(b->d_blk); cudaFree(b->d_ncrow); cudaFree(b->d_flags); cudaFree(b->d_krange);
    cudaFree(b->d_angD); cudaFree(b->d_angX);
    cudaFree(b->d_row_n1); cudaFree(b->d_row_n2); cudaFree(b->d_row_blk);
    cudaFree(b->d_site_key); cudaFree(b->d_site_ptr); cudaFree(b->d_site_rows);
    cudaFree(b->d_cntH); cudaFree(b->d_cntS); cudaFree(b->d_Hptr); cudaFree(b->d_Sptr);
    cudaFree(b->d_Hidx); cudaFree(b->d_Sidx); cudaFree(b->d_Hdat); cudaFree(b->d_Sdat);
    cudaFree(b->d_scan_tmp);
    delete b;
}

}  // namespace bs2e
