// block.cu -- stage C: assembly of one (L, parity) symmetry block of the
// coupled two-electron Hamiltonian H and overlap S as CSR matrices.
//
// Stands in for count_nnz (src/mat_els/hamiltonian.f90:348-416) and
// construct_block_tensor (:106-283) with the element formulas of
// src/mat_els/mat_els.f90:552-571,608-633,664-715.  The reference scans all
// n_config^2 configuration pairs twice and evaluates 3j/6j symbols for every
// pair; here
//   * the angular factors are tabulated once per pair of (l1,l2) groups on
//     the host (exact arithmetic, wigner.cpp) and kept on the device,
//   * the group structure, the radial sites and the row counts are built on
//     the device (plan_dev.cu),
//   * the band partners of a row are GENERATED from that structure (no pair
//     scan): site_count_kernel counts them in closed form per radial site, an
//     exclusive scan builds index_ptr, and site_fill_kernel (one CTA per
//     radial site, one thread per candidate column) writes indices and
//     values; block_count_kernel / block_fill_kernel (one thread / one warp
//     per row) remain as the general fallback.
// Bound: HBM (24 B written per stored element + R^k read once per site).
#include <cub/device/device_scan.cuh>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "ctx.h"
#include "plan.h"
#include "site_core.h"
#include "dev_async.h"
#include "site_mma.h"

namespace bs2e {

// ---------------------------------------------------------------------------
// row-wise fallback kernels (max_k beyond the site kernel's instantiations, or
// tables that do not fit shared memory); they need the per-row tables of the plan
// ---------------------------------------------------------------------------
__global__ void block_count_kernel(Geom g, Plan pl, long long nrows,
                                   long long* __restrict__ cntH, long long* __restrict__ cntS)
{
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx > nrows) return;
    long long h = 0, s = 0;
    if (idx < nrows) row_count(g, pl, row_of_local(pl.rr, (int)idx), &h, &s);
    cntH[idx] = h;  // slot nrows holds 0 so that the scan yields the total
    cntS[idx] = s;
}

constexpr int kFillWarps = 8;

__global__ void __launch_bounds__(kFillWarps * 32)
block_fill_kernel(Geom g, Plan pl, OneBody ob, const double* __restrict__ R,
                  long long nrows, const long long* __restrict__ Hptr,
                  const long long* __restrict__ Sptr, long long* __restrict__ Hidx,
                  double2* __restrict__ Hdat, long long* __restrict__ Sidx,
                  double2* __restrict__ Sdat)
{
    const long long wrow = (long long)blockIdx.x * kFillWarps + (threadIdx.x >> 5);
    if (wrow >= nrows) return;
    const int lane = threadIdx.x & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    const RowInfo r = row_info(pl, row_of_local(pl.rr, (int)wrow));
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
// All rows (l_a,l_b; n_a,n_b) of the site read the same R^k values; only the
// angular factors and the clipping by the column group differ.
//   phase 0  the rows of the site, derived from the group tables (one row per
//            (l1,l2) group that holds the configuration (n_a,n_b))
//   phase 1  per column group bj and n_c slot: clipped windows T[bj][q]
//   phase 2  per (bj, storage mode): prefix over the n_c slots of the number of
//            stored entries, hp[bj][mode][q] (mode: D, X, D+X, diagonal pair)
//   phase 3  per row of the site: storage mode, multipole parity and offset
//            inside the row of each of its column groups; the (row, column
//            group) pairs are filed as 16-byte records by (parity, mode) and
//            cut into CHUNKS whose packed angular factors fit one staging buffer
//   phase 4  THREAD t OWNS CANDIDATE COLUMN t of the site.  The multipoles of a
//            pair have one parity (wigner_tools.f90:131), so the work is done in
//            two parity passes: the thread loads the R^k values of its column
//            for that parity (both windows) into registers and walks the records
//            of that parity; per (bj, mode) it works out once whether / where
//            the column is stored, then per record sum_k ang_k R^k with the
//            factors read as broadcast loads from the staged chunk -- two
//            records at a time so that their FP64 chains overlap.
//   The factor chunks are moved by the bulk-copy engine (cp.async.bulk +
//   mbarrier, one copy per record) into a double buffer, the next chunk while
//   the current one is consumed; a site whose records fit two chunks runs
//   phase 4 without any block-wide barrier.  R^k is read once per site.
// ---------------------------------------------------------------------------
struct SiteList {
    const unsigned long long* key;  // sorted site keys (plan.h: site_sort_key): n_a << 16 | n_b in the low word
    const int* count;               // [0] number of sites, [1] those with exchange windows
};

struct alignas(16) RowCache {  // per row of the group
    long long hbase, sbase;    // 0-based position of the first H / S entry of the row
    int bi, la, lb, pad;
};

struct alignas(16) RowRec {    // one (row, column group) pair of the group, filed under (bj, parity, mode)
    long long hpos;            // 0-based position of the pair's first H entry
    int cf;                    // index of the pair's packed factors (units of 2*NKP doubles)
    int meta;                  // ri << 8 | kDirAny / kExAny of the pair
};

struct alignas(8) SiteChunk {  // consecutive column groups of one parity whose records share a staging buffer
    short par, bj0, bj1, nrec;
};

struct SiteSmem {   // element counts of the dynamic shared memory carve-up
    int ncmax;      // n_c slots
    int G;          // rows per group (<= 32)
    int chrec;      // records per staging buffer
    int nl;         // l_max + 1 of the one-particle matrices
    size_t bytes;
};

__host__ __device__ inline size_t align16(size_t b) { return (b + 15) & ~(size_t)15; }

__host__ __device__ inline int site_ncmax(const Geom& g, bool wx) { return wx ? site_max_nc(g) : 2 * g.w + 1; }

__host__ __device__ inline size_t site_smem_bytes(const Geom& g, int nblk, int G, int nkp, int chrec, int nl, bool wx)
{
    const size_t ncmax = (size_t)site_ncmax(g, wx);
    const size_t cfs = (size_t)(wx ? 2 : 1) * nkp;
    size_t b = 0;
    b += align16(sizeof(double) * 2 * (size_t)chrec * cfs);                      // two staging buffers
    b += align16(sizeof(double) * (size_t)site_1p_doubles(g, nl));               // band rows of H_l and S
    b += align16(sizeof(SiteEntry) * (size_t)nblk * ncmax);                      // T
    b += align16(sizeof(RowRec) * (size_t)nblk * G);                             // rlist
    b += align16(sizeof(RowCache) * (size_t)G);                                  // rcache
    b += align16(sizeof(BlockDesc) * (size_t)nblk);                              // sblk
    b += align16((size_t)G * nblk);                                              // sfl
    b += align16(sizeof(SiteChunk) * (size_t)(2 * nblk + 2));                    // chunk table
    b += align16(sizeof(uchar4) * 2 * (size_t)nblk);                             // gcnt[2][nblk]
    b += align16(sizeof(unsigned) * (size_t)G * nblk);                           // pm
    b += align16(sizeof(int) * (ncmax + 1));                                     // cprefix
    b += align16(sizeof(int) * 2 * (size_t)nblk);                                // srow_bi, srow_local
    b += align16(sizeof(unsigned short) * 2 * (size_t)nblk);                     // cfoff[2][nblk]
    b += align16(sizeof(unsigned short) * (size_t)nblk * kModes * (ncmax + 1));  // hp
    b += align16(sizeof(unsigned short) * (size_t)nblk * (ncmax + 1));           // sp
    b += 64;                                                                     // mbarriers, counters
    return b;
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

constexpr size_t kSiteSmemLimit = 200 * 1024;

// RB dot products at once, each over the NKH multipoles of one parity in ascending k
// (mat_els.f90:566-570); factors in shared memory, read as 16-byte broadcasts; the RB
// accumulation chains are independent, which hides the FP64 latency
#ifndef BS2E_RB
#define BS2E_RB 2
#endif
constexpr int kSiteRB = BS2E_RB;   // records per step of the inner loop

template <int NKH, int RB>
__device__ __forceinline__ void site_dot_n(const double* const (&cf)[RB], int shift, const double (&R)[NKH], double (&acc)[RB])
{
#pragma unroll
    for (int r = 0; r < RB; ++r) acc[r] = 0.0;
#pragma unroll
    for (int i = 0; i < NKH; i += 2) {
        double2 c[RB];
#pragma unroll
        for (int r = 0; r < RB; ++r) c[r] = *reinterpret_cast<const double2*>(cf[r] + shift + i);
#pragma unroll
        for (int r = 0; r < RB; ++r) acc[r] += c[r].x * R[i];
        if (i + 1 < NKH) {
#pragma unroll
            for (int r = 0; r < RB; ++r) acc[r] += c[r].y * R[i + 1];
        }
    }
}

template <int NKH>
__device__ __forceinline__ double site_dot(const double* __restrict__ cf, const double (&R)[NKH])
{
    const double* const one[1] = {cf};
    double acc[1];
    site_dot_n<NKH, 1>(one, 0, R, acc);
    return acc[0];
}

#ifdef BS2E_PHASE_TIMING
__device__ unsigned long long g_site_phase_cycles[8];   // debug build: cycles per phase, summed over CTAs (thread 0)
#define BS2E_PHASE_MARK(q)                                                      \
    do {                                                                        \
        if (threadIdx.x == 0) {                                                 \
            const long long now_ = clock64();                                   \
            atomicAdd(&g_site_phase_cycles[q], (unsigned long long)(now_ - t_phase_)); \
            t_phase_ = now_;                                                    \
        }                                                                       \
    } while (0)
#else
#define BS2E_PHASE_MARK(q) do { } while (0)
#endif

// launch shapes (macros: kernel-variant builds for A/B measurements)
#ifndef BS2E_D_NT
#define BS2E_D_NT 256   // threads per CTA, sites without exchange windows
#endif
#ifndef BS2E_D_RC
#define BS2E_D_RC 1     // candidate columns per thread
#endif
#ifndef BS2E_X_NT
#define BS2E_X_NT 128   // sites with exchange windows
#endif
#ifndef BS2E_X_RC
#define BS2E_X_RC 1
#endif
#ifndef BS2E_D_REGS
#define BS2E_D_REGS 128  // register budget per thread (sets the CTAs per SM of the launch bounds)
#endif
#ifndef BS2E_X_REGS
#define BS2E_X_REGS 128
#endif
template <int KMAX, bool WX>
struct SiteLaunch {   // threads per CTA, candidates per thread and CTAs per SM of an instantiation
    static constexpr int NT = WX ? BS2E_X_NT : BS2E_D_NT;
    static constexpr int RC = WX ? (KMAX <= 13 ? 2 : BS2E_X_RC) : BS2E_D_RC;
    static constexpr int regs = WX ? (KMAX <= 21 ? BS2E_X_REGS : 168) : BS2E_D_REGS;
    static constexpr int min_blocks = 65536 / (regs * NT);
};

// WX: the sites of this launch have exchange windows (site_wants_X); the sites that
// have none run a leaner instantiation (no exchange registers, half the n_c slots).
template <int KMAX, bool WX>
__global__ void __launch_bounds__(SiteLaunch<KMAX, WX>::NT, SiteLaunch<KMAX, WX>::min_blocks)
site_fill_kernel(Geom g, Plan pl, OneBody ob, SiteList sl, SiteSmem lay, int site_off, const double* __restrict__ R,
                 const long long* __restrict__ Hptr,
                 const long long* __restrict__ Sptr, long long* __restrict__ Hidx,
                 double2* __restrict__ Hdat, long long* __restrict__ Sidx,
                 double2* __restrict__ Sdat)
{
    constexpr int NT = SiteLaunch<KMAX, WX>::NT;
    constexpr int RC = SiteLaunch<KMAX, WX>::RC;
    constexpr int NW = NT / 32;
    constexpr int NKP = (((KMAX + 1) / 2) + 1) & ~1;  // = site_nkp(KMAX): packed factors per window
    constexpr int NKH = (KMAX + 1) / 2;               // multipoles of one parity
    constexpr int CFS = WX ? 2 * NKP : NKP;           // staged doubles per record (direct half only without X)
    extern __shared__ __align__(16) unsigned char smraw[];
    const int K1 = g.K1, ncmax = lay.ncmax, G = lay.G, chrec = lay.chrec;
    const int nblk = pl.nblk;
    unsigned char* sp_ = smraw;
    auto carve = [&](size_t bytes) { unsigned char* p = sp_; sp_ += align16(bytes); return p; };
    double* cfs = reinterpret_cast<double*>(carve(sizeof(double) * 2 * (size_t)chrec * CFS));
    double* ob_s = reinterpret_cast<double*>(carve(sizeof(double) * (size_t)site_1p_doubles(g, lay.nl)));
    SiteEntry* T = reinterpret_cast<SiteEntry*>(carve(sizeof(SiteEntry) * (size_t)nblk * ncmax));
    RowRec* rlist = reinterpret_cast<RowRec*>(carve(sizeof(RowRec) * (size_t)nblk * G));
    RowCache* rcache = reinterpret_cast<RowCache*>(carve(sizeof(RowCache) * (size_t)G));
    BlockDesc* sblk = reinterpret_cast<BlockDesc*>(carve(sizeof(BlockDesc) * (size_t)nblk));
    SiteChunk* chunks = reinterpret_cast<SiteChunk*>(carve(sizeof(SiteChunk) * (size_t)(2 * nblk + 2)));
    uchar4* gcnt = reinterpret_cast<uchar4*>(carve(sizeof(uchar4) * 2 * (size_t)nblk));
    unsigned* pm = reinterpret_cast<unsigned*>(carve(sizeof(unsigned) * (size_t)G * nblk));
    int* cprefix = reinterpret_cast<int*>(carve(sizeof(int) * (ncmax + 1)));
    int* srow_bi = reinterpret_cast<int*>(carve(sizeof(int) * 2 * (size_t)nblk));
    int* srow_local = srow_bi + nblk;
    unsigned short* cfoff = reinterpret_cast<unsigned short*>(carve(sizeof(unsigned short) * 2 * (size_t)nblk));
    unsigned short* hp = reinterpret_cast<unsigned short*>(carve(sizeof(unsigned short) * (size_t)nblk * kModes * (ncmax + 1)));
    unsigned short* sp = reinterpret_cast<unsigned short*>(carve(sizeof(unsigned short) * (size_t)nblk * (ncmax + 1)));
    unsigned char* sfl = reinterpret_cast<unsigned char*>(carve((size_t)G * nblk));   // flags of the group's (row, column group) pairs
    unsigned long long* mbar = reinterpret_cast<unsigned long long*>(carve(16));
    int* misc = reinterpret_cast<int*>(carve(16));   // [0] rows of the site, [1] chunks of the group

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
#ifdef BS2E_PHASE_TIMING
    long long t_phase_ = clock64();
#endif
    const int sidx = blockIdx.x + site_off;
    const unsigned key = (unsigned)(sl.key[sidx] & 0xffffffffull);
    const Site s = make_site(g, (int)(key >> 16), (int)(key & 0xffffu), WX);
    const int nnc = s.nnc;
    constexpr bool wantX = WX;
    const size_t plane = (size_t)g.P * g.ldP;

    // The RC candidate columns this thread owns (candidate pass*NT*RC + rc*NT + tid) and
    // their R^k values of one multipole parity (both windows), read once and streaming.
    // Direct terms of parity `par` pair with exchange terms of parity par ^ pi.
    OwnCand c[RC];
    bool act[RC];
    double Rd[RC][NKH], Rx[RC][WX ? NKH : 1];
    auto load_R = [&](int par, int pi) {
#pragma unroll
        for (int rc = 0; rc < RC; ++rc) {
            const double* pD = R + (size_t)c[rc].rowD * g.ldP + c[rc].colD;
            const double* pX = R + (size_t)c[rc].rowX * g.ldP + c[rc].colX;
            const bool onD = act[rc] && c[rc].inD, onX = act[rc] && c[rc].inX;
            const int parx = par ^ pi;
#pragma unroll
            for (int i = 0; i < NKH; ++i) {
                const int k = 2 * i + par, kx = 2 * i + parx;
                Rd[rc][i] = (onD && k < K1) ? __ldcs(pD + (size_t)k * plane) : 0.0;
                if constexpr (WX) Rx[rc][i] = (onX && kx < K1) ? __ldcs(pX + (size_t)kx * plane) : 0.0;
            }
            if constexpr (!WX) Rx[rc][0] = 0.0;
        }
    };
    // without exchange windows the candidate list is known up front: the loads of the
    // first pass (parity 0) are issued here and complete behind phases 0-3
    const bool prefetched = !WX && s.nD <= NT * RC;
    if (prefetched) {
#pragma unroll
        for (int rc = 0; rc < RC; ++rc) {
            act[rc] = rc * NT + tid < s.nD;
            c[rc] = site_own_cand(g, s, cprefix, false, act[rc] ? rc * NT + tid : 0);
        }
        load_R(0, 0);
    }

    // ---- phase 0: rows of the site; phase 1: clipped windows, candidate prefix, band rows, group table ----
    if (warp == 0) {
        int run = 0;
        for (int b0 = 0; b0 < nblk; b0 += 32) {
            const int bi = b0 + lane;
            int local = -1;
            if (bi < nblk) {
                const int row = config_index(g, pl, bi, s.na, s.nb);
                if (row > 0) local = row_local_of(pl.rr, row);
            }
            const unsigned m = __ballot_sync(0xffffffffu, local >= 0);
            if (local >= 0) {
                const int pos = run + __popc(m & ((1u << lane) - 1u));
                srow_bi[pos] = bi;
                srow_local[pos] = local;
            }
            run += __popc(m);
        }
        if (lane == 0) misc[0] = run;
    }
    if (tid == 0) {
        mbar_init(&mbar[0], 1);
        mbar_init(&mbar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == NW - 1 && wantX) {
        int run = 0;
        for (int q0 = 0; q0 < nnc; q0 += 32) {
            const int q = q0 + lane;
            const int cq = q < nnc ? site_cand_DX_count(s, q) : 0;
            const int inc = warp_incl_scan(cq, lane);
            if (q < nnc) cprefix[q] = run + inc - cq;
            run += __shfl_sync(0xffffffffu, inc, 31);
        }
        if (lane == 0) cprefix[nnc] = run;
    }
    for (int idx = tid; idx < nblk * nnc; idx += NT) {
        const int bj = idx / nnc, q = idx - bj * nnc;
        T[bj * ncmax + q] = site_entry(g, pl, s, bj, q);
    }
    for (int bj = tid; bj < nblk; bj += NT) sblk[bj] = pl.blk[bj];
    for (int idx = tid; idx < site_1p_doubles(g, lay.nl) / 2; idx += NT) {
        const Cplx v = site_1p_source(g, ob, s, lay.nl, idx);
        reinterpret_cast<double2*>(ob_s)[idx] = make_double2(v.re, v.im);
    }
    const SiteOneBody so{ob_s, ob_s + (size_t)lay.nl * 2 * (2 * g.w + 1) * 2};
    __syncthreads();
    BS2E_PHASE_MARK(0);
    const int nr = misc[0];
    const int pi = (sblk[0].l1 + sblk[0].l2) & 1;   // parity of l1+l2, the same for every group of a symmetry
    // ---- phase 2: prefix of stored entries over the n_c slots, per (bj, mode);
    //      one thread per (bj, mode), serial over the slots ----
    for (int task = tid; task < nblk * kModes; task += NT) {
        const int bj = task / kModes, mode = task - bj * kModes;
        if (!wantX && (mode == kModeX || mode == kModeDX)) continue;  // never read
        const bool useD = mode_useD(mode), useX = mode_useX(mode);
        const bool diag = mode == kModeDiag;
        const bool samex = diag && sblk[bj].l1 == sblk[bj].l2;
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

    double* const Hd = reinterpret_cast<double*>(Hdat);
    double* const Sd = reinterpret_cast<double*>(Sdat);
    unsigned mphase = 0;          // phase parity of the two staging barriers (bit b)
    int resident0 = -1, resident1 = -1;   // chunk held by each staging buffer (of the current group)
    bool pending0 = false, pending1 = false;

    for (int g0 = 0; g0 < nr; g0 += G) {
        const int gr = imin(G, nr - g0);
        // the group's rows: first entries of the rows, flags of their (row, column group) pairs -- the
        // only global loads of phase 3, all independent
        for (int ri = tid; ri < gr; ri += NT) {
            const int bi = srow_bi[g0 + ri];
            const long long wrow = srow_local[g0 + ri];
            rcache[ri] = RowCache{Hptr[wrow] - 1, Sptr[wrow] - 1, bi, sblk[bi].l1, sblk[bi].l2, 0};
        }
        for (int idx = tid; idx < gr * nblk; idx += NT) {
            const int ri = idx / nblk, bj = idx - ri * nblk;
            sfl[idx] = pl.flags[(size_t)srow_bi[g0 + ri] * nblk + bj];
        }
        __syncthreads();  // phase 2 / the group's tables
        BS2E_PHASE_MARK(1);
        // ---- phase 3a: coupled column groups of each row and their offsets inside the row ----
        for (int ri = warp; ri < gr; ri += NW) {
            const RowCache rc = rcache[ri];
            int run = 0;
            for (int b0 = 0; b0 < nblk; b0 += 32) {
                const int bj = b0 + lane;
                int cnt = 0, mode = -1;
                if (bj < nblk && (pl.full || bj >= rc.bi)) {
                    const unsigned f = sfl[ri * nblk + bj];
                    mode = bj == rc.bi ? kModeDiag
                           : (f & kDirAny) ? ((f & kExAny) ? kModeDX : kModeD) : ((f & kExAny) ? kModeX : -1);   // pair_mode
                    if (mode >= 0) {
                        const unsigned short* hb = hp + (bj * kModes) * (ncmax + 1) + nnc;
                        mode = effective_mode(mode, hb[kModeD * (ncmax + 1)], wantX ? hb[kModeX * (ncmax + 1)] : 0);
                        cnt = (wantX || mode != kModeX) ? hb[mode * (ncmax + 1)] : 0;
                    }
                }
                const int inc = warp_incl_scan(cnt, lane);
                if (bj < nblk)
                    pm[ri * nblk + bj] = cnt > 0 ? pm_pack(run + inc - cnt, mode, (rc.la + sblk[bj].l1) & 1) : 0u;
                run += __shfl_sync(0xffffffffu, inc, 31);
            }
        }
        __syncthreads();
        BS2E_PHASE_MARK(2);
        // ---- phase 3b: the pairs of each column group, filed by (parity, storage mode);
        //      one warp per column group, lane = row of the group ----
        for (int bj = warp; bj < nblk; bj += NW) {
            const unsigned v = lane < gr ? pm[lane * nblk + bj] : 0u;
            const bool valid = pm_valid(v);
            const int slot = valid ? pm_pd(v) * kModes + pm_mode(v) : -1;
            int cnt[2 * kModes];
            int where = 0, base = 0;
#pragma unroll
            for (int q = 0; q < 2 * kModes; ++q) {
                const unsigned m = __ballot_sync(0xffffffffu, slot == q);
                if (slot == q) where = base + __popc(m & ((1u << lane) - 1u));
                cnt[q] = __popc(m);
                base += cnt[q];
            }
            if (lane == 0) {
                gcnt[bj] = make_uchar4((unsigned char)cnt[0], (unsigned char)cnt[1], (unsigned char)cnt[2], (unsigned char)cnt[3]);
                gcnt[nblk + bj] = make_uchar4((unsigned char)cnt[4], (unsigned char)cnt[5], (unsigned char)cnt[6], (unsigned char)cnt[7]);
            }
            if (valid) {
                const RowCache rc = rcache[lane];
                const int fl = sfl[lane * nblk + bj] & (kDirAny | kExAny);
                rlist[bj * G + where] = RowRec{rc.hbase + pm_off(v), rc.bi * nblk + bj, fl | (lane << 8)};
            }
        }
        __syncthreads();
        BS2E_PHASE_MARK(3);
        // ---- phase 3c: chunks of consecutive column groups whose factors fit a staging buffer ----
        if (warp == 0) {
            int nch = 0;
            for (int par = 0; par < 2; ++par) {
                int run = 0;   // records of this parity; cfoff = offset of each column group inside ONE chunk
                for (int b0 = 0; b0 < nblk; b0 += 32) {
                    const int bj = b0 + lane;
                    int n = 0;
                    if (bj < nblk) { const uchar4 gc = gcnt[par * nblk + bj]; n = gc.x + gc.y + gc.z + gc.w; }
                    const int inc = warp_incl_scan(n, lane);
                    if (bj < nblk) cfoff[par * nblk + bj] = (unsigned short)(run + inc - n);
                    run += __shfl_sync(0xffffffffu, inc, 31);
                }
                if (run == 0) continue;
                if (run <= chrec) {
                    if (lane == 0) chunks[nch] = SiteChunk{(short)par, 0, (short)nblk, (short)run};
                    ++nch;
                } else {   // greedy split, serial
                    int add = 0;
                    if (lane == 0) {
                        int cur = 0, bj0 = 0;
                        for (int bj = 0; bj < nblk; ++bj) {
                            const uchar4 gc = gcnt[par * nblk + bj];
                            const int n = gc.x + gc.y + gc.z + gc.w;
                            if (cur + n > chrec) {
                                chunks[nch + add++] = SiteChunk{(short)par, (short)bj0, (short)bj, (short)cur};
                                cur = 0;
                                bj0 = bj;
                            }
                            cfoff[par * nblk + bj] = (unsigned short)cur;
                            cur += n;
                        }
                        if (cur > 0) chunks[nch + add++] = SiteChunk{(short)par, (short)bj0, (short)nblk, (short)cur};
                    }
                    nch += __shfl_sync(0xffffffffu, add, 0);
                }
            }
            if (lane == 0) misc[1] = nch;
        }
        __syncthreads();
        BS2E_PHASE_MARK(4);
        const int nch = misc[1];
        resident0 = resident1 = -1;   // the buffers hold chunks of the previous group
        // chunk ch -> staging buffer bsel: one bulk copy per record, issued by all threads
        auto stage = [&](int ch, int bsel) {
            const SiteChunk cc = chunks[ch];
            unsigned long long* bar = &mbar[bsel];
            if (tid == 0) mbar_expect_tx(bar, (unsigned)cc.nrec * CFS * 8u);
            double* dst = cfs + (size_t)bsel * chrec * CFS;
            for (int bj = cc.bj0 + warp; bj < cc.bj1; bj += NW) {
                const uchar4 g0c = gcnt[bj], gpc = gcnt[cc.par * nblk + bj];
                const int first = cc.par ? g0c.x + g0c.y + g0c.z + g0c.w : 0;
                const int n = gpc.x + gpc.y + gpc.z + gpc.w;
                const int off = cfoff[cc.par * nblk + bj];
                for (int i = lane; i < n; i += 32) {
                    const RowRec rec = rlist[bj * G + first + i];
                    bulk_g2s(dst + (size_t)(off + i) * CFS, pl.angP + (size_t)rec.cf * (2 * NKP), CFS * 8u, bar);
                }
            }
            if (bsel) { resident1 = ch; pending1 = true; } else { resident0 = ch; pending0 = true; }
        };
        auto wait_buf = [&](int bsel) {
            if (!(bsel ? pending1 : pending0)) return;
            mbar_wait(&mbar[bsel], (mphase >> bsel) & 1u);
            mphase ^= 1u << bsel;
            if (bsel) pending1 = false; else pending0 = false;
        };
        int cur = 0;   // staging buffer of the chunk being consumed; the sequence of chunks alternates buffers
        if (nch > 0) stage(0, 0);
        // ---- phase 4: fill ----
        const int nc_all = site_num_cand(s, cprefix, wantX);
        const int npass = nch > 0 ? (nc_all + NT * RC - 1) / (NT * RC) : 0;
        for (int pass = 0; pass < npass; ++pass) {
            const bool keepR = prefetched && pass == 0 && g0 == 0;   // parity-0 values already in registers
            if (!keepR) {
#pragma unroll
                for (int rc = 0; rc < RC; ++rc) {
                    const int t = (pass * RC + rc) * NT + tid;
                    act[rc] = t < nc_all;
                    c[rc] = site_own_cand(g, s, cprefix, wantX, act[rc] ? t : 0);
                }
            }
            const bool warp_has_cand = __ballot_sync(0xffffffffu, act[0]) != 0u;   // candidates of slot 0 come first
            int rpar = keepR ? 0 : -1;   // parity of the values in Rd / Rx
            for (int ch = 0; ch < nch; ++ch) {
                // next chunk of the cyclic sequence into the other buffer (all threads take the same branch)
                const int nxt = ch + 1 < nch ? ch + 1 : (pass + 1 < npass ? 0 : -1);
                const bool flip = nxt >= 0 && nxt != ch;
                if (flip && ((cur ^ 1) ? resident1 : resident0) != nxt) {
                    __syncthreads();   // every thread is done with the chunk that buffer held
                    stage(nxt, cur ^ 1);
                }
                wait_buf(cur);
                const double* cbuf = cfs + (size_t)cur * chrec * CFS;
                if (flip) cur ^= 1;
                if (!warp_has_cand) continue;
                const SiteChunk cc = chunks[ch];
                const int par = cc.par;
                if (rpar != par) { load_R(par, pi); rpar = par; }
                for (int bj = cc.bj0; bj < cc.bj1; ++bj) {
                    const uchar4 gc = gcnt[par * nblk + bj];
                    if ((gc.x | gc.y | gc.z | gc.w) == 0) continue;
                    const uchar4 g0c = gcnt[bj];
                    const RowRec* rl = rlist + bj * G + (par ? g0c.x + g0c.y + g0c.z + g0c.w : 0);
                    const double* cfm = cbuf + (size_t)cfoff[par * nblk + bj] * CFS;
                    SiteEntry e[RC];
#pragma unroll
                    for (int rc = 0; rc < RC; ++rc) e[rc] = T[bj * ncmax + c[rc].q];
#pragma unroll
                    for (int mode = 0; mode < kModes; ++mode) {
                        if (!WX && (mode == kModeX || mode == kModeDX)) continue;  // no such pairs without exchange windows
                        const int nrow = mode == 0 ? gc.x : mode == 1 ? gc.y : mode == 2 ? gc.z : gc.w;
                        if (nrow == 0) continue;
                        const RowRec* rm = rl;
                        const double* cf = cfm;
                        rl += nrow;
                        cfm += (size_t)nrow * CFS;
                        const bool diag = mode == kModeDiag;
                        ModeSlot ms[RC];
                        bool stored[RC];
                        long long* pI[RC];   // &Hidx[rank], &Hd[2 rank]: the record adds its row offset
                        double* pV[RC];
                        bool any = false;
#pragma unroll
                        for (int rc = 0; rc < RC; ++rc) {
                            ms[rc] = site_mode_slot(s, c[rc], e[rc], hp + (bj * kModes + mode) * (ncmax + 1), mode, diag && !pl.full);
                            stored[rc] = act[rc] && (ms[rc].sup || ms[rc].sup_ex);
                            pI[rc] = Hidx + ms[rc].rank;
                            pV[rc] = Hd + 2 * (long long)ms[rc].rank;
                            any = any || stored[rc];
                        }
                        if (__ballot_sync(0xffffffffu, any) == 0u) continue;
                        if (!diag) {
                            // RB records per step for each of the RC candidates: RB*RC independent FP64 chains
                            // (a short tail repeats its last record, stored once)
                            constexpr int RB = kSiteRB;
                            for (int i = 0; i < nrow; i += RB) {
                                const double* cfr[RB];
                                long long hpos[RB];
#pragma unroll
                                for (int r = 0; r < RB; ++r) {
                                    const int ii = imin(i + r, nrow - 1);
                                    hpos[r] = rm[ii].hpos;
                                    cfr[r] = cf + (size_t)ii * CFS;
                                }
#pragma unroll
                                for (int rc = 0; rc < RC; ++rc) {
                                    double res[RB];
                                    if (mode == kModeD) {
                                        site_dot_n<NKH, RB>(cfr, 0, Rd[rc], res);   // stored == sup
                                    } else {
                                        double d[RB], x[RB];
#pragma unroll
                                        for (int r = 0; r < RB; ++r) d[r] = x[r] = 0.0;
                                        if (mode != kModeX) site_dot_n<NKH, RB>(cfr, 0, Rd[rc], d);
                                        if constexpr (WX) site_dot_n<NKH, RB>(cfr, NKP, Rx[rc], x);
#pragma unroll
                                        for (int r = 0; r < RB; ++r) res[r] = (ms[rc].sup ? d[r] : 0.0) + (ms[rc].sup_ex ? x[r] : 0.0);
                                    }
#ifdef BS2E_EXP_NOSTORE
                                    if (stored[rc] && hpos[0] < 0) {   // experiment: never true
#else
                                    if (stored[rc]) {
#endif
#pragma unroll
                                        for (int r = 0; r < RB; ++r) {
                                            if (r > 0 && i + r >= nrow) break;
                                            pI[rc][hpos[r]] = ms[rc].jcol;
                                            *reinterpret_cast<double2*>(pV[rc] + 2 * hpos[r]) = make_double2(res[r], 0.0);
                                        }
                                    }
                                }
                            }
                        } else {  // exactly one row: the row whose own group is bj (parity 0)
                            const RowRec rec = rm[0];
                            const RowCache rc_ = rcache[rec.meta >> 8];
#pragma unroll
                            for (int rc = 0; rc < RC; ++rc) {
                                const double d = site_dot<NKH>(cf, Rd[rc]);
                                double res = ms[rc].sup ? d : 0.0;
                                if constexpr (WX) {
                                    const double x = site_dot<NKH>(cf + NKP, Rx[rc]);
                                    res += ms[rc].sup_ex ? x : 0.0;
                                }
                                // hamiltonian.f90:183: the r_12 sums enter only when a supported window has a factor above 5e-15
                                const bool allowed = (ms[rc].sup && (rec.meta & kDirAny)) || (ms[rc].sup_ex && (rec.meta & kExAny));
                                if (!allowed) res = 0.0;
                                if (stored[rc]) {
                                    double re = res, im = 0.0;
                                    site_diag_terms(g, pl, so, s, rc_.la, rc_.lb, c[rc], ms[rc], rc_.la == rc_.lb,
                                                    sp + bj * (ncmax + 1), rc_.sbase, &re, &im, Sidx, Sd);
                                    pI[rc][rec.hpos] = ms[rc].jcol;
                                    *reinterpret_cast<double2*>(pV[rc] + 2 * rec.hpos) = make_double2(re, im);
                                }
                            }
                        }
                    }
                }
            }
        }
        // a staging copy that was issued must land before the buffers are reused or the CTA exits
        wait_buf(0);
        wait_buf(1);
        BS2E_PHASE_MARK(5);
        if (g0 + G < nr) __syncthreads();   // the next group rewrites the tables
    }
}

// Count pass on the site tables.  The row count of a row is the sum over its coupled
// column groups of the stored entries of the pair's storage mode, which depend on the
// site and the column group only.  One WARP per site, no block-wide barrier: lane = column
// group, one pass over the n_c slots gives the totals of all four storage modes; then
// lane = column group again for each row of the site.  (block_count_kernel, one thread
// per row, is kept for plans without a site list.)
constexpr int kCountWarps = 4;

__host__ __device__ inline size_t count_smem_bytes(int nblk)
{
    return sizeof(unsigned short) * (size_t)kCountWarps * (kModes + 1) * nblk + 16;
}

__global__ void __launch_bounds__(kCountWarps * 32)
site_count_kernel(Geom g, Plan pl, SiteList sl, long long* __restrict__ cntH, long long* __restrict__ cntS)
{
    extern __shared__ __align__(16) unsigned char smraw[];
    const int nblk = pl.nblk;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned short* tot = reinterpret_cast<unsigned short*>(smraw) + (size_t)warp * (kModes + 1) * nblk;
    unsigned short* stot = tot + (size_t)kModes * nblk;   // S entries of the diagonal pair of group bj
    const int sidx = blockIdx.x * kCountWarps + warp;
    if (blockIdx.x == 0 && threadIdx.x == 0) { cntH[pl.nrows] = 0; cntS[pl.nrows] = 0; }  // the scan yields the totals there
    if (sidx >= sl.count[0]) return;
    const unsigned key = (unsigned)(sl.key[sidx] & 0xffffffffull);
    const bool wantX = site_wants_X(g, pl.max_nd, (int)(key >> 16));
    const Site s = make_site(g, (int)(key >> 16), (int)(key & 0xffffu), wantX);
    const int nnc = s.nnc;
    for (int bj = lane; bj < nblk; bj += 32) {
        const bool samex = pl.blk[bj].l1 == pl.blk[bj].l2;
        int cD = 0, cX = 0, cDX = 0, cG = 0, cS = 0;
        for (int q = 0; q < nnc; ++q) {
            const SiteEntry e = site_entry(g, pl, s, bj, q);
            cD += entry_count(e, true, false);
            cX += entry_count(e, false, true);
            cDX += entry_count(e, true, true);
            const SiteEntry ec = pl.full ? e : entry_cut(e, s, site_nc(s, q));
            cG += entry_count(ec, true, true);
            cS += entry_count(ec, true, samex);
        }
        tot[kModeD * nblk + bj] = (unsigned short)cD;
        tot[kModeX * nblk + bj] = (unsigned short)cX;
        tot[kModeDX * nblk + bj] = (unsigned short)cDX;
        tot[kModeDiag * nblk + bj] = (unsigned short)cG;
        stot[bj] = (unsigned short)cS;
    }
    __syncwarp();
    for (int bi = 0; bi < nblk; ++bi) {   // the rows of the site: one per group that holds (n_a, n_b)
        const int rowi = config_index(g, pl, bi, s.na, s.nb);
        if (rowi == 0) continue;
        const int wrow = row_local_of(pl.rr, rowi);
        if (wrow < 0) continue;
        const BlockDesc bd = pl.blk[bi];
        const RowInfo r{rowi, bi, s.na, s.nb, bd.l1, bd.l2};
        int run = 0, srun = 0;
        for (int bj = lane; bj < nblk; bj += 32) {
            int mode = pair_mode(pl, r, bj);
            if (mode >= 0) {
                mode = effective_mode(mode, tot[kModeD * nblk + bj], wantX ? tot[kModeX * nblk + bj] : 0);
                run += (wantX || mode != kModeX) ? tot[mode * nblk + bj] : 0;
                if (mode == kModeDiag) srun += stot[bj];
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            run += __shfl_xor_sync(0xffffffffu, run, o);
            srun += __shfl_xor_sync(0xffffffffu, srun, o);
        }
        if (lane == 0) {
            cntH[wrow] = run;
            cntS[wrow] = srun;
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

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
namespace {
// staging-buffer size of the site kernel: large enough for the records of a typical
// whole site, bounded so that the launch keeps its CTAs per SM (BS2E_SITE_CHUNK_KB overrides)
int site_chunk_records(const Geom& g, int nblk, int G, int nkp, bool wx)
{
    size_t cap_kb = 8;
    if (const char* e = getenv("BS2E_SITE_CHUNK_KB")) cap_kb = (size_t)std::max(1, atoi(e));
    const size_t per = sizeof(double) * (wx ? 2 : 1) * nkp;
    size_t rec = std::min((size_t)G * nblk, cap_kb * 1024 / per);
    rec = std::max(rec, (size_t)G);   // a column group's records (<= G) must fit one buffer
    return (int)rec;
}

SiteSmem site_layout(const bs2e_ctx* c, int nblk, bool wx)
{
    const Geom& g = c->hg;
    SiteSmem lay{};
    const int kmax = site_kmax_for(g.K1);
    const int nkp = site_nkp(kmax);
    lay.ncmax = site_ncmax(g, wx);
    lay.G = std::min(32, nblk);   // a site has at most one row per (l1,l2) group
    lay.nl = c->lmax_1p + 1;
    lay.chrec = site_chunk_records(g, nblk, lay.G, nkp, wx);
    lay.bytes = site_smem_bytes(g, nblk, lay.G, nkp, lay.chrec, lay.nl > 0 ? lay.nl : 1, wx);
    return lay;
}
}  // namespace

// site kernels unless max_k exceeds their largest instantiation or their tables do
// not fit shared memory; BS2E_FILL=row asks for the row kernels (A/B measurements)
bool site_kernel_usable(const bs2e_ctx* c, int nblk, int lmax)
{
    const Geom& g = c->hg;
    const char* mode = getenv("BS2E_FILL");
    if (mode && strcmp(mode, "row") == 0) return false;
    const int kmax = site_kmax_for(g.K1);
    if (kmax <= 0) return false;
    if (site_max_slots(g) > 65535 || (size_t)nblk * site_max_slots(g) >= (1u << 24)) return false;
    if (nblk > 255) return false;   // per-group record counts are bytes
    if (count_smem_bytes(nblk) > 48 * 1024) return false;
    const int G = std::min(32, nblk), nkp = site_nkp(kmax);
    const int nl = std::max(c->lmax_1p, lmax) + 1;   // band rows of H_l for every l of the one-particle input
    return site_smem_bytes(g, nblk, G, nkp, site_chunk_records(g, nblk, G, nkp, true), nl, true) <= kSiteSmemLimit;
}

BlockStreams default_streams(bs2e_ctx* c) { return BlockStreams{c->stream, c->side, c->ev_fork, c->ev_join}; }

// count pass + exclusive scan -> 1-based index_ptr of the planned rows
void block_count_scan(bs2e_block* b, bool read_totals, const BlockStreams* bsp)
{
    bs2e_ctx* c = b->ctx;
    cudaStream_t st = bsp ? bsp->main : c->stream;
    const long long nrows = b->nrows;
    if (b->n_config == 0) return;
    if (b->use_site) {
        const SiteList sl{b->d_site_key, b->d_counters};
        const size_t cbytes = count_smem_bytes(b->dplan.nblk);
        site_count_kernel<<<(unsigned)((b->site_cap + kCountWarps - 1) / kCountWarps), kCountWarps * 32, cbytes, st>>>(
            c->dg, b->dplan, sl, b->d_cntH, b->d_cntS);
    } else {
        block_count_kernel<<<(unsigned)((nrows + 1 + 127) / 128), 128, 0, st>>>(
            c->dg, b->dplan, nrows, b->d_cntH, b->d_cntS);
    }
    BS2E_LAUNCHED();
    size_t tmp = b->scan_tmp_bytes;
    // exclusive scan seeded with 1: the 1-based index_ptr of the reference (sparse_array_tools.f90:63-89)
    BS2E_CUDA(cub::DeviceScan::ExclusiveScan(b->d_scan_tmp, tmp, b->d_cntH, b->d_Hptr, cub::Sum(), 1LL, nrows + 1, st));
    g_launches.fetch_add(2);  // DeviceScanInitKernel + DeviceScanKernel
    BS2E_CUDA(cub::DeviceScan::ExclusiveScan(b->d_scan_tmp, tmp, b->d_cntS, b->d_Sptr, cub::Sum(), 1LL, nrows + 1, st));
    g_launches.fetch_add(2);
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

template <int KMAX, bool WX>
static void launch_site_fill(bs2e_block* b, const SiteSmem& lay, cudaStream_t st, int first, int count)
{
    if (count <= 0) return;
    bs2e_ctx* c = b->ctx;
    auto kern = site_fill_kernel<KMAX, WX>;
    const SiteList sl{b->d_site_key, b->d_counters};
    BS2E_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lay.bytes));
    kern<<<(unsigned)count, SiteLaunch<KMAX, WX>::NT, lay.bytes, st>>>(
        c->dg, b->dplan, c->one_body(), sl, lay, first, c->d_R, b->d_Hptr, b->d_Sptr, b->d_Hidx,
        reinterpret_cast<double2*>(b->d_Hdat), b->d_Sidx, reinterpret_cast<double2*>(b->d_Sdat));
    BS2E_LAUNCHED();
}

void block_assemble(bs2e_block* b, const BlockStreams* bsp)
{
    bs2e_ctx* c = b->ctx;
    const BlockStreams bs = bsp ? *bsp : default_streams(c);
    if (b->n_config == 0 || b->nrows == 0) { b->assembled = true; return; }
    if (!c->have_R) throw Error("block_assemble: call bs2e_rk_build first");
    if (b->a_need_lo < c->R_lo || b->a_need_hi > c->R_hi)
        throw Error("block_assemble: the planned rows read R^k rows outside the slice that was built (bs2e_rk_rows)");
    if (!c->have_1p) throw Error("block_assemble: call bs2e_set_one_particle first");
    if (b->lmax > c->lmax_1p) throw Error("block_assemble: configuration l exceeds max_l_1p of H_vec");
    if (!b->d_Hidx) {   // largest first: each takes the smallest cached array that is large enough
        b->d_Hdat = static_cast<double*>(out_take(c, sizeof(double) * 2 * (size_t)b->nnzH, bs.main, &b->cap_Hdat));
        b->d_Hidx = static_cast<long long*>(out_take(c, sizeof(long long) * (size_t)b->nnzH, bs.main, &b->cap_Hidx));
        b->d_Sdat = static_cast<double*>(out_take(c, sizeof(double) * 2 * (size_t)b->nnzS, bs.main, &b->cap_Sdat));
        b->d_Sidx = static_cast<long long*>(out_take(c, sizeof(long long) * (size_t)b->nnzS, bs.main, &b->cap_Sidx));
    }
    const long long nrows = b->nrows;
    const Geom& g = c->dg;
    const int nblk = b->dplan.nblk;
    if (b->use_site) {
        const int kmax = site_kmax_for(g.K1);
        const SiteSmem layX = site_layout(c, nblk, true), layD = site_layout(c, nblk, false);
        if (layX.bytes > kSiteSmemLimit || layD.bytes > kSiteSmemLimit)
            throw Error("block_assemble: site tables exceed shared memory (max_l_1p above the planned bound)");
        // the launch without exchange windows goes to the side stream so that it fills
        // the SMs the other launch leaves idle in its last wave
        const bool fork = b->nsites_x > 0 && b->nsites_x < b->nsites && !getenv("BS2E_NOFORK");   // BS2E_NOFORK: A/B measurements
        if (fork) {
            BS2E_CUDA(cudaEventRecord(bs.fork, bs.main));
            BS2E_CUDA(cudaStreamWaitEvent(bs.side, bs.fork, 0));
        }
        cudaStream_t stD = fork ? bs.side : bs.main;
        const int nx = b->nsites_x, nd = b->nsites - b->nsites_x;
        if (b->use_mma) launch_site_mma(b, bs.main, stD);
        else switch (kmax) {
        case 7: launch_site_fill<7, true>(b, layX, bs.main, 0, nx); launch_site_fill<7, false>(b, layD, stD, nx, nd); break;
        case 13: launch_site_fill<13, true>(b, layX, bs.main, 0, nx); launch_site_fill<13, false>(b, layD, stD, nx, nd); break;
        case 21: launch_site_fill<21, true>(b, layX, bs.main, 0, nx); launch_site_fill<21, false>(b, layD, stD, nx, nd); break;
        case 31: launch_site_fill<31, true>(b, layX, bs.main, 0, nx); launch_site_fill<31, false>(b, layD, stD, nx, nd); break;
        default: throw Error("block_assemble: no site kernel for this max_k");
        }
        if (fork) {
            BS2E_CUDA(cudaEventRecord(bs.join, bs.side));
            BS2E_CUDA(cudaStreamWaitEvent(bs.main, bs.join, 0));
        }
    } else {
        block_fill_kernel<<<(unsigned)((nrows + kFillWarps - 1) / kFillWarps), kFillWarps * 32, 0,
                            bs.main>>>(c->dg, b->dplan, c->one_body(), c->d_R, nrows,
                                         b->d_Hptr, b->d_Sptr, b->d_Hidx,
                                         reinterpret_cast<double2*>(b->d_Hdat), b->d_Sidx,
                                         reinterpret_cast<double2*>(b->d_Sdat));
        BS2E_LAUNCHED();
    }
    b->assembled = true;
}

// count pass (optional) + fill of several blocks of one context, consecutive blocks on
// different stream pairs; everything is ordered after the work already queued on the
// context's stream and the context's stream waits for all of it.
static void ensure_lanes(bs2e_ctx* c)
{
    if (c->have_lanes) return;
    for (auto& ln : c->lanes) {
        BS2E_CUDA(cudaStreamCreateWithFlags(&ln.main, cudaStreamNonBlocking));
        BS2E_CUDA(cudaStreamCreateWithFlags(&ln.side, cudaStreamNonBlocking));
        BS2E_CUDA(cudaEventCreateWithFlags(&ln.fork, cudaEventDisableTiming));
        BS2E_CUDA(cudaEventCreateWithFlags(&ln.join, cudaEventDisableTiming));
        BS2E_CUDA(cudaEventCreateWithFlags(&ln.done, cudaEventDisableTiming));
    }
    BS2E_CUDA(cudaEventCreateWithFlags(&c->ev_start, cudaEventDisableTiming));
    c->have_lanes = true;
}

void blocks_run(bs2e_ctx* c, long long n, bs2e_block** blks, bool recount)
{
    if (n <= 0) return;
    ensure_lanes(c);
    for (long long i = 0; i < n; ++i)
        if (!blks[i] || blks[i]->ctx != c) throw Error("bs2e_blocks_run: block of another context");
    BS2E_CUDA(cudaEventRecord(c->ev_start, c->stream));
    const int nl = (int)std::min<long long>(n, bs2e_ctx::kLanes);
    for (int l = 0; l < nl; ++l) BS2E_CUDA(cudaStreamWaitEvent(c->lanes[l].main, c->ev_start, 0));
    for (long long i = 0; i < n; ++i) {
        bs2e_ctx::Lane& ln = c->lanes[i % nl];
        const BlockStreams bs{ln.main, ln.side, ln.fork, ln.join};
        if (recount) block_count_scan(blks[i], false, &bs);
        block_assemble(blks[i], &bs);
    }
    for (int l = 0; l < nl; ++l) {
        BS2E_CUDA(cudaEventRecord(c->lanes[l].done, c->lanes[l].main));
        BS2E_CUDA(cudaStreamWaitEvent(c->stream, c->lanes[l].done, 0));
    }
}

void block_download(bs2e_block* b, int64_t* H_ptr, int64_t* H_idx, double* H_dat, int64_t* S_ptr,
                    int64_t* S_idx, double* S_dat)
{
    if (!b->assembled) throw Error("block_download: call bs2e_block_assemble first");
    bs2e_ctx* c = b->ctx;
    const long long nrows = b->nrows;
    if (b->n_config == 0 || nrows == 0) {   // empty block: index_ptr = [1]
        if (H_ptr) H_ptr[0] = 1;
        if (S_ptr) S_ptr[0] = 1;
        return;
    }
    auto d2h = [&](void* dst, const void* src, size_t bytes) {
        if (dst && bytes) download_to_host(c, dst, src, bytes);
    };
    d2h(H_ptr, b->d_Hptr, sizeof(long long) * (nrows + 1));
    d2h(S_ptr, b->d_Sptr, sizeof(long long) * (nrows + 1));
    d2h(H_idx, b->d_Hidx, sizeof(long long) * b->nnzH);
    d2h(S_idx, b->d_Sidx, sizeof(long long) * b->nnzS);
    d2h(H_dat, b->d_Hdat, sizeof(double) * 2 * b->nnzH);
    d2h(S_dat, b->d_Sdat, sizeof(double) * 2 * b->nnzS);
    download_flush(c);
}

void block_row_counts(bs2e_block* b, int64_t* cH, int64_t* cS)
{
    cudaStream_t st = b->ctx->stream;
    const long long nrows = b->nrows;
    if (nrows == 0) return;
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

// debug builds (-DBS2E_PHASE_TIMING): cycles spent per phase of site_fill_kernel, summed over CTAs
void site_phase_cycles(unsigned long long* out8, bool reset)
{
#ifdef BS2E_PHASE_TIMING
    BS2E_CUDA(cudaDeviceSynchronize());
    BS2E_CUDA(cudaMemcpyFromSymbol(out8, g_site_phase_cycles, sizeof(unsigned long long) * 8));
    if (reset) {
        unsigned long long z[8] = {};
        BS2E_CUDA(cudaMemcpyToSymbol(g_site_phase_cycles, z, sizeof(z)));
    }
#else
    (void)reset;
    for (int q = 0; q < 8; ++q) out8[q] = 0;
#endif
}

// ---------------------------------------------------------------------------
// CSR output arrays.  They are multi-GB and live for one block; cudaMalloc / cudaFree cost 5-500 ms each, and the
// driver's stream-ordered pool, which took their place first, was measured to stall a cudaMallocAsync for up to
// 1.06 s now and then (gpurun_out/r03g: one step of 21 ms became 1081 ms; the pool regroups its free blocks when
// requests of many sizes alternate).  So arrays given back by bs2e_block_free wait in a per-context cache and the
// next block takes the smallest one that is large enough; only a request no cached array can serve goes to
// cudaMallocAsync.  The cache is bounded (45 % of the device memory, 24 arrays): the smallest arrays leave first,
// so that after one pass over the blocks of a configuration it holds arrays that serve every block.
// ---------------------------------------------------------------------------
void* out_take(bs2e_ctx* c, size_t bytes, cudaStream_t st, size_t* cap)
{
    if (bytes == 0) bytes = 16;
    bs2e_ctx::OutBuf pick{nullptr, 0, nullptr};
    {
        std::lock_guard<std::mutex> lk(c->out_mu);
        int at = -1;
        for (int i = 0; i < (int)c->out_free.size(); ++i)
            if (c->out_free[i].bytes >= bytes && (at < 0 || c->out_free[i].bytes < c->out_free[at].bytes)) at = i;
        if (at >= 0) {
            pick = c->out_free[at];
            c->out_free.erase(c->out_free.begin() + at);
            c->out_cached -= pick.bytes;
        }
    }
    if (pick.p) {
        BS2E_CUDA(cudaStreamWaitEvent(st, pick.ev, 0));   // the work that last used the array
        cudaEventDestroy(pick.ev);
        *cap = pick.bytes;
        return pick.p;
    }
    void* p = nullptr;
    const size_t want = (bytes + ((size_t)2 << 20) - 1) & ~(((size_t)2 << 20) - 1);
    if (getenv("BS2E_TRACE")) fprintf(stderr, "[bs2e] out_take: no cached array for %.2f GB (cache holds %zu arrays, %.1f GB)\n",
                                      1e-9 * (double)bytes, c->out_free.size(), 1e-9 * (double)c->out_cached);
    if (cudaMallocAsync(&p, want, st) != cudaSuccess) {   // out of memory with arrays parked in the cache: drop them
        cudaGetLastError();
        out_release_all(c);
        BS2E_CUDA(cudaStreamSynchronize(c->stream));
        BS2E_CUDA(cudaMallocAsync(&p, want, st));
    }
    *cap = want;
    return p;
}

void out_give(bs2e_ctx* c, void* p, size_t cap, cudaStream_t st)
{
    if (!p) return;
    if (c->out_cap == 0) {
        size_t fr = 0, tot = 0;
        cudaMemGetInfo(&fr, &tot);
        double frac = 0.45;
        if (const char* e = getenv("BS2E_OUT_CACHE_FRAC")) frac = std::min(0.9, std::max(0.0, atof(e)));   // A/B measurements
        c->out_cap = (size_t)(frac * (double)tot);
    }
    bs2e_ctx::OutBuf b{p, cap, nullptr};
    if (cap > c->out_cap || cudaEventCreateWithFlags(&b.ev, cudaEventDisableTiming) != cudaSuccess) {
        cudaGetLastError();
        cudaFreeAsync(p, st);
        return;
    }
    cudaEventRecord(b.ev, st);
    std::vector<bs2e_ctx::OutBuf> drop;
    {
        std::lock_guard<std::mutex> lk(c->out_mu);
        c->out_free.push_back(b);
        c->out_cached += cap;
        while (c->out_cached > c->out_cap || c->out_free.size() > 24) {   // the smallest arrays leave first
            int at = 0;
            for (int i = 1; i < (int)c->out_free.size(); ++i)
                if (c->out_free[i].bytes < c->out_free[at].bytes) at = i;
            drop.push_back(c->out_free[at]);
            c->out_cached -= c->out_free[at].bytes;
            c->out_free.erase(c->out_free.begin() + at);
        }
    }
    for (auto& d : drop) {
        cudaStreamWaitEvent(st, d.ev, 0);
        cudaFreeAsync(d.p, st);
        cudaEventDestroy(d.ev);
    }
}

void out_release_all(bs2e_ctx* c)
{
    std::vector<bs2e_ctx::OutBuf> all;
    {
        std::lock_guard<std::mutex> lk(c->out_mu);
        all.swap(c->out_free);
        c->out_cached = 0;
    }
    for (auto& d : all) {
        cudaStreamWaitEvent(c->stream, d.ev, 0);
        cudaFreeAsync(d.p, c->stream);
        cudaEventDestroy(d.ev);
    }
}

void block_free(bs2e_block* b)
{
    if (!b) return;
    if (b->ctx) {   // stream-ordered: returns at once, the arrays wait in the context's cache for the next block
        cudaStream_t st = b->ctx->stream;
        if (b->d_Hdat) out_give(b->ctx, b->d_Hdat, b->cap_Hdat, st);
        if (b->d_Hidx) out_give(b->ctx, b->d_Hidx, b->cap_Hidx, st);
        if (b->d_Sdat) out_give(b->ctx, b->d_Sdat, b->cap_Sdat, st);
        if (b->d_Sidx) out_give(b->ctx, b->d_Sidx, b->cap_Sidx, st);
    }
    if (b->ctx) {   // the plan tables go back to the context's pool, reusable once the queued work has read them
        arena_give(b->ctx, b->arena1, b->ctx->stream);
        arena_give(b->ctx, b->arena0, b->ctx->stream);
    }
    delete b;
}

}  // namespace bs2e
