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
//     the configuration list (no pair scan): site_count_kernel counts them
//     in closed form per radial site, an exclusive scan builds index_ptr, and
//     site_fill_kernel (one CTA per radial site, one thread per candidate
//     column) writes indices and values; block_count_kernel / block_fill_kernel
//     (one thread / one warp per row) remain as the general fallback.
// Bound: HBM (24 B written per stored element + R^k gathers).
#include <cub/device/device_scan.cuh>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <set>
#include <unordered_map>

#include "ctx.h"
#include "plan.h"
#include "site_core.h"

namespace bs2e {

// ---------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------
__global__ void block_count_kernel(Geom g, Plan pl, long long nrows,
                                   long long* __restrict__ cntH, long long* __restrict__ cntS)
{
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx > nrows) return;
    long long h = 0, s = 0;
    if (idx < nrows) row_count(g, pl, pl.rows[idx], &h, &s);
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
    const RowInfo r = row_info(pl, pl.rows[wrow]);
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
// angular factors and the clipping by the column block differ.
//   phase 1  per column block bj and n_c slot: clipped windows T[bj][q]
//   phase 2  per (bj, storage mode): prefix over the n_c slots of the number of
//            stored entries, hp[bj][mode][q] (mode: D, X, D+X, diagonal pair)
//   phase 3  per row of the site: storage mode, k parity and offset inside the
//            row of each of its column blocks (pm); per (bj, mode) the bit mask
//            of the rows that store that list (gmask)
//   phase 4  THREAD t OWNS CANDIDATE COLUMN t of the site: it loads the R^k
//            values of the column for all multipoles into registers once (both
//            windows) and then walks the column blocks; per (bj, mode) it works
//            out once whether / where the column is stored, then for every row
//            of the mask sum_k ang_k R^k with the packed factors read as
//            broadcast loads -- per stored element ~K1/2 FP64 FMAs, K1/4 loads,
//            two stores.  R^k is read from HBM exactly once per site.
// ---------------------------------------------------------------------------
struct SiteList {
    const unsigned* key;  // [nsites]  n_a << 16 | n_b
    const int* ptr;       // [nsites+1] into rows
    const int* rows;      // 1-based configuration (row) indices, ascending per site
    int nsites;
};

struct alignas(16) RowCache {  // per row of the group
    long long hbase, sbase;    // 0-based position of the first H / S entry of the row
    int bi, la, lb, pad;
};

struct alignas(16) RowRec {    // one (row, column block) pair of the group, filed under (bj, mode)
    long long hpos;            // 0-based position of the pair's first H entry
    int cf;                    // index of the pair's packed factors (units of 2*NKP doubles)
    int meta;                  // pd | px << 1 | ri << 8
};

constexpr int kSiteThreads = 128;   // sites with exchange windows
constexpr int kSiteThreadsD = 256;  // sites without: one pass over the <= (2w+1)^2 candidates

struct SiteSmem {   // element counts of the dynamic shared memory carve-up
    int ncmax;      // n_c slots
    int G;          // rows per group (<= 32)
    int cfsm;       // packed factors of the group's pairs staged in shared memory
    int nl;         // l_max + 1 of the one-particle matrices
    size_t bytes;
};

__host__ __device__ inline size_t site_smem_bytes(const Geom& g, int nblk, int G, int nkp, bool cfsm, int nl, bool wx)
{
    const size_t ncmax = (size_t)site_max_nc(g);
    size_t b = 0;
    b += sizeof(double) * (size_t)site_1p_doubles(g, nl);               // band rows of H_l and S
    b += sizeof(SiteEntry) * (size_t)nblk * ncmax;                      // T
    b += sizeof(RowRec) * (size_t)nblk * G;                             // rlist
    b += sizeof(RowCache) * (size_t)G;                                  // rcache
    // cfs: packed factors (direct half only without X) of all pairs of the group, or, when those do
    // not fit, of the rows of one column block at a time
    b += sizeof(double) * (size_t)G * (cfsm ? nblk : 1) * (wx ? 2 : 1) * nkp;
    b += sizeof(uchar4) * (size_t)((nblk + 3) & ~3);                    // gcnt
    b += sizeof(unsigned) * (size_t)((G * nblk + 3) & ~3);              // pm
    b += sizeof(int) * ((ncmax + 1 + 3) & ~3);                          // cprefix
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

constexpr size_t kSiteSmemLimit = 160 * 1024;
constexpr size_t kSiteCoefSmem = 48 * 1024;  // stage the packed factors when they fit this

// the packed factors of one (row block, column block) pair and window: NKP doubles
template <int NKP, bool SM>
__device__ __forceinline__ void load_coefs(const double* __restrict__ cf, double* out)
{
#pragma unroll
    for (int i = 0; i < NKP; i += 2) {
        const double2 v = SM ? *reinterpret_cast<const double2*>(cf + i)
                             : __ldg(reinterpret_cast<const double2*>(cf + i));
        out[i] = v.x;
        out[i + 1] = v.y;
    }
}

// WX: the sites of this launch have exchange windows (site_wants_X); the sites that
// have none run a leaner instantiation (no exchange registers, half the n_c slots).
template <int NT, int KMAX, bool CFSM, bool WX>
__global__ void __launch_bounds__(NT, (WX ? (KMAX <= 13 ? 512 : KMAX <= 21 ? 384 : 256) : (KMAX <= 13 ? 768 : 512)) / NT)
site_fill_kernel(Geom g, Plan pl, OneBody ob, SiteList sl, SiteSmem lay, int site_off, const double* __restrict__ R,
                 const long long* __restrict__ Hptr,
                 const long long* __restrict__ Sptr, long long* __restrict__ Hidx,
                 double2* __restrict__ Hdat, long long* __restrict__ Sidx,
                 double2* __restrict__ Sdat)
{
    constexpr int NW = NT / 32;
    constexpr int NKP = (((KMAX + 1) / 2) + 1) & ~1;  // = site_nkp(KMAX)
    extern __shared__ __align__(16) unsigned char smraw[];
    const int K1 = g.K1, ncmax = lay.ncmax, G = lay.G;
    const int nblk = pl.nblk;
    double* ob_s = reinterpret_cast<double*>(smraw);
    SiteEntry* T = reinterpret_cast<SiteEntry*>(ob_s + site_1p_doubles(g, lay.nl));
    RowRec* rlist = reinterpret_cast<RowRec*>(T + (size_t)nblk * ncmax);
    RowCache* rcache = reinterpret_cast<RowCache*>(rlist + (size_t)nblk * G);
    double* cfs = reinterpret_cast<double*>(rcache + G);
    constexpr int CFS = WX ? 2 * NKP : NKP;  // staged doubles per pair (direct half only without X)
    uchar4* gcnt = reinterpret_cast<uchar4*>(cfs + (size_t)G * (CFSM ? nblk : 1) * CFS);
    unsigned* pm = reinterpret_cast<unsigned*>(gcnt + ((nblk + 3) & ~3));
    int* cprefix = reinterpret_cast<int*>(pm + ((G * nblk + 3) & ~3));
    unsigned short* hp = reinterpret_cast<unsigned short*>(cprefix + ((ncmax + 1 + 3) & ~3));
    unsigned short* sp = hp + (size_t)nblk * kModes * (ncmax + 1);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int sidx = blockIdx.x + site_off;
    const unsigned key = sl.key[sidx];
    const Site s = make_site(g, (int)(key >> 16), (int)(key & 0xffffu), WX);
    const int* srows = sl.rows + sl.ptr[sidx];
    const int nr = sl.ptr[sidx + 1] - sl.ptr[sidx];
    const int nnc = s.nnc;
    constexpr bool wantX = WX;
    const size_t plane = (size_t)g.P * g.ldP;

    // The candidate column this thread owns and its R^k values (all multipoles, both
    // windows), read once and streaming.  Without exchange windows the candidate list
    // is known up front, so the loads of the first pass are issued here and complete
    // behind phases 1-3.
    OwnCand c;
    double Rd[KMAX], Rx[WX ? KMAX : 1];
    auto load_cand = [&](int t, bool act) {
        c = site_own_cand(g, s, cprefix, wantX, act ? t : 0);
        const double* pD = R + (size_t)c.rowD * g.ldP + c.colD;
        const double* pX = R + (size_t)c.rowX * g.ldP + c.colX;
        const bool onD = act && c.inD, onX = act && c.inX;
#pragma unroll
        for (int k = 0; k < KMAX; ++k) {
            Rd[k] = (onD && k < K1) ? __ldcs(pD + (size_t)k * plane) : 0.0;
            if constexpr (WX) Rx[k] = (onX && k < K1) ? __ldcs(pX + (size_t)k * plane) : 0.0;
        }
        if constexpr (!WX) Rx[0] = 0.0;
    };
    const bool prefetched = !WX && s.nD <= NT && nr <= G;
    if (prefetched) load_cand(tid, tid < s.nD);

    // ---- phase 1: clipped windows per column block; slot prefix of the candidate list ----
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
    for (int idx = tid; idx < site_1p_doubles(g, lay.nl) / 2; idx += NT) {
        const Cplx v = site_1p_source(g, ob, s, lay.nl, idx);
        reinterpret_cast<double2*>(ob_s)[idx] = make_double2(v.re, v.im);
    }
    const SiteOneBody so{ob_s, ob_s + (size_t)lay.nl * 2 * (2 * g.w + 1) * 2};
    __syncthreads();
    // ---- phase 2: prefix of stored entries over the n_c slots, per (bj, mode);
    //      one thread per (bj, mode), serial over the slots ----
    for (int task = tid; task < nblk * kModes; task += NT) {
        const int bj = task / kModes, mode = task - bj * kModes;
        if (!wantX && (mode == kModeX || mode == kModeDX)) continue;  // never read
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

    double* const Hd = reinterpret_cast<double*>(Hdat);
    double* const Sd = reinterpret_cast<double*>(Sdat);

    for (int g0 = 0; g0 < nr; g0 += G) {
        const int gr = imin(G, nr - g0);
        __syncthreads();  // phase 2 / previous group finished
        // ---- phase 3a: coupled column blocks of each row and their offsets inside the row ----
        for (int ri = warp; ri < gr; ri += NW) {
            const int rowi = srows[g0 + ri];
            const RowInfo r = row_info(pl, rowi);
            if (lane == 0) {
                const long long wrow = pl.row_local[rowi - 1];
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
                if (bj < nblk)
                    pm[ri * nblk + bj] = c > 0 ? pm_pack(run + inc - c, mode, (r.la + pl.blk[bj].l1) & 1) : 0u;
                run += __shfl_sync(0xffffffffu, inc, 31);
            }
        }
        __syncthreads();
        // ---- phase 3b: the pairs of each column block, filed by storage mode ----
        for (int bj = tid; bj < nblk; bj += NT) {
            int cnt[kModes] = {0, 0, 0, 0};
            for (int ri = 0; ri < gr; ++ri) {
                const unsigned v = pm[ri * nblk + bj];
                if (pm_valid(v)) ++cnt[pm_mode(v)];
            }
            gcnt[bj] = make_uchar4((unsigned char)cnt[0], (unsigned char)cnt[1], (unsigned char)cnt[2],
                                   (unsigned char)cnt[3]);
            int pos[kModes] = {0, cnt[0], cnt[0] + cnt[1], cnt[0] + cnt[1] + cnt[2]};
            for (int ri = 0; ri < gr; ++ri) {
                const unsigned v = pm[ri * nblk + bj];
                if (!pm_valid(v)) continue;
                const int mode = pm_mode(v);
                int where = 0;  // pos[mode]++ without dynamic indexing of a register array
#pragma unroll
                for (int q = 0; q < kModes; ++q)
                    if (q == mode) where = pos[q]++;
                const RowCache rc = rcache[ri];
                const int px = pm_pd(v) ^ ((pl.blk[bj].l1 + pl.blk[bj].l2) & 1);
                rlist[bj * G + where] = RowRec{rc.hbase + pm_off(v), CFSM ? ri * nblk + bj : rc.bi * nblk + bj,
                                               pm_pd(v) | (px << 1) | (ri << 8)};
            }
        }
        if (CFSM) {  // packed factors of the group's pairs -> shared memory (16-byte units)
            const int per = CFS / 2;  // double2 per pair
            for (int idx = tid; idx < gr * nblk * per; idx += NT) {
                const int pair = idx / per, l = idx - pair * per;
                const int ri = pair / nblk, bj = pair - ri * nblk;
                if (!pm_valid(pm[pair])) continue;
                const double2* src = reinterpret_cast<const double2*>(
                    pl.angP + ((size_t)rcache[ri].bi * nblk + bj) * (2 * NKP));
                reinterpret_cast<double2*>(cfs)[(size_t)pair * per + l] = __ldg(src + l);
            }
        }
        __syncthreads();
        // ---- phase 4: fill ----
        const int nc_all = site_num_cand(s, cprefix, wantX);
        for (int t0 = 0; t0 < nc_all; t0 += NT) {
            const int t = t0 + tid;
            const bool act = t < nc_all;
            if (!(prefetched && t0 == 0)) load_cand(t, act);
            const bool warp_has_cand = __ballot_sync(0xffffffffu, act) != 0u;
            if (CFSM && !warp_has_cand) continue;  // (with per-block staging every warp must reach the barriers)
            for (int bj = 0; bj < nblk; ++bj) {
                const uchar4 gc = gcnt[bj];
                if ((gc.x | gc.y | gc.z | gc.w) == 0) continue;
                const RowRec* rl = rlist + bj * G;
                if (!CFSM) {  // stage the packed factors of the rows of this column block (list order)
                    const int nlist = gc.x + gc.y + gc.z + gc.w;
                    __syncthreads();   // the previous tile has been consumed
                    for (int idx = tid; idx < nlist * (CFS / 2); idx += NT) {
                        const int row = idx / (CFS / 2), l = idx - row * (CFS / 2);
                        const double2* src = reinterpret_cast<const double2*>(pl.angP + (size_t)rl[row].cf * (2 * NKP));
                        reinterpret_cast<double2*>(cfs)[idx] = __ldg(src + l);
                    }
                    __syncthreads();
                    if (!warp_has_cand) continue;
                }
                const SiteEntry e = T[bj * ncmax + c.q];
                const RowRec* const rl0 = rl;
#pragma unroll
                for (int mode = 0; mode < kModes; ++mode) {
                    if (!WX && (mode == kModeX || mode == kModeDX)) continue;  // no such pairs without exchange windows
                    const int nrow = mode == 0 ? gc.x : mode == 1 ? gc.y : mode == 2 ? gc.z : gc.w;
                    if (nrow == 0) continue;
                    const RowRec* rm = rl;
                    rl += nrow;
                    const bool diag = mode == kModeDiag;
                    const ModeSlot ms = site_mode_slot(s, c, e, hp + (bj * kModes + mode) * (ncmax + 1), mode,
                                                       diag && !pl.full);
                    const bool stored = act && (ms.sup || ms.sup_ex);
                    if (__ballot_sync(0xffffffffu, stored) == 0u) continue;
                    const long long jcol = ms.jcol;
                    if (!diag) {
#pragma unroll 2
                        for (int i = 0; i < nrow; ++i) {
                            const RowRec rec = rm[i];
                            const int pd = rec.meta & 1, px = (rec.meta >> 1) & 1;
                            const double* cf = cfs + (size_t)(CFSM ? rec.cf : (int)(rm - rl0) + i) * CFS;
                            double res = 0.0;
                            if (mode != kModeX) {
                                double cD[NKP];
                                load_coefs<NKP, true>(cf, cD);
                                const double d = site_dot_par<KMAX>(cD, Rd, pd);
                                res += ms.sup ? d : 0.0;
                            }
                            if constexpr (WX) {
                                if (mode != kModeD) {
                                    double cX[NKP];
                                    load_coefs<NKP, true>(cf + NKP, cX);
                                    const double x = site_dot_par<KMAX>(cX, Rx, px);
                                    res += ms.sup_ex ? x : 0.0;
                                }
                            }
                            if (stored) {
                                const long long pos = rec.hpos + ms.rank;
                                Hidx[pos] = jcol;
                                *reinterpret_cast<double2*>(Hd + 2 * pos) = make_double2(res, 0.0);
                            }
                        }
                    } else {  // exactly one row: the row whose own block is bj
                        const RowRec rec = rm[0];
                        const RowCache rc = rcache[rec.meta >> 8];
                        const int pd = rec.meta & 1, px = (rec.meta >> 1) & 1;
                        const double* cf = cfs + (size_t)(CFSM ? rec.cf : (int)(rm - rl0)) * CFS;
                        double cD[NKP];
                        load_coefs<NKP, true>(cf, cD);
                        const double d = site_dot_par<KMAX>(cD, Rd, pd);
                        double res = 0.0;
                        res += ms.sup ? d : 0.0;
                        if constexpr (WX) {
                            double cX[NKP];
                            load_coefs<NKP, true>(cf + NKP, cX);
                            const double x = site_dot_par<KMAX>(cX, Rx, px);
                            res += ms.sup_ex ? x : 0.0;
                        }
                        if (stored) {
                            double re = res, im = 0.0;
                            site_diag_terms(g, pl, so, s, rc.la, rc.lb, c, ms, rc.la == rc.lb, sp + bj * (ncmax + 1),
                                            rc.sbase, &re, &im, Sidx, Sd);
                            const long long pos = rec.hpos + ms.rank;
                            Hidx[pos] = jcol;
                            *reinterpret_cast<double2*>(Hd + 2 * pos) = make_double2(re, im);
                        }
                    }
                }
            }
        }
    }
}

// Count pass on the site tables.  The row count of a row is the sum over its coupled
// column blocks of the stored entries of the pair's storage mode, which depend on the
// site and the column block only.  One WARP per site, no block-wide barrier: lane = column
// block, one pass over the n_c slots gives the totals of all four storage modes; then
// lane = column block again for each row of the site.  (block_count_kernel, one thread
// per row, is kept for plans without a site list.)
struct CountSmem { int stride; size_t bytes; };
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
    unsigned short* stot = tot + (size_t)kModes * nblk;   // S entries of the diagonal pair of block bj
    const int sidx = blockIdx.x * kCountWarps + warp;
    if (blockIdx.x == 0 && threadIdx.x == 0) { cntH[pl.nrows] = 0; cntS[pl.nrows] = 0; }  // the scan yields the totals there
    if (sidx >= sl.nsites) return;
    const unsigned key = sl.key[sidx];
    const bool wantX = site_wants_X(g, pl.max_nd, (int)(key >> 16));
    const Site s = make_site(g, (int)(key >> 16), (int)(key & 0xffffu), wantX);
    const int* srows = sl.rows + sl.ptr[sidx];
    const int nr = sl.ptr[sidx + 1] - sl.ptr[sidx];
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
    for (int ri = 0; ri < nr; ++ri) {
        const int rowi = srows[ri];
        const RowInfo r = row_info(pl, rowi);
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
            const int wrow = pl.row_local[rowi - 1];
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

// count pass + exclusive scan -> 1-based index_ptr of the planned rows
BlockStreams default_streams(bs2e_ctx* c) { return BlockStreams{c->stream, c->side, c->ev_fork, c->ev_join}; }

void block_count_scan(bs2e_block* b, bool read_totals, const BlockStreams* bsp)
{
    bs2e_ctx* c = b->ctx;
    cudaStream_t st = bsp ? bsp->main : c->stream;
    const long long nrows = b->nrows;
    const char* cmode = getenv("BS2E_COUNT");
    const size_t cbytes = count_smem_bytes(b->dplan.nblk);
    if (b->nsites > 0 && cbytes <= 48 * 1024 && !(cmode && strcmp(cmode, "row") == 0)) {
        const SiteList sl{b->d_site_key, b->d_site_ptr, b->d_site_rows, b->nsites};
        site_count_kernel<<<(unsigned)((b->nsites + kCountWarps - 1) / kCountWarps), kCountWarps * 32, cbytes, st>>>(
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

// ---------------------------------------------------------------------------
// plan: derive the block structure from the configuration list, count, scan
// ---------------------------------------------------------------------------
bs2e_block* block_plan(bs2e_ctx* c, int L, long long n_config, const int64_t* conf_n,
                       const int64_t* conf_l, int full, long long n_ranges, const int64_t* range_lo,
                       const int64_t* range_hi)
{
    HostPlan hp;
    try {
        hp = build_host_plan(c->hg, L, n_config, conf_n, conf_l, full, n_ranges, range_lo, range_hi);
    } catch (const std::invalid_argument& e) {
        throw Error(e.what());
    }
    const int nblk = hp.nblk;
    bs2e_block* b = new bs2e_block();
    b->ctx = c;
    b->L = L;
    b->full = full ? 1 : 0;
    b->n_config = n_config;
    b->nrows = (long long)hp.rows.size();
    b->lmax = hp.lmax;
    try {
        cudaStream_t st = c->stream;
        b->d_blk = dev_upload(hp.blocks, st);
        b->d_ncrow = dev_upload(hp.ncrow, st);
        b->d_flags = dev_upload(hp.flags, st);
        b->d_krange = dev_upload(hp.krange, st);
        b->d_angD = dev_upload(hp.angD, st);
        b->d_angX = dev_upload(hp.angX, st);
        if (hp.nkp > 0) b->d_angP = dev_upload(hp.angP, st);
        b->d_row_n1 = dev_upload(hp.row_n1, st);
        b->d_row_n2 = dev_upload(hp.row_n2, st);
        b->d_row_blk = dev_upload(hp.row_blk, st);
        b->d_rows = dev_upload(hp.rows, st);
        b->d_row_local = dev_upload(hp.row_local, st);
        b->nsites = (int)hp.site_key.size();
        b->nsites_x = hp.nsites_x;
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
        pl.angP = b->d_angP;
        pl.nkp = hp.nkp;
        pl.row_n1 = b->d_row_n1;
        pl.row_n2 = b->d_row_n2;
        pl.row_blk = b->d_row_blk;
        pl.nrows = (int)hp.rows.size();
        pl.rows = b->d_rows;
        pl.row_local = b->d_row_local;
        pl.max_nd = hp.max_nd;

        const long long nrows = b->nrows;
        b->d_cntH = dev_alloc<long long>(nrows + 1);
        b->d_cntS = dev_alloc<long long>(nrows + 1);
        b->d_Hptr = dev_alloc<long long>(nrows + 1);
        b->d_Sptr = dev_alloc<long long>(nrows + 1);
        size_t tmp = 0;
        BS2E_CUDA(cub::DeviceScan::ExclusiveScan(nullptr, tmp, b->d_cntH, b->d_Hptr, cub::Sum(), 1LL, nrows + 1, st));
        b->scan_tmp_bytes = tmp;
        BS2E_CUDA(cudaMalloc(&b->d_scan_tmp, tmp ? tmp : 1));
        block_count_scan(b, true);
    } catch (...) {
        block_free(b);
        throw;
    }
    return b;
}

void block_assemble(bs2e_block* b, const BlockStreams* bsp)
{
    bs2e_ctx* c = b->ctx;
    const BlockStreams bs = bsp ? *bsp : default_streams(c);
    if (!c->have_R) throw Error("block_assemble: call bs2e_rk_build first");
    if (!c->have_1p) throw Error("block_assemble: call bs2e_set_one_particle first");
    if (b->lmax > c->lmax_1p) throw Error("block_assemble: configuration l exceeds max_l_1p of H_vec");
    if (!b->d_Hidx) {
        b->d_Hidx = dev_alloc_async<long long>(b->nnzH, bs.main);
        b->d_Hdat = dev_alloc_async<double>(2 * (size_t)b->nnzH, bs.main);
        b->d_Sidx = dev_alloc_async<long long>(b->nnzS, bs.main);
        b->d_Sdat = dev_alloc_async<double>(2 * (size_t)b->nnzS, bs.main);
    }
    const long long nrows = b->nrows;
    // site kernel unless max_k exceeds its largest instantiation or its tables do
    // not fit shared memory; BS2E_FILL=row asks for the row kernel (A/B measurements)
    const Geom& g = c->dg;
    const char* mode = getenv("BS2E_FILL");
    const int nblk = b->dplan.nblk;
    const int kmax = site_kmax_for(g.K1);
    bool use_site = b->nsites > 0 && b->d_angP && !(mode && strcmp(mode, "row") == 0) && kmax > 0 &&
                    site_max_slots(g) <= 65535 && (size_t)nblk * site_max_slots(g) < (1u << 24);
    SiteSmem lay{};
    if (use_site) {
        const int nkp = site_nkp(kmax);
        lay.ncmax = site_max_nc(g);
        lay.G = std::min(32, nblk);   // a site has at most one row per (l1,l2) block
        lay.cfsm = sizeof(double) * (size_t)lay.G * nblk * 2 * nkp <= kSiteCoefSmem;
        const char* cmode = getenv("BS2E_SITE_COEFS");   // "block": force the per-column-block staging (tests)
        if (cmode && strcmp(cmode, "block") == 0) lay.cfsm = 0;
        lay.nl = c->lmax_1p + 1;
        lay.bytes = site_smem_bytes(g, nblk, lay.G, nkp, lay.cfsm != 0, lay.nl, true);
        if (lay.bytes > kSiteSmemLimit) use_site = false;
    }
    if (use_site) {
        const SiteList sl{b->d_site_key, b->d_site_ptr, b->d_site_rows, b->nsites};
        // the launch without exchange windows goes to the side stream so that it fills
        // the SMs the other launch leaves idle in its last wave
        const bool fork = b->nsites_x > 0 && b->nsites_x < b->nsites;
        if (fork) {
            BS2E_CUDA(cudaEventRecord(bs.fork, bs.main));
            BS2E_CUDA(cudaStreamWaitEvent(bs.side, bs.fork, 0));
        }
        auto launch = [&](auto kern, int nt, bool wx, int first, int count) {
            if (count <= 0) return;
            cudaStream_t st = (fork && !wx) ? bs.side : bs.main;
            const size_t bytes = site_smem_bytes(g, nblk, lay.G, site_nkp(kmax), lay.cfsm != 0, lay.nl, wx);
            BS2E_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
            kern<<<(unsigned)count, nt, bytes, st>>>(
                c->dg, b->dplan, c->one_body(), sl, lay, first, c->d_R, b->d_Hptr, b->d_Sptr,
                b->d_Hidx, reinterpret_cast<double2*>(b->d_Hdat), b->d_Sidx, reinterpret_cast<double2*>(b->d_Sdat));
            BS2E_LAUNCHED();
        };
        constexpr int NT = kSiteThreads, NTD = kSiteThreadsD;
        const int nx = b->nsites_x, nd = b->nsites - b->nsites_x;
#define BS2E_SITE(KM)                                                          \
    case KM:                                                                   \
        if (lay.cfsm) {                                                        \
            launch(site_fill_kernel<NT, KM, true, true>, NT, true, 0, nx);           \
            launch(site_fill_kernel<NTD, KM, true, false>, NTD, false, nx, nd);         \
        } else {                                                               \
            launch(site_fill_kernel<NT, KM, false, true>, NT, true, 0, nx);          \
            launch(site_fill_kernel<NTD, KM, false, false>, NTD, false, nx, nd);        \
        }                                                                      \
        break;
        switch (kmax) {
            BS2E_SITE(7) BS2E_SITE(13) BS2E_SITE(21) BS2E_SITE(31)
        default: throw Error("block_assemble: no site kernel for this max_k");
        }
#undef BS2E_SITE
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
void blocks_run(bs2e_ctx* c, long long n, bs2e_block** blks, bool recount)
{
    if (n <= 0) return;
    if (!c->have_lanes) {
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
    cudaStream_t st = b->ctx->stream;
    const long long nrows = b->nrows;
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
    const long long nrows = b->nrows;
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
    cudaFree(b->d_angD); cudaFree(b->d_angX); cudaFree(b->d_angP);
    cudaFree(b->d_row_n1); cudaFree(b->d_row_n2); cudaFree(b->d_row_blk);
    cudaFree(b->d_rows); cudaFree(b->d_row_local);
    cudaFree(b->d_site_key); cudaFree(b->d_site_ptr); cudaFree(b->d_site_rows);
    cudaFree(b->d_cntH); cudaFree(b->d_cntS); cudaFree(b->d_Hptr); cudaFree(b->d_Sptr);
    {   // stream-ordered: returns at once, the pool keeps the pages for the next block
        cudaStream_t st = b->ctx ? b->ctx->stream : nullptr;
        if (b->d_Hidx) cudaFreeAsync(b->d_Hidx, st);
        if (b->d_Sidx) cudaFreeAsync(b->d_Sidx, st);
        if (b->d_Hdat) cudaFreeAsync(b->d_Hdat, st);
        if (b->d_Sdat) cudaFreeAsync(b->d_Sdat, st);
    }
    cudaFree(b->d_scan_tmp);
    delete b;
}

}  // namespace bs2e
