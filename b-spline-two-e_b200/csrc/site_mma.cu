// site_mma.cu -- stage C fill on the FP64 tensor cores: one CTA per radial site.
//
// Stands in for construct_block_tensor (src/mat_els/hamiltonian.f90:106-283) with the
// element formulas of src/mat_els/mat_els.f90:552-571,608-633,664-715.
//
// All rows (l_a,l_b; n_a,n_b) of a symmetry block that sit on the radial site (n_a,n_b)
// couple to the same CANDIDATE columns (n_c,n_d) -- the union of the direct and the
// exchange window, site_core.h -- and read the same R^k values; a stored entry is
//     H = sum_k angD_k R^k(n_a n_b; n_c n_d) + sum_k angX_k R^k(n_a n_b; n_d n_c)
// with factors that depend on the (row group, column group) pair only.  Over a site this
// is a dense product   C[candidate][record] = Rmat[candidate][k] * F[k][record]
// (record = one coupled (row, column group) pair; k = the multipoles of one parity,
// wigner_tools.f90:131), run as mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4):
//     A fragment  R^k values of 8 candidates x 4 multipoles, in registers for the whole site
//     B fragment  angular factors of 8 records x 4 multipoles, staged in shared memory by
//                 the bulk-copy engine (cp.async.bulk + mbarrier, double buffered)
//     C fragment  thread (lane) holds candidate lane/4 of records 2*(lane%4), 2*(lane%4)+1:
//                 the 8 lanes that share a record write 8 consecutive entries of one CSR row
// The instruction accumulates k in ascending order (measured bit for bit,
// scripts/microbench/dmma_probe.cu), i.e. the sum order of mat_els.f90:566-570.
// Where a candidate is stored inside the row, and whether, comes from bit masks over the
// candidate list per (column group, storage mode) (site_core.h: MaskWord).
// Bound: HBM (24 B written per stored element); R^k is read once per site.
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "ctx.h"
#include "dev_async.h"
#include "plan.h"
#include "site_core.h"
#include "site_mma.h"

namespace bs2e {

namespace {

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b, double c0, double c1)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%4,%5};"
                 : "=d"(d0), "=d"(d1)
                 : "d"(a), "d"(b), "d"(c0), "d"(c1));
}

__device__ __forceinline__ int warp_scan_incl(int v, int lane)
{
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += u;
    }
    return v;
}

// predicated store of one complex value (re, 0) into shared memory at byte address base + 16*off
__device__ __forceinline__ void sts_value_if(unsigned on, unsigned base, unsigned off, double re)
{
    asm volatile(
        "{\n.reg .pred p;\n.reg .u32 a;\nsetp.ne.u32 p, %0, 0;\nmad.lo.u32 a, %1, 16, %2;\n"
        "@p st.shared.v2.f64 [a], {%3, %4};\n}\n" ::"r"(on), "r"(off), "r"(base), "d"(re), "d"(0.0)
        : "memory");
}
// predicated streaming store of a staged complex value: global byte address base + 16*off
__device__ __forceinline__ void st_value2_if(unsigned on, const char* base, unsigned off, double2 v)
{
    asm volatile(
        "{\n.reg .pred p;\n.reg .u64 a;\nsetp.ne.u32 p, %0, 0;\nmad.wide.u32 a, %1, 16, %2;\n"
        "@p st.global.cs.v2.f64 [a], {%3, %4};\n}\n" ::"r"(on), "r"(off), "l"(base), "d"(v.x), "d"(v.y));
}
// predicated streaming stores of one CSR entry at element offset `off` behind a byte base address
__device__ __forceinline__ void st_index_if(unsigned on, const char* base, unsigned off, int jcol)
{
    asm volatile(
        "{\n.reg .pred p;\n.reg .u64 a;\nsetp.ne.u32 p, %0, 0;\nmad.wide.u32 a, %1, 8, %2;\n"
        "@p st.global.cs.v2.u32 [a], {%3, %4};\n}\n" ::"r"(on), "r"(off), "l"(base), "r"(jcol), "r"(0));
}
__device__ __forceinline__ void st_value_if(unsigned on, const char* base, unsigned off, double re)
{
    asm volatile(
        "{\n.reg .pred p;\n.reg .u64 a;\nsetp.ne.u32 p, %0, 0;\nmad.wide.u32 a, %1, 16, %2;\n"
        "@p st.global.cs.v2.f64 [a], {%3, %4};\n}\n" ::"r"(on), "r"(off), "l"(base), "d"(re), "d"(0.0));
}

// one elected lane of the (converged) warp
__device__ __forceinline__ bool elect_one()
{
    unsigned pred;
    asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
    return pred != 0;
}
// shared -> global bulk copy (TMA engine, linear form) as part of the thread's current bulk group
__device__ __forceinline__ void bulk_s2g(void* dst, unsigned src_smem, unsigned bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src_smem), "r"(bytes)
                 : "memory");
}

struct alignas(16) RowC {   // per row of the group
    long long hbase, sbase; // 0-based position of the first H / S entry of the row
    int bi, la, lb, pad;
};

}  // namespace

// CTAs per SM: two (128 registers).  Three CTAs of 80 registers were measured for the sites without exchange
// windows (-DBS2E_MMA_MINB_D=3; their shared memory fits with the half staging tiles): the compiler then keeps the
// A fragments in local memory and reloads them per tile -- 5.70 instead of 4.61 ms on the L=6 block of cfg4
// (profiles/r02v_fill_ncu_summary.md, section r02w), long-scoreboard stalls 1.2 -> 6.4 per issue.
#ifndef BS2E_MMA_MINB_X
#define BS2E_MMA_MINB_X 2
#endif
#ifndef BS2E_MMA_MINB_D
#define BS2E_MMA_MINB_D 2
#endif
#ifndef BS2E_MMA_HALF
#define BS2E_MMA_HALF 1
#endif
// Measured and not kept (profiles/r02v_fill_ncu_summary.md, gpurun_out/r03k): an L2 prefetch of the next (pass, parity)'s
// R^k values one record sweep ahead (41.25 against 41.28 ms over cfg4: the fetches are not what the kernel waits for),
// and a warp-uniform skip of the store sequence of rows whose mask is empty in the warp's segment (41.68 ms: slower).
template <int KMAX, bool WX, bool BULK>
__global__ void __launch_bounds__(kMmaThreads, WX ? BS2E_MMA_MINB_X : BS2E_MMA_MINB_D)
site_mma_kernel(const __grid_constant__ Geom g, const __grid_constant__ Plan pl, const __grid_constant__ OneBody ob,
                const unsigned long long* __restrict__ site_key, const __grid_constant__ MmaSmem lay, int site_off,
                const double* __restrict__ R, const long long* __restrict__ Hptr, const long long* __restrict__ Sptr,
                long long* __restrict__ Hidx, double2* __restrict__ Hdat, long long* __restrict__ Sidx,
                double2* __restrict__ Sdat)
{
    constexpr int NT = kMmaThreads, NW = NT / 32;
    constexpr int NKH = (KMAX + 1) / 2;                // multipoles of one parity
    constexpr int NKP = ((NKH + 1) & ~1);              // = site_nkp(KMAX): packed factors per window
    constexpr int KS = (NKH + 3) / 4;                  // k steps of the 8x8x4 instruction
    constexpr int CFS = WX ? 2 * NKP : NKP;            // doubles staged per record (direct half only without X)
    constexpr int STR = mma_cf_stride(CFS);
    constexpr int CT = kSegCand / 8;                   // candidate tiles per segment
    constexpr int CTH = WX ? CT / 2 : CT;              // tiles whose products are in flight together
    constexpr bool HALF = BS2E_MMA_HALF && !WX && !BULK;   // values staged four rows at a time (half the staging tile)
    static_assert(4 * KS <= NKP, "k steps read inside the packed factors");
    extern __shared__ __align__(16) unsigned char smraw[];
    // the carve-up is computed on the host (mma_layout); the offsets live in the constant bank, not in registers
#define BS2E_SM(type, off) (reinterpret_cast<type*>(smraw + lay.off))
    const int nblk = pl.nblk, ncmax = lay.ncmax, G = lay.G, nsegS = lay.nseg;
    const int mstr = lay.mstr;   // words per row of the mask table (odd: rows of different records fall into different banks)

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned key = (unsigned)(site_key[blockIdx.x + site_off] & 0xffffffffull);
    int nr, pi, nc_all, nseg;
    {
        // ---- phases 0-2: tables of the site ----
        const Site s = make_site(g, (int)(key >> 16), (int)(key & 0xffffu), WX);
        const int nnc = s.nnc;
        int* cprefix = BS2E_SM(int, off_cprefix);
        int* misc = BS2E_SM(int, off_misc);
        unsigned* mraw = BS2E_SM(unsigned, off_mraw);
        // phase 0: rows of the site, candidate prefix, cleared masks
        if (warp == 0) {
            int* srow_bi = BS2E_SM(int, off_srow);
            int* srow_local = srow_bi + nblk;
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
            unsigned long long* mbar = BS2E_SM(unsigned long long, off_mbar);
            mbar_init(&mbar[0], 1);
            mbar_init(&mbar[1], 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        if (warp == NW - 1 && WX) {
            int run = 0;
            for (int q0 = 0; q0 < nnc; q0 += 32) {
                const int q = q0 + lane;
                const int cq = q < nnc ? site_cand_DX_count(s, q) : 0;
                const int inc = warp_scan_incl(cq, lane);
                if (q < nnc) cprefix[q] = run + inc - cq;
                run += __shfl_sync(0xffffffffu, inc, 31);
            }
            if (lane == 0) cprefix[nnc] = run;
        }
        for (int i = tid; i < 2 * nblk * nsegS; i += NT) mraw[i] = 0u;
        {
            BlockDesc* sblk = BS2E_SM(BlockDesc, off_sblk);
            for (int bj = tid; bj < nblk; bj += NT) sblk[bj] = pl.blk[bj];
            int* ncq = BS2E_SM(int, off_ncq);
            for (int q = tid; q < nnc; q += NT) ncq[q] = site_nc(s, q);
            double* ob_s = BS2E_SM(double, off_ob);
            for (int idx = tid; idx < site_1p_doubles(g, lay.nl) / 2; idx += NT) {
                const Cplx v = site_1p_source(g, ob, s, lay.nl, idx);
                reinterpret_cast<double2*>(ob_s)[idx] = make_double2(v.re, v.im);
            }
        }
        __syncthreads();
        nr = misc[0];
        pi = (BS2E_SM(BlockDesc, off_sblk)[0].l1 + BS2E_SM(BlockDesc, off_sblk)[0].l2) & 1;   // parity of l1+l2 of the symmetry
        nc_all = site_num_cand(s, cprefix, WX);
        nseg = (nc_all + kSegCand - 1) / kSegCand;

        // phase 1: clipped windows of every (column group, n_c slot) and their candidate masks
        {
            int* jb = BS2E_SM(int, off_jb);
            unsigned* mD = mraw;                          // [nblk][nsegS] direct-window masks of the column groups
            unsigned* mX = mraw + (size_t)nblk * nsegS;   // exchange-window masks
            for (int idx = tid; idx < nblk * nnc; idx += NT) {
                const int bj = idx / nnc, q = idx - bj * nnc;
                const SiteEntry e = site_entry(g, pl, s, bj, q);
                jb[bj * ncmax + q] = e.jbase;
                const CandSlot cs = cand_slot(s, cprefix, WX, q);
                entry_cand_ranges(
                    cs, e,
                    [&](int t0, int t1) { mask_range(t0, t1, [&](int seg, unsigned bits) { atomicOr(&mD[bj * nsegS + seg], bits); }); },
                    [&](int t0, int t1) { mask_range(t0, t1, [&](int seg, unsigned bits) { atomicOr(&mX[bj * nsegS + seg], bits); }); });
            }
            // the candidates themselves: where their R^k values sit, n_c slot and n_d
            {
                int4* cand = BS2E_SM(int4, off_cand);
                for (int t = tid; t < nc_all; t += NT) {
                    const OwnCand c = site_own_cand(g, s, cprefix, WX, t);
                    cand[t] = make_int4(c.rowD * g.ldP + c.colD, c.rowX * g.ldP + c.colX, (c.q * 4) | (c.nd << 12),
                                        (c.inD ? 1 : 0) | (c.inX ? 2 : 0));
                }
            }
            __syncthreads();
            // phase 2: mask table rows of the (column group, mode) pairs: mask + entries before the segment
            MaskWord* mtab = BS2E_SM(MaskWord, off_mtab);
            unsigned short* tot = BS2E_SM(unsigned short, off_tot);
            for (int task = tid; task < nblk * mask_modes(WX) + 1; task += NT) {
                if (task == nblk * mask_modes(WX)) {   // the all-zero row of the padding records
                    const int row = mask_row_zero(nblk, G, WX);
                    for (int seg = 0; seg < nseg; ++seg) mtab[row * mstr + seg] = MaskWord{0u, 0u};
                    tot[row] = 0;
                    continue;
                }
                const int bj = WX ? task / 3 : task, mode = WX ? task - bj * 3 : kModeD;
                unsigned run = 0;
                for (int seg = 0; seg < nseg; ++seg) {
                    const unsigned d = mD[bj * nsegS + seg], x = mX[bj * nsegS + seg];
                    const unsigned m = mode == kModeD ? d : (mode == kModeX ? x : (d | x));
                    mtab[task * mstr + seg] = MaskWord{m, run};
                    run += __popc(m);
                }
                tot[task] = (unsigned short)run;
            }
        }
    }

    // ---- this thread's candidates and their R^k values (A fragments) ----
    // candidate lane/4 of every candidate tile ct of the warp's segment; multipoles par + 2*(4 ks + lane%4)
    // (exchange window: parity par ^ pi); where the values sit comes from the candidate table of phase 1.
    const int l4 = lane >> 2, l3 = lane & 3;
    char* const stile = reinterpret_cast<char*>(smraw) + lay.off_stage + warp * (HALF ? kMmaHalfBytes : kMmaStageBytes);   // this warp's staging tile
    int qn[CT];   // n_c slot as byte offset into a jbase row (low 12 bits) and n_d of the candidate of tile ct
    int ql = 0;   // the same for candidate `lane` of the segment (row-wise role)
    double Ad[CT][KS], Ax[CT][WX ? KS : 1];
    auto fetch_A = [&](int pass, int par) {
        const int4* cand = BS2E_SM(int4, off_cand);
        const size_t plane = (size_t)g.P * g.ldP;
        const int t0 = (pass * NW + warp) * kSegCand + l4;
        {
            const int tl = (pass * NW + warp) * kSegCand + lane;
            ql = cand[tl < nc_all ? tl : 0].z;
        }
#pragma unroll
        for (int ct = 0; ct < CT; ++ct) {
            const int t = t0 + ct * 8;
            const bool ok = t < nc_all;
            const int4 cd = cand[ok ? t : 0];
            qn[ct] = cd.z;
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) {
                const int i = 4 * ks + l3;
                const int k = 2 * i + par, kx = 2 * i + (par ^ pi);
                Ad[ct][ks] = (ok && (cd.w & 1) && k < g.K1) ? __ldcs(R + (size_t)k * plane + cd.x) : 0.0;
                if constexpr (WX) Ax[ct][ks] = (ok && (cd.w & 2) && kx < g.K1) ? __ldcs(R + (size_t)kx * plane + cd.y) : 0.0;
            }
            if constexpr (!WX) Ax[ct][0] = 0.0;
        }
    };
    // the values of the first (pass, parity) are requested here, before the record lists are built, so that
    // their latency hides behind phase 3
    if (nr > 0) fetch_A(0, 0);

    unsigned tctr = 0;                        // staged tiles so far (row buffer = tctr & 1)
    unsigned mphase = 0;                      // phase parity of the two staging barriers (bit b)
    bool pending0 = false, pending1 = false;  // a copy into buffer b is in flight

    for (int g0 = 0; g0 < nr; g0 += G) {
        const int gr = imin(G, nr - g0);
        {
            const Site s = make_site(g, (int)(key >> 16), (int)(key & 0xffffu), WX);
            const int nnc = s.nnc;
            int* misc = BS2E_SM(int, off_misc);
            RowC* rcache = BS2E_SM(RowC, off_rcache);
            const BlockDesc* sblk = BS2E_SM(BlockDesc, off_sblk);
            unsigned* mgH = BS2E_SM(unsigned, off_mraw) + (size_t)2 * nblk * nsegS;   // [G][nsegS] H mask of the diagonal pair of each row
            unsigned* mgS = mgH + (size_t)G * nsegS;                                   // S mask
            MaskWord* mtab = BS2E_SM(MaskWord, off_mtab);
            unsigned short* tot = BS2E_SM(unsigned short, off_tot);
            MmaRec* recs = BS2E_SM(MmaRec, off_recs);
            const int capP = lay.cap;
            // ---- phase 3a: the group's rows; masks of their diagonal pairs ----
            {
                const int* srow_bi = BS2E_SM(int, off_srow);
                const int* srow_local = srow_bi + nblk;
                for (int ri = tid; ri < gr; ri += NT) {
                    const int bi = srow_bi[g0 + ri];
                    const long long wrow = srow_local[g0 + ri];
                    rcache[ri] = RowC{Hptr[wrow] - 1, Sptr[wrow] - 1, bi, sblk[bi].l1, sblk[bi].l2, 0};
                }
            }
            for (int i = tid; i < 2 * G * nsegS; i += NT) mgH[i] = 0u;
            if (tid < 4) misc[1 + tid] = 0;
            __syncthreads();   // also closes phase 2
            {
                const int* cprefix = BS2E_SM(int, off_cprefix);
                for (int idx = tid; idx < gr * nnc; idx += NT) {
                    const int ri = idx / nnc, q = idx - ri * nnc;
                    const RowC rc = rcache[ri];
                    SiteEntry e = site_entry(g, pl, s, rc.bi, q);
                    if (!pl.full) e = entry_cut(e, s, site_nc(s, q));   // j >= i inside the diagonal pair
                    const bool samex = rc.la == rc.lb;
                    const CandSlot cs = cand_slot(s, cprefix, WX, q);
                    entry_cand_ranges(
                        cs, e,
                        [&](int t0, int t1) {
                            mask_range(t0, t1, [&](int seg, unsigned bits) {
                                atomicOr(&mgH[ri * nsegS + seg], bits);
                                atomicOr(&mgS[ri * nsegS + seg], bits);
                            });
                        },
                        [&](int t0, int t1) {
                            mask_range(t0, t1, [&](int seg, unsigned bits) {
                                atomicOr(&mgH[ri * nsegS + seg], bits);
                                if (samex) atomicOr(&mgS[ri * nsegS + seg], bits);
                            });
                        });
                }
            }
            __syncthreads();
            for (int task = tid; task < 2 * gr; task += NT) {
                const int ri = task >> 1, isS = task & 1;
                const unsigned* src = (isS ? mgS : mgH) + ri * nsegS;
                const int row = isS ? mask_row_diagS(nblk, G, ri, WX) : mask_row_diagH(nblk, ri, WX);
                unsigned run = 0;
                for (int seg = 0; seg < nseg; ++seg) {
                    mtab[row * mstr + seg] = MaskWord{src[seg], run};
                    run += __popc(src[seg]);
                }
                tot[row] = (unsigned short)run;
            }
            __syncthreads();
            // ---- phase 3b: the coupled column groups of each row (one warp per row, lane = column group):
            //      offset of the pair inside the row, record filed under the parity of its multipoles ----
            for (int ri = warp; ri < gr; ri += NW) {
                const RowC rc = rcache[ri];
                int run = 0;
                for (int b0 = 0; b0 < nblk; b0 += 32) {
                    const int bj = b0 + lane;
                    int cnt = 0, row = 0;
                    if (bj < nblk && (pl.full || bj >= rc.bi)) {
                        if (bj == rc.bi) {
                            row = mask_row_diagH(nblk, ri, WX);
                            cnt = tot[row];
                        } else {
                            const unsigned f = pl.flags[(size_t)rc.bi * nblk + bj];
                            int mode = (f & kDirAny) ? ((f & kExAny) ? kModeDX : kModeD) : ((f & kExAny) ? kModeX : -1);
                            if (!WX && mode == kModeDX) mode = kModeD;   // no exchange windows on this site
                            if (mode >= 0 && (WX || mode == kModeD)) {
                                row = mask_row_pair(bj, mode, WX);
                                cnt = tot[row];
                            }
                        }
                    }
                    const int inc = warp_scan_incl(cnt, lane);
                    const bool diag = cnt > 0 && bj == rc.bi;
                    const int par = diag ? 2 : ((rc.la + (bj < nblk ? sblk[bj].l1 : 0)) & 1);   // list: parity 0, parity 1, diagonal
                    const MmaRec rec{rc.hbase + run + inc - cnt, rc.bi * nblk + bj, (unsigned short)row, (unsigned char)bj,
                                     (unsigned char)ri};
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        const unsigned m = __ballot_sync(0xffffffffu, cnt > 0 && par == c);
                        if (m == 0u) continue;
                        int base = 0;
                        if (lane == 0) base = atomicAdd(&misc[1 + c], __popc(m));
                        base = __shfl_sync(0xffffffffu, base, 0);
                        if (cnt > 0 && par == c) {
                            const int at = base + __popc(m & ((1u << lane) - 1u));
                            // parity 0 from the front of the first half, parity 1 from the front of the second half,
                            // diagonal records from the END of the first half (moved behind the parity-0 records below)
                            recs[c == 0 ? at : (c == 1 ? capP + at : capP - 1 - at)] = rec;
                        }
                    }
                    run += __shfl_sync(0xffffffffu, inc, 31);
                }
            }
            __syncthreads();
            // ---- phase 3c: parity-0 list = records, padding to a multiple of 8, diagonal records, padding ----
            {
                const int n0 = misc[1], n1 = misc[2], ndg = misc[3];
                const int n0p = (n0 + 7) & ~7, dg0 = n0p, n0all = n0p + ((ndg + 7) & ~7), n1p = (n1 + 7) & ~7;
                const MmaRec padrec{0, -1, (unsigned short)mask_row_zero(nblk, G, WX), 0, 0};
                MmaRec mine{};
                const bool mv = tid < ndg;
                if (mv) mine = recs[capP - 1 - tid];
                __syncthreads();   // diagonal records read before their slots are overwritten
                if (mv) recs[dg0 + tid] = mine;
                for (int i = n0 + tid; i < n0p; i += NT) recs[i] = padrec;
                for (int i = dg0 + ndg + tid; i < n0all; i += NT) recs[i] = padrec;
                for (int i = n1 + tid; i < n1p; i += NT) recs[capP + i] = padrec;
            }
            __syncthreads();
        }

        // ---- phase 4: fill ----
        // chunks of CH records; one sequence over (pass, parity, chunk): parity-0 chunks, then parity-1 chunks
        const int CH = lay.chrec;
        volatile const int* nrec = BS2E_SM(int, off_misc) + 1;   // records of parity 0 / parity 1 / diagonal pairs
        auto n0p_f = [&]() { return (nrec[0] + 7) & ~7; };
        auto n0all_f = [&]() { return n0p_f() + ((nrec[2] + 7) & ~7); };
        auto n1p_f = [&]() { return (nrec[1] + 7) & ~7; };
        const int nch0 = (n0all_f() + CH - 1) / CH, nchd = nch0 + (n1p_f() + CH - 1) / CH;
        const int total = nchd > 0 ? ((nseg + NW - 1) / NW) * nchd : 0;
        const bool resident = nchd <= 2;   // both buffers keep their chunk for all passes
        auto stage = [&](int gi) {   // chunk gi of the sequence -> buffer gi & 1: one bulk copy per record
            const int l = gi % nchd;
            const int par = l >= nch0;
            const int first = (par ? l - nch0 : l) * CH;
            const int count = imin(CH, (par ? n1p_f() : n0all_f()) - first);
            const int bsel = gi & 1;
            const MmaRec* rl = BS2E_SM(MmaRec, off_recs) + (par ? lay.cap : 0) + first;
            // the padding records (after the real ones of every list section) are not copied
            auto overlap = [](int a0, int a1, int b0, int b1) { return imax(0, imin(a1, b1) - imax(a0, b0)); };
            const int real = par ? overlap(first, first + count, 0, nrec[1])
                                 : overlap(first, first + count, 0, nrec[0]) +
                                       overlap(first, first + count, n0p_f(), n0p_f() + nrec[2]);
            unsigned long long* bar = BS2E_SM(unsigned long long, off_mbar) + bsel;
            if (tid == 0) mbar_expect_tx(bar, (unsigned)real * CFS * 8u);
            double* dst = BS2E_SM(double, off_cfs) + (size_t)bsel * CH * STR;
            for (int i = tid; i < count; i += NT) {
                const int cf = rl[i].cf;
                if (cf >= 0) bulk_g2s(dst + (size_t)i * STR, pl.angP + (size_t)cf * (2 * NKP), CFS * 8u, bar);
            }
            if (bsel) pending1 = true; else pending0 = true;
        };
        auto wait_buf = [&](int bsel) {
            if (!(bsel ? pending1 : pending0)) return;
            mbar_wait(BS2E_SM(unsigned long long, off_mbar) + bsel, (mphase >> bsel) & 1u);
            mphase ^= 1u << bsel;
            if (bsel) pending1 = false; else pending0 = false;
        };
        if (total > 0) { stage(0); }
        if (resident && nchd == 2) { stage(1); }

#pragma unroll 1
        for (int gi = 0; gi < total; ++gi) {
            const int l = gi % nchd, pass = gi / nchd;
            const int par = l >= nch0;
            const int first = (par ? l - nch0 : l) * CH;
            const int seg = pass * NW + warp;
            // a new (pass, parity): its R^k values (those of the very first one were requested before phase 3)
            if ((l == 0 || l == nch0) && !(g0 == 0 && gi == 0)) fetch_A(pass, par);
            // next chunk of the sequence into the other buffer
            if (!resident && gi + 1 < total) {
                __syncthreads();   // every thread is done with the chunk that buffer held
                stage(gi + 1);
            }
            const int bsel = resident ? l : (gi & 1);
            wait_buf(bsel);
            const bool has = seg < nseg;
            if (!BULK && !has) continue;   // (with CTA-level row staging every warp takes part in its barriers)
            const int nin = has ? imin(kSegCand, nc_all - seg * kSegCand) : 0;   // candidates of this warp's segment
            const int count = imin(CH, (par ? n1p_f() : n0all_f()) - first);
            const int dgt = par ? count : imax(0, n0p_f() - first);   // tiles from here on hold diagonal records
            const double* cfr = BS2E_SM(double, off_cfs) + (size_t)bsel * CH * STR + (size_t)l4 * STR + l3;
            const MmaRec* rl = BS2E_SM(MmaRec, off_recs) + (par ? lay.cap : 0) + first + 2 * l3;
            const MaskWord* mseg = BS2E_SM(MaskWord, off_mtab) + (has ? seg : 0);
#pragma unroll 1
            for (int rt = 0; rt < count; rt += 8, cfr += 8 * STR, rl += 8) {
                // the two records this thread stores for: 2*(lane%4) and the next
                const MmaRec ra = rl[0], rb = rl[1];
                const MaskWord wa = mseg[ra.tbl * mstr], wb = mseg[rb.tbl * mstr];
                // B fragments: factors of record lane/4 of the tile, multipole slot 4 ks + lane%4
                double bd[KS], bx[WX ? KS : 1];
#pragma unroll
                for (int ks = 0; ks < KS; ++ks) {
                    bd[ks] = cfr[4 * ks];
                    if constexpr (WX) bx[ks] = cfr[NKP + 4 * ks];
                }
                if constexpr (!WX) bx[0] = 0.0;
                if (rt < dgt) {
                    const bool any = has && !__all_sync(0xffffffffu, (wa.mask | wb.mask) == 0u);   // something stored in this segment
                    if (!BULK && !any) continue;
                    const unsigned ma = wa.mask >> l4, mb = wb.mask >> l4;   // bit 8 ct: this thread's candidate of tile ct
                    // Stores.  The C fragment holds 8 candidates x 4 rows per instruction: stored directly that is four
                    // 128-byte runs per store, which the L1 -> L2 path takes at a third of its rate; row-wise stores
                    // (one 512-byte run per instruction) reach 0.64 of the HBM write rate, bulk copies of whole rows
                    // from shared memory 0.9 and more (scripts/microbench/store_probe.cu).  Hence
                    //   default: the values are transposed through the warp's staging tile and stored row-wise;
                    //   BULK (BS2E_MMA_STORE=bulk, sites without exchange windows): the values of the 8 rows of a tile
                    //     are collected by all warps in a CTA-level buffer (row r at r*rowb, entry at 16*rank) and leave
                    //     as ONE bulk copy per row (cp.async.bulk shared -> global, issued by warp r), double buffered,
                    //     one barrier per tile.  Measured 7 % slower than the default at cfg4: the barrier per tile and
                    //     the issue slots of the copies cost more than the store path gains;
                    //   column indices: row-wise (lane = candidate of the segment), no staging.
                    if constexpr (HALF) {
                        // records 0,2,4,6 of the tile (first C register of every thread), then 1,3,5,7: four staged rows
                        double c0[CT], c1[CT];
#pragma unroll
                        for (int ct = 0; ct < CT; ++ct) c0[ct] = c1[ct] = 0.0;
#pragma unroll
                        for (int ks = 0; ks < KS; ++ks)
#pragma unroll
                            for (int ct = 0; ct < CT; ++ct)
                                dmma884(c0[ct], c1[ct], Ad[ct][ks], bd[ks], c0[ct], c1[ct]);   // (tiles past the last candidate hold zeros)
                        const MmaRec* rt8 = rl - 2 * l3;   // first record of the tile
                        const unsigned vh = smem_u32(stile) + l3 * kMmaHalfRow;
                        const unsigned lt = (1u << lane) - 1u;
                        const char* jbl = reinterpret_cast<const char*>(BS2E_SM(int, off_jb)) + (ql & 0xfff);
                        const int ndl = ql >> 12;
                        const double2* srow = reinterpret_cast<const double2*>(stile) + lane;
#pragma unroll
                        for (int half = 0; half < 2; ++half) {
                            const unsigned wm = half ? wb.mask : wa.mask;
                            __syncwarp();   // the rows staged before have been read
#pragma unroll
                            for (int ct = 0; ct < CT; ++ct) {
                                const int bit = ct * 8 + l4;
                                sts_value_if((wm >> bit) & 1u, vh, __popc(wm & ((1u << bit) - 1u)), half ? c1[ct] : c0[ct]);
                            }
                            __syncwarp();
                            MmaRec rr[4];
                            MaskWord wr[4];
                            double2 v[4];
                            int jc[4];
#pragma unroll
                            for (int r = 0; r < 4; ++r) rr[r] = rt8[2 * r + half];
#pragma unroll
                            for (int r = 0; r < 4; ++r) {
                                wr[r] = mseg[rr[r].tbl * mstr];
                                v[r] = srow[r * (kMmaHalfRow / 16)];
                                jc[r] = *reinterpret_cast<const int*>(jbl + rr[r].bj * ncmax * 4) + ndl;
                            }
#pragma unroll
                            for (int r = 0; r < 4; ++r) {
                                const long long start = rr[r].hpos + wr[r].pre;
                                st_value2_if(lane < __popc(wr[r].mask) ? 1u : 0u, reinterpret_cast<const char*>(Hdat + start), lane, v[r]);
                                st_index_if((wr[r].mask >> lane) & 1u, reinterpret_cast<const char*>(Hidx + start), __popc(wr[r].mask & lt), jc[r]);
                            }
                        }
                        continue;
                    }
                    unsigned va, vb, offa0, offb0;
                    if constexpr (!BULK) {
                        va = smem_u32(stile) + (2 * l3) * kMmaStageRow;
                        vb = va + kMmaStageRow;
                        offa0 = offb0 = 0;
                        __syncwarp();   // the rows of the previous tile have been read
                    } else {
                        const unsigned vbuf = smem_u32(smraw) + lay.off_stage + (tctr & 1) * 8 * lay.rowb;
                        va = vbuf + (2 * l3) * lay.rowb;
                        vb = va + lay.rowb;
                        const int seg0 = pass * NW;   // first segment of the pass: rows are staged from its first entry on
                        offa0 = wa.pre - BS2E_SM(MaskWord, off_mtab)[ra.tbl * mstr + seg0].pre;
                        offb0 = wb.pre - BS2E_SM(MaskWord, off_mtab)[rb.tbl * mstr + seg0].pre;
                    }
                    if (any) {
                        // products of CTH tiles together: CTH (2 CTH with exchange windows) independent chains
#pragma unroll
                        for (int h = 0; h < CT; h += CTH) {
                            if (h * 8 < nin) {
                                double c0[CTH], c1[CTH];
#pragma unroll
                                for (int u = 0; u < CTH; ++u) c0[u] = c1[u] = 0.0;
#pragma unroll
                                for (int ks = 0; ks < KS; ++ks)
#pragma unroll
                                    for (int u = 0; u < CTH; ++u) dmma884(c0[u], c1[u], Ad[h + u][ks], bd[ks], c0[u], c1[u]);
                                if constexpr (WX) {
                                    double x0[CTH], x1[CTH];
#pragma unroll
                                    for (int u = 0; u < CTH; ++u) x0[u] = x1[u] = 0.0;
#pragma unroll
                                    for (int ks = 0; ks < KS; ++ks)
#pragma unroll
                                        for (int u = 0; u < CTH; ++u) dmma884(x0[u], x1[u], Ax[h + u][ks], bx[ks], x0[u], x1[u]);
#pragma unroll
                                    for (int u = 0; u < CTH; ++u) { c0[u] += x0[u]; c1[u] += x1[u]; }
                                }
#pragma unroll
                                for (int u = 0; u < CTH; ++u) {
                                    const int ct = h + u;
                                    const unsigned below = (1u << (ct * 8 + l4)) - 1u;
                                    sts_value_if((ma >> (8 * ct)) & 1u, va, offa0 + __popc(wa.mask & below), c0[u]);
                                    sts_value_if((mb >> (8 * ct)) & 1u, vb, offb0 + __popc(wb.mask & below), c1[u]);
                                }
                            }
                        }
                    }
                    const MmaRec* rt8 = rl - 2 * l3;   // first record of the tile
                    if constexpr (!BULK) {
                        __syncwarp();
                    } else {
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the staged values, for the copy engine
                        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // this warp's copy of the tile before the last
                        __syncthreads();
                        {   // warp r hands row r of the tile to the bulk-copy engine
                            const MmaRec rr = rt8[warp];
                            const MaskWord* mrow = BS2E_SM(MaskWord, off_mtab) + rr.tbl * mstr;
                            const int seg0 = pass * NW;
                            const unsigned pre0 = mrow[seg0].pre;
                            const unsigned pre1 = seg0 + NW < nseg ? mrow[seg0 + NW].pre : BS2E_SM(unsigned short, off_tot)[rr.tbl];
                            if (pre1 > pre0 && elect_one())
                                bulk_s2g(Hdat + rr.hpos + pre0, smem_u32(smraw) + lay.off_stage + ((tctr & 1) * 8 + warp) * lay.rowb,
                                         (pre1 - pre0) * 16u);
                            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                        }
                        ++tctr;
                    }
                    // row-wise: lane = entry of the row (values, with exchange windows), lane = candidate of the segment (indices)
                    if (any) {
                        const unsigned lt = (1u << lane) - 1u;
                        const char* jbl = reinterpret_cast<const char*>(BS2E_SM(int, off_jb)) + (ql & 0xfff);
                        const int ndl = ql >> 12;
                        const double2* srow = reinterpret_cast<const double2*>(stile) + lane;
#pragma unroll
                        for (int r0 = 0; r0 < 8; r0 += 4) {
                            MmaRec rr[4];
                            MaskWord wr[4];
                            double2 v[BULK ? 1 : 4];
                            int jc[4];
#pragma unroll
                            for (int r = 0; r < 4; ++r) rr[r] = rt8[r0 + r];
#pragma unroll
                            for (int r = 0; r < 4; ++r) {
                                wr[r] = mseg[rr[r].tbl * mstr];
                                if constexpr (!BULK) v[r] = srow[(r0 + r) * (kMmaStageRow / 16)];
                                jc[r] = *reinterpret_cast<const int*>(jbl + rr[r].bj * ncmax * 4) + ndl;
                            }
#pragma unroll
                            for (int r = 0; r < 4; ++r) {
                                const long long start = rr[r].hpos + wr[r].pre;
                                if constexpr (!BULK)
                                    st_value2_if(lane < __popc(wr[r].mask) ? 1u : 0u, reinterpret_cast<const char*>(Hdat + start), lane, v[r]);
                                st_index_if((wr[r].mask >> lane) & 1u, reinterpret_cast<const char*>(Hidx + start), __popc(wr[r].mask & lt), jc[r]);
                            }
                        }
                    }
                } else if (has) {
                    // diagonal pairs (column group == row group): one-body terms and the S entry
                    const RowC* rcache = BS2E_SM(RowC, off_rcache);
                    const RowC rca = rcache[ra.ri], rcb = rcache[rb.ri];
                    const MaskWord sa = mseg[(ra.tbl + G) * mstr], sb = mseg[(rb.tbl + G) * mstr];
                    const double* ob_s = BS2E_SM(double, off_ob);
                    const SiteOneBody so{ob_s, ob_s + (size_t)lay.nl * 2 * (2 * g.w + 1) * 2};
                    const int na = (int)(key >> 16), nb = (int)(key & 0xffffu);
#pragma unroll
                    for (int ct = 0; ct < CT; ++ct) {
                        if (ct * 8 < nin) {
                            double c0 = 0.0, c1 = 0.0;
#pragma unroll
                            for (int ks = 0; ks < KS; ++ks) dmma884(c0, c1, Ad[ct][ks], bd[ks], c0, c1);
                            if constexpr (WX) {
                                double x0 = 0.0, x1 = 0.0;
#pragma unroll
                                for (int ks = 0; ks < KS; ++ks) dmma884(x0, x1, Ax[ct][ks], bx[ks], x0, x1);
                                c0 += x0;
                                c1 += x1;
                            }
                            const int bit = ct * 8 + l4;
                            const unsigned below = (1u << bit) - 1u;
                            const int q4 = qn[ct] & 0xfff, nd = qn[ct] >> 12;
                            const int nc = *reinterpret_cast<const int*>(reinterpret_cast<const char*>(BS2E_SM(int, off_ncq)) + q4);
#pragma unroll
                            for (int e = 0; e < 2; ++e) {
                                const MmaRec& rr = e ? rb : ra;
                                const MaskWord& wh = e ? wb : wa;
                                const MaskWord& ws = e ? sb : sa;
                                const RowC& rc = e ? rcb : rca;
                                if (rr.cf < 0 || !((wh.mask >> bit) & 1u)) continue;
                                double re = e ? c1 : c0, im = 0.0;
                                const int jcol = *reinterpret_cast<const int*>(reinterpret_cast<const char*>(BS2E_SM(int, off_jb) + rr.bj * ncmax) + q4) + nd;
                                if ((ws.mask >> bit) & 1u) {
                                    Cplx h, sv;
                                    site_onebody_at(g, pl.L, so, na, nb, rc.la, rc.lb, nc, nd, rc.la == rc.lb, &h, &sv);
                                    re += h.re;
                                    im += h.im;
                                    const long long spos = rc.sbase + ws.pre + __popc(ws.mask & below);
                                    __stcs(Sidx + spos, (long long)jcol);
                                    __stcs(Sdat + spos, make_double2(sv.re, sv.im));
                                }
                                const long long pos = rr.hpos + wh.pre + __popc(wh.mask & below);
                                __stcs(Hidx + pos, (long long)jcol);
                                __stcs(Hdat + pos, make_double2(re, im));
                            }
                        }
                    }
                }
            }
        }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // the output copies are complete
        // a staging copy that was issued must land before the buffers are reused or the CTA exits
        wait_buf(0);
        wait_buf(1);
        if (g0 + G < nr) __syncthreads();   // the next group rewrites the tables
    }
#undef BS2E_SM
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
namespace {
size_t al16(size_t b) { return (b + 15) & ~(size_t)15; }
}

MmaSmem mma_layout(const bs2e_ctx* c, int nblk, int maxc, int maxrec, bool wx, int lmax, bool bulk)
{
    const Geom& g = c->hg;
    MmaSmem lay{};
    const int kmax = site_kmax_for(g.K1);
    const int nkp = site_nkp(kmax);
    const int cfsn = (wx ? 2 : 1) * nkp, str = mma_cf_stride(cfsn);
    lay.ncmax = wx ? site_max_nc(g) : 2 * g.w + 1;
    lay.nseg = ((wx ? site_max_slots(g) : (2 * g.w + 1) * (2 * g.w + 1)) + kSegCand - 1) / kSegCand;
    lay.mstr = lay.nseg | 1;
    lay.nl = std::max(c->lmax_1p, lmax) + 1;
    size_t cap_kb = 8;
    if (const char* e = getenv("BS2E_SITE_CHUNK_KB")) cap_kb = (size_t)std::max(1, atoi(e));
    lay.chrec = (int)std::max<size_t>(8, (cap_kb * 1024 / (sizeof(double) * str)) & ~(size_t)7);
    // rows per group: as many as the record lists allow within ~48 KB
    const int per_row = std::max(1, std::min(maxc, nblk));
    // rows per group: all rows of a site when its record lists fit ~24 KB per parity, else as many as fit
    int G = std::min(nblk, 255);
    auto cap_of = [&](int rows) {   // parity-0 half also holds the diagonal records
        const int rec = rows >= nblk ? std::min(rows * per_row, std::max(maxrec, 1)) : rows * per_row;
        return ((rec + 7) & ~7) + ((rows + 7) & ~7) + 8;
    };
    while (G > 1 && sizeof(MmaRec) * (size_t)cap_of(G) > 24 * 1024) --G;
    lay.G = G;
    lay.cap = cap_of(G);
    size_t b = 0;
    auto put = [&](int& off, size_t bytes) { off = (int)b; b += al16(bytes); };
    put(lay.off_cfs, sizeof(double) * 2 * (size_t)lay.chrec * str);
    put(lay.off_ob, sizeof(double) * (size_t)site_1p_doubles(g, lay.nl));
    put(lay.off_recs, sizeof(MmaRec) * 2 * (size_t)lay.cap);
    put(lay.off_rcache, sizeof(RowC) * (size_t)G);
    put(lay.off_sblk, sizeof(BlockDesc) * (size_t)nblk);
    put(lay.off_mtab, sizeof(MaskWord) * (size_t)mask_rows(nblk, G, wx) * lay.mstr);
    put(lay.off_mraw, sizeof(unsigned) * (size_t)(2 * nblk + 2 * G) * lay.nseg);
    put(lay.off_jb, sizeof(int) * (size_t)nblk * lay.ncmax);
    put(lay.off_ncq, sizeof(int) * (size_t)lay.ncmax);
    put(lay.off_cand, sizeof(int4) * (size_t)lay.nseg * kSegCand);
    put(lay.off_cprefix, sizeof(int) * (lay.ncmax + 1));
    put(lay.off_srow, sizeof(int) * 2 * (size_t)nblk);
    put(lay.off_tot, sizeof(unsigned short) * (size_t)mask_rows(nblk, G, wx));
    put(lay.off_mbar, 16);
    put(lay.off_misc, 32);
    b = (b + 127) & ~(size_t)127;
    // staging of the values: with exchange windows one tile per warp; without, two CTA-level buffers of 8 whole rows
    lay.rowb = (int)((std::min<size_t>((size_t)(kMmaThreads / 32) * kSegCand, (size_t)lay.nseg * kSegCand) * 16 + 127) & ~(size_t)127);
    put(lay.off_stage, !bulk ? (size_t)(kMmaThreads / 32) * ((wx || !BS2E_MMA_HALF) ? kMmaStageBytes : kMmaHalfBytes) : (size_t)2 * 8 * lay.rowb);
    lay.bytes = b;
    return lay;
}

bool site_mma_usable(const bs2e_ctx* c, int nblk, int maxc, int maxrec, int lmax)
{
    const Geom& g = c->hg;
    if (site_kmax_for(g.K1) <= 0) return false;
    if (nblk > 255 || 2 * (2 * g.w + 1) > 255) return false;   // column group and n_c slot are bytes
    if (site_max_slots(g) > 65535) return false;
    return mma_layout(c, nblk, maxc, maxrec, true, lmax, false).bytes <= kMmaSmemLimit &&
           mma_layout(c, nblk, maxc, maxrec, false, lmax, true).bytes <= kMmaSmemLimit;
}

template <int KMAX, bool WX, bool BULK = false>
static void launch_one(bs2e_block* b, const MmaSmem& lay, cudaStream_t st, int first, int count)
{
    if (count <= 0) return;
    bs2e_ctx* c = b->ctx;
    auto kern = site_mma_kernel<KMAX, WX, BULK>;
    BS2E_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lay.bytes));
    kern<<<(unsigned)count, kMmaThreads, lay.bytes, st>>>(
        c->dg, b->dplan, c->one_body(), b->d_site_key, lay, first, c->d_R, b->d_Hptr, b->d_Sptr, b->d_Hidx,
        reinterpret_cast<double2*>(b->d_Hdat), b->d_Sidx, reinterpret_cast<double2*>(b->d_Sdat));
    BS2E_LAUNCHED();
}

void launch_site_mma(bs2e_block* b, cudaStream_t stX, cudaStream_t stD)
{
    bs2e_ctx* c = b->ctx;
    const int nblk = b->dplan.nblk;
    const int kmax = site_kmax_for(c->hg.K1);
    const char* store = getenv("BS2E_MMA_STORE");
    const bool bulk = store && strcmp(store, "bulk") == 0;
    const MmaSmem layX = mma_layout(c, nblk, b->ang->host.maxc, b->ang->host.maxrec, true, b->lmax, false);
    const MmaSmem layD = mma_layout(c, nblk, b->ang->host.maxc, b->ang->host.maxrec, false, b->lmax, bulk);
    if (layX.bytes > kMmaSmemLimit || layD.bytes > kMmaSmemLimit)
        throw Error("block_assemble: site tables exceed shared memory");
    const int nx = b->nsites_x, nd = b->nsites - b->nsites_x;
    switch (kmax) {
    case 7:
        launch_one<7, true>(b, layX, stX, 0, nx);
        if (bulk) launch_one<7, false, true>(b, layD, stD, nx, nd); else launch_one<7, false>(b, layD, stD, nx, nd);
        break;
    case 13:
        launch_one<13, true>(b, layX, stX, 0, nx);
        if (bulk) launch_one<13, false, true>(b, layD, stD, nx, nd); else launch_one<13, false>(b, layD, stD, nx, nd);
        break;
    case 21:
        launch_one<21, true>(b, layX, stX, 0, nx);
        if (bulk) launch_one<21, false, true>(b, layD, stD, nx, nd); else launch_one<21, false>(b, layD, stD, nx, nd);
        break;
    case 31:
        launch_one<31, true>(b, layX, stX, 0, nx);
        if (bulk) launch_one<31, false, true>(b, layD, stD, nx, nd); else launch_one<31, false>(b, layD, stD, nx, nd);
        break;
    default: throw Error("block_assemble: no site kernel for this max_k");
    }
}

}  // namespace bs2e
