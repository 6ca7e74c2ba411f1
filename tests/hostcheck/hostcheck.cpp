// hostcheck.cpp -- TEST INFRASTRUCTURE.  Executes the thread-level bodies of
// the CUDA kernels (core.h / slater_core.h, compiled here as plain C++) in
// serial loops on the CPU, so the index logic of the GPU path can be checked
// against the oracle in the GPU-less CI container.  Never linked into, loaded
// by, or shipped with the product library.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "../../b-spline-two-e_b200/csrc/core.h"
#include "../../b-spline-two-e_b200/csrc/geom_host.h"
#include "../../b-spline-two-e_b200/csrc/plan.h"
#include "../../b-spline-two-e_b200/csrc/slater_core.h"
#include "../../b-spline-two-e_b200/csrc/site_core.h"
#include "../../b-spline-two-e_b200/csrc/dip_plan.h"
#include "../../b-spline-two-e_b200/csrc/onebody_core.h"

using namespace bs2e;

static std::string g_err;

struct HcCtx {
    HostGeom hg;
    std::vector<double> mom_rk, mom_rmk, pre, sufx, rd, R;
    std::vector<RkRow> rkrow;
    std::vector<double> Hb, Sb;
    int lmax_1p = -1;
    std::vector<double> dipA, dipB;
    int dip_gauge = 0;
};

extern "C" {

const char* hc_last_error() { return g_err.c_str(); }

HcCtx* hc_create(int64_t ks, int64_t nt, const double* knots, int64_t max_k, int64_t kgl,
                 const double* glx, const double* glw)
{
    try {
        HcCtx* c = new HcCtx();
        build_host_geom(c->hg, (int)ks, (int)nt, knots, (int)max_k, (int)kgl, glx, glw);
        return c;
    } catch (const std::exception& e) {
        g_err = e.what();
        return nullptr;
    }
}
void hc_destroy(HcCtx* c) { delete c; }

void hc_sizes(HcCtx* c, int64_t* nb, int64_t* cells, int64_t* P, int64_t* ldP, int64_t* nnz4,
              int64_t* nnz6)
{
    *nb = c->hg.g.nb; *cells = c->hg.g.cells; *P = c->hg.g.P; *ldP = c->hg.g.ldP;
    *nnz4 = c->hg.nnz_4d; *nnz6 = c->hg.nnz_6d;
}

// emulates run_slater_cells: one "CTA" per cell, nthr threads, phases separated
// where the kernel has __syncthreads()
void hc_slater_cells(HcCtx* c, int64_t nthr_mom, int64_t nthr_diag, int64_t ksplit)
{
    const Geom& g = c->hg.g;
    const size_t ks2 = (size_t)g.ks * g.ks;
    c->mom_rk.assign((size_t)g.K1 * g.P * g.ks, 0.0);
    c->mom_rmk.assign((size_t)g.K1 * g.P * g.ks, 0.0);
    c->pre.assign((size_t)g.K1 * g.P * (g.ks + 1), -1.0);
    c->sufx.assign((size_t)g.K1 * g.P * (g.ks + 1), -1.0);
    c->rd.assign((size_t)g.cells * g.K1 * ks2 * ks2, -7.0);
    std::vector<double> sm(mom_smem_doubles(g));
    for (int v = 1; v <= g.cells; ++v) {
        MomSmem m = mom_smem_carve(g, sm.data());
        for (int t = 0; t < nthr_mom; ++t) mom_phase_tables(g, v, m, t, (int)nthr_mom);
        for (int t = 0; t < nthr_mom; ++t)
            mom_phase_integrate(g, v, m, t, (int)nthr_mom, c->mom_rk.data(), c->mom_rmk.data());
    }
    c->rkrow.assign((size_t)g.K1 * g.P, rk_row_empty());
    for (size_t idx = 0; idx < (size_t)g.K1 * g.P; ++idx)
        pair_prefix_item(g, idx, c->mom_rk.data(), c->mom_rmk.data(), c->pre.data(), c->sufx.data(), c->rkrow.data());
    std::vector<double> sd(diag_smem_doubles(g));
    for (int v = 1; v <= g.cells; ++v)
        for (int by = 0; by < ksplit; ++by) {
            DiagSmem d = diag_smem_carve(g, sd.data());
            for (int t = 0; t < nthr_diag; ++t) diag_phase_tables(g, v, d, t, (int)nthr_diag);
            for (int k = by; k < g.K1; k += (int)ksplit) {
                for (int t = 0; t < nthr_diag; ++t) diag_phase_powers(g, k, d, t, (int)nthr_diag);
                for (int t = 0; t < nthr_diag; ++t) diag_phase_inner(g, k, d, t, (int)nthr_diag);
                for (int t = 0; t < nthr_diag; ++t)
                    diag_phase_outer(g, v, k, d, t, (int)nthr_diag, c->rd.data());
            }
        }
}

void hc_get_moments(HcCtx* c, double* mom_rk, double* mom_rmk, double* rd)
{
    if (mom_rk) memcpy(mom_rk, c->mom_rk.data(), sizeof(double) * c->mom_rk.size());
    if (mom_rmk) memcpy(mom_rmk, c->mom_rmk.data(), sizeof(double) * c->mom_rmk.size());
    if (rd) memcpy(rd, c->rd.data(), sizeof(double) * c->rd.size());
}

// emulates run_rk_build with the same decomposition: band tiles (one thread per column of the band) and streaming
// tiles (one thread per column pair), every entry written exactly once
void hc_rk_build(HcCtx* c)
{
    const Geom& g = c->hg.g;
    const double unset = -3.0;
    c->R.assign((size_t)g.K1 * g.P * g.ldP, unset);
    CellData cd{c->mom_rk.data(), c->mom_rmk.data(), c->pre.data(), c->sufx.data(), c->rd.data()};
    const int row_lo = 0, row_hi = g.P;
    for (int k = 0; k < g.K1; ++k) {
        const RkRow* rk = c->rkrow.data() + (size_t)k * g.P;
        for (int p = 0; p < g.P; ++p) {   // the packed records are what rk_row_data derives from the prefix tables
            const RkRow a = rk[p], b = rk_row_data(g, cd, k, p);
            if (a.lo != b.lo || a.hi != b.hi || a.trk != b.trk || a.trmk != b.trmk) throw std::logic_error("RkRow mismatch");
        }
        std::vector<double> before;
        for (int r0 = row_lo; r0 < row_hi; r0 += kRkRowsB) {          // band role
            const int nrows = std::min(kRkRowsB, row_hi - r0);
            int c0, c1;
            rk_band_columns(g, r0, r0 + nrows, &c0, &c1);
            // even row tiles through the staged body (vectors handed in), odd ones through the direct one
            for (int p2 = c0; p2 < c1; ++p2) {
                if (((r0 - row_lo) / kRkRowsB) & 1) { rk_band_column(g, cd, c->R.data(), rk + r0, r0, nrows, k, p2, rk[p2]); continue; }
                const size_t rb = ((size_t)k * g.P + r0) * (g.ks + 1), cb = ((size_t)k * g.P + p2) * g.ks;
                rk_band_column_staged(g, cd, c->R.data(), rk + r0, r0, nrows, k, p2, rk[p2], c->pre.data() + rb,
                                      c->sufx.data() + rb, c->mom_rk.data() + cb, c->mom_rmk.data() + cb);
            }
        }
        const double* plane = c->R.data() + (size_t)k * g.P * g.ldP;
        before.assign(plane, plane + (size_t)g.P * g.ldP);
        for (int r0 = row_lo; r0 < row_hi; r0 += kRkRowsS) {          // streaming role
            const int nrows = std::min(kRkRowsS, row_hi - r0);
            for (int p2 = 0; p2 < g.ldP; p2 += 2) {
                const RkRow c0 = p2 < g.P ? rk[p2] : rk_row_empty(), c1 = p2 + 1 < g.P ? rk[p2 + 1] : rk_row_empty();
                rk_stream_columns(g, c->R.data(), rk + r0, r0, nrows, k, p2, c0, c1);
            }
        }
        for (size_t i = 0; i < before.size(); ++i) {   // the two roles write disjoint entries and together all of them
            if (before[i] != unset && plane[i] != before[i]) throw std::logic_error("stage B: entry written by both roles");
            if (plane[i] == unset) throw std::logic_error("stage B: entry not written");
        }
    }
}

void hc_get_R(HcCtx* c, double* R) { memcpy(R, c->R.data(), sizeof(double) * c->R.size()); }
// install an externally computed R (layout [K1][P][ldP]) for stage-C-only checks
void hc_set_R(HcCtx* c, const double* R)
{
    const Geom& g = c->hg.g;
    c->R.assign(R, R + (size_t)g.K1 * g.P * g.ldP);
}

int hc_set_one_particle(HcCtx* c, int64_t max_l_1p, const double* H_vec, const double* S)
{
    try {
        const Geom& g = c->hg.g;
        const size_t per_l = band_doubles(g);
        c->Hb.assign((size_t)(max_l_1p + 1) * per_l, 0.0);
        c->Sb.assign(per_l, 0.0);
        pack_band(g, S, c->Sb.data(), "S");
        for (int l = 0; l <= max_l_1p; ++l)
            pack_band(g, H_vec + (size_t)l * g.nb * g.nb * 2, c->Hb.data() + l * per_l, "H_vec");
        c->lmax_1p = (int)max_l_1p;
        return 0;
    } catch (const std::exception& e) {
        g_err = e.what();
        return 1;
    }
}

// emulates block_plan (count kernel + scan) for the union of the given row ranges
int hc_block_count(HcCtx* c, int64_t L, int64_t n_config, const int64_t* conf_n,
                   const int64_t* conf_l, int64_t full, int64_t n_ranges, const int64_t* range_lo,
                   const int64_t* range_hi, int64_t* H_ptr, int64_t* S_ptr)
{
    try {
        HostPlan hp = build_host_plan(c->hg.g, (int)L, n_config, conf_n, conf_l, (int)full, n_ranges, range_lo, range_hi);
        const Plan pl = hp.view();
        long long accH = 1, accS = 1;
        for (int idx = 0; idx < pl.nrows; ++idx) {
            long long h, s;
            row_count(c->hg.g, pl, row_of_local(pl.rr, idx), &h, &s);
            H_ptr[idx] = accH;
            S_ptr[idx] = accS;
            accH += h;
            accS += s;
        }
        H_ptr[pl.nrows] = accH;
        S_ptr[pl.nrows] = accS;
        return 0;
    } catch (const std::exception& e) {
        g_err = e.what();
        return 1;
    }
}

// emulates block_fill_kernel: one "warp" per row, 32 lanes, ballot emulated
int hc_block_fill(HcCtx* c, int64_t L, int64_t n_config, const int64_t* conf_n,
                  const int64_t* conf_l, int64_t full, int64_t n_ranges, const int64_t* range_lo,
                  const int64_t* range_hi, const int64_t* H_ptr, const int64_t* S_ptr, int64_t* H_idx, double* H_dat,
                  int64_t* S_idx, double* S_dat)
{
    try {
        const Geom& g = c->hg.g;
        HostPlan hp = build_host_plan(g, (int)L, n_config, conf_n, conf_l, (int)full, n_ranges, range_lo, range_hi);
        if (hp.lmax > c->lmax_1p) throw std::invalid_argument("l exceeds max_l_1p");
        const Plan pl = hp.view();
        const OneBody ob{c->Hb.data(), c->Sb.data()};
        const double* R = c->R.data();
        for (int wrow = 0; wrow < pl.nrows; ++wrow) {
            const long long i = row_of_local(pl.rr, wrow);
            const RowInfo r = row_info(pl, (int)i);
            long long hpos = H_ptr[wrow] - 1, spos = S_ptr[wrow] - 1;
            for_each_chunk(g, pl, r, [&](int bj, int nc, const Segment& s, const Coupling& cp, int base, int hi) {
                bool storeS[32];
                Element el[32];
                for (int lane = 0; lane < 32; ++lane) {
                    const int nd = base + lane;
                    const bool act = nd <= hi;
                    storeS[lane] = false;
                    if (act) {
                        const bool sup = nd >= s.dlo && nd <= s.dhi;
                        const bool sup_ex = nd >= s.xlo && nd <= s.xhi;
                        el[lane] = element_value(g, pl, ob, R, r, cp, bj, nc, nd, sup, sup_ex);
                        storeS[lane] = el[lane].storeS;
                        H_idx[hpos + lane] = (long long)s.jbase + nd;
                        H_dat[2 * (hpos + lane)] = el[lane].H.re;
                        H_dat[2 * (hpos + lane) + 1] = el[lane].H.im;
                    }
                }
                hpos += imin(32, hi - base + 1);
                int cnt = 0;
                for (int lane = 0; lane < 32; ++lane)
                    if (storeS[lane]) {
                        const long long pos = spos + cnt++;
                        S_idx[pos] = (long long)s.jbase + base + lane;
                        S_dat[2 * pos] = el[lane].S.re;
                        S_dat[2 * pos + 1] = el[lane].S.im;
                    }
                spos += cnt;
            });
            if (hpos != H_ptr[wrow + 1] - 1 || spos != S_ptr[wrow + 1] - 1)
                throw std::logic_error("fill wrote a different number of entries than counted, row " +
                                       std::to_string(i));
        }
        return 0;
    } catch (const std::exception& e) {
        g_err = e.what();
        return 1;
    }
}

}  // extern "C"

// emulates site_fill_kernel<NT,KMAX> phase by phase: one "CTA" of nthreads threads per
// radial site; warp scans / ballots become serial loops, everything lane-level comes
// from site_core.h.  group_rows > 0 overrides the rows-per-group of the launcher (to
// exercise the multi-group path on small bases).
template <int KMAX>
static void site_fill_emulate(HcCtx* c, const HostPlan& hp_,
                              int group_rows, int nthreads, const int64_t* H_ptr, const int64_t* S_ptr,
                              int64_t* H_idx, double* H_dat, int64_t* S_idx, double* S_dat)
{
    constexpr int NKP = (((KMAX + 1) / 2) + 1) & ~1;
    const Geom& g = c->hg.g;
    const Plan pl = hp_.view();
    if (pl.nkp != NKP || site_nkp(KMAX) != NKP) throw std::logic_error("packed factor stride mismatch");
    const OneBody ob{c->Hb.data(), c->Sb.data()};
    const double* R = c->R.data();
    const int K1 = g.K1, ncmax = site_max_nc(g), nblk = pl.nblk;
    const int G = group_rows > 0 ? std::min(group_rows, 32) : std::min(32, nblk);
    const int NT = nthreads;
    const size_t plane = (size_t)g.P * g.ldP;
    std::vector<SiteEntry> T((size_t)nblk * ncmax);
    std::vector<unsigned short> hp((size_t)nblk * kModes * (ncmax + 1)), sp((size_t)nblk * (ncmax + 1));
    struct Rec { long long hpos; int cf, meta; };
    std::vector<unsigned> pm((size_t)G * nblk);
    std::vector<int> gcnt((size_t)nblk * kModes);
    std::vector<Rec> rlist((size_t)nblk * G);
    std::vector<double> cfs((size_t)G * nblk * 2 * NKP);
    std::vector<int> cprefix(ncmax + 1);
    struct RowC { long long hbase, sbase; int bi, la, lb; };
    std::vector<RowC> rcache(G);
    long long* Hi = reinterpret_cast<long long*>(H_idx);
    long long* Si = reinterpret_cast<long long*>(S_idx);
    const int nsites = (int)hp_.site_key.size();
    long long rows_seen = 0;
    for (int sidx = 0; sidx < nsites; ++sidx) {
        const unsigned key = (unsigned)(hp_.site_key[sidx] & 0xffffffffull);
        const bool wantX = site_wants_X(g, pl.max_nd, (int)(key >> 16));
        if (wantX != (sidx < hp_.nsites_x)) throw std::logic_error("site filed in the wrong launch class");
        const Site s = make_site(g, (int)(key >> 16), (int)(key & 0xffffu), wantX);
        const int nnc = s.nnc;
        if (nnc > ncmax) throw std::logic_error("site exceeds smem bounds");
        // phase 1
        const int nl = c->lmax_1p + 1;
        std::vector<double> ob_s(site_1p_doubles(g, nl));
        for (int idx = 0; idx < site_1p_doubles(g, nl) / 2; ++idx) {
            const Cplx v = site_1p_source(g, ob, s, nl, idx);
            ob_s[2 * idx] = v.re;
            ob_s[2 * idx + 1] = v.im;
        }
        const SiteOneBody so{ob_s.data(), ob_s.data() + (size_t)nl * 2 * (2 * g.w + 1) * 2};
        std::fill(cprefix.begin(), cprefix.end(), -12345);
        if (wantX) {
            int run = 0;
            for (int q = 0; q < nnc; ++q) { cprefix[q] = run; run += site_cand_DX_count(s, q); }
            cprefix[nnc] = run;
            if (run > site_max_slots(g)) throw std::logic_error("candidate list exceeds its bound");
        }
        for (int bj = 0; bj < nblk; ++bj)
            for (int q = 0; q < nnc; ++q) T[bj * ncmax + q] = site_entry(g, pl, s, bj, q);
        // phase 2
        std::fill(hp.begin(), hp.end(), (unsigned short)0xdead);
        for (int task = 0; task < nblk * kModes; ++task) {
            const int bj = task / kModes, mode = task - bj * kModes;
            if (!wantX && (mode == kModeX || mode == kModeDX)) continue;
            const bool useD = mode_useD(mode), useX = mode_useX(mode), diag = mode == kModeDiag;
            const bool samex = diag && pl.blk[bj].l1 == pl.blk[bj].l2;
            int hrun = 0, srun = 0;
            for (int q = 0; q < nnc; ++q) {
                SiteEntry e = T[bj * ncmax + q];
                if (diag && !pl.full) e = entry_cut(e, s, site_nc(s, q));
                hp[(size_t)task * (ncmax + 1) + q] = (unsigned short)hrun;
                hrun += entry_count(e, useD, useX);
                if (diag) {
                    sp[(size_t)bj * (ncmax + 1) + q] = (unsigned short)srun;
                    srun += entry_count(e, true, samex);
                }
            }
            hp[(size_t)task * (ncmax + 1) + nnc] = (unsigned short)hrun;
            if (diag) sp[(size_t)bj * (ncmax + 1) + nnc] = (unsigned short)srun;   // site_count_kernel
        }
        // phase 0: the rows of the site, one per (l1,l2) group that holds (n_a,n_b) among the planned rows
        std::vector<int> srows;
        for (int bi = 0; bi < nblk; ++bi) {
            const int row = config_index(g, pl, bi, s.na, s.nb);
            if (row > 0 && row_local_of(pl.rr, row) >= 0) srows.push_back(row);
        }
        const int nr = (int)srows.size();
        if (nr != 1023 - (int)((hp_.site_key[sidx] >> 32) & 1023)) throw std::logic_error("site key row count");
        for (int g0 = 0; g0 < nr; g0 += G) {
            const int gr = std::min(G, nr - g0);
            // phase 3a
            for (int ri = 0; ri < gr; ++ri) {
                const int rowi = srows[g0 + ri];
                ++rows_seen;
                const RowInfo r = row_info(pl, rowi);
                if (r.na != s.na || r.nb != s.nb) throw std::logic_error("row filed under the wrong site");
                const long long wrow = row_local_of(pl.rr, rowi);
                if (wrow < 0 || row_of_local(pl.rr, (int)wrow) != rowi) throw std::logic_error("row map inconsistent");
                rcache[ri] = RowC{H_ptr[wrow] - 1, S_ptr[wrow] - 1, r.bi, r.la, r.lb};
                int run = 0, srun = 0;
                for (int bj = 0; bj < nblk; ++bj) {
                    int cnt = 0, mode = pair_mode(pl, r, bj);
                    if (mode >= 0) {
                        const unsigned short* hb = hp.data() + (bj * kModes) * (ncmax + 1) + nnc;
                        mode = effective_mode(mode, hb[kModeD * (ncmax + 1)], wantX ? hb[kModeX * (ncmax + 1)] : 0);
                        cnt = (wantX || mode != kModeX) ? hb[mode * (ncmax + 1)] : 0;
                        if (mode == kModeDiag) srun += sp[(size_t)bj * (ncmax + 1) + nnc];
                    }
                    pm[ri * nblk + bj] = cnt > 0 ? pm_pack(run, mode, (r.la + pl.blk[bj].l1) & 1) : 0u;
                    run += cnt;
                }
                // the site-table count (site_count_kernel) must agree with the row-by-row count
                if (run != H_ptr[wrow + 1] - H_ptr[wrow] || srun != S_ptr[wrow + 1] - S_ptr[wrow])
                    throw std::logic_error("site fill: row " + std::to_string(rowi) + " counted differently");
            }
            // phase 3b: the pairs of each column block, filed by storage mode
            for (int bj = 0; bj < nblk; ++bj) {
                int cnt[kModes] = {0, 0, 0, 0};
                for (int ri = 0; ri < gr; ++ri) {
                    const unsigned v = pm[ri * nblk + bj];
                    if (pm_valid(v)) ++cnt[pm_mode(v)];
                }
                for (int mode = 0; mode < kModes; ++mode) gcnt[bj * kModes + mode] = cnt[mode];
                if (cnt[kModeDiag] > 1) throw std::logic_error("diagonal pair with several rows");
                int pos[kModes] = {0, cnt[0], cnt[0] + cnt[1], cnt[0] + cnt[1] + cnt[2]};
                for (int ri = 0; ri < gr; ++ri) {
                    const unsigned v = pm[ri * nblk + bj];
                    if (!pm_valid(v)) continue;
                    const int where = pos[pm_mode(v)]++;
                    const int px = pm_pd(v) ^ ((pl.blk[bj].l1 + pl.blk[bj].l2) & 1);
                    rlist[bj * G + where] = Rec{rcache[ri].hbase + pm_off(v), ri * nblk + bj, pm_pd(v) | (px << 1) | (ri << 8)};
                }
            }
            // packed factors of the group's pairs ("shared memory" copy)
            std::fill(cfs.begin(), cfs.end(), std::nan(""));
            for (int pair = 0; pair < gr * nblk; ++pair) {
                if (!pm_valid(pm[pair])) continue;
                const int ri = pair / nblk, bj = pair - ri * nblk;
                const double* src = pl.angP + ((size_t)rcache[ri].bi * nblk + bj) * (2 * NKP);
                std::copy(src, src + 2 * NKP, cfs.begin() + (size_t)pair * 2 * NKP);
            }
            // phase 4: thread t0+tid owns candidate column t
            const int nc_all = site_num_cand(s, cprefix.data(), wantX);
            for (int t0 = 0; t0 < nc_all; t0 += NT)
                for (int tid = 0; tid < NT; ++tid) {
                    const int t = t0 + tid;
                    if (t >= nc_all) continue;
                    const OwnCand cd = site_own_cand(g, s, cprefix.data(), wantX, t);
                    double Rd[KMAX], Rx[KMAX];
                    for (int k = 0; k < KMAX; ++k) {
                        Rd[k] = (cd.inD && k < K1) ? R[(size_t)k * plane + (size_t)cd.rowD * g.ldP + cd.colD] : 0.0;
                        Rx[k] = (cd.inX && k < K1) ? R[(size_t)k * plane + (size_t)cd.rowX * g.ldP + cd.colX] : 0.0;
                    }
                    for (int bj = 0; bj < nblk; ++bj) {
                        const SiteEntry e = T[bj * ncmax + cd.q];
                        const Rec* rl = rlist.data() + bj * G;
                        for (int mode = 0; mode < kModes; ++mode) {
                            const int nrow = gcnt[bj * kModes + mode];
                            if (!nrow) continue;
                            const Rec* rm = rl;
                            rl += nrow;
                            const bool diag = mode == kModeDiag;
                            const ModeSlot ms = site_mode_slot(s, cd, e, hp.data() + (bj * kModes + mode) * (ncmax + 1),
                                                               mode, diag && !pl.full);
                            if (!(ms.sup || ms.sup_ex)) continue;
                            for (int i = 0; i < nrow; ++i) {
                                const Rec rec = rm[i];
                                const RowC rc = rcache[rec.meta >> 8];
                                const int pd = rec.meta & 1, px = (rec.meta >> 1) & 1;
                                const double* cf = cfs.data() + (size_t)rec.cf * (2 * NKP);
                                double res = 0.0;
                                if (mode != kModeX) {
                                    const double d = site_dot_par<KMAX>(cf, Rd, pd);
                                    res += ms.sup ? d : 0.0;
                                }
                                if (mode != kModeD) {
                                    const double x = site_dot_par<KMAX>(cf + NKP, Rx, px);
                                    res += ms.sup_ex ? x : 0.0;
                                }
                                double re = res, im = 0.0;
                                if (diag) {
                                    if (rc.bi != bj) throw std::logic_error("diagonal pair on a foreign block");
                                    site_diag_terms(g, pl, so, s, rc.la, rc.lb, cd, ms, rc.la == rc.lb,
                                                    sp.data() + bj * (ncmax + 1), rc.sbase, &re, &im, Si, S_dat);
                                }
                                const long long pos = rec.hpos + ms.rank;
                                Hi[pos] = ms.jcol;
                                H_dat[2 * pos] = re;
                                H_dat[2 * pos + 1] = im;
                            }
                        }
                    }
                }
        }
    }
    if (rows_seen != pl.nrows) throw std::logic_error("site list does not cover the rows");
}


// emulates site_mma_kernel<KMAX,WX> (csrc/site_mma.cu) phase by phase: one "CTA" of nthreads threads per
// radial site, warps = nthreads/32; atomics, ballots and warp scans become serial loops; the 8x8x4 tensor
// instruction becomes its measured arithmetic (ascending-k FMA chain per output element); the staging of the
// factor rows is a plain copy.  Everything lane-level comes from site_core.h.
template <int KMAX>
static void site_mma_emulate(HcCtx* c, const HostPlan& hp_, int group_rows, int nthreads, int chunk_rec,
                             const int64_t* H_ptr, const int64_t* S_ptr, int64_t* H_idx, double* H_dat,
                             int64_t* S_idx, double* S_dat)
{
    constexpr int NKH = (KMAX + 1) / 2, NKP = ((NKH + 1) & ~1), KS = (NKH + 3) / 4, CT = kSegCand / 8;
    const Geom& g = c->hg.g;
    const Plan pl = hp_.view();
    if (pl.nkp != NKP) throw std::logic_error("packed factor stride mismatch");
    const OneBody ob{c->Hb.data(), c->Sb.data()};
    const double* R = c->R.data();
    const int K1 = g.K1, nblk = pl.nblk;
    const int NT = nthreads, NW = std::max(1, NT / 32);
    const size_t plane = (size_t)g.P * g.ldP;
    const int maxc = hp_.ang.maxc;
    const int per_row = std::max(1, std::min(maxc, nblk));
    const int G = group_rows > 0 ? std::min(group_rows, nblk) : std::min(nblk, 255);
    const int capP = ((G * per_row + 7) & ~7) + ((G + 7) & ~7) + 8;
    const int CH = std::max(8, chunk_rec & ~7);
    long long* Hi = reinterpret_cast<long long*>(H_idx);
    long long* Si = reinterpret_cast<long long*>(S_idx);
    const int nsites = (int)hp_.site_key.size();
    long long rows_seen = 0;
    struct RowC { long long hbase, sbase; int bi, la, lb; };
    for (int sidx = 0; sidx < nsites; ++sidx) {
        const unsigned key = (unsigned)(hp_.site_key[sidx] & 0xffffffffull);
        const bool WX = site_wants_X(g, pl.max_nd, (int)(key >> 16));
        if (WX != (sidx < hp_.nsites_x)) throw std::logic_error("site filed in the wrong launch class");
        const Site s = make_site(g, (int)(key >> 16), (int)(key & 0xffffu), WX);
        const int nnc = s.nnc;
        const int ncmax = WX ? site_max_nc(g) : 2 * g.w + 1;
        const int nsegS = ((WX ? site_max_slots(g) : (2 * g.w + 1) * (2 * g.w + 1)) + kSegCand - 1) / kSegCand;
        const int CFS = WX ? 2 * NKP : NKP, STR = mma_cf_stride(CFS);
        if (nnc > ncmax) throw std::logic_error("site exceeds smem bounds");
        // phase 0
        std::vector<int> srow_bi, srow_local;
        for (int bi = 0; bi < nblk; ++bi) {
            const int row = config_index(g, pl, bi, s.na, s.nb);
            const int local = row > 0 ? row_local_of(pl.rr, row) : -1;
            if (local >= 0) { srow_bi.push_back(bi); srow_local.push_back(local); }
        }
        const int nr = (int)srow_bi.size();
        if (nr != 1023 - (int)((hp_.site_key[sidx] >> 32) & 1023)) throw std::logic_error("site key row count");
        std::vector<int> cprefix(ncmax + 1, -12345);
        if (WX) {
            int run = 0;
            for (int q = 0; q < nnc; ++q) { cprefix[q] = run; run += site_cand_DX_count(s, q); }
            cprefix[nnc] = run;
        }
        const int nl = c->lmax_1p + 1;
        std::vector<double> ob_s(site_1p_doubles(g, nl));
        for (int idx = 0; idx < site_1p_doubles(g, nl) / 2; ++idx) {
            const Cplx v = site_1p_source(g, ob, s, nl, idx);
            ob_s[2 * idx] = v.re;
            ob_s[2 * idx + 1] = v.im;
        }
        const SiteOneBody so{ob_s.data(), ob_s.data() + (size_t)nl * 2 * (2 * g.w + 1) * 2};
        const int pi = (pl.blk[0].l1 + pl.blk[0].l2) & 1;
        const int nc_all = site_num_cand(s, cprefix.data(), WX);
        const int nseg = (nc_all + kSegCand - 1) / kSegCand;
        if (nseg > nsegS || nc_all > site_max_slots(g)) throw std::logic_error("candidate list exceeds its bound");
        // phase 1
        std::vector<SiteEntry> T((size_t)nblk * ncmax);
        std::vector<int> jb((size_t)nblk * ncmax, -999999);
        std::vector<unsigned> mD((size_t)nblk * nsegS, 0u), mX((size_t)nblk * nsegS, 0u);
        for (int idx = 0; idx < nblk * nnc; ++idx) {
            const int bj = idx / nnc, q = idx - bj * nnc;
            const SiteEntry e = site_entry(g, pl, s, bj, q);
            T[bj * ncmax + q] = e;
            jb[bj * ncmax + q] = e.jbase;
            const CandSlot cs = cand_slot(s, cprefix.data(), WX, q);
            entry_cand_ranges(
                cs, e,
                [&](int t0, int t1) { mask_range(t0, t1, [&](int seg, unsigned bits) { mD[bj * nsegS + seg] |= bits; }); },
                [&](int t0, int t1) { mask_range(t0, t1, [&](int seg, unsigned bits) { mX[bj * nsegS + seg] |= bits; }); });
        }
        // phase 2
        const int nrows_tab = mask_rows(nblk, G, WX);
        std::vector<MaskWord> mtab((size_t)nrows_tab * nsegS, MaskWord{0xdeadbeefu, 0xdeadu});
        std::vector<unsigned short> tot(nrows_tab, 0xdead);
        for (int task = 0; task < nblk * mask_modes(WX) + 1; ++task) {
            if (task == nblk * mask_modes(WX)) {
                const int row = mask_row_zero(nblk, G, WX);
                for (int seg = 0; seg < nseg; ++seg) mtab[row * nsegS + seg] = MaskWord{0u, 0u};
                tot[row] = 0;
                continue;
            }
            const int bj = WX ? task / 3 : task, mode = WX ? task - bj * 3 : kModeD;
            unsigned run = 0;
            for (int seg = 0; seg < nseg; ++seg) {
                const unsigned d = mD[bj * nsegS + seg], x = mX[bj * nsegS + seg];
                const unsigned m = mode == kModeD ? d : (mode == kModeX ? x : (d | x));
                mtab[task * nsegS + seg] = MaskWord{m, run};
                run += popc32(m);
            }
            tot[task] = (unsigned short)run;
        }
        for (int g0 = 0; g0 < nr; g0 += G) {
            const int gr = std::min(G, nr - g0);
            // phase 3a
            std::vector<RowC> rcache(G);
            for (int ri = 0; ri < gr; ++ri) {
                const int bi = srow_bi[g0 + ri];
                const long long wrow = srow_local[g0 + ri];
                rcache[ri] = RowC{H_ptr[wrow] - 1, S_ptr[wrow] - 1, bi, pl.blk[bi].l1, pl.blk[bi].l2};
                ++rows_seen;
            }
            std::vector<unsigned> mgH((size_t)G * nsegS, 0u), mgS((size_t)G * nsegS, 0u);
            for (int idx = 0; idx < gr * nnc; ++idx) {
                const int ri = idx / nnc, q = idx - ri * nnc;
                const RowC rc = rcache[ri];
                SiteEntry e = T[rc.bi * ncmax + q];
                if (!pl.full) e = entry_cut(e, s, site_nc(s, q));
                const bool samex = rc.la == rc.lb;
                const CandSlot cs = cand_slot(s, cprefix.data(), WX, q);
                entry_cand_ranges(
                    cs, e,
                    [&](int t0, int t1) {
                        mask_range(t0, t1, [&](int seg, unsigned bits) { mgH[ri * nsegS + seg] |= bits; mgS[ri * nsegS + seg] |= bits; });
                    },
                    [&](int t0, int t1) {
                        mask_range(t0, t1, [&](int seg, unsigned bits) {
                            mgH[ri * nsegS + seg] |= bits;
                            if (samex) mgS[ri * nsegS + seg] |= bits;
                        });
                    });
            }
            for (int task = 0; task < 2 * gr; ++task) {
                const int ri = task >> 1, isS = task & 1;
                const unsigned* src = (isS ? mgS.data() : mgH.data()) + ri * nsegS;
                const int row = isS ? mask_row_diagS(nblk, G, ri, WX) : mask_row_diagH(nblk, ri, WX);
                unsigned run = 0;
                for (int seg = 0; seg < nseg; ++seg) {
                    mtab[row * nsegS + seg] = MaskWord{src[seg], run};
                    run += popc32(src[seg]);
                }
                tot[row] = (unsigned short)run;
            }
            // phase 3b
            std::vector<MmaRec> recs((size_t)2 * capP, MmaRec{-777, -5, 0xffff, 0xff, 0xff});
            int cnts[3] = {0, 0, 0};
            for (int ri = 0; ri < gr; ++ri) {
                const RowC rc = rcache[ri];
                int run = 0, srun = 0;
                for (int bj = 0; bj < nblk; ++bj) {
                    int cnt = 0, row = 0;
                    if (pl.full || bj >= rc.bi) {
                        if (bj == rc.bi) {
                            row = mask_row_diagH(nblk, ri, WX);
                            cnt = tot[row];
                            srun += tot[mask_row_diagS(nblk, G, ri, WX)];
                        } else {
                            const unsigned f = pl.flags[(size_t)rc.bi * nblk + bj];
                            int mode = (f & kDirAny) ? ((f & kExAny) ? kModeDX : kModeD) : ((f & kExAny) ? kModeX : -1);
                            if (!WX && mode == kModeDX) mode = kModeD;
                            if (mode >= 0 && (WX || mode == kModeD)) {
                                row = mask_row_pair(bj, mode, WX);
                                cnt = tot[row];
                            }
                        }
                    }
                    if (cnt > 0) {
                        const bool diag = bj == rc.bi;
                        const int par = diag ? 2 : ((rc.la + pl.blk[bj].l1) & 1);
                        const MmaRec rec{rc.hbase + run, rc.bi * nblk + bj, (unsigned short)row, (unsigned char)bj, (unsigned char)ri};
                        const int at = cnts[par]++;
                        if (par != 2 && at >= capP - ((G + 7) & ~7) - 8) throw std::logic_error("record list overflow");
                        recs[par == 0 ? at : (par == 1 ? capP + at : capP - 1 - at)] = rec;
                    }
                    run += cnt;
                }
                // the mask-table counts must agree with the count pass (site_count_kernel / row_count)
                const long long wrow = srow_local[g0 + ri];
                if (run != H_ptr[wrow + 1] - H_ptr[wrow] || srun != S_ptr[wrow + 1] - S_ptr[wrow])
                    throw std::logic_error("site fill: row of group " + std::to_string(rc.bi) + " counted differently");
            }
            // phase 3c
            const int n0 = cnts[0], n1 = cnts[1], ndg = cnts[2];
            const int n0p = (n0 + 7) & ~7, dg0 = n0p, n0all = n0p + ((ndg + 7) & ~7), n1p = (n1 + 7) & ~7;
            if (n0all > capP || n1p > capP || dg0 + ndg > capP - ndg) throw std::logic_error("record list overflow");
            {
                const MmaRec padrec{0, -1, (unsigned short)mask_row_zero(nblk, G, WX), 0, 0};
                std::vector<MmaRec> mine(ndg);
                for (int t = 0; t < ndg; ++t) mine[t] = recs[capP - 1 - t];
                for (int t = 0; t < ndg; ++t) recs[dg0 + t] = mine[t];
                for (int i = n0; i < n0p; ++i) recs[i] = padrec;
                for (int i = dg0 + ndg; i < n0all; ++i) recs[i] = padrec;
                for (int i = n1; i < n1p; ++i) recs[capP + i] = padrec;
            }
            // phase 4
            const int nch0 = (n0all + CH - 1) / CH, nch1 = (n1p + CH - 1) / CH, nchd = nch0 + nch1;
            const int npass = nchd > 0 ? (nseg + NW - 1) / NW : 0;
            std::vector<double> cbuf((size_t)CH * STR);
            for (int pass = 0; pass < npass; ++pass)
                for (int warp = 0; warp < NW; ++warp) {
                    const int seg = pass * NW + warp;
                    if (seg >= nseg) continue;
                    const int nin = std::min(kSegCand, nc_all - seg * kSegCand);
                    for (int par = 0; par < 2; ++par) {
                        const int nchp = par ? nch1 : nch0;
                        // A fragments as a dense table [ct][candidate of the tile][multipole slot]
                        double Ad[CT][8][4 * KS], Ax[CT][8][4 * KS];
                        int qn[CT][8];
                        for (int ct = 0; ct < CT; ++ct)
                            for (int m = 0; m < 8; ++m) {
                                const int t = seg * kSegCand + ct * 8 + m;
                                const bool ok = t < nc_all;
                                const OwnCand cd = site_own_cand(g, s, cprefix.data(), WX, ok ? t : 0);
                                qn[ct][m] = cd.q | (cd.nd << 8);
                                for (int i = 0; i < 4 * KS; ++i) {
                                    const int k = 2 * i + par, kx = 2 * i + (par ^ pi);
                                    Ad[ct][m][i] = (ok && cd.inD && k < K1) ? R[(size_t)k * plane + (size_t)cd.rowD * g.ldP + cd.colD] : 0.0;
                                    Ax[ct][m][i] = (WX && ok && cd.inX && kx < K1) ? R[(size_t)kx * plane + (size_t)cd.rowX * g.ldP + cd.colX] : 0.0;
                                }
                            }
                        for (int cc = 0; cc < nchp; ++cc) {
                            const int first = cc * CH;
                            const int count = std::min(CH, (par ? n1p : n0all) - first);
                            const MmaRec* rl = recs.data() + (par ? capP : 0) + first;
                            std::fill(cbuf.begin(), cbuf.end(), std::nan(""));
                            int real = 0;
                            for (int i = 0; i < count; ++i) {
                                if (rl[i].cf < 0) continue;
                                ++real;
                                const double* src = pl.angP + (size_t)rl[i].cf * (2 * NKP);
                                std::copy(src, src + CFS, cbuf.begin() + (size_t)i * STR);
                            }
                            auto overlap = [](int a0, int a1, int b0, int b1) { return imax(0, imin(a1, b1) - imax(a0, b0)); };
                            const int real2 = par ? overlap(first, first + count, 0, n1)
                                                  : overlap(first, first + count, 0, n0) + overlap(first, first + count, dg0, dg0 + ndg);
                            if (real != real2 || real == 0) throw std::logic_error("staged byte count of a chunk");
                            for (int rt = 0; rt < count; rt += 8) {
                                const bool diag = !par && first + rt >= dg0;
                                for (int n = 0; n < 8; ++n) {          // record of the tile
                                    const MmaRec rr = rl[rt + n];
                                    const MaskWord wh = mtab[rr.tbl * nsegS + seg];
                                    const double* cf = cbuf.data() + (size_t)(rt + n) * STR;
                                    for (int ct = 0; ct < CT; ++ct) {
                                        if (ct * 8 >= nin) break;
                                        for (int m = 0; m < 8; ++m) {  // candidate of the tile
                                            const int bit = ct * 8 + m;
                                            if (!mask_has(wh, bit)) continue;
                                            if (rr.cf < 0) throw std::logic_error("padding record stores");
                                            double c0 = 0.0;
                                            for (int i = 0; i < 4 * KS; ++i) c0 += cf[i] * Ad[ct][m][i];   // on the device: one FMA chain, like the FMA kernels
                                            if (WX) {
                                                double x0 = 0.0;
                                                for (int i = 0; i < 4 * KS; ++i) x0 += cf[NKP + i] * Ax[ct][m][i];
                                                c0 += x0;
                                            }
                                            const int q = qn[ct][m] & 0xff, nd = qn[ct][m] >> 8;
                                            const int jcol = jb[rr.bj * ncmax + q] + nd;
                                            double re = c0, im = 0.0;
                                            if (diag) {
                                                const RowC rc = rcache[rr.ri];
                                                if (rc.bi != rr.bj) throw std::logic_error("diagonal pair on a foreign block");
                                                const MaskWord ws = mtab[(rr.tbl + G) * nsegS + seg];
                                                if (mask_has(ws, bit)) {
                                                    Cplx h, sv;
                                                    site_onebody(g, pl, so, s, rc.la, rc.lb, site_nc(s, q), nd, rc.la == rc.lb, &h, &sv);
                                                    re += h.re;
                                                    im += h.im;
                                                    const long long spos = rc.sbase + mask_rank(ws, bit);
                                                    Si[spos] = jcol;
                                                    S_dat[2 * spos] = sv.re;
                                                    S_dat[2 * spos + 1] = sv.im;
                                                }
                                            }
                                            const long long pos = rr.hpos + mask_rank(wh, bit);
                                            Hi[pos] = jcol;
                                            H_dat[2 * pos] = re;
                                            H_dat[2 * pos + 1] = im;
                                        }
                                    }
                                }
                            }
                        }
                    }
                }
        }
    }
    if (rows_seen != pl.nrows) throw std::logic_error("site list does not cover the rows");
}

extern "C" {

int hc_site_fill(HcCtx* c, int64_t L, int64_t n_config, const int64_t* conf_n,
                 const int64_t* conf_l, int64_t full, int64_t n_ranges, const int64_t* range_lo,
                 const int64_t* range_hi, const int64_t* H_ptr, const int64_t* S_ptr, int64_t* H_idx, double* H_dat,
                 int64_t* S_idx, double* S_dat, int64_t group_rows, int64_t nthreads, int64_t mma, int64_t chunk_rec)
{
    try {
        const Geom& g = c->hg.g;
        HostPlan hp_ = build_host_plan(g, (int)L, n_config, conf_n, conf_l, (int)full, n_ranges, range_lo, range_hi);
        if (hp_.lmax > c->lmax_1p) throw std::invalid_argument("l exceeds max_l_1p");
        if (hp_.site_key.empty()) throw std::logic_error("no site list");
#define HC_SITE(KM)                                                                                  \
    case KM:                                                                                         \
        if (mma)                                                                                     \
            site_mma_emulate<KM>(c, hp_, (int)group_rows, (int)nthreads, (int)chunk_rec, H_ptr, S_ptr, \
                                 H_idx, H_dat, S_idx, S_dat);                                        \
        else                                                                                         \
            site_fill_emulate<KM>(c, hp_, (int)group_rows, (int)nthreads, H_ptr, S_ptr,              \
                                  H_idx, H_dat, S_idx, S_dat);                                       \
        break;
        switch (site_kmax_for(g.K1)) {
            HC_SITE(7) HC_SITE(13) HC_SITE(21) HC_SITE(31)
        default: throw std::logic_error("no site kernel instantiation for this max_k");
        }
#undef HC_SITE
        return 0;
    } catch (const std::exception& e) {
        g_err = e.what();
        return 1;
    }
}

}  // extern "C"

// ---- dipole blocks: emulates dip_count_kernel / dip_fill_kernel (dip.cu) ----
extern "C" {

int hc_set_radial_dipole(HcCtx* c, int64_t gauge, const double* A, const double* B)
{
    try {
        const Geom& g = c->hg.g;
        c->dipA.assign(band_doubles(g), 0.0);
        c->dipB.assign(band_doubles(g), 0.0);
        pack_band(g, A, c->dipA.data(), "A");
        if (gauge == 'v') pack_band(g, B, c->dipB.data(), "B");
        c->dip_gauge = (int)gauge;
        return 0;
    } catch (const std::exception& e) {
        g_err = e.what();
        return 1;
    }
}

// index_ptr == NULL: count only; returns nnz or -1
int64_t hc_dip_block(HcCtx* c, int64_t q, const int64_t* sym1, int64_t n1, const int64_t* conf_n1,
                     const int64_t* conf_l1, const int64_t* sym2, int64_t n2, const int64_t* conf_n2,
                     const int64_t* conf_l2, int64_t compute, int64_t* index_ptr, int64_t* indices, double* data)
{
    try {
        const Geom& g = c->hg.g;
        HostDipPlan hp = build_dip_plan(g, c->dip_gauge, (int)q, sym1, n1, conf_n1, conf_l1, sym2, n2, conf_n2,
                                        conf_l2, compute != 0);
        if (hp.empty) return 0;
        Plan plC = hp.cols.view();
        if (plC.full != 1) throw std::logic_error("column plan must be built without the triangle cut");
        const DipTables dt = hp.tables();
        const DipBand bd{c->dipA.data(), c->dipB.data(), c->Sb.data()};
        long long nnz = 0;
        std::vector<long long> ptr(n1 + 1);
        for (long long i = 0; i < n1; ++i) {     // count kernel + scan
            ptr[i] = nnz + 1;
            nnz += dip_row_count(g, plC, dt, (int)i + 1);
        }
        ptr[n1] = nnz + 1;
        if (!index_ptr) return nnz;
        for (long long i = 0; i <= n1; ++i) index_ptr[i] = ptr[i];
        for (long long wrow = 0; wrow < n1; ++wrow) {   // fill kernel: one "warp" per row
            const RowInfo r = dip_row(dt, (int)wrow + 1);
            long long pos = ptr[wrow] - 1;
            dip_for_each_chunk(g, plC, dt, r, [&](int bj, int nc, const Segment& s, int base, int hi) {
                for (int lane = 0; lane < 32; ++lane) {
                    const int nd = base + lane;
                    if (nd > hi) continue;
                    const Cplx v = dip_value(g, dt, bd, r, bj, nc, nd);
                    indices[pos + lane] = (long long)s.jbase + nd;
                    data[2 * (pos + lane)] = v.re;
                    data[2 * (pos + lane) + 1] = v.im;
                }
                pos += imin(32, hi - base + 1);
            });
            if (pos != ptr[wrow + 1] - 1) throw std::logic_error("dipole fill wrote a different number of entries than counted");
        }
        return nnz;
    } catch (const std::exception& e) {
        g_err = e.what();
        return -1;
    }
}

}  // extern "C"

// ---- one-particle matrices and radial dipole integrals: emulates one_body_kernel (onebody.cu), one
//      "thread" per band entry; output as dense complex n_b x n_b column-major matrices ----
extern "C" int hc_one_body(HcCtx* c, int64_t Z, int64_t lmax, int64_t cap_order, double cap_r0, double eta_re,
                           double eta_im, int64_t want_1p, int64_t gauge, double* H_vec, double* S, double* A, double* B)
{
    try {
        const Geom& g = c->hg.g;
        const size_t per = band_doubles(g);
        std::vector<double> Sb(per, 0.0), Hb((size_t)(lmax + 1) * per, 0.0), Ab(per, 0.0), Bb(per, 0.0);
        const OneBodyParams p{(int)Z, (int)lmax, (int)cap_order, cap_r0, eta_re, eta_im, (int)want_1p, (int)gauge};
        const OneBodyOut o{Sb.data(), Hb.data(), Ab.data(), Bb.data()};
        const int bw = 2 * g.w + 1;
        for (int idx = 0; idx < g.nb * bw; ++idx) one_body_entry(g, p, o, idx / bw + 1, idx % bw);
        auto dense = [&](const double* band, double* out) {
            if (!out) return;
            for (size_t q = 0; q < (size_t)g.nb * g.nb * 2; ++q) out[q] = 0.0;
            for (int n = 1; n <= g.nb; ++n)
                for (int d = 0; d < bw; ++d) {
                    const int np = n + d - g.w;
                    if (np < 1 || np > g.nb) continue;
                    const size_t at = 2 * ((size_t)(n - 1) + (size_t)g.nb * (np - 1));
                    out[at] = band[((size_t)n * bw + d) * 2];
                    out[at + 1] = band[((size_t)n * bw + d) * 2 + 1];
                }
        };
        dense(Sb.data(), S);
        if (H_vec) for (int l = 0; l <= lmax; ++l) dense(Hb.data() + l * per, H_vec + (size_t)l * g.nb * g.nb * 2);
        dense(Ab.data(), A);
        dense(Bb.data(), B);
        return 0;
    } catch (const std::exception& e) {
        g_err = e.what();
        return 1;
    }
}
