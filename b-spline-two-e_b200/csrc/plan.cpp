// plan.cpp -- host part of stage C.
//   * build_ang_tables: the angular factors ang_k_LS (src/tools/wigner_tools.f90:126-138)
//     per pair of (l1,l2) groups with the thresholds of src/mat_els/hamiltonian.f90:174
//     (5e-15, sparsity pattern) and src/mat_els/mat_els.f90:568 (5e-16, terms of the k sum).
//     This is the only host arithmetic of a block plan in the product (plan_dev.cu).
//   * build_host_plan: the (l1,l2) group structure of a configuration list that
//     count_configs generated (src/tools/orbital_tools.f90:157-193), built on the host.
//     The product builds the same tables on the device (plan_dev.cu); this host version
//     serves the dipole plan and the CPU emulation of the kernels (tests/hostcheck).
#include "plan.h"

#include <algorithm>
#include <cmath>
#include <set>
#include <stdexcept>
#include <thread>

#include "site_core.h"
#include "wigner.h"

namespace bs2e {

AngTables build_ang_tables(const std::vector<BlockDesc>& blocks, int L, int K1)
{
    AngTables t;
    const int nblk = (int)blocks.size();
    t.nblk = nblk;
    t.K1 = K1;
    const int kmax = site_kmax_for(K1);
    const int nkp = kmax > 0 ? site_nkp(kmax) : 0;
    t.nkp = nkp;
    t.flags.assign((size_t)nblk * nblk, 0);
    t.krange.assign((size_t)nblk * nblk, KRange{1, 0, 1, 0});
    t.angD.assign((size_t)nblk * nblk * K1, 0.0);
    t.angX.assign((size_t)nblk * nblk * K1, 0.0);
    t.angP.assign((size_t)nblk * nblk * 2 * nkp, 0.0);
    int lmax = 0;
    for (const BlockDesc& b : blocks) lmax = std::max(lmax, std::max(b.l1, b.l2));
    // <a || C^k || b> once per (k, a, b)
    const int nl = lmax + 1;
    std::vector<double> cred((size_t)K1 * nl * nl);
    for (int k = 0; k < K1; ++k)
        for (int a = 0; a < nl; ++a)
            for (int b = 0; b < nl; ++b) cred[((size_t)k * nl + a) * nl + b] = C_red_mat(k, a, b);
    // wigner_tools.f90:126-138 with the reduced matrix elements from the table
    auto ang = [&](int k, int la, int lb, int lc, int ld) -> double {
        if (((la + k + lc) & 1) || ((lb + k + ld) & 1)) return 0.0;
        const double c1 = cred[((size_t)k * nl + la) * nl + lc], c2 = cred[((size_t)k * nl + lb) * nl + ld];
        if (c1 == 0.0 || c2 == 0.0) return 0.0;  // exact structural zeros of the 3j symbols
        const double sgn = ((lb + lc + L) & 1) ? -1.0 : 1.0;
        return sgn * six_j(la, lb, L, ld, lc, k) * c1 * c2;
    };
    auto row_task = [&](int bi) {
        for (int bj = 0; bj < nblk; ++bj) {
            const int la = blocks[bi].l1, lb = blocks[bi].l2;
            const int lc = blocks[bj].l1, ld = blocks[bj].l2;
            const size_t o = (size_t)bi * nblk + bj;
            const double sgn = ((lc + ld + L) & 1) ? -1.0 : 1.0;
            unsigned f = 0;
            KRange kr{1, 0, 1, 0};
            bool dfirst = true, xfirst = true;
            for (int k = 0; k < K1; ++k) {
                const double ad = ang(k, la, lb, lc, ld);
                const double ax = ang(k, la, lb, ld, lc);
                if (std::fabs(ad) > 5.e-15) f |= kDirAny;
                if (std::fabs(ax) > 5.e-15) f |= kExAny;
                if (!(std::fabs(ad) < 5.e-16)) {
                    t.angD[o * K1 + k] = ad;
                    if (dfirst) { kr.dlo = (signed char)k; dfirst = false; }
                    kr.dhi = (signed char)k;
                }
                if (!(std::fabs(ax) < 5.e-16)) {
                    t.angX[o * K1 + k] = sgn * ax;
                    if (xfirst) { kr.xlo = (signed char)k; xfirst = false; }
                    kr.xhi = (signed char)k;
                }
            }
            t.flags[o] = (unsigned char)f;
            t.krange[o] = kr;
            // the same factors packed by multipole parity for the site kernels.  A window whose pattern flag
            // is off contributes nothing (hamiltonian.f90:174,183: no factor above 5e-15; with exact 3j/6j
            // arithmetic such a window has no non-zero factor at all): its half stays zero, so the site
            // kernels add both windows' sums without looking at the flags.
            if (nkp > 0) {
                const int pd = (la + lc) & 1, px = (la + ld) & 1;
                for (int k = 0; k < K1; ++k) {
                    if (!(f & kDirAny) && bi != bj) continue;
                    if (t.angD[o * K1 + k] != 0.0) {
                        if ((k ^ pd) & 1) throw std::logic_error("block_plan: direct factor off parity");
                        t.angP[o * 2 * nkp + (k - pd) / 2] = t.angD[o * K1 + k];
                    }
                }
                for (int k = 0; k < K1; ++k) {
                    if (!(f & kExAny)) continue;
                    if (t.angX[o * K1 + k] != 0.0) {
                        if ((k ^ px) & 1) throw std::logic_error("block_plan: exchange factor off parity");
                        t.angP[o * 2 * nkp + nkp + (k - px) / 2] = t.angX[o * K1 + k];
                    }
                }
            }
        }
    };
    unsigned nthr = std::thread::hardware_concurrency();
    nthr = std::max(1u, std::min(nthr, 16u));
    if ((size_t)nblk * nblk * K1 < 4096 || nthr == 1) {
        for (int bi = 0; bi < nblk; ++bi) row_task(bi);
    } else {
        std::vector<std::thread> pool;
        std::vector<std::string> errs(nthr);
        for (unsigned w = 0; w < nthr; ++w)
            pool.emplace_back([&, w] {
                try {
                    for (int bi = (int)w; bi < nblk; bi += (int)nthr) row_task(bi);
                } catch (const std::exception& e) {
                    errs[w] = e.what();
                }
            });
        for (auto& th : pool) th.join();
        for (auto& e : errs)
            if (!e.empty()) throw std::logic_error(e);
    }
    for (int bi = 0; bi < nblk; ++bi) {
        int c = 0;
        for (int bj = 0; bj < nblk; ++bj) c += bj == bi || t.flags[(size_t)bi * nblk + bj] != 0;
        t.maxc = std::max(t.maxc, c);
    }
    {   // records a site can hold per multipole parity (wigner_tools.f90:131: parity of l_a + l_c)
        int cnt[2] = {0, 0};
        for (int bi = 0; bi < nblk; ++bi)
            for (int bj = 0; bj < nblk; ++bj)
                if (bj != bi && t.flags[(size_t)bi * nblk + bj] != 0) ++cnt[(blocks[bi].l1 + blocks[bj].l1) & 1];
        t.maxrec = std::max(cnt[0], cnt[1]);
    }
    return t;
}

void check_row_ranges(long long n_config, long long n_ranges, const int64_t* range_lo, const int64_t* range_hi,
                      std::vector<int>& lo, std::vector<int>& hi, std::vector<int>& off, int* nrows)
{
    typedef std::invalid_argument Error;
    if (n_ranges < 0 || (n_ranges > 0 && (!range_lo || !range_hi))) throw Error("block_plan: no row range given");
    lo.clear(); hi.clear(); off.clear();
    long long prev = 0, run = 0;
    for (long long q = 0; q < n_ranges; ++q) {
        const long long a = range_lo[q], b = range_hi[q];
        if (a < 1 || b > n_config || b < a) throw Error("block_plan: row range outside 1..n_config");
        if (a <= prev) throw Error("block_plan: row ranges must be ascending and disjoint");
        lo.push_back((int)a);
        hi.push_back((int)b);
        off.push_back((int)run);
        run += b - a + 1;
        prev = b;
    }
    *nrows = (int)run;
}

HostPlan build_host_plan(const Geom& hg, int L, long long n_config, const int64_t* conf_n,
                         const int64_t* conf_l, int full, long long n_ranges, const int64_t* range_lo,
                         const int64_t* range_hi, unsigned parts)
{
    typedef std::invalid_argument Error;
    if (n_config < 0) throw Error("block_plan: negative n_config");
    if (n_config > 2147483000LL) throw Error("block_plan: n_config exceeds 32-bit row indices");
    if (hg.nb > 65535) throw Error("block_plan: n_b exceeds 16-bit storage");
    if (L < 0 || L > 255) throw Error("block_plan: L out of range");
    HostPlan hp;
    hp.L = L;
    hp.full = full ? 1 : 0;
    hp.n_config = n_config;
    if (n_config == 0) {  // a symmetry without configurations: empty CSR blocks (hamiltonian.f90:137-139 with nnz = 0)
        hp.blk_start.assign(1, 0);
        return hp;
    }
    if (n_ranges < 1) throw Error("block_plan: no row range given");
    check_row_ranges(n_config, n_ranges, range_lo, range_hi, hp.range_lo, hp.range_hi, hp.range_off, &hp.nrows);

    std::vector<BlockDesc> blocks;
    std::vector<int> blk_start;
    std::vector<NcRow> ncrow;
    std::vector<unsigned short> rn1(n_config), rn2(n_config), rblk(n_config);
    std::set<std::pair<int, int>> seen;
    const int stride = hg.nb + 1;
    int prev_n1 = 0, prev_n2 = 0;
    for (long long i = 0; i < n_config; ++i) {
        const long long n1 = conf_n[2 * i], n2 = conf_n[2 * i + 1];
        const long long l1 = conf_l[2 * i], l2 = conf_l[2 * i + 1];
        if (n1 < 1 || n1 > hg.nb || n2 < 1 || n2 > hg.nb)
            throw Error("block_plan: configuration n outside 1..n_b");
        if (l2 < 0 || l1 < l2 || l1 > 120)
            throw Error("block_plan: configurations must have l(1) >= l(2) >= 0 (count_configs order)");
        const bool newblk = blocks.empty() || blocks.back().l1 != l1 || blocks.back().l2 != l2;
        if (newblk) {
            if (!seen.insert({(int)l1, (int)l2}).second)
                throw Error("block_plan: configurations of one (l1,l2) pair are not contiguous");
            blocks.push_back(BlockDesc{(int)l1, (int)l2, (int)n1, (int)n1});
            blk_start.push_back((int)i);
            ncrow.resize(blocks.size() * (size_t)stride, NcRow{1, 0, 0, 0});
        }
        BlockDesc& b = blocks.back();
        NcRow& row = ncrow[(blocks.size() - 1) * (size_t)stride + n1];
        if (newblk || n1 != prev_n1) {
            if (!newblk && n1 < prev_n1)
                throw Error("block_plan: n(1) not ascending inside an (l1,l2) block");
            if (row.nd_hi >= row.nd_lo) throw Error("block_plan: duplicate n(1) row");
            row.nd_lo = (int)n2;
            row.nd_hi = (int)n2;
            row.start = (int)(i + 1);
            b.nc_hi = (int)n1;
        } else {
            if (n2 != prev_n2 + 1)
                throw Error("block_plan: n(2) not consecutive inside an n(1) row");
            row.nd_hi = (int)n2;
        }
        prev_n1 = (int)n1;
        prev_n2 = (int)n2;
        rn1[i] = (unsigned short)n1;
        rn2[i] = (unsigned short)n2;
        rblk[i] = (unsigned short)(blocks.size() - 1);
    }
    blk_start.push_back((int)n_config);
    const int nblk = (int)blocks.size();
    if (nblk > kMaxBlocks) throw Error("block_plan: too many (l1,l2) blocks");
    hp.nblk = nblk;
    hp.lmax = 0;
    for (auto& d : blocks) hp.lmax = d.l1 > hp.lmax ? d.l1 : hp.lmax;
    hp.max_nd = 0;
    for (auto v : rn2) hp.max_nd = v > hp.max_nd ? v : hp.max_nd;

    if (parts & kPlanAngular) hp.ang = build_ang_tables(blocks, L, hg.K1);

    hp.blocks = std::move(blocks);
    hp.blk_start = std::move(blk_start);
    hp.ncrow = std::move(ncrow);
    hp.row_n1 = std::move(rn1);
    hp.row_n2 = std::move(rn2);
    hp.row_blk = std::move(rblk);

    // radial sites that carry planned rows, in the order of the device build (plan_dev.cu)
    if (parts & kPlanSites) {
        const Plan pl = hp.view();
        for (int na = 1; na <= hg.nb; ++na)
            for (int nb = 1; nb <= hg.nb; ++nb) {
                int cnt = 0;
                for (int bi = 0; bi < nblk; ++bi) {
                    const int row = config_index(hg, pl, bi, na, nb);
                    if (row > 0 && row_local_of(pl.rr, row) >= 0) ++cnt;
                }
                if (cnt > 0) hp.site_key.push_back(site_sort_key(site_wants_X(hg, hp.max_nd, na), cnt, na, nb));
            }
        std::sort(hp.site_key.begin(), hp.site_key.end());
        hp.nsites_x = 0;
        for (auto k : hp.site_key)
            if (((k >> 42) & 1) == 0) ++hp.nsites_x;
    }
    return hp;
}

Plan HostPlan::view() const
{
    Plan pl{};
    pl.nblk = nblk;
    pl.n_config = (int)n_config;
    pl.full = full;
    pl.L = L;
    pl.max_nd = max_nd;
    pl.blk = blocks.data();
    pl.ncrow = ncrow.data();
    pl.flags = ang.flags.data();
    pl.krange = ang.krange.data();
    pl.angD = ang.angD.data();
    pl.angX = ang.angX.data();
    pl.angP = ang.angP.data();
    pl.nkp = ang.nkp;
    pl.row_n1 = row_n1.data();
    pl.row_n2 = row_n2.data();
    pl.row_blk = row_blk.data();
    pl.nrows = nrows;
    pl.rr = RowRanges{(int)range_lo.size(), range_lo.data(), range_hi.data(), range_off.data(),
                      range_lo.empty() ? 1 : range_lo[0], range_hi.empty() ? 0 : range_hi[0]};
    return pl;
}

}  // namespace bs2e
