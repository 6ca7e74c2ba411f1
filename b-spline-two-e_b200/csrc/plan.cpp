// plan.cpp -- host part of stage C: derive the (l1,l2) block structure from
// the configuration list that count_configs generated
// (src/tools/orbital_tools.f90:157-193) and tabulate the angular factors
// ang_k_LS (src/tools/wigner_tools.f90:126-138) per pair of blocks with the
// thresholds of src/mat_els/hamiltonian.f90:174 (5e-15, sparsity pattern) and
// src/mat_els/mat_els.f90:568 (5e-16, terms of the k sum).
#include "plan.h"

#include <cmath>
#include <mutex>
#include <set>
#include <stdexcept>
#include <unordered_map>

#include "site_core.h"
#include "wigner.h"

namespace bs2e {

// ---------------------------------------------------------------------------
// angular tables, memoised per process ("tabulated on the host, uploaded once")
// ---------------------------------------------------------------------------
namespace {
std::mutex g_ang_mu;
std::unordered_map<uint64_t, double> g_ang_memo;

double ang_cached(int k, int la, int lb, int lc, int ld, int L)
{
    const uint64_t key = ((uint64_t)k << 40) | ((uint64_t)la << 32) | ((uint64_t)lb << 24) |
                         ((uint64_t)lc << 16) | ((uint64_t)ld << 8) | (uint64_t)L;
    {
        std::lock_guard<std::mutex> lk(g_ang_mu);
        auto it = g_ang_memo.find(key);
        if (it != g_ang_memo.end()) return it->second;
    }
    const double v = ang_k_LS(k, la, lb, lc, ld, L);
    std::lock_guard<std::mutex> lk(g_ang_mu);
    g_ang_memo.emplace(key, v);
    return v;
}
}  // namespace

HostPlan build_host_plan(const Geom& hg, int L, long long n_config, const int64_t* conf_n,
                         const int64_t* conf_l, int full, long long n_ranges, const int64_t* range_lo,
                         const int64_t* range_hi)
{
    typedef std::invalid_argument Error;
    if (n_config <= 0) throw Error("block_plan: n_config must be positive");
    if (n_config > 2147483000LL) throw Error("block_plan: n_config exceeds 32-bit row indices");
    if (n_ranges < 1 || !range_lo || !range_hi) throw Error("block_plan: no row range given");
    std::vector<int> rows;
    std::vector<int> row_local((size_t)n_config, -1);
    {
        long long prev = 0;
        for (long long q = 0; q < n_ranges; ++q) {
            const long long lo = range_lo[q], hi = range_hi[q];
            if (lo < 1 || hi > n_config || hi < lo) throw Error("block_plan: row range outside 1..n_config");
            if (lo <= prev) throw Error("block_plan: row ranges must be ascending and disjoint");
            for (long long i = lo; i <= hi; ++i) {
                row_local[(size_t)i - 1] = (int)rows.size();
                rows.push_back((int)i);
            }
            prev = hi;
        }
    }
    if (hg.nb > 65535) throw Error("block_plan: n_b exceeds 16-bit storage");
    if (L < 0 || L > 255) throw Error("block_plan: L out of range");

    std::vector<BlockDesc> blocks;
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
    const int nblk = (int)blocks.size();
    if (nblk > 65535) throw Error("block_plan: too many (l1,l2) blocks");

    // angular tables (hamiltonian.f90:171-178 pattern test, mat_els.f90:566-570 sum)
    const int K1 = hg.K1;
    std::vector<double> angD((size_t)nblk * nblk * K1, 0.0), angX((size_t)nblk * nblk * K1, 0.0);
    std::vector<unsigned char> flags((size_t)nblk * nblk, 0);
    std::vector<KRange> krange((size_t)nblk * nblk);
    for (int bi = 0; bi < nblk; ++bi)
        for (int bj = 0; bj < nblk; ++bj) {
            const int la = blocks[bi].l1, lb = blocks[bi].l2;
            const int lc = blocks[bj].l1, ld = blocks[bj].l2;
            const size_t o = (size_t)bi * nblk + bj;
            const double sgn = ((lc + ld + L) & 1) ? -1.0 : 1.0;
            unsigned f = 0;
            KRange kr{1, 0, 1, 0};
            bool dfirst = true, xfirst = true;
            for (int k = 0; k < K1; ++k) {
                const double ad = ang_cached(k, la, lb, lc, ld, L);
                const double ax = ang_cached(k, la, lb, ld, lc, L);
                if (fabs(ad) > 5.e-15) f |= kDirAny;
                if (fabs(ax) > 5.e-15) f |= kExAny;
                if (!(fabs(ad) < 5.e-16)) {
                    angD[o * K1 + k] = ad;
                    if (dfirst) { kr.dlo = (signed char)k; dfirst = false; }
                    kr.dhi = (signed char)k;
                }
                if (!(fabs(ax) < 5.e-16)) {
                    angX[o * K1 + k] = sgn * ax;
                    if (xfirst) { kr.xlo = (signed char)k; xfirst = false; }
                    kr.xhi = (signed char)k;
                }
            }
            flags[o] = (unsigned char)f;
            krange[o] = kr;
        }

    // the same factors packed by multipole parity for the site kernel
    const int kmax = site_kmax_for(K1);
    const int nkp = kmax > 0 ? site_nkp(kmax) : 0;
    std::vector<double> angP((size_t)nblk * nblk * 2 * nkp, 0.0);
    if (nkp > 0)
        for (int bi = 0; bi < nblk; ++bi)
            for (int bj = 0; bj < nblk; ++bj) {
                const size_t o = (size_t)bi * nblk + bj;
                const int pd = (blocks[bi].l1 + blocks[bj].l1) & 1, px = (blocks[bi].l1 + blocks[bj].l2) & 1;
                for (int k = 0; k < K1; ++k) {
                    if (angD[o * K1 + k] != 0.0) {
                        if ((k ^ pd) & 1) throw std::logic_error("block_plan: direct factor off parity");
                        angP[o * 2 * nkp + (k - pd) / 2] = angD[o * K1 + k];
                    }
                    if (angX[o * K1 + k] != 0.0) {
                        if ((k ^ px) & 1) throw std::logic_error("block_plan: exchange factor off parity");
                        angP[o * 2 * nkp + nkp + (k - px) / 2] = angX[o * K1 + k];
                    }
                }
            }

    // rows of the requested range grouped by radial site (counting sort on the
    // site key, then sites ordered by descending row count so that the heaviest
    // CTAs of the site kernel start first)
    std::vector<unsigned> site_key;
    std::vector<int> site_ptr, site_rows;
    int nsites_x = 0;
    if ((size_t)stride * stride <= ((size_t)1 << 26)) {  // else: no site list, the row kernel is used
        const long long nrows = (long long)rows.size();
        const size_t nkeys = (size_t)stride * stride;
        std::vector<int> kcount(nkeys + 1, 0);
        for (const int r1 : rows) ++kcount[(size_t)rn1[r1 - 1] * stride + rn2[r1 - 1] + 1];
        // distinct sites, bucketed by (has exchange windows, number of rows)
        int max_nd_all = 0;
        for (auto v : rn2) max_nd_all = v > max_nd_all ? v : max_nd_all;
        auto cls = [&](size_t kq) { return site_wants_X(hg, max_nd_all, (int)(kq / stride)) ? 0 : 1; };
        std::vector<int> per_n(2 * (nblk + 2), 0);
        for (size_t kq = 0; kq < nkeys; ++kq)
            if (kcount[kq + 1] > 0) ++per_n[cls(kq) * (nblk + 2) + kcount[kq + 1]];
        std::vector<int> first_of_n(2 * (nblk + 2), 0);  // first site index of (class, n rows)
        int nsites = 0;
        for (int cl = 0; cl < 2; ++cl) {
            for (int n = nblk; n >= 1; --n) { first_of_n[cl * (nblk + 2) + n] = nsites; nsites += per_n[cl * (nblk + 2) + n]; }
            if (cl == 0) nsites_x = nsites;
        }
        site_key.resize(nsites);
        site_ptr.assign(nsites + 1, 0);
        std::vector<int> site_of_key(nkeys, -1);
        {
            std::vector<int> fill = first_of_n;
            for (size_t kq = 0; kq < nkeys; ++kq) {
                const int n = kcount[kq + 1];
                if (n <= 0) continue;
                const int sidx = fill[cls(kq) * (nblk + 2) + n]++;
                site_of_key[kq] = sidx;
                site_key[sidx] = ((unsigned)(kq / stride) << 16) | (unsigned)(kq % stride);
                site_ptr[sidx + 1] = n;
            }
        }
        for (int q = 0; q < nsites; ++q) site_ptr[q + 1] += site_ptr[q];
        site_rows.resize((size_t)nrows);
        std::vector<int> cursor(site_ptr.begin(), site_ptr.end() - 1);
        for (const int r1 : rows) {
            const int sidx = site_of_key[(size_t)rn1[r1 - 1] * stride + rn2[r1 - 1]];
            site_rows[cursor[sidx]++] = r1;
        }
    }

    HostPlan hp;
    hp.site_key = std::move(site_key);
    hp.site_ptr = std::move(site_ptr);
    hp.site_rows = std::move(site_rows);
    hp.nsites_x = nsites_x;
    hp.nblk = nblk;
    hp.L = L;
    hp.full = full ? 1 : 0;
    hp.n_config = n_config;
    hp.lmax = 0;
    for (auto& d : blocks) hp.lmax = d.l1 > hp.lmax ? d.l1 : hp.lmax;
    hp.max_nd = 0;
    for (auto v : rn2) hp.max_nd = v > hp.max_nd ? v : hp.max_nd;
    hp.blocks = std::move(blocks);
    hp.ncrow = std::move(ncrow);
    hp.flags = std::move(flags);
    hp.krange = std::move(krange);
    hp.angD = std::move(angD);
    hp.angX = std::move(angX);
    hp.angP = std::move(angP);
    hp.nkp = nkp;
    hp.row_n1 = std::move(rn1);
    hp.row_n2 = std::move(rn2);
    hp.row_blk = std::move(rblk);
    hp.rows = std::move(rows);
    hp.row_local = std::move(row_local);
    return hp;
}

Plan HostPlan::view() const
{
    Plan pl;
    pl.nblk = nblk;
    pl.n_config = (int)n_config;
    pl.full = full;
    pl.L = L;
    pl.max_nd = max_nd;
    pl.blk = blocks.data();
    pl.ncrow = ncrow.data();
    pl.flags = flags.data();
    pl.krange = krange.data();
    pl.angD = angD.data();
    pl.angX = angX.data();
    pl.angP = angP.data();
    pl.nkp = nkp;
    pl.row_n1 = row_n1.data();
    pl.row_n2 = row_n2.data();
    pl.row_blk = row_blk.data();
    pl.nrows = (int)rows.size();
    pl.rows = rows.data();
    pl.row_local = row_local.data();
    return pl;
}

}  // namespace bs2e
