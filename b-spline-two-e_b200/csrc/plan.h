// plan.h -- structure of one symmetry block (see plan.cpp for the host build and
// plan_dev.cu for the device build that the product uses)
#pragma once
#include <cstdint>
#include <vector>

#include "core.h"

namespace bs2e {

// angular factors of every pair of (l1,l2) groups of a block (wigner_tools.f90:126-138),
// with the thresholds of hamiltonian.f90:174 (pattern) and mat_els.f90:568 (sum) folded in
struct AngTables {
    int nblk = 0, K1 = 0, nkp = 0;
    int maxc = 0;                      // most column groups any row group couples to (its own included)
    int maxrec = 0;                    // most off-diagonal (row group, column group) pairs of one multipole parity
    std::vector<unsigned char> flags;  // [nblk][nblk]
    std::vector<KRange> krange;        // [nblk][nblk]
    std::vector<double> angD, angX;    // [nblk][nblk][K1]
    std::vector<double> angP;          // [nblk][nblk][2*nkp], packed by multipole parity (site_core.h)
};
// exact 3j/6j arithmetic (wigner.cpp), spread over the host cores
AngTables build_ang_tables(const std::vector<BlockDesc>& blocks, int L, int K1);

// 64-bit sort key of a radial site: sites with exchange windows first, inside each
// class the sites with most rows first (the heaviest CTAs of the site kernel start first)
BS2E_HD unsigned long long site_sort_key(bool wantX, int nrows, int na, int nb)
{
    return ((unsigned long long)(wantX ? 0 : 1) << 42) | ((unsigned long long)(1023 - nrows) << 32) |
           ((unsigned long long)na << 16) | (unsigned long long)nb;
}
constexpr int kMaxBlocks = 1023;  // (l1,l2) groups per symmetry block (site_sort_key)

struct HostPlan {
    int nblk = 0, L = 0, full = 0, lmax = 0, max_nd = 0;
    long long n_config = 0;
    std::vector<BlockDesc> blocks;
    std::vector<int> blk_start;        // [nblk+1] 0-based first configuration of each group
    std::vector<NcRow> ncrow;
    AngTables ang;
    std::vector<unsigned short> row_n1, row_n2, row_blk;
    // planned rows: ascending disjoint ranges, and the local index of the first row of each
    std::vector<int> range_lo, range_hi, range_off;
    int nrows = 0;
    // radial sites (n1,n2) that carry planned rows, in site_sort_key order
    std::vector<unsigned long long> site_key;
    int nsites_x = 0;                  // sites 0..nsites_x-1 have exchange windows (site_wants_X)
    Plan view() const;  // Plan over the HOST arrays
};

constexpr unsigned kPlanAngular = 1u;  // angular tables (exact 3j/6j: the expensive part)
constexpr unsigned kPlanSites = 2u;    // radial-site list of the planned rows
// rows = union of the n_ranges ascending, disjoint, inclusive 1-based ranges [range_lo[q], range_hi[q]]
HostPlan build_host_plan(const Geom& hg, int L, long long n_config, const int64_t* conf_n,
                         const int64_t* conf_l, int full, long long n_ranges, const int64_t* range_lo,
                         const int64_t* range_hi, unsigned parts = kPlanAngular | kPlanSites);
inline HostPlan build_host_plan(const Geom& hg, int L, long long n_config, const int64_t* conf_n,
                                const int64_t* conf_l, int full, long long row_lo, long long row_hi)
{
    const int64_t lo = row_lo, hi = row_hi;
    return build_host_plan(hg, L, n_config, conf_n, conf_l, full, 1, &lo, &hi);
}

// checks shared by the host and the device build
void check_row_ranges(long long n_config, long long n_ranges, const int64_t* range_lo, const int64_t* range_hi,
                      std::vector<int>& lo, std::vector<int>& hi, std::vector<int>& off, int* nrows);

}  // namespace bs2e
