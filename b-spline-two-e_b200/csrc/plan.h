// plan.h -- host-side plan of one symmetry block (see plan.cpp)
#pragma once
#include <cstdint>
#include <vector>

#include "core.h"

namespace bs2e {

struct HostPlan {
    int nblk = 0, L = 0, full = 0, lmax = 0, max_nd = 0;
    long long n_config = 0;
    std::vector<BlockDesc> blocks;
    std::vector<NcRow> ncrow;
    std::vector<unsigned char> flags;
    std::vector<KRange> krange;
    std::vector<double> angD, angX;
    std::vector<double> angP;  // packed by parity for the site kernel, [nblk][nblk][2*nkp]
    int nkp = 0;
    std::vector<unsigned short> row_n1, row_n2, row_blk;
    std::vector<int> rows, row_local;  // planned rows: local -> configuration index, and back (-1: not planned)
    // the planned rows grouped by radial site (n1,n2): sites with exchange windows
    // first, inside each class the sites with most rows first
    std::vector<unsigned> site_key;  // n1 << 16 | n2
    std::vector<int> site_ptr;       // [nsites+1]
    std::vector<int> site_rows;      // 1-based row indices, ascending inside a site
    int nsites_x = 0;                // sites 0..nsites_x-1 have exchange windows (site_wants_X), the rest do not
    Plan view() const;  // Plan over the HOST arrays
};

// rows = union of the n_ranges ascending, disjoint, inclusive 1-based ranges [range_lo[q], range_hi[q]]
HostPlan build_host_plan(const Geom& hg, int L, long long n_config, const int64_t* conf_n,
                         const int64_t* conf_l, int full, long long n_ranges, const int64_t* range_lo,
                         const int64_t* range_hi);
inline HostPlan build_host_plan(const Geom& hg, int L, long long n_config, const int64_t* conf_n,
                                const int64_t* conf_l, int full, long long row_lo, long long row_hi)
{
    const int64_t lo = row_lo, hi = row_hi;
    return build_host_plan(hg, L, n_config, conf_n, conf_l, full, 1, &lo, &hi);
}

}  // namespace bs2e
