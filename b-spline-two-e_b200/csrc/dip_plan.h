// dip_plan.h -- host plan of one dipole block (see dip_plan.cpp)
#pragma once
#include <cstdint>
#include <vector>

#include "dip_core.h"
#include "plan.h"

namespace bs2e {

struct HostDipPlan {
    bool empty = true;          // forbidden transition or compute == false: nnz = 0
    double ang = 0.0;           // (-1)^(L1-M1) (L1 1 L2; -M1 q M2)
    HostPlan rows, cols;        // structure of sym1's / sym2's configuration lists (cols: full = 1)
    std::vector<unsigned char> flag;
    std::vector<double> coef;
    DipTables tables() const;   // view over the HOST arrays
};

// sym = (l, m, pi); gauge 'l' or 'v'
HostDipPlan build_dip_plan(const Geom& hg, int gauge, int q, const int64_t* sym1, long long n1,
                           const int64_t* conf_n1, const int64_t* conf_l1, const int64_t* sym2, long long n2,
                           const int64_t* conf_n2, const int64_t* conf_l2, bool compute);

}  // namespace bs2e
