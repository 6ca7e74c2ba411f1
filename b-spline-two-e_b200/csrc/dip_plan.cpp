// dip_plan.cpp -- host part of the dipole blocks: Wigner-Eckart factor, pattern flags
// and folded angular coefficients per pair of (l1,l2) groups.
//   construct_dip_block_tensor  src/mat_els/dipole.f90:8-47
//   dip_mat_neq / dip_red_mat   src/mat_els/mat_els.f90:717-770
//   dip_red_1p_len / _vel       src/mat_els/mat_els.f90:772-811
//   ang_dip_red                 src/mat_els/mat_els.f90:813-831
#include "dip_plan.h"

#include <cmath>
#include <stdexcept>

#include "wigner.h"

namespace bs2e {

namespace {
// angular part of dip_red_1p for (l -> l_p): value = alpha * A(n,n') + beta * B(n,n')
void red_1p(int gauge, int l, int lp, double* alpha, double* beta)
{
    *alpha = *beta = 0.0;
    if (gauge == 'l') {
        *alpha = C_red_mat(1, l, lp);
        return;
    }
    if (std::abs(l - lp) != 1) return;
    if (l > lp) { *alpha = std::sqrt((double)l); *beta = -std::sqrt((double)l) * (lp + 1); }
    else { *alpha = -std::sqrt((double)lp); *beta = -std::sqrt((double)lp) * lp; }
}
}  // namespace

HostDipPlan build_dip_plan(const Geom& hg, int gauge, int q, const int64_t* sym1, long long n1,
                           const int64_t* conf_n1, const int64_t* conf_l1, const int64_t* sym2, long long n2,
                           const int64_t* conf_n2, const int64_t* conf_l2, bool compute)
{
    if (gauge != 'l' && gauge != 'v') throw std::invalid_argument("dipole: gauge must be 'l' or 'v'");
    if (q < -1 || q > 1) throw std::invalid_argument("dipole: q must be -1, 0 or 1");
    HostDipPlan dp;
    const int L1 = (int)sym1[0], M1 = (int)sym1[1], L2 = (int)sym2[0], M2 = (int)sym2[1];
    const bool parity_allowed = (sym1[2] != 0) != (sym2[2] != 0);
    dp.ang = three_j(L1, 1, L2, -M1, q, M2);
    if ((L1 - M1) & 1) dp.ang = -dp.ang;
    if (std::fabs(dp.ang) < 5.e-16 || !parity_allowed || !compute) return dp;   // dipole.f90:26-30
    if (n1 <= 0 || n2 <= 0) return dp;   // a symmetry without configurations: empty block
    dp.empty = false;
    const int64_t n1r = n1, n2r = n2;
    const int64_t one = 1;
    dp.rows = build_host_plan(hg, L1, n1, conf_n1, conf_l1, 1, 1, &one, &n1r, 0u);   // structure only
    dp.cols = build_host_plan(hg, L2, n2, conf_n2, conf_l2, 1, 1, &one, &n2r, 0u);
    const int nR = dp.rows.nblk, nC = dp.cols.nblk;
    dp.flag.assign((size_t)nR * nC, 0);
    dp.coef.assign((size_t)nR * nC * 8, 0.0);
    const double rt = std::sqrt((double)((2 * L1 + 1) * (2 * L2 + 1)));
    auto sgn = [](int e) { return (e & 1) ? -1.0 : 1.0; };
    for (int bi = 0; bi < nR; ++bi)
        for (int bj = 0; bj < nC; ++bj) {
            const int la = dp.rows.blocks[bi].l1, lb = dp.rows.blocks[bi].l2;
            const int lc = dp.cols.blocks[bj].l1, ld = dp.cols.blocks[bj].l2;
            const size_t o = (size_t)bi * nC + bj;
            double red = 0.0;   // ang_dip_red on the direct pairing
            if (lb == ld) red += std::fabs(six_j(L1, 1, L2, lc, lb, la) * C_red_mat(1, la, lc));
            if (la == lc) red += std::fabs(six_j(L1, 1, L2, ld, la, lb) * C_red_mat(1, lb, ld));
            dp.flag[o] = red > 5.e-16 ? 1 : 0;
            double* cf = dp.coef.data() + o * 8;
            double al, be;
            const double sx = sgn(L2 + lc + ld);   // exchange sign of dip_mat_neq
            if (lb == ld) {                        // res_1: d(la,lc; na,nc) S(nb,nd)
                red_1p(gauge, la, lc, &al, &be);
                const double kap = dp.ang * sgn(la + lb + L2 + 1) * rt * six_j(L1, 1, L2, lc, lb, la);
                cf[0] = kap * al; cf[1] = kap * be;
            }
            if (la == lc) {                        // res_2: d(lb,ld; nb,nd) S(na,nc)
                red_1p(gauge, lb, ld, &al, &be);
                const double kap = dp.ang * sgn(la + ld + L1 + 1) * rt * six_j(L1, 1, L2, ld, la, lb);
                cf[2] = kap * al; cf[3] = kap * be;
            }
            if (lb == lc) {                        // exchange res_1: d(la,ld; na,nd) S(nb,nc)
                red_1p(gauge, la, ld, &al, &be);
                const double kap = dp.ang * sx * sgn(la + lb + L2 + 1) * rt * six_j(L1, 1, L2, ld, lb, la);
                cf[4] = kap * al; cf[5] = kap * be;
            }
            if (la == ld) {                        // exchange res_2: d(lb,lc; nb,nc) S(na,nd)
                red_1p(gauge, lb, lc, &al, &be);
                const double kap = dp.ang * sx * sgn(la + lc + L1 + 1) * rt * six_j(L1, 1, L2, lc, la, lb);
                cf[6] = kap * al; cf[7] = kap * be;
            }
        }
    return dp;
}

DipTables HostDipPlan::tables() const
{
    DipTables t;
    t.nblkR = rows.nblk;
    t.nblkC = cols.nblk;
    t.flag = flag.data();
    t.coef = coef.data();
    t.row_n1 = rows.row_n1.data();
    t.row_n2 = rows.row_n2.data();
    t.row_blk = rows.row_blk.data();
    t.nrows = (int)rows.n_config;
    return t;
}

}  // namespace bs2e
