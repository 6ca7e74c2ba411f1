// dip_core.h -- thread-level bodies of the dipole block kernels (dip.cu), shared with
// the CPU logic checker like core.h.
//
// construct_dip_block_tensor / init_dip_block (src/mat_els/dipole.f90:8-47,87-146):
// block <sym1| d_q |sym2>, rows = configurations (l_a,l_b; n_a,n_b) of sym1, columns =
// configurations (l_c,l_d; n_c,n_d) of sym2, no triangle cut.  A column is stored iff
//   (support or support_ex)  and  ang_dip_red(L1, L2, l's) > 5e-16
// (the angular test is made on the DIRECT pairing also when only the exchange pairing
// has support, dipole.f90:105-109 -- mirrored), i.e. the union of the direct and the
// exchange window of the row's radial site inside every column group (l_c,l_d) whose
// flag is set.  The value (mat_els.f90:717-770) is a sum of four products
//   (alpha_t A(n,n') + beta_t B(n,n')) S(m,m')
// of band matrices: A = r_mat (length gauge) or dr_mat, B = r_inv_mat (velocity), with
// the 6j / phase / Wigner-Eckart factors folded into (alpha_t, beta_t) per pair of
// (l1,l2) groups on the host (dip_plan.cpp).
#pragma once
#include "core.h"

namespace bs2e {

struct DipTables {
    int nblkR, nblkC;
    const unsigned char* flag;  // [nblkR][nblkC]  ang_dip_red > 5e-16
    const double* coef;         // [nblkR][nblkC][8] (alpha_t, beta_t), t = 0..3
    const unsigned short* row_n1;   // rows = configurations of sym1
    const unsigned short* row_n2;
    const unsigned short* row_blk;
    int nrows;
};

struct DipBand {    // band storage like OneBody: M[n][n'-n+w], complex interleaved
    const double* A;
    const double* B;  // may alias A when the gauge has no second matrix (beta = 0)
    const double* S;
};

BS2E_HD Cplx band_at(const Geom& g, const double* M, int n, int np)
{
    const int d = np - n + g.w;
    if (d < 0 || d > 2 * g.w) return Cplx{0.0, 0.0};
    const double* q = M + ((size_t)n * (2 * g.w + 1) + d) * 2;
    return Cplx{q[0], q[1]};
}

// the stored n_d intervals of one (row, column group, n_c): plC is the Plan of sym2's
// configuration list built with full = 1 (no triangle cut)
BS2E_HD RowInfo dip_row(const DipTables& dt, int i /*1-based*/)
{
    RowInfo r;
    r.i = i;
    r.bi = dt.row_blk[i - 1];
    r.na = dt.row_n1[i - 1];
    r.nb = dt.row_n2[i - 1];
    r.la = r.lb = 0;
    return r;
}

// walk the stored columns of row r in ascending order in chunks of <= 32 consecutive n_d
template <class F>
BS2E_HD void dip_for_each_chunk(const Geom& g, const Plan& plC, const DipTables& dt, const RowInfo& r, F&& f)
{
    RowInfo rr = r;
    rr.bi = -1;  // never the "own" group of the column list: no triangle cut
    for (int bj = 0; bj < plC.nblk; ++bj) {
        if (!dt.flag[(size_t)r.bi * dt.nblkC + bj]) continue;
        const Union2 win = nc_windows(g, plC, rr, bj);
        for (int q = 0; q < win.n; ++q)
            for (int nc = win.lo[q]; nc <= win.hi[q]; ++nc) {
                const Segment s = segment(g, plC, rr, bj, nc);
                const Union2 u = union2(s.dlo, s.dhi, s.xlo, s.xhi);
                for (int z = 0; z < u.n; ++z)
                    for (int base = u.lo[z]; base <= u.hi[z]; base += 32) f(bj, nc, s, base, u.hi[z]);
            }
    }
}

BS2E_HD long long dip_row_count(const Geom& g, const Plan& plC, const DipTables& dt, int i)
{
    long long c = 0;
    dip_for_each_chunk(g, plC, dt, dip_row(dt, i),
                       [&](int, int, const Segment&, int base, int hi) { c += imin(32, hi - base + 1); });
    return c;
}

// value of the entry (row r, column (bj; nc, nd))
BS2E_HD Cplx dip_value_cf(const Geom& g, const double* cf, const DipBand& bd, const RowInfo& r, int nc, int nd);
BS2E_HD Cplx dip_value(const Geom& g, const DipTables& dt, const DipBand& bd, const RowInfo& r, int bj, int nc, int nd)
{
    return dip_value_cf(g, dt.coef + ((size_t)r.bi * dt.nblkC + bj) * 8, bd, r, nc, nd);
}
// cf: the eight folded coefficients (alpha_t, beta_t) of the (row group, column group) pair
BS2E_HD Cplx dip_value_cf(const Geom& g, const double* cf, const DipBand& bd, const RowInfo& r, int nc, int nd)
{
    Cplx acc = Cplx{0.0, 0.0};
    // t: (n, n') of the one-particle dipole, (m, m') of the overlap
    const int n_[4] = {r.na, r.nb, r.na, r.nb}, np_[4] = {nc, nd, nd, nc};
    const int m_[4] = {r.nb, r.na, r.nb, r.na}, mp_[4] = {nd, nc, nc, nd};
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int t = 0; t < 4; ++t) {
        const double al = cf[2 * t], be = cf[2 * t + 1];
        if (al == 0.0 && be == 0.0) continue;
        const Cplx a = band_at(g, bd.A, n_[t], np_[t]);
        Cplx d = Cplx{al * a.re, al * a.im};
        if (be != 0.0) {   // (length gauge: no second matrix; the products with beta = 0 are skipped, their value is +0)
            const Cplx b = band_at(g, bd.B, n_[t], np_[t]);
            d = Cplx{al * a.re + be * b.re, al * a.im + be * b.im};
        }
        acc = cadd(acc, cmul(d, band_at(g, bd.S, m_[t], mp_[t])));
    }
    return acc;
}

}  // namespace bs2e
