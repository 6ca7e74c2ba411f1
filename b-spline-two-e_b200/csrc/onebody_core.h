// onebody_core.h -- one-particle matrices and radial dipole integrals, element by element.
//
// Stands in for setup_S / setup_H_one_particle (src/mat_els/mat_els.f90:47-118 with compute_H / compute_S,
// :294-346; hydrogenic potential src/mat_els/potentials.f90:35-43, CAP src/tools/CAP_tools.f90:24-34) and
// setup_radial_dip (:120-170 with :348-390).  One thread owns one band entry (n, n') of ALL matrices: it walks
// the cells where both splines live in ascending order and the Gauss-Legendre nodes of each cell in ascending
// order -- the summation order of the reference -- and accumulates S, H_l for every l, r_mat, dr_mat and
// r_inv_mat from one evaluation of the splines per node.
// __host__ __device__ so that the CPU checker (tests/hostcheck) runs the same statements.
#pragma once
#include "core.h"

namespace bs2e {

// values (m = 0) or m-th derivatives of the ks splines living on cell v at x: out[s] <-> full index v+s.
// Derivatives by differencing the coefficient vector (de Boor), as BVALUE_D does (bspline_tools.f90:151-224).
BS2E_HD void bspline_derivs(const double* t, int ks, int v, double x, int m, double* out)
{
    const int left = ks - 1 + v;  // 1-based
    if (m == 0) { bspline_values_left(t, ks, left, x, out); return; }
    if (m >= ks) { for (int s = 0; s < ks; ++s) out[s] = 0.0; return; }
    double low[kMaxOrder];
    bspline_values_left(t, ks - m, left, x, low);  // low[q] <-> full index v+m+q
    for (int s = 0; s < ks; ++s) {
        double c[kMaxOrder + 1];   // coefficients of the unit vector e_{v+s}, differenced m times
        for (int j = 0; j < ks; ++j) c[j] = (j == s) ? 1.0 : 0.0;
        int lo = 0;
        for (int d = 1; d <= m; ++d) {
            for (int j = ks - 1; j >= lo + 1; --j) {
                const int full = v + j;
                c[j] = (ks - d) * (c[j] - c[j - 1]) / (t[full + ks - d - 1] - t[full - 1]);
            }
            lo += 1;
        }
        double acc = 0.0;
        for (int j = m; j < ks; ++j) acc += c[j] * low[j - m];
        out[s] = acc;
    }
}

constexpr int kMaxL1p = 48;   // H_l accumulators per thread

struct OneBodyParams {
    int Z, lmax;          // nuclear charge, l = 0..lmax
    int cap_order;
    double cap_r0, eta_re, eta_im;
    int want_1p;          // S and H_l
    int gauge;            // 0: no dipole integrals; 'l': r_mat; 'v': dr_mat and r_inv_mat
};

struct OneBodyOut {       // band storage [n][n'-n+w] complex (core.h: OneBody), device or host pointers
    double* Sb;           // [nb+1][2w+1][2]
    double* Hb;           // [lmax+1][nb+1][2w+1][2]
    double* A;            // r_mat (gauge 'l') or dr_mat (gauge 'v')
    double* B;            // r_inv_mat (gauge 'v')
};

// entry (n, np = n + d - w), 1 <= n, np <= nb
BS2E_HD void one_body_entry(const Geom& g, const OneBodyParams& p, const OneBodyOut& o, int n, int d)
{
    const int ks = g.ks, w = g.w, bw = 2 * w + 1;
    const int np = n + d - w;
    if (np < 1 || np > g.nb) return;
    const int lo = pair_lo_cell(g, n, np), hi = pair_hi_cell(g, n, np);
    double S = 0.0, rl = 0.0, dr = 0.0, ri = 0.0;
    double Hre[kMaxL1p], Him[kMaxL1p];
    for (int l = 0; l <= p.lmax; ++l) Hre[l] = Him[l] = 0.0;
    double Bv[kMaxOrder], D1[kMaxOrder], D2[kMaxOrder];
    for (int v = lo; v <= hi; ++v) {
        const double a = g.bp[v - 1], b = g.bp[v];
        const double scale = 0.5 * (b - a), mid = 0.5 * (b + a);
        const int si = n + 1 - v, sj = np + 1 - v;   // local slots: full index n+1 = v + s
        for (int q = 0; q < g.kgl; ++q) {
            const double r = scale * g.glx[q] + mid, wq = scale * g.glw[q];
            bspline_values(g.t, ks, v, r, Bv);
            const double Bi = Bv[si], Bj = Bv[sj];
            if (p.want_1p) {
                bspline_derivs(g.t, ks, v, r, 2, D2);
                S += wq * Bi * Bj;
                // V(r,l) = l(l+1)/(2 r^2) - Z/r; CAP = -i eta (r - r0)^order for r >= r0
                double cre = 0.0, cim = 0.0;
                if (r >= p.cap_r0) {
                    const double pw = powi(r - p.cap_r0, p.cap_order);
                    cre = p.eta_im * pw;      // (-i)(eta_re + i eta_im) = eta_im - i eta_re
                    cim = -p.eta_re * pw;
                }
                const double kin = -0.5 * Bi * D2[sj], BB = Bi * Bj;
                for (int l = 0; l <= p.lmax; ++l) {
                    const double V = 0.5 * (double)l * (double)(l + 1) / (r * r) - (double)p.Z / r;
                    Hre[l] += wq * (kin + (V + cre) * BB);
                    Him[l] += wq * (cim * BB);
                }
            }
            if (p.gauge == 'l') rl += wq * r * Bi * Bj;
            if (p.gauge == 'v') {
                bspline_derivs(g.t, ks, v, r, 1, D1);
                dr += wq * Bi * D1[sj];
                ri += wq * Bi * Bj / r;
            }
        }
    }
    const size_t at = ((size_t)n * bw + d) * 2;
    const size_t per = (size_t)(g.nb + 1) * bw * 2;
    if (p.want_1p) {
        o.Sb[at] = S;
        o.Sb[at + 1] = 0.0;
        for (int l = 0; l <= p.lmax; ++l) {
            o.Hb[l * per + at] = Hre[l];
            o.Hb[l * per + at + 1] = Him[l];
        }
    }
    if (p.gauge == 'l') { o.A[at] = rl; o.A[at + 1] = 0.0; }
    if (p.gauge == 'v') {   // dr_mat = -i int B_i B_j', r_inv_mat = -i int B_i B_j / r
        o.A[at] = 0.0; o.A[at + 1] = -dr;
        o.B[at] = 0.0; o.B[at + 1] = -ri;
    }
}

}  // namespace bs2e
