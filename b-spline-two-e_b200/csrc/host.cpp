// host.cpp -- host-side companions of the hot path: the cheap reference
// routines whose OUTPUT feeds the GPU path (knot grid, Gauss-Legendre rule,
// one-particle matrices, configuration lists).  In the Fortran program these
// stay Fortran (out of scope as GPU targets); this file
// lets the C++/Python drivers of this repository produce the same inputs.
//
// Mirrors: src/tools/grid_tools.f90:6-55, src/tools/quad_tools.f90:14-27,
// src/tools/bspline_tools.f90:364-373, src/mat_els/mat_els.f90:47-118,
// src/mat_els/potentials.f90:35-43, src/tools/CAP_tools.f90:24-34,
// src/tools/orbital_tools.f90:46-72,119-216,245-343.
#include "host.h"

#include <cmath>
#include <complex>
#include <stdexcept>

#include "core.h"

namespace bs2e {
namespace host {

// grid_tools.f90:6-55: k-fold knot at 0, m linear steps 2^-m, geometric growth
// by (1+2^-m) until the step reaches h_max*Z, linear steps h_max*Z up to
// Z*r_max, (k-1) repeated end knots; everything divided by Z.
std::vector<double> generate_grid(int k, int m, int Z, double h_max, double r_max)
{
    if (k < 2 || k > kMaxOrder || m < 0 || Z < 1 || !(h_max > 0) || !(r_max > 0))
        throw std::invalid_argument("generate_grid: bad parameters");
    const double h = std::ldexp(1.0, -m);
    std::vector<double> g(k, 0.0);
    for (int i = 0; i < m; ++i) g.push_back(g.back() + h);
    for (;;) {
        const double next = g.back() * (1.0 + h);
        if (next - g.back() >= h_max * Z) break;
        g.push_back(next);
    }
    while (g.back() < Z * r_max) g.push_back(g.back() + Z * h_max);
    const double last = g.back();
    for (int i = 0; i < k - 1; ++i) g.push_back(last);
    for (auto& x : g) x /= Z;
    return g;
}

// Gauss-Legendre rule on [a,b] (quad_tools.f90:14-27 -> stdlib gauss_legendre):
// Newton iteration on P_N from Chebyshev-like guesses, ascending nodes.
void gauss_legendre(int N, double a, double b, double* x, double* w)
{
    if (N < 1) throw std::invalid_argument("gauss_legendre: N < 1");
    const long double pi = 3.14159265358979323846264338327950288L;
    for (int i = 0; i < (N + 1) / 2; ++i) {
        long double z = std::cos(pi * (i + 0.75L) / (N + 0.5L));
        long double dp = 1.0L;
        for (int it = 0; it < 100; ++it) {
            long double p0 = 1.0L, p1 = z;
            for (int j = 2; j <= N; ++j) {
                const long double p2 = ((2 * j - 1) * z * p1 - (j - 1) * p0) / j;
                p0 = p1;
                p1 = p2;
            }
            if (N == 1) { p0 = 1.0L; p1 = z; }
            dp = N * (z * p1 - p0) / (z * z - 1.0L);
            const long double dz = p1 / dp;
            z -= dz;
            if (std::fabs((double)dz) < 1e-19) break;
        }
        // recompute derivative at the converged node for the weight
        long double p0 = 1.0L, p1 = z;
        for (int j = 2; j <= N; ++j) {
            const long double p2 = ((2 * j - 1) * z * p1 - (j - 1) * p0) / j;
            p0 = p1;
            p1 = p2;
        }
        dp = N * (z * p1 - p0) / (z * z - 1.0L);
        const long double wt = 2.0L / ((1.0L - z * z) * dp * dp);
        x[i] = (double)(-z);
        x[N - 1 - i] = (double)z;
        w[i] = (double)wt;
        w[N - 1 - i] = (double)wt;
    }
    if (N & 1) x[N / 2] = 0.0;
    if (!(a == -1.0 && b == 1.0))
        for (int i = 0; i < N; ++i) {
            x[i] = 0.5 * (b - a) * x[i] + 0.5 * (b + a);
            w[i] = 0.5 * (b - a) * w[i];
        }
}

// bspline_tools.f90:364-373: index of the first breakpoint >= x, minus one
int find_max_n_b(int k, const std::vector<double>& knots, double x)
{
    const int nbp = (int)knots.size() - 2 * k + 2;
    for (int q = 0; q < nbp; ++q)
        if (knots[k - 1 + q] >= x) return q;
    return -1;
}

namespace {

// values and m-th derivatives of the ks splines living on cell v
void spline_derivs(const std::vector<double>& t, int ks, int v, double x, int m, double* out)
{
    const int left = ks - 1 + v;  // 1-based
    if (m == 0) {
        bspline_values_left(t.data(), ks, left, x, out);
        return;
    }
    if (m >= ks) { for (int s = 0; s < ks; ++s) out[s] = 0.0; return; }
    double low[kMaxOrder];
    bspline_values_left(t.data(), ks - m, left, x, low);  // low[q] <-> full index v+m+q
    auto T = [&](int j) { return t[j - 1]; };             // 1-based knot access
    for (int s = 0; s < ks; ++s) {
        // coefficients of the unit vector e_{v+s} differenced m times
        double c[kMaxOrder + 1];
        for (int j = 0; j < ks; ++j) c[j] = (j == s) ? 1.0 : 0.0;  // c[j] <-> full index v+j
        int lo = 0;  // first valid entry after each differencing step
        for (int d = 1; d <= m; ++d) {
            // order after this step: ks-d; c_d[j] = (ks-d)(c[j]-c[j-1])/(t_{j+ks-d}-t_j)
            for (int j = ks - 1; j >= lo + 1; --j) {
                const int full = v + j;
                c[j] = (ks - d) * (c[j] - c[j - 1]) / (T(full + ks - d) - T(full));
            }
            lo += 1;
        }
        double acc = 0.0;
        for (int j = m; j < ks; ++j) acc += c[j] * low[j - m];
        out[s] = acc;
    }
}

}  // namespace

// mat_els.f90:85-118,329-346
void setup_S(int ks, const std::vector<double>& t, int k_GL, std::complex<double>* S)
{
    const int n = (int)t.size() - ks, nb = n - 2, cells = n - ks + 1;
    for (long long q = 0; q < (long long)nb * nb; ++q) S[q] = 0.0;
    std::vector<double> x(k_GL), w(k_GL);
    double B[kMaxOrder];
    for (int v = 1; v <= cells; ++v) {
        gauss_legendre(k_GL, t[ks - 1 + v - 1], t[ks - 1 + v], x.data(), w.data());
        for (int q = 0; q < k_GL; ++q) {
            bspline_values(t.data(), ks, v, x[q], B);
            for (int s2 = 0; s2 < ks; ++s2)
                for (int s = 0; s < ks; ++s) {
                    const int i = v + s - 1, j = v + s2 - 1;
                    if (i < 1 || i > nb || j < 1 || j > nb) continue;
                    S[(i - 1) + (size_t)nb * (j - 1)] += w[q] * B[s] * B[s2];
                }
        }
    }
}

// mat_els.f90:120-170 setup_radial_dip with :348-390: gauge 'l': A = r_mat = int B_i r B_j;
// gauge 'v': A = dr_mat = -i int B_i B_j', B = r_inv_mat = -i int B_i B_j / r
void setup_radial_dip(int ks, const std::vector<double>& t, int k_GL, int gauge, std::complex<double>* A,
                      std::complex<double>* Bm)
{
    if (gauge != 'l' && gauge != 'v') throw std::invalid_argument("setup_radial_dip: gauge must be 'l' or 'v'");
    const int n = (int)t.size() - ks, nb = n - 2, cells = n - ks + 1;
    for (long long q = 0; q < (long long)nb * nb; ++q) A[q] = 0.0;
    if (gauge == 'v') for (long long q = 0; q < (long long)nb * nb; ++q) Bm[q] = 0.0;
    std::vector<double> x(k_GL), w(k_GL);
    double B[kMaxOrder], D1[kMaxOrder];
    const std::complex<double> mi(0.0, -1.0);
    for (int v = 1; v <= cells; ++v) {
        gauss_legendre(k_GL, t[ks - 1 + v - 1], t[ks - 1 + v], x.data(), w.data());
        for (int q = 0; q < k_GL; ++q) {
            const double r = x[q];
            bspline_values(t.data(), ks, v, r, B);
            if (gauge == 'v') spline_derivs(t, ks, v, r, 1, D1);
            for (int s2 = 0; s2 < ks; ++s2)
                for (int s = 0; s < ks; ++s) {
                    const int i = v + s - 1, j = v + s2 - 1;
                    if (i < 1 || i > nb || j < 1 || j > nb) continue;
                    const size_t at = (size_t)(i - 1) + (size_t)nb * (j - 1);
                    if (gauge == 'l') {
                        A[at] += w[q] * r * B[s] * B[s2];
                    } else {
                        A[at] += mi * (w[q] * B[s] * D1[s2]);
                        Bm[at] += mi * (w[q] * B[s] * B[s2] / r);
                    }
                }
        }
    }
}

// mat_els.f90:47-83,294-327; V(r,l) = l(l+1)/(2 r^2) - Z/r; CAP = -i eta (r-r0)^order
void setup_H_one_particle(int ks, const std::vector<double>& t, int Z, int l, int CAP_order,
                          double CAP_r_0, std::complex<double> CAP_eta, int k_GL,
                          std::complex<double>* H)
{
    const int n = (int)t.size() - ks, nb = n - 2, cells = n - ks + 1;
    for (long long q = 0; q < (long long)nb * nb; ++q) H[q] = 0.0;
    std::vector<double> x(k_GL), w(k_GL);
    double B[kMaxOrder], D2[kMaxOrder];
    const std::complex<double> mi(0.0, -1.0);
    for (int v = 1; v <= cells; ++v) {
        gauss_legendre(k_GL, t[ks - 1 + v - 1], t[ks - 1 + v], x.data(), w.data());
        for (int q = 0; q < k_GL; ++q) {
            const double r = x[q];
            bspline_values(t.data(), ks, v, r, B);
            spline_derivs(t, ks, v, r, 2, D2);
            const double V = 0.5 * l * (l + 1) / (r * r) - (double)Z / r;
            std::complex<double> Vc = 0.0;
            if (r >= CAP_r_0) Vc = mi * CAP_eta * powi(r - CAP_r_0, CAP_order);
            for (int s2 = 0; s2 < ks; ++s2)
                for (int s = 0; s < ks; ++s) {
                    const int i = v + s - 1, j = v + s2 - 1;
                    if (i < 1 || i > nb || j < 1 || j > nb) continue;
                    H[(i - 1) + (size_t)nb * (j - 1)] +=
                        w[q] * (-0.5 * B[s] * D2[s2] + (V + Vc) * B[s] * B[s2]);
                }
        }
    }
}

// orbital_tools.f90:46-72
static bool consistent(int l1, int l2, int L, bool pi, bool eqv)
{
    if (eqv && (L % 2 != 0)) return false;
    if (!(std::abs(l1 - l2) <= L && L <= l1 + l2)) return false;
    return (((l1 + l2) % 2) != 0) == pi;
}

// orbital_tools.f90:119-216 (the unused one-particle-energy bookkeeping dropped)
void count_configs(int L, bool pi, int max_l_1p, int n_b, int k_spline, int max_n_b, int n_all_l,
                   int l_2_max, std::vector<int64_t>& conf_n, std::vector<int64_t>& conf_l,
                   std::vector<int64_t>& conf_eqv)
{
    conf_n.clear();
    conf_l.clear();
    conf_eqv.clear();
    for (int li = 0; li <= max_l_1p; ++li)
        for (int lj = 0; lj <= li; ++lj)
            for (int ni = std::min(li + 1, k_spline - 1); ni <= n_b; ++ni) {
                if (ni > n_all_l && lj > l_2_max) continue;
                const int hi = (lj == li) ? std::min(ni, max_n_b) : std::min(n_b, max_n_b);
                for (int nj = std::min(lj + 1, k_spline - 1); nj <= hi; ++nj) {
                    const bool eqv = (li == lj) && (ni == nj);
                    if (!consistent(li, lj, L, pi, eqv)) continue;
                    conf_n.push_back(ni);
                    conf_n.push_back(nj);
                    conf_l.push_back(li);
                    conf_l.push_back(lj);
                    conf_eqv.push_back(eqv ? 1 : 0);
                }
            }
}

// orbital_tools.f90:245-343, two_el = .true.
std::vector<SymLabel> basis_syms(int max_L, bool z_pol)
{
    std::vector<SymLabel> s;
    s.push_back(SymLabel{0, 0, false});
    for (int l = 1; l <= max_L; ++l) {
        if (z_pol) {
            s.push_back(SymLabel{l, 0, (l % 2) != 0});
        } else {
            for (int p = 0; p <= 1; ++p)
                for (int m = -l; m <= l; ++m) {
                    if (std::abs(m % 2) != p) continue;
                    s.push_back(SymLabel{l, m, p == 1});
                }
        }
    }
    return s;
}

}  // namespace host
}  // namespace bs2e
