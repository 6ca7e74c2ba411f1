/*
 * bs2e_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * A loop-by-loop restatement in C of the two-electron hot path of
 * edvinolo/b-spline-two-e (reference files cited per function, paths relative
 * to the reference root).  See bs2e_oracle.h for the usage rules and for the
 * "parity unpinned" statement.
 *
 * Third-party arithmetic that is not under the reference tree:
 *   - fortran-lang/stdlib (unpinned, git HEAD): stdlib_quadrature::gauss_legendre
 *     -> restated from its published algorithm (Newton on P_n from Chebyshev
 *     guesses), call site src/tools/quad_tools.f90:26.
 *   - GSL (unpinned distro package): gsl_sf_coupling_3j / _6j -> replaced by
 *     exact integer arithmetic (Racah single-sum formula), call sites
 *     src/tools/wigner_tools.f90:43,59.
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off -fopenmp).
 */
#include "bs2e_oracle.h"

#include <complex.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef double complex zcplx;

static int64_t iabs64(int64_t a) { return a < 0 ? -a : a; }
static int64_t imin64(int64_t a, int64_t b) { return a < b ? a : b; }
static int64_t imax64(int64_t a, int64_t b) { return a > b ? a : b; }

/* gfortran lowers r**k (integer k) to __builtin_powi == libgcc __powidf2:
 * square-and-multiply from the low bit.  Restated so rounding matches.      */
static double powi(double x, int64_t m)
{
    uint64_t n = (uint64_t)(m < 0 ? -m : m);
    double y = (n & 1) ? x : 1.0;
    while (n >>= 1) {
        x = x * x;
        if (n & 1) y *= x;
    }
    return m < 0 ? 1.0 / y : y;
}

int64_t orc_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* The launcher of a multi-rank job (torchrun) exports OMP_NUM_THREADS=1; the timed
 * CPU baseline states and sets its own thread count instead of inheriting that.  */
void orc_set_threads(int64_t n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads((int)n);
#else
    (void)n;
#endif
}

/* ======================================================================== */
/* grid_tools.f90:6-55  generate_grid                                        */
/* ======================================================================== */
int64_t orc_generate_grid(int64_t k, int64_t m, int64_t Z, double h_max,
                          double r_max, double *out, int64_t cap)
{
    double h = pow(2.0, (double)(-m));
    /* the reference sizes its work array 2k+m+ceil(r_max/h); we only need
     * "big enough" and grow on demand */
    int64_t cap_w = 2 * k + m + (int64_t)ceil(r_max / h) + 64;
    double *g = (double *)malloc(sizeof(double) * (size_t)(cap_w + 1)); /* 1-based */
    int64_t i;
    for (i = 1; i <= k; ++i) g[i] = 0.0;
    for (i = k + 1; i <= k + m; ++i) g[i] = g[i - 1] + h;
    i = k + m;
    for (;;) {
        double next = g[i] * (1.0 + h);
        if ((next - g[i]) >= h_max * (double)Z) break;
        g[i + 1] = next;
        i = i + 1;
    }
    while (g[i] < (double)Z * r_max) {
        g[i + 1] = g[i] + (double)Z * h_max;
        i = i + 1;
    }
    for (int64_t q = i + 1; q <= i + k - 1; ++q) g[q] = g[i];
    int64_t N = i + k - 1;
    if (N <= cap)
        for (int64_t q = 1; q <= N; ++q) out[q - 1] = g[q] / (double)Z;
    free(g);
    return N;
}

/* ======================================================================== */
/* stdlib_quadrature::gauss_legendre (published algorithm), used through     */
/* quad_tools.f90:14-27 setup_GL                                             */
/* ======================================================================== */
static double legendre_p(int64_t n, double x)
{
    if (n == 0) return 1.0;
    if (n == 1) return x;
    double d1 = x, d2 = 1.0, leg = 0.0;
    for (int64_t i = 2; i <= n; ++i) {
        leg = (2 * i - 1) * x * d1 / i - (i - 1) * d2 / i;
        d2 = d1;
        d1 = leg;
    }
    return leg;
}
static double dlegendre_p(int64_t n, double x)
{
    if (n == 0) return 0.0;
    if (n == 1) return 1.0;
    return n * (x * legendre_p(n, x) - legendre_p(n - 1, x)) / (x * x - 1.0);
}

void orc_gauss_legendre(int64_t N, double a, double b, double *x, double *w)
{
    const double pi = 3.14159265358979323846264338327950288;
    const double tol = 4.0 * 2.220446049250313e-16;
    int64_t n = N - 1;
    if (n == 0) {
        x[0] = 0.0; w[0] = 2.0;
    } else if (n == 1) {
        x[0] = -sqrt(1.0 / 3.0); x[1] = -x[0];
        w[0] = 1.0; w[1] = 1.0;
    } else {
        for (int64_t i = 0; i <= (n + 1) / 2 - 1; ++i) {
            double xi = -cos((2 * i + 1) / (2.0 * n + 2.0) * pi);
            for (int it = 0; it < 100; ++it) {
                double leg = legendre_p(n + 1, xi);
                double dleg = dlegendre_p(n + 1, xi);
                double delta = -leg / dleg;
                xi = xi + delta;
                if (fabs(delta) <= tol * fabs(xi)) break;
            }
            x[i] = xi;
            x[n - i] = -xi;
            double dleg = dlegendre_p(n + 1, xi);
            w[i] = 2.0 / ((1.0 - xi * xi) * dleg * dleg);
            w[n - i] = w[i];
        }
        if (n % 2 == 0) {
            x[n / 2] = 0.0;
            double dleg = dlegendre_p(n + 1, 0.0);
            w[n / 2] = 2.0 / (dleg * dleg);
        }
    }
    if (!(a == -1.0 && b == 1.0)) {
        for (int64_t i = 0; i < N; ++i) {
            x[i] = 0.5 * (b - a) * x[i] + 0.5 * (b + a);
            w[i] = 0.5 * (b - a) * w[i];
        }
    }
}

/* ======================================================================== */
/* bspline_tools.f90:27-56  init / init_support                              */
/* ======================================================================== */
orc_bspline *orc_bspline_new(int64_t k, int64_t nt, const double *t)
{
    orc_bspline *bs = (orc_bspline *)calloc(1, sizeof(orc_bspline));
    bs->k = k;
    bs->nt = nt;
    bs->n = nt - k;
    bs->n_b = bs->n - 2;
    bs->t = (double *)malloc(sizeof(double) * (size_t)nt);
    memcpy(bs->t, t, sizeof(double) * (size_t)nt);
    /* breakpoints = t(k : size(t)-k+1)  (1-based, inclusive) */
    bs->nbp = (nt - k + 1) - k + 1;
    bs->bp = (double *)malloc(sizeof(double) * (size_t)bs->nbp);
    for (int64_t q = 0; q < bs->nbp; ++q) bs->bp[q] = t[k - 1 + q];
    return bs;
}
void orc_bspline_free(orc_bspline *bs)
{
    if (!bs) return;
    free(bs->t);
    free(bs->bp);
    free(bs);
}
int64_t orc_bspline_cells(const orc_bspline *bs) { return bs->nbp - 1; }
int64_t orc_bspline_nb(const orc_bspline *bs) { return bs->n_b; }

/* support(iv, j) is .true. for j = iv .. iv+k-1 (bspline_tools.f90:50-54);
 * iv is a 1-based cell, j a 1-based FULL spline index.                      */
static int support(const orc_bspline *bs, int64_t iv, int64_t j)
{
    return (j >= iv) && (j <= iv + bs->k - 1);
}

/* bspline_tools.f90:151-224  BVALUE_D, branch "present(iv)" (m_flag = 0).
 * a(1:n) coefficient vector (0-based here), 1-based knot arithmetic kept.   */
double orc_bvalue(const orc_bspline *bs, const double *a, double x,
                  int64_t i_deriv, int64_t iv)
{
    const int64_t k = bs->k;
    const double *t = bs->t - 1; /* t[1..nt] */
    double a_j[21], d_p[21], d_m[21];
    int64_t k_m_i_der = k - i_deriv;
    if (k_m_i_der <= 0) return 0.0;
    int64_t k_m_1 = k - 1;
    int64_t i = k - 1 + iv;
    int64_t i_m_k = i - k;
    for (int64_t j = 1; j <= k; ++j) a_j[j] = a[i_m_k + j - 1];
    for (int64_t j = 1; j <= i_deriv; ++j) {
        int64_t k_m_j = k - j;
        for (int64_t jj = 1; jj <= k_m_j; ++jj) {
            int64_t i_h_i = i + jj;
            int64_t i_h_m = i_h_i - k_m_j;
            a_j[jj] = (a_j[jj + 1] - a_j[jj]) / (t[i_h_i] - t[i_h_m]) * (double)k_m_j;
        }
    }
    int64_t i_p_1 = i + 1;
    for (int64_t j = 1; j <= k_m_i_der; ++j) {
        d_p[j] = t[i + j] - x;
        d_m[j] = x - t[i_p_1 - j];
    }
    for (int64_t j = i_deriv + 1; j <= k_m_1; ++j) {
        int64_t k_m_j = k - j;
        int64_t ii = k_m_j;
        for (int64_t jj = 1; jj <= k_m_j; ++jj) {
            a_j[jj] = (a_j[jj + 1] * d_m[ii] + a_j[jj] * d_p[jj]) / (d_m[ii] + d_p[jj]);
            ii = ii - 1;
        }
    }
    return a_j[1];
}

/* value of the single B-spline with b-index i_b (unit coefficient at full
 * index i_b+1, mat_els.f90:204,209) */
static double bspl_unit(const orc_bspline *bs, double *cwork, int64_t i_b,
                        double x, int64_t i_deriv, int64_t iv)
{
    cwork[i_b] = 1.0; /* c(i_b+1) 1-based == cwork[i_b] 0-based */
    double v = orc_bvalue(bs, cwork, x, i_deriv, iv);
    cwork[i_b] = 0.0;
    return v;
}

/* bspline_tools.f90:364-373  find_max_n_b: findloc(breakpoints>=x,.true.)-1 */
int64_t orc_find_max_n_b(const orc_bspline *bs, double x)
{
    for (int64_t q = 0; q < bs->nbp; ++q)
        if (bs->bp[q] >= x) return (q + 1) - 1;
    return -1; /* findloc returns 0 -> -1 */
}

/* ======================================================================== */
/* mat_els.f90:85-118,329-346  setup_S / compute_S                           */
/* ======================================================================== */
void orc_setup_S(const orc_bspline *bs, int64_t k_GL, double *S_out)
{
    const int64_t nb = bs->n_b, cells = bs->nbp - 1;
    zcplx *S = (zcplx *)S_out;
    for (int64_t q = 0; q < nb * nb; ++q) S[q] = 0.0;
    double *x = (double *)malloc(sizeof(double) * (size_t)(k_GL * cells));
    double *w = (double *)malloc(sizeof(double) * (size_t)(k_GL * cells));
    for (int64_t c = 0; c < cells; ++c) /* quad_tools.f90:29-52 gau_leg%init */
        orc_gauss_legendre(k_GL, bs->bp[c], bs->bp[c + 1], x + c * k_GL, w + c * k_GL);
    double *cw = (double *)calloc((size_t)bs->n, sizeof(double));
    for (int64_t j_b = 1; j_b <= nb; ++j_b)
        for (int64_t i_b = 1; i_b <= nb; ++i_b) {
            if (iabs64(j_b - i_b) >= bs->k) continue;
            for (int64_t i_r = 1; i_r <= cells; ++i_r) {
                if (!(support(bs, i_r, i_b + 1) && support(bs, i_r, j_b + 1))) continue;
                const double *r = x + (i_r - 1) * k_GL, *ww = w + (i_r - 1) * k_GL;
                for (int64_t q = 0; q < k_GL; ++q) {
                    double B_i = bspl_unit(bs, cw, i_b, r[q], 0, i_r);
                    double B_j = bspl_unit(bs, cw, j_b, r[q], 0, i_r);
                    S[(i_b - 1) + nb * (j_b - 1)] += ww[q] * B_i * B_j;
                }
            }
        }
    free(cw); free(x); free(w);
}

/* mat_els.f90:47-83,294-327 setup_H_one_particle / compute_H with
 * potentials.f90:35-43 (hydrogenic) and CAP_tools.f90:24-34               */
void orc_setup_H_one_particle(const orc_bspline *bs, int64_t Z, int64_t l,
                              int64_t CAP_order, double CAP_r_0,
                              double CAP_eta_re, double CAP_eta_im,
                              int64_t k_GL, double *H_out)
{
    const int64_t nb = bs->n_b, cells = bs->nbp - 1;
    zcplx *H = (zcplx *)H_out;
    const zcplx eta = CAP_eta_re + I * CAP_eta_im;
    for (int64_t q = 0; q < nb * nb; ++q) H[q] = 0.0;
    double *x = (double *)malloc(sizeof(double) * (size_t)(k_GL * cells));
    double *w = (double *)malloc(sizeof(double) * (size_t)(k_GL * cells));
    for (int64_t c = 0; c < cells; ++c)
        orc_gauss_legendre(k_GL, bs->bp[c], bs->bp[c + 1], x + c * k_GL, w + c * k_GL);
    double *cw = (double *)calloc((size_t)bs->n, sizeof(double));
    for (int64_t j_b = 1; j_b <= nb; ++j_b)
        for (int64_t i_b = 1; i_b <= nb; ++i_b) {
            if (iabs64(j_b - i_b) >= bs->k) continue;
            for (int64_t i_r = 1; i_r <= cells; ++i_r) {
                if (!(support(bs, i_r, i_b + 1) && support(bs, i_r, j_b + 1))) continue;
                const double *r = x + (i_r - 1) * k_GL, *ww = w + (i_r - 1) * k_GL;
                for (int64_t q = 0; q < k_GL; ++q) {
                    double B_i = bspl_unit(bs, cw, i_b, r[q], 0, i_r);
                    double B_j = bspl_unit(bs, cw, j_b, r[q], 0, i_r);
                    double D_B_j = bspl_unit(bs, cw, j_b, r[q], 2, i_r);
                    double V = 0.5 * (double)l * (double)(l + 1) / (r[q] * r[q]) - (double)Z / r[q];
                    zcplx Vc = 0.0;
                    if (r[q] >= CAP_r_0)
                        Vc = (0.0 - 1.0 * I) * eta * powi(r[q] - CAP_r_0, CAP_order);
                    H[(i_b - 1) + nb * (j_b - 1)] +=
                        ww[q] * (-0.5 * B_i * D_B_j + (V + Vc) * B_i * B_j);
                }
            }
        }
    free(cw); free(x); free(w);
}

/* ======================================================================== */
/* sparse_array_tools.f90:276-366  counts                                    */
/* ======================================================================== */
int64_t orc_count_nnz_4d(const orc_bspline *bs)
{
    const int64_t nb = bs->n_b, cells = bs->nbp - 1;
    int64_t nnz = 0;
    for (int64_t j = 1; j <= nb; ++j)
        for (int64_t i = 1; i <= nb; ++i) {
            if (iabs64(i - j) >= bs->k) continue;
            for (int64_t iv = 1; iv <= cells; ++iv)
                if (support(bs, iv, i + 1) && support(bs, iv, j + 1)) nnz++;
        }
    return nnz;
}

/* number of cells where all of the given full indices have support */
static int64_t common_cells(const orc_bspline *bs, int64_t fmin, int64_t fmax)
{
    const int64_t cells = bs->nbp - 1;
    int64_t lo = imax64(1, fmax - bs->k + 1), hi = imin64(cells, fmin);
    return hi >= lo ? hi - lo + 1 : 0;
}

int64_t orc_count_nnz_6d(const orc_bspline *bs)
{
    /* same loop nest as count_nnz_6d; the innermost cell loop is replaced by
     * its closed form (the count is an integer, no arithmetic to mirror)    */
    const int64_t nb = bs->n_b;
    int64_t nnz = 0;
    for (int64_t j_p = 1; j_p <= nb; ++j_p)
        for (int64_t j = 1; j <= nb; ++j) {
            if (iabs64(j - j_p) >= bs->k) continue;
            for (int64_t i_p = 1; i_p <= nb; ++i_p)
                for (int64_t i = imax64(1, i_p - bs->k + 1); i <= imin64(nb, i_p + bs->k - 1); ++i) {
                    int64_t fmin = imin64(imin64(i, i_p), imin64(j, j_p)) + 1;
                    int64_t fmax = imax64(imax64(i, i_p), imax64(j, j_p)) + 1;
                    nnz += common_cells(bs, fmin, fmax);
                }
        }
    return nnz;
}

int64_t orc_num_pairs(const orc_bspline *bs)
{
    const int64_t nb = bs->n_b;
    int64_t P = 0;
    for (int64_t i_p = 1; i_p <= nb; ++i_p)
        for (int64_t i = 1; i <= nb; ++i)
            if (iabs64(i - i_p) < bs->k) P++;
    return P;
}

int64_t orc_count_nnz_R_k(const orc_bspline *bs)
{
    int64_t P = orc_num_pairs(bs);
    return P * P;
}

/* 0-based index of the ordered band pair (a,c), |a-c| < k; pairs are numbered
 * a-major, c ascending.  (Only an addressing scheme for the dense stand-in of
 * the reference hash map.)                                                   */
int64_t orc_pair_index(const orc_bspline *bs, int64_t a, int64_t c)
{
    const int64_t nb = bs->n_b, w = bs->k - 1;
    if (a < 1 || a > nb || c < 1 || c > nb || iabs64(a - c) > w) return -1;
    int64_t off = 0;
    for (int64_t q = 1; q < a; ++q)
        off += imin64(nb, q + w) - imax64(1, q - w) + 1;
    return off + (c - imax64(1, a - w));
}

/* ======================================================================== */
/* mat_els.f90:184-228 setup_Slater_off_diag + :392-439 compute_..._off_diag */
/* ======================================================================== */
void orc_setup_Slater_off_diag(const orc_bspline *bs, int64_t max_k, int64_t k_GL,
                               double *r_k, double *r_m_k,
                               int64_t *iv_a, int64_t *i_a, int64_t *j_a)
{
    const int64_t nb = bs->n_b, cells = bs->nbp - 1;
    const int64_t nnz = orc_count_nnz_4d(bs);
    double *x = (double *)malloc(sizeof(double) * (size_t)k_GL);
    double *w = (double *)malloc(sizeof(double) * (size_t)k_GL);
    orc_gauss_legendre(k_GL, -1.0, 1.0, x, w);
    memset(r_k, 0, sizeof(double) * (size_t)(nnz * (max_k + 1)));
    memset(r_m_k, 0, sizeof(double) * (size_t)(nnz * (max_k + 1)));

#pragma omp parallel for schedule(static)
    for (int64_t k = 0; k <= max_k; ++k) {
        double *cw = (double *)calloc((size_t)bs->n, sizeof(double));
        double *r = (double *)malloc(sizeof(double) * (size_t)k_GL);
        double *B_i = (double *)malloc(sizeof(double) * (size_t)k_GL);
        double *B_j = (double *)malloc(sizeof(double) * (size_t)k_GL);
        int64_t ptr = 0;
        for (int64_t j_b = 1; j_b <= nb; ++j_b)
            for (int64_t i_b = 1; i_b <= nb; ++i_b) {
                if (iabs64(j_b - i_b) >= bs->k) continue;
                for (int64_t i_r = 1; i_r <= cells; ++i_r) {
                    if (!(support(bs, i_r, i_b + 1) && support(bs, i_r, j_b + 1))) continue;
                    ptr = ptr + 1;
                    double lim1 = bs->bp[i_r - 1], lim2 = bs->bp[i_r];
                    double scale = 0.5 * (lim2 - lim1);
                    double translate = 0.5 * (lim2 + lim1);
                    for (int64_t q = 0; q < k_GL; ++q) {
                        r[q] = scale * x[q] + translate;
                        B_i[q] = bspl_unit(bs, cw, i_b, r[q], 0, i_r);
                        B_j[q] = bspl_unit(bs, cw, j_b, r[q], 0, i_r);
                    }
                    double acc = 0.0, accm = 0.0;
                    for (int64_t q = 0; q < k_GL; ++q) {
                        acc = acc + w[q] * B_i[q] * B_j[q] * powi(r[q], k);
                        accm = accm + w[q] * B_i[q] * B_j[q] / powi(r[q], k + 1);
                    }
                    r_k[(ptr - 1) + nnz * k] = scale * acc;
                    r_m_k[(ptr - 1) + nnz * k] = scale * accm;
                    if (k == 0) {
                        iv_a[ptr - 1] = i_r;
                        i_a[ptr - 1] = i_b;
                        j_a[ptr - 1] = j_b;
                    }
                }
            }
        free(cw); free(r); free(B_i); free(B_j);
    }
    free(x); free(w);
}

/* ======================================================================== */
/* mat_els.f90:230-292 setup_Slater_diag + :441-491 compute_Slater_diag      */
/* ======================================================================== */
typedef struct {
    const orc_bspline *bs;
    int64_t k_GL;
    const double *x, *w;
    /* optional tables (tabulate != 0): values for local spline s (full index
     * iv+s) on cell iv at outer node q / inner node (q,p)                   */
    const double *Bout;  /* [cell][q][s]      */
    const double *Bin;   /* [cell][q][p][s]   */
} diag_ctx;

static double diag_entry(const diag_ctx *c, double *cw, int64_t i_r,
                         const int64_t idx[4], int64_t k)
{
    const orc_bspline *bs = c->bs;
    const int64_t k_GL = c->k_GL, ks = bs->k;
    const double lim1 = bs->bp[i_r - 1], lim2 = bs->bp[i_r];
    const double scale_i = 0.5 * (lim2 - lim1);
    const double translate_i = 0.5 * (lim2 + lim1);
    double acc = 0.0;
    /* local slots: full index f = b+1, slot s = f - i_r */
    const int64_t s1 = idx[0] + 1 - i_r, s2 = idx[1] + 1 - i_r;
    const int64_t s3 = idx[2] + 1 - i_r, s4 = idx[3] + 1 - i_r;
    for (int64_t q = 0; q < k_GL; ++q) {
        double r = scale_i * c->x[q] + translate_i;
        double B_i, B_i_p;
        if (c->Bout) {
            const double *bo = c->Bout + ((i_r - 1) * k_GL + q) * ks;
            B_i = bo[s1]; B_i_p = bo[s2];
        } else {
            B_i = bspl_unit(bs, cw, idx[0], r, 0, i_r);
            B_i_p = bspl_unit(bs, cw, idx[1], r, 0, i_r);
        }
        double scale_j = 0.5 * (r - lim1);
        double translate_j = 0.5 * (r + lim1);
        double int_j = 0.0;
        for (int64_t p = 0; p < k_GL; ++p) {
            double r_j = scale_j * c->x[p] + translate_j;
            double B_j, B_j_p;
            if (c->Bin) {
                const double *bi = c->Bin + (((i_r - 1) * k_GL + q) * k_GL + p) * ks;
                B_j = bi[s3]; B_j_p = bi[s4];
            } else {
                B_j = bspl_unit(bs, cw, idx[2], r_j, 0, i_r);
                B_j_p = bspl_unit(bs, cw, idx[3], r_j, 0, i_r);
            }
            int_j = int_j + c->w[p] * B_j * B_j_p * powi(r_j, k);
        }
        int_j = scale_j * int_j;
        acc = acc + c->w[q] * B_i * B_i_p * int_j / powi(r, k + 1);
    }
    return scale_i * acc;
}

void orc_setup_Slater_diag(const orc_bspline *bs, int64_t max_k, int64_t k_GL,
                           double *r_d_k, int64_t *iv_a, int64_t *i_a, int64_t *j_a,
                           int64_t *ip_a, int64_t *jp_a,
                           int64_t tabulate, int64_t par_mode)
{
    const int64_t nb = bs->n_b, cells = bs->nbp - 1, ks = bs->k;
    const int64_t nnz = orc_count_nnz_6d(bs);
    double *x = (double *)malloc(sizeof(double) * (size_t)k_GL);
    double *w = (double *)malloc(sizeof(double) * (size_t)k_GL);
    orc_gauss_legendre(k_GL, -1.0, 1.0, x, w);

    double *Bout = NULL, *Bin = NULL;
    if (tabulate) {
        Bout = (double *)calloc((size_t)(cells * k_GL * ks), sizeof(double));
        Bin = (double *)calloc((size_t)(cells * k_GL * k_GL * ks), sizeof(double));
#pragma omp parallel for schedule(dynamic)
        for (int64_t i_r = 1; i_r <= cells; ++i_r) {
            double *cw = (double *)calloc((size_t)bs->n, sizeof(double));
            const double lim1 = bs->bp[i_r - 1], lim2 = bs->bp[i_r];
            const double scale_i = 0.5 * (lim2 - lim1), translate_i = 0.5 * (lim2 + lim1);
            for (int64_t q = 0; q < k_GL; ++q) {
                double r = scale_i * x[q] + translate_i;
                double scale_j = 0.5 * (r - lim1), translate_j = 0.5 * (r + lim1);
                for (int64_t s = 0; s < ks; ++s) {
                    int64_t b = i_r + s - 1; /* b-index of full index i_r+s */
                    if (b < 1 || b > nb) continue;
                    Bout[((i_r - 1) * k_GL + q) * ks + s] = bspl_unit(bs, cw, b, r, 0, i_r);
                    for (int64_t p = 0; p < k_GL; ++p) {
                        double r_j = scale_j * x[p] + translate_j;
                        Bin[(((i_r - 1) * k_GL + q) * k_GL + p) * ks + s] =
                            bspl_unit(bs, cw, b, r_j, 0, i_r);
                    }
                }
            }
            free(cw);
        }
    }
    diag_ctx ctx = { bs, k_GL, x, w, Bout, Bin };

    /* ptr at the start of every j_b_p iteration (needed only to let more
     * than max_k+1 threads work; the sequence of ptr values is unchanged)   */
    int64_t *ptr0 = (int64_t *)malloc(sizeof(int64_t) * (size_t)(nb + 2));
    {
        int64_t ptr = 0;
        for (int64_t j_b_p = 1; j_b_p <= nb; ++j_b_p) {
            ptr0[j_b_p] = ptr;
            for (int64_t j_b = 1; j_b <= nb; ++j_b) {
                if (iabs64(j_b - j_b_p) >= ks) continue;
                for (int64_t i_b_p = 1; i_b_p <= nb; ++i_b_p)
                    for (int64_t i_b = imax64(1, i_b_p - ks + 1); i_b <= imin64(nb, i_b_p + ks - 1); ++i_b) {
                        int64_t fmin = imin64(imin64(i_b, i_b_p), imin64(j_b, j_b_p)) + 1;
                        int64_t fmax = imax64(imax64(i_b, i_b_p), imax64(j_b, j_b_p)) + 1;
                        ptr += common_cells(bs, fmin, fmax);
                    }
            }
        }
        ptr0[nb + 1] = ptr;
    }

    const int64_t n_outer = par_mode ? (max_k + 1) * nb : (max_k + 1);
#pragma omp parallel for schedule(dynamic)
    for (int64_t task = 0; task < n_outer; ++task) {
        const int64_t k = par_mode ? task / nb : task;
        const int64_t jp_lo = par_mode ? task % nb + 1 : 1;
        const int64_t jp_hi = par_mode ? jp_lo : nb;
        double *cw = (double *)calloc((size_t)bs->n, sizeof(double));
        int64_t idx[4];
        int64_t ptr = ptr0[jp_lo];
        for (int64_t j_b_p = jp_lo; j_b_p <= jp_hi; ++j_b_p) {
            idx[3] = j_b_p;
            for (int64_t j_b = 1; j_b <= nb; ++j_b) {
                if (iabs64(j_b - j_b_p) >= ks) continue;
                idx[2] = j_b;
                for (int64_t i_b_p = 1; i_b_p <= nb; ++i_b_p) {
                    idx[1] = i_b_p;
                    for (int64_t i_b = 1; i_b <= nb; ++i_b) {
                        idx[0] = i_b;
                        if (iabs64(i_b - i_b_p) >= ks) continue;
                        for (int64_t i_r = 1; i_r <= cells; ++i_r) {
                            int i_sup = support(bs, i_r, i_b + 1) && support(bs, i_r, i_b_p + 1);
                            int j_sup = support(bs, i_r, j_b + 1) && support(bs, i_r, j_b_p + 1);
                            if (!(i_sup && j_sup)) continue;
                            ptr = ptr + 1;
                            r_d_k[(ptr - 1) + nnz * k] = diag_entry(&ctx, cw, i_r, idx, k);
                            if (k == 0) {
                                iv_a[ptr - 1] = i_r;
                                i_a[ptr - 1] = idx[0];
                                ip_a[ptr - 1] = idx[1];
                                j_a[ptr - 1] = idx[2];
                                jp_a[ptr - 1] = idx[3];
                            }
                        }
                    }
                }
            }
        }
        free(cw);
    }
    free(ptr0); free(x); free(w); free(Bout); free(Bin);
}

/* Timing helper for the CPU baseline: the reference-faithful evaluation
 * (de Boor call per point, threads over k only) of setup_Slater_diag for every
 * jp_step-th value of the outermost index j_b_p.  Returns a checksum so the
 * work cannot be optimised away; *entries receives the number of (entry,k)
 * values computed.                                                          */
double orc_time_Slater_diag_sample(const orc_bspline *bs, int64_t max_k, int64_t k_GL,
                                   int64_t jp_step, int64_t *entries)
{
    const int64_t nb = bs->n_b, cells = bs->nbp - 1, ks = bs->k;
    double *x = (double *)malloc(sizeof(double) * (size_t)k_GL);
    double *w = (double *)malloc(sizeof(double) * (size_t)k_GL);
    orc_gauss_legendre(k_GL, -1.0, 1.0, x, w);
    diag_ctx ctx = { bs, k_GL, x, w, NULL, NULL };
    double total = 0.0;
    int64_t count = 0;
#pragma omp parallel for schedule(static) reduction(+ : total, count)
    for (int64_t k = 0; k <= max_k; ++k) {
        double *cw = (double *)calloc((size_t)bs->n, sizeof(double));
        int64_t idx[4];
        for (int64_t j_b_p = 1; j_b_p <= nb; j_b_p += jp_step) {
            idx[3] = j_b_p;
            for (int64_t j_b = 1; j_b <= nb; ++j_b) {
                if (iabs64(j_b - j_b_p) >= ks) continue;
                idx[2] = j_b;
                for (int64_t i_b_p = 1; i_b_p <= nb; ++i_b_p) {
                    idx[1] = i_b_p;
                    for (int64_t i_b = 1; i_b <= nb; ++i_b) {
                        idx[0] = i_b;
                        if (iabs64(i_b - i_b_p) >= ks) continue;
                        for (int64_t i_r = 1; i_r <= cells; ++i_r) {
                            int i_sup = support(bs, i_r, i_b + 1) && support(bs, i_r, i_b_p + 1);
                            int j_sup = support(bs, i_r, j_b + 1) && support(bs, i_r, j_b_p + 1);
                            if (!(i_sup && j_sup)) continue;
                            total += diag_entry(&ctx, cw, i_r, idx, k);
                            count++;
                        }
                    }
                }
            }
        }
        free(cw);
    }
    free(x); free(w);
    *entries = count;
    return total;
}

/* ======================================================================== */
/* sparse_array_tools.f90:452-493 compute_R_K_map  (+ :495-555 Nd_DOK)       */
/* ======================================================================== */
void orc_compute_R_k_map(const orc_bspline *bs, int64_t max_k,
                         int64_t nnz4, const double *r_k, const double *r_m_k,
                         const int64_t *iv4, const int64_t *i4, const int64_t *j4,
                         int64_t nnz6, const double *r_d_k,
                         const int64_t *i6, const int64_t *j6,
                         const int64_t *ip6, const int64_t *jp6,
                         double *R)
{
    const int64_t P = orc_num_pairs(bs), K1 = max_k + 1;
    memset(R, 0, sizeof(double) * (size_t)(P * P * K1));
    int64_t *pn = (int64_t *)malloc(sizeof(int64_t) * (size_t)nnz4);
    for (int64_t n = 0; n < nnz4; ++n) pn[n] = orc_pair_index(bs, i4[n], j4[n]);

    /* key [i(n), i(m), j(n), j(m)]: electron-1 pair = entry n, electron-2
     * pair = entry m.  Serial in the reference; serial here.                */
    for (int64_t n = 0; n < nnz4; ++n)
        for (int64_t m = 0; m < nnz4; ++m) {
            double *dst = R + (pn[n] * P + pn[m]) * K1;
            if (iv4[n] < iv4[m]) {
                for (int64_t k = 0; k < K1; ++k)
                    dst[k] = dst[k] + r_k[n + nnz4 * k] * r_m_k[m + nnz4 * k];
            } else if (iv4[n] > iv4[m]) {
                for (int64_t k = 0; k < K1; ++k)
                    dst[k] = dst[k] + r_k[m + nnz4 * k] * r_m_k[n + nnz4 * k];
            }
        }
    free(pn);

    for (int64_t n = 0; n < nnz6; ++n) {
        /* set_val([i, j, i_p, j_p]) then set_val([j, i, j_p, i_p]) */
        int64_t p_i = orc_pair_index(bs, i6[n], ip6[n]);
        int64_t p_j = orc_pair_index(bs, j6[n], jp6[n]);
        double *d1 = R + (p_i * P + p_j) * K1;
        for (int64_t k = 0; k < K1; ++k) d1[k] = d1[k] + r_d_k[n + nnz6 * k];
        double *d2 = R + (p_j * P + p_i) * K1;
        for (int64_t k = 0; k < K1; ++k) d2[k] = d2[k] + r_d_k[n + nnz6 * k];
    }
}

int orc_R_get_val(const orc_bspline *bs, int64_t max_k, const double *R,
                  int64_t a, int64_t b, int64_t c, int64_t d, double *vals)
{
    const int64_t P = orc_num_pairs(bs), K1 = max_k + 1;
    int64_t p1 = orc_pair_index(bs, a, c), p2 = orc_pair_index(bs, b, d);
    if (p1 < 0 || p2 < 0) return -1;
    memcpy(vals, R + (p1 * P + p2) * K1, sizeof(double) * (size_t)K1);
    return 0;
}

/* ======================================================================== */
/* Exact Wigner 3j(000) / 6j  (stand-in for GSL behind wigner_tools.f90)     */
/* ======================================================================== */
#define BIG_LIMBS 48 /* 1536 bits: enough for every factorial ratio with
                        argument sums up to ~300 */
typedef struct { uint32_t d[BIG_LIMBS]; int n; } bigu;

static void big_set(bigu *a, uint32_t v) { memset(a, 0, sizeof(*a)); a->d[0] = v; a->n = v ? 1 : 0; }
static void big_mul_small(bigu *a, uint32_t m)
{
    uint64_t carry = 0;
    for (int i = 0; i < a->n; ++i) {
        uint64_t v = (uint64_t)a->d[i] * m + carry;
        a->d[i] = (uint32_t)v;
        carry = v >> 32;
    }
    if (carry) {
        if (a->n >= BIG_LIMBS) { fprintf(stderr, "oracle bigint overflow\n"); abort(); }
        a->d[a->n++] = (uint32_t)carry;
    }
}
static void big_div_small(bigu *a, uint32_t m) /* exact division expected */
{
    uint64_t rem = 0;
    for (int i = a->n - 1; i >= 0; --i) {
        uint64_t v = (rem << 32) | a->d[i];
        a->d[i] = (uint32_t)(v / m);
        rem = v % m;
    }
    if (rem != 0) { fprintf(stderr, "oracle bigint inexact division\n"); abort(); }
    while (a->n > 0 && a->d[a->n - 1] == 0) a->n--;
}
static int big_cmp(const bigu *a, const bigu *b)
{
    if (a->n != b->n) return a->n < b->n ? -1 : 1;
    for (int i = a->n - 1; i >= 0; --i)
        if (a->d[i] != b->d[i]) return a->d[i] < b->d[i] ? -1 : 1;
    return 0;
}
static void big_add(bigu *a, const bigu *b)
{
    uint64_t carry = 0;
    int n = a->n > b->n ? a->n : b->n;
    for (int i = 0; i < n; ++i) {
        uint64_t v = (uint64_t)(i < a->n ? a->d[i] : 0) + (i < b->n ? b->d[i] : 0) + carry;
        a->d[i] = (uint32_t)v;
        carry = v >> 32;
    }
    a->n = n;
    if (carry) {
        if (a->n >= BIG_LIMBS) { fprintf(stderr, "oracle bigint overflow\n"); abort(); }
        a->d[a->n++] = (uint32_t)carry;
    }
}
static void big_sub(bigu *a, const bigu *b) /* a >= b */
{
    int64_t borrow = 0;
    for (int i = 0; i < a->n; ++i) {
        int64_t v = (int64_t)a->d[i] - (i < b->n ? b->d[i] : 0) - borrow;
        if (v < 0) { v += ((int64_t)1 << 32); borrow = 1; } else borrow = 0;
        a->d[i] = (uint32_t)v;
    }
    while (a->n > 0 && a->d[a->n - 1] == 0) a->n--;
}
static long double big_to_ld(const bigu *a)
{
    long double v = 0.0L;
    for (int i = a->n - 1; i >= 0; --i) v = v * 4294967296.0L + (long double)a->d[i];
    return v;
}
static void big_mul_fact(bigu *a, int64_t n) { for (int64_t q = 2; q <= n; ++q) big_mul_small(a, (uint32_t)q); }
static void big_div_fact(bigu *a, int64_t n) { for (int64_t q = 2; q <= n; ++q) big_div_small(a, (uint32_t)q); }

/* sqrt( prod num_i! / prod den_i! ) via prime exponents, in long double */
static long double sqrt_fact_ratio(const int64_t *num, int nn, const int64_t *den, int nd)
{
    int64_t maxn = 1;
    for (int i = 0; i < nn; ++i) maxn = imax64(maxn, num[i]);
    for (int i = 0; i < nd; ++i) maxn = imax64(maxn, den[i]);
    long double res = 1.0L;
    for (int64_t p = 2; p <= maxn; ++p) {
        int is_p = 1;
        for (int64_t q = 2; q * q <= p; ++q) if (p % q == 0) { is_p = 0; break; }
        if (!is_p) continue;
        int64_t e = 0;
        for (int i = 0; i < nn; ++i) for (int64_t pp = p; pp <= num[i]; pp *= p) e += num[i] / pp;
        for (int i = 0; i < nd; ++i) for (int64_t pp = p; pp <= den[i]; pp *= p) e -= den[i] / pp;
        int64_t ae = e < 0 ? -e : e;
        long double f = powl((long double)p, (long double)(ae / 2));
        if (ae & 1) f *= sqrtl((long double)p);
        res = e < 0 ? res / f : res * f;
    }
    return res;
}

static int triangle_ok(int64_t a, int64_t b, int64_t c)
{
    return (a + b >= c) && (a + c >= b) && (b + c >= a);
}

/* ( ja jb jc ; 0 0 0 ) for integer j: zero unless triangle and J even;
 * (-1)^g sqrt(Delta(ja jb jc)) g!/((g-ja)!(g-jb)!(g-jc)!),  g = J/2          */
double orc_three_j0(int64_t ja, int64_t jb, int64_t jc)
{
    if (ja < 0 || jb < 0 || jc < 0 || !triangle_ok(ja, jb, jc)) return 0.0;
    int64_t J = ja + jb + jc;
    if (J & 1) return 0.0;
    int64_t g = J / 2;
    /* value^2 = Delta * (g!/((g-a)!(g-b)!(g-c)!))^2, all under one sqrt */
    int64_t num[5] = { J - 2 * jc, J - 2 * jb, J - 2 * ja, g, g };
    int64_t den[7] = { J + 1, g - ja, g - ja, g - jb, g - jb, g - jc, g - jc };
    long double v = sqrt_fact_ratio(num, 5, den, 7);
    return (double)((g & 1) ? -v : v);
}

/* { ja jb jc ; jd je jf } for integer j, Racah's single sum with exact
 * integer terms; zero unless the four triads (abc)(aef)(dbf)(dec) close.    */
double orc_six_j(int64_t ja, int64_t jb, int64_t jc, int64_t jd, int64_t je, int64_t jf)
{
    if (ja < 0 || jb < 0 || jc < 0 || jd < 0 || je < 0 || jf < 0) return 0.0;
    if (!triangle_ok(ja, jb, jc) || !triangle_ok(ja, je, jf) ||
        !triangle_ok(jd, jb, jf) || !triangle_ok(jd, je, jc)) return 0.0;
    int64_t a1 = ja + jb + jc, a2 = ja + je + jf, a3 = jd + jb + jf, a4 = jd + je + jc;
    int64_t b1 = ja + jb + jd + je, b2 = jb + jc + je + jf, b3 = jc + ja + jf + jd;
    int64_t tmin = imax64(imax64(a1, a2), imax64(a3, a4));
    int64_t tmax = imin64(b1, imin64(b2, b3));
    if (tmax < tmin) return 0.0;
    /* term(t) = (t+1)! / [(t-a1)!(t-a2)!(t-a3)!(t-a4)!(b1-t)!(b2-t)!(b3-t)!] */
    bigu term, pos, neg;
    big_set(&term, 1);
    big_mul_fact(&term, tmin + 1);
    big_div_fact(&term, tmin - a1); big_div_fact(&term, tmin - a2);
    big_div_fact(&term, tmin - a3); big_div_fact(&term, tmin - a4);
    big_div_fact(&term, b1 - tmin); big_div_fact(&term, b2 - tmin);
    big_div_fact(&term, b3 - tmin);
    big_set(&pos, 0); big_set(&neg, 0);
    for (int64_t t = tmin;; ++t) {
        if (t & 1) big_add(&neg, &term); else big_add(&pos, &term);
        if (t == tmax) break;
        big_mul_small(&term, (uint32_t)(t + 2));
        big_mul_small(&term, (uint32_t)(b1 - t));
        big_mul_small(&term, (uint32_t)(b2 - t));
        big_mul_small(&term, (uint32_t)(b3 - t));
        big_div_small(&term, (uint32_t)(t + 1 - a1));
        big_div_small(&term, (uint32_t)(t + 1 - a2));
        big_div_small(&term, (uint32_t)(t + 1 - a3));
        big_div_small(&term, (uint32_t)(t + 1 - a4));
    }
    int c = big_cmp(&pos, &neg);
    if (c == 0) return 0.0;
    long double s;
    if (c > 0) { big_sub(&pos, &neg); s = big_to_ld(&pos); }
    else { big_sub(&neg, &pos); s = -big_to_ld(&neg); }
    int64_t num[12] = {
        ja + jb - jc, ja - jb + jc, -ja + jb + jc,
        ja + je - jf, ja - je + jf, -ja + je + jf,
        jd + jb - jf, jd - jb + jf, -jd + jb + jf,
        jd + je - jc, jd - je + jc, -jd + je + jc };
    int64_t den[4] = { a1 + 1, a2 + 1, a3 + 1, a4 + 1 };
    return (double)(s * sqrt_fact_ratio(num, 12, den, 4));
}

/* wigner_tools.f90:107-112 */
double orc_C_red_mat(int64_t k, int64_t a, int64_t b)
{
    double sgn = (a & 1) ? -1.0 : 1.0;
    return sgn * sqrt((double)((2 * a + 1) * (2 * b + 1))) * orc_three_j0(a, k, b);
}

/* wigner_tools.f90:126-138 */
double orc_ang_k_LS(int64_t k, int64_t la, int64_t lb, int64_t lc, int64_t ld, int64_t L)
{
    if (((la + k + lc) % 2 != 0) || ((lb + k + ld) % 2 != 0)) return 0.0;
    double sgn = ((lb + lc + L) & 1) ? -1.0 : 1.0;
    return sgn * orc_six_j(la, lb, L, ld, lc, k) * orc_C_red_mat(k, la, lc) * orc_C_red_mat(k, lb, ld);
}

/* ======================================================================== */
/* orbital_tools.f90:46-72 consistent, :119-216 count_configs                */
/* ======================================================================== */
static int consistent(int64_t l1, int64_t l2, int64_t term_l, int64_t term_pi, int eqv)
{
    if (eqv && (term_l % 2 != 0)) return 0;
    int tri = (iabs64(l1 - l2) <= term_l) && (term_l <= l1 + l2);
    if (!tri) return 0;
    int pi = ((l1 % 2 != 0) != (l2 % 2 != 0)); /* parity([pi1,pi2]) */
    if (pi != (term_pi != 0)) return 0;
    return 1;
}

int64_t orc_count_configs(int64_t term_l, int64_t term_pi,
                          int64_t max_l_1p, int64_t n_b, int64_t k_spline,
                          int64_t max_n_b, int64_t n_all_l, int64_t l_2_max,
                          int64_t *conf_n, int64_t *conf_l, int64_t *conf_eqv,
                          int64_t cap)
{
    int64_t ptr = 1;
    for (int64_t i = 0; i <= max_l_1p; ++i)
        for (int64_t j = 0; j <= i; ++j)
            for (int64_t n_i = imin64(i + 1, k_spline - 1); n_i <= n_b; ++n_i) {
                if ((n_i > n_all_l) && (j > l_2_max)) continue;
                int64_t hi = (j == i) ? imin64(n_i, max_n_b) : imin64(n_b, max_n_b);
                for (int64_t n_j = imin64(j + 1, k_spline - 1); n_j <= hi; ++n_j) {
                    int eqv = (i == j) && (n_i == n_j);
                    if (consistent(i, j, term_l, term_pi, eqv)) {
                        if (ptr <= cap) {
                            conf_n[2 * (ptr - 1)] = n_i;
                            conf_n[2 * (ptr - 1) + 1] = n_j;
                            conf_l[2 * (ptr - 1)] = i;
                            conf_l[2 * (ptr - 1) + 1] = j;
                            conf_eqv[ptr - 1] = eqv;
                        }
                        ptr = ptr + 1;
                    }
                }
            }
    return ptr - 1;
}

/* orbital_tools.f90:245-343  count_terms + the symmetry enumeration of
 * init_basis (two_el = .true. branch)                                       */
int64_t orc_init_basis_syms(int64_t max_L, int64_t z_pol,
                            int64_t *sym_l, int64_t *sym_m, int64_t *sym_pi)
{
    int64_t ptr = 0;
    sym_l[ptr] = 0; sym_m[ptr] = 0; sym_pi[ptr] = 0; ptr++;
    if (z_pol) {
        for (int64_t l = 1; l <= max_L; ++l) {
            sym_l[ptr] = l; sym_m[ptr] = 0; sym_pi[ptr] = (l % 2 != 0); ptr++;
        }
    } else {
        for (int64_t l = 1; l <= max_L; ++l)
            for (int64_t p = 0; p <= 1; ++p)
                for (int64_t m = -l; m <= l; ++m) {
                    if (iabs64(m % 2) != p) continue;
                    sym_l[ptr] = l; sym_m[ptr] = m; sym_pi[ptr] = p; ptr++;
                }
    }
    return ptr;
}

/* ======================================================================== */
/* memo of ang_k_LS for one L (pure function of small integers; the          */
/* reference re-evaluates GSL each time, the values are the same)            */
/* ======================================================================== */
typedef struct {
    int64_t lmax, K1, L;
    double *val;
    unsigned char *have;
} ang_memo;

static void memo_init(ang_memo *m, int64_t lmax, int64_t max_k, int64_t L)
{
    m->lmax = lmax; m->K1 = max_k + 1; m->L = L;
    size_t n = (size_t)((lmax + 1) * (lmax + 1) * (lmax + 1) * (lmax + 1) * m->K1);
    m->val = (double *)malloc(sizeof(double) * n);
    m->have = (unsigned char *)calloc(n, 1);
}
static void memo_free(ang_memo *m) { free(m->val); free(m->have); }
static double memo_ang(ang_memo *m, int64_t k, int64_t la, int64_t lb, int64_t lc, int64_t ld)
{
    int64_t s = m->lmax + 1;
    size_t idx = (size_t)(((((la * s + lb) * s + lc) * s + ld) * m->K1) + k);
    if (!m->have[idx]) {
        m->val[idx] = orc_ang_k_LS(k, la, lb, lc, ld, m->L);
        m->have[idx] = 1;
    }
    return m->val[idx];
}
static int64_t conf_lmax(int64_t n_config, const int64_t *conf_l)
{
    int64_t lm = 0;
    for (int64_t q = 0; q < 2 * n_config; ++q) lm = imax64(lm, conf_l[q]);
    return lm;
}

/* ======================================================================== */
/* hamiltonian.f90:348-416  count_nnz                                        */
/* ======================================================================== */
void orc_count_nnz_rows(int64_t k_spline, int64_t term_l, int64_t n_config,
                        const int64_t *conf_n, const int64_t *conf_l,
                        int64_t max_k, int64_t full, int64_t row_lo, int64_t row_hi,
                        int64_t *res);

void orc_count_nnz(int64_t k_spline, int64_t term_l, int64_t n_config,
                   const int64_t *conf_n, const int64_t *conf_l,
                   int64_t max_k, int64_t full, int64_t *res)
{
    orc_count_nnz_rows(k_spline, term_l, n_config, conf_n, conf_l, max_k, full, 1, n_config, res);
}

/* the same scan restricted to outer rows row_lo..row_hi (sampled timing) */
void orc_count_nnz_rows(int64_t k_spline, int64_t term_l, int64_t n_config,
                        const int64_t *conf_n, const int64_t *conf_l,
                        int64_t max_k, int64_t full, int64_t row_lo, int64_t row_hi,
                        int64_t *res)
{
    ang_memo memo;
    memo_init(&memo, conf_lmax(n_config, conf_l), max_k, term_l);
    res[0] = 0; res[1] = 0;
    for (int64_t i = row_lo; i <= row_hi; ++i) {
        int64_t n_a = conf_n[2 * (i - 1)], n_b = conf_n[2 * (i - 1) + 1];
        int64_t l_a = conf_l[2 * (i - 1)], l_b = conf_l[2 * (i - 1) + 1];
        for (int64_t j = (full ? 1 : i); j <= n_config; ++j) {
            int64_t n_c = conf_n[2 * (j - 1)], n_d = conf_n[2 * (j - 1) + 1];
            int64_t l_c = conf_l[2 * (j - 1)], l_d = conf_l[2 * (j - 1) + 1];
            int sup = (iabs64(n_a - n_c) < k_spline) && (iabs64(n_b - n_d) < k_spline);
            int sup_ex = (iabs64(n_a - n_d) < k_spline) && (iabs64(n_b - n_c) < k_spline);
            if (sup) {
                int l_eq = (l_a == l_c) && (l_b == l_d);
                if (l_eq) res[1]++;
                for (int64_t k = 0; k <= max_k; ++k) {
                    double ang = memo_ang(&memo, k, l_a, l_b, l_c, l_d);
                    if (fabs(ang) > 5.e-15) { res[0]++; break; }
                }
            } else if (sup_ex) {
                int l_eq_ex = (l_a == l_d) && (l_b == l_c);
                if (l_eq_ex) res[1]++;
                for (int64_t k = 0; k <= max_k; ++k) {
                    double ang = memo_ang(&memo, k, l_a, l_b, l_d, l_c);
                    if (fabs(ang) > 5.e-15) { res[0]++; break; }
                }
            }
        }
    }
    memo_free(&memo);
}

/* ======================================================================== */
/* mat_els.f90:552-571 r_12_tens, :608-633 c_mat_neq_tens,                   */
/* :664-678 S_mat_neq, :697-715 H_1p_neq                                     */
/* ======================================================================== */
typedef struct {
    const orc_bspline *bs;
    int64_t max_k, P;
    const double *R;
    const int64_t *rowoff; /* pair-index row offsets, [1..n_b] */
    ang_memo *memo;
} r12_ctx;

static const double *R_lookup(const r12_ctx *c, int64_t a, int64_t b, int64_t cc, int64_t d)
{
    const int64_t w = c->bs->k - 1;
    int64_t p1 = c->rowoff[a] + (cc - imax64(1, a - w));
    int64_t p2 = c->rowoff[b] + (d - imax64(1, b - w));
    return c->R + (p1 * c->P + p2) * (c->max_k + 1);
}

static double r_12_tens(const r12_ctx *c, int64_t la, int64_t lb, int64_t lc, int64_t ld,
                        int64_t a, int64_t b, int64_t cc, int64_t d)
{
    double res = 0.0;
    const double *vals = R_lookup(c, a, b, cc, d);
    for (int64_t k = 0; k <= c->max_k; ++k) {
        double ang = memo_ang(c->memo, k, la, lb, lc, ld);
        if (fabs(ang) < 5.e-16) continue;
        res = res + vals[k] * ang;
    }
    return res;
}

static double c_mat_neq_tens(const r12_ctx *c, int64_t L,
                             int64_t la, int64_t lb, int64_t lc, int64_t ld,
                             int64_t a, int64_t b, int64_t cc, int64_t d,
                             int sup, int sup_ex)
{
    double res = 0.0;
    if (sup) res = res + r_12_tens(c, la, lb, lc, ld, a, b, cc, d);
    if (sup_ex) {
        double sgn = ((lc + ld + L) & 1) ? -1.0 : 1.0;
        res = res + sgn * r_12_tens(c, la, lb, ld, lc, a, b, d, cc);
    }
    return res;
}

int orc_construct_block_tensor(const orc_bspline *bs, int64_t max_l_1p,
                               const double *H_vec, const double *S_in,
                               int64_t term_l, int64_t n_config,
                               const int64_t *conf_n, const int64_t *conf_l,
                               int64_t max_k, const double *R, int64_t full,
                               int64_t row_lo, int64_t row_hi,
                               int64_t cap_H, int64_t *H_ptr, int64_t *H_idx, double *H_dat_,
                               int64_t cap_S, int64_t *S_ptr, int64_t *S_idx, double *S_dat_,
                               int64_t *emitted)
{
    (void)max_l_1p;
    const int64_t nb = bs->n_b, ks = bs->k, L = term_l;
    const zcplx *S = (const zcplx *)S_in;
    const zcplx *Hv = (const zcplx *)H_vec;
    zcplx *H_dat = (zcplx *)H_dat_, *S_dat = (zcplx *)S_dat_;
#define S_(n, np) S[((n) - 1) + nb * ((np) - 1)]
#define H_(l, n, np) Hv[(size_t)(l) * (size_t)(nb * nb) + (size_t)(((n) - 1) + nb * ((np) - 1))]

    ang_memo memo;
    memo_init(&memo, conf_lmax(n_config, conf_l), max_k, L);
    int64_t *rowoff = (int64_t *)malloc(sizeof(int64_t) * (size_t)(nb + 2));
    {
        int64_t off = 0;
        for (int64_t q = 1; q <= nb; ++q) {
            rowoff[q] = off;
            off += imin64(nb, q + ks - 1) - imax64(1, q - ks + 1) + 1;
        }
    }
    r12_ctx rc = { bs, max_k, orc_num_pairs(bs), R, rowoff, &memo };

    /* init_CS (sparse_array_tools.f90:557-569): zero-filled arrays */
    for (int64_t q = 0; q < cap_H; ++q) { H_idx[q] = 0; H_dat[q] = 0.0; }
    for (int64_t q = 0; q < cap_S; ++q) { S_idx[q] = 0; S_dat[q] = 0.0; }
    int64_t row_ptr_H = 1, row_ptr_S = 1;
    H_ptr[row_lo - 1] = 1;
    S_ptr[row_lo - 1] = 1;
    int overflow = 0;
    /* the odd-L and even-L branches of the reference (:149-205 / :206-279)
     * execute the same statements; one loop restates both                   */
    for (int64_t i = row_lo; i <= row_hi && !overflow; ++i) {
        int64_t n_a = conf_n[2 * (i - 1)], n_b = conf_n[2 * (i - 1) + 1];
        int64_t l_a = conf_l[2 * (i - 1)], l_b = conf_l[2 * (i - 1) + 1];
        for (int64_t j = (full ? 1 : i); j <= n_config; ++j) {
            int64_t n_c = conf_n[2 * (j - 1)], n_d = conf_n[2 * (j - 1) + 1];
            int sup = (iabs64(n_a - n_c) < ks) && (iabs64(n_b - n_d) < ks);
            int sup_ex = (iabs64(n_a - n_d) < ks) && (iabs64(n_b - n_c) < ks);
            if (!(sup || sup_ex)) continue;
            int64_t l_c = conf_l[2 * (j - 1)], l_d = conf_l[2 * (j - 1) + 1];
            int r_12_allowed = 0;
            for (int64_t k = 0; k <= max_k; ++k) {
                double ang = memo_ang(&memo, k, l_a, l_b, l_c, l_d);
                double ang_ex = memo_ang(&memo, k, l_a, l_b, l_d, l_c);
                if (((fabs(ang) > 5.e-15) && sup) || ((fabs(ang_ex) > 5.e-15) && sup_ex)) {
                    r_12_allowed = 1;
                    break;
                }
            }
            int l_eq = (l_a == l_c) && (l_b == l_d);
            int l_eq_ex = (l_a == l_d) && (l_b == l_c);
            int store_S = (sup && l_eq) || (sup_ex && l_eq_ex);
            int store_H = r_12_allowed || l_eq || l_eq_ex;
            if ((store_H && row_ptr_H > cap_H) || (store_S && row_ptr_S > cap_S)) {
                overflow = 1;
                break;
            }
            if (r_12_allowed)
                H_dat[row_ptr_H - 1] = H_dat[row_ptr_H - 1] +
                    c_mat_neq_tens(&rc, L, l_a, l_b, l_c, l_d, n_a, n_b, n_c, n_d, sup, sup_ex);
            if (store_S) {
                /* H_1p_neq (mat_els.f90:697-715) */
                zcplx h = 0.0;
                if (l_eq)
                    h = h + H_(l_a, n_a, n_c) * S_(n_b, n_d) + H_(l_b, n_b, n_d) * S_(n_a, n_c);
                if (l_eq_ex) {
                    double sgn = ((L + l_c + l_d) & 1) ? -1.0 : 1.0;
                    h = h + (H_(l_a, n_a, n_d) * S_(n_b, n_c) + H_(l_b, n_b, n_c) * S_(n_a, n_d)) * sgn;
                }
                H_dat[row_ptr_H - 1] = H_dat[row_ptr_H - 1] + h;
                /* S_mat_neq (mat_els.f90:664-678) */
                zcplx s = 0.0;
                if (l_eq) s = s + S_(n_a, n_c) * S_(n_b, n_d);
                if (l_eq_ex) {
                    double sgn = ((L + l_c + l_d) & 1) ? -1.0 : 1.0;
                    s = s + sgn * S_(n_a, n_d) * S_(n_b, n_c);
                }
                S_dat[row_ptr_S - 1] = S_dat[row_ptr_S - 1] + s;
                S_idx[row_ptr_S - 1] = j;
                row_ptr_S = row_ptr_S + 1;
            }
            if (store_H) {
                H_idx[row_ptr_H - 1] = j;
                row_ptr_H = row_ptr_H + 1;
            }
        }
        H_ptr[i] = row_ptr_H;
        S_ptr[i] = row_ptr_S;
    }
    emitted[0] = row_ptr_H - 1;
    emitted[1] = row_ptr_S - 1;
    free(rowoff);
    memo_free(&memo);
#undef S_
#undef H_
    return overflow ? -1 : 0;
}

/* ======================================================================== */
/* Dipole blocks                                       */
/* ======================================================================== */

/* wigner_tools.f90:30-45 three_j -> gsl_sf_coupling_3j, integer j and m.
 * Racah's formula in long double (all j <= ~30 here: factorials <= 100!
 * stay inside the long double range and the sum has no severe cancellation
 * for the small arguments of this path).                                    */
static long double lfact(int64_t n)
{
    long double f = 1.0L;
    for (int64_t q = 2; q <= n; ++q) f *= (long double)q;
    return f;
}

double orc_three_j(int64_t ja, int64_t jb, int64_t jc, int64_t ma, int64_t mb, int64_t mc)
{
    if (ja < 0 || jb < 0 || jc < 0) return 0.0;
    if (ma + mb + mc != 0) return 0.0;
    if (iabs64(ma) > ja || iabs64(mb) > jb || iabs64(mc) > jc) return 0.0;
    if (!triangle_ok(ja, jb, jc)) return 0.0;
    const int64_t t1 = jb - jc - ma, t2 = ja + mb - jc;       /* lower limits -t1, -t2 */
    const int64_t t3 = ja + jb - jc, t4 = ja - ma, t5 = jb + mb;
    const int64_t tmin = imax64(0, imax64(t1, t2)), tmax = imin64(t3, imin64(t4, t5));
    if (tmax < tmin) return 0.0;
    long double sum = 0.0L;
    for (int64_t t = tmin; t <= tmax; ++t) {
        long double d = lfact(t) * lfact(t - t1) * lfact(t - t2) * lfact(t3 - t) * lfact(t4 - t) * lfact(t5 - t);
        sum += ((t & 1) ? -1.0L : 1.0L) / d;
    }
    long double delta = lfact(ja + jb - jc) * lfact(ja - jb + jc) * lfact(-ja + jb + jc) / lfact(ja + jb + jc + 1);
    long double pref = sqrtl(delta * lfact(ja + ma) * lfact(ja - ma) * lfact(jb + mb) * lfact(jb - mb) *
                             lfact(jc + mc) * lfact(jc - mc));
    long double v = pref * sum;
    if ((ja - jb - mc) & 1) v = -v;
    return (double)v;
}

/* mat_els.f90:120-170 setup_radial_dip with :348-390 compute_radial_dip_len /
 * compute_radial_dip_vel.  gauge 'l': A = r_mat, B untouched (may be NULL);
 * gauge 'v': A = dr_mat, B = r_inv_mat.  Dense complex n_b x n_b, column-major. */
int orc_setup_radial_dip(const orc_bspline *bs, int64_t k_GL, int gauge, double *A_out, double *B_out)
{
    if (gauge != 'l' && gauge != 'v') return -1;
    const int64_t nb = bs->n_b, cells = bs->nbp - 1;
    zcplx *A = (zcplx *)A_out, *B = (zcplx *)B_out;
    for (int64_t q = 0; q < nb * nb; ++q) A[q] = 0.0;
    if (gauge == 'v') for (int64_t q = 0; q < nb * nb; ++q) B[q] = 0.0;
    double *x = (double *)malloc(sizeof(double) * (size_t)(k_GL * cells));
    double *w = (double *)malloc(sizeof(double) * (size_t)(k_GL * cells));
    for (int64_t c = 0; c < cells; ++c)
        orc_gauss_legendre(k_GL, bs->bp[c], bs->bp[c + 1], x + c * k_GL, w + c * k_GL);
    double *cw = (double *)calloc((size_t)bs->n, sizeof(double));
    const zcplx mi = -I;   /* dcmplx(0.d0,-1.d0) */
    for (int64_t j_b = 1; j_b <= nb; ++j_b)
        for (int64_t i_b = 1; i_b <= nb; ++i_b) {
            if (iabs64(j_b - i_b) >= bs->k) continue;
            for (int64_t i_r = 1; i_r <= cells; ++i_r) {
                if (!(support(bs, i_r, i_b + 1) && support(bs, i_r, j_b + 1))) continue;
                const double *r = x + (i_r - 1) * k_GL, *ww = w + (i_r - 1) * k_GL;
                const size_t at = (size_t)((i_b - 1) + nb * (j_b - 1));
                for (int64_t q = 0; q < k_GL; ++q) {
                    double B_i = bspl_unit(bs, cw, i_b, r[q], 0, i_r);
                    double B_j = bspl_unit(bs, cw, j_b, r[q], 0, i_r);
                    if (gauge == 'l') {
                        A[at] = A[at] + ww[q] * r[q] * B_i * B_j;
                    } else {
                        double D_B_j = bspl_unit(bs, cw, j_b, r[q], 1, i_r);
                        A[at] = A[at] + mi * ww[q] * B_i * D_B_j;
                        B[at] = B[at] + mi * ww[q] * B_i * B_j / r[q];
                    }
                }
            }
        }
    free(cw); free(x); free(w);
    return 0;
}

typedef struct {
    int gauge;
    int64_t nb;
    const zcplx *A, *B, *S;
} dip_ctx;

/* mat_els.f90:772-811 dip_red_1p_len / dip_red_1p_vel for the pair (n,l) -> (n_p,l_p) */
static zcplx dip_red_1p(const dip_ctx *c, int64_t n, int64_t n_p, int64_t l, int64_t l_p)
{
    const size_t at = (size_t)((n - 1) + c->nb * (n_p - 1));
    if (c->gauge == 'l') return c->A[at] * orc_C_red_mat(1, l, l_p);
    if (iabs64(l - l_p) != 1) return 0.0;
    if (l > l_p) return sqrt((double)l) * (c->A[at] - (double)(l_p + 1) * c->B[at]);
    return -sqrt((double)l_p) * (c->A[at] + (double)l_p * c->B[at]);
}

/* mat_els.f90:739-770 dip_red_mat; n = (n1,n2), l = (l1,l2) of both configurations */
static zcplx dip_red_mat(const dip_ctx *c, int64_t L_1, int64_t L_2, const int64_t *n1, const int64_t *l1,
                         const int64_t *n2, const int64_t *l2)
{
    zcplx res_1 = 0.0, res_2 = 0.0;
    const double rt = sqrt((double)((2 * L_1 + 1) * (2 * L_2 + 1)));
    if (l1[1] == l2[1]) {
        double sg = ((l1[0] + l1[1] + L_2 + 1) & 1) ? -1.0 : 1.0;
        res_1 = sg * rt * orc_six_j(L_1, 1, L_2, l2[0], l1[1], l1[0]) * dip_red_1p(c, n1[0], n2[0], l1[0], l2[0]) *
                c->S[(size_t)((n1[1] - 1) + c->nb * (n2[1] - 1))];
    }
    if (l1[0] == l2[0]) {
        double sg = ((l1[0] + l2[1] + L_1 + 1) & 1) ? -1.0 : 1.0;
        res_2 = sg * rt * orc_six_j(L_1, 1, L_2, l2[1], l1[0], l1[1]) * dip_red_1p(c, n1[1], n2[1], l1[1], l2[1]) *
                c->S[(size_t)((n1[0] - 1) + c->nb * (n2[0] - 1))];
    }
    return res_1 + res_2;
}

/* mat_els.f90:813-831 ang_dip_red */
static double ang_dip_red(int64_t L_1, int64_t L_2, const int64_t *l1, const int64_t *l2)
{
    double res = 0.0;
    if (l1[1] == l2[1]) res = res + fabs(orc_six_j(L_1, 1, L_2, l2[0], l1[1], l1[0]) * orc_C_red_mat(1, l1[0], l2[0]));
    if (l1[0] == l2[0]) res = res + fabs(orc_six_j(L_1, 1, L_2, l2[1], l1[0], l1[1]) * orc_C_red_mat(1, l1[1], l2[1]));
    return res;
}

/* dipole.f90:8-47 construct_dip_block_tensor with :87-146 init_dip_block.
 * sym = (l, m, pi).  Returns nnz (>= 0); when index_ptr is NULL only counts.
 * An empty block (forbidden, or compute false) has nnz 0 and no arrays.      */
int64_t orc_dip_block(const orc_bspline *bs, int gauge, const double *A, const double *B, const double *S,
                      int64_t q, const int64_t *sym1, int64_t n1c, const int64_t *conf_n1, const int64_t *conf_l1,
                      const int64_t *sym2, int64_t n2c, const int64_t *conf_n2, const int64_t *conf_l2,
                      int64_t compute, int64_t *index_ptr, int64_t *indices, double *data_)
{
    const int64_t ks = bs->k;
    const int64_t L_1 = sym1[0], L_2 = sym2[0];
    const int parity_allowed = (sym1[2] != 0) != (sym2[2] != 0);
    double ang = orc_three_j(L_1, 1, L_2, -sym1[1], q, sym2[1]);
    if ((L_1 - sym1[1]) & 1) ang = -ang;
    if (fabs(ang) < 5.e-16 || !parity_allowed || !compute) return 0;
    dip_ctx c = { gauge, bs->n_b, (const zcplx *)A, (const zcplx *)B, (const zcplx *)S };
    zcplx *data = (zcplx *)data_;
    int64_t ptr = 1;
    for (int64_t i = 1; i <= n1c; ++i) {
        const int64_t *na = conf_n1 + 2 * (i - 1), *la = conf_l1 + 2 * (i - 1);
        if (index_ptr) index_ptr[i - 1] = ptr;
        for (int64_t j = 1; j <= n2c; ++j) {
            const int64_t *nc = conf_n2 + 2 * (j - 1), *lc = conf_l2 + 2 * (j - 1);
            int sup = (iabs64(na[0] - nc[0]) < ks) && (iabs64(na[1] - nc[1]) < ks);
            int sup_ex = (iabs64(na[0] - nc[1]) < ks) && (iabs64(na[1] - nc[0]) < ks);
            int nz = 0;
            if (sup) { if (ang_dip_red(L_1, L_2, la, lc) > 5.e-16) nz = 1; }
            else if (sup_ex) { if (ang_dip_red(L_1, L_2, la, lc) > 5.e-16) nz = 1; }
            if (!nz) continue;
            if (index_ptr) {
                /* dipole.f90:36-42 + mat_els.f90:717-737 dip_mat_neq */
                const int64_t nx[2] = { nc[1], nc[0] }, lx[2] = { lc[1], lc[0] };
                zcplx red = dip_red_mat(&c, L_1, L_2, na, la, nc, lc);
                double sg = ((L_2 + lc[0] + lc[1]) & 1) ? -1.0 : 1.0;
                zcplx red_ex = sg * dip_red_mat(&c, L_1, L_2, na, la, nx, lx);
                indices[ptr - 1] = j;
                data[ptr - 1] = ang * (red + red_ex);
            }
            ptr = ptr + 1;
        }
    }
    if (index_ptr) index_ptr[n1c] = ptr;
    return ptr - 1;
}
