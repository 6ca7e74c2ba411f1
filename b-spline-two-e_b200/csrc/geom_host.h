// geom_host.h -- host-side construction of the basis geometry (Geom) and of
// the band-packed one-particle matrices; shared by abi.cu and tests/hostcheck.
#pragma once
#include <stdexcept>
#include <string>
#include <vector>

#include "core.h"

namespace bs2e {

struct HostGeom {
    Geom g{};  // pointers refer to the vectors below
    int max_k = 0;
    std::vector<double> knots, bp, glx, glw;
    std::vector<int> rowoff;
    std::vector<PairAC> pairs;
    long long nnz_4d = 0, nnz_6d = 0;
};

// type(b_spline) init (bspline_tools.f90:27-56) + pair numbering + the entry
// counts of sparse_4d / sparse_6d (sparse_array_tools.f90:276-324)
inline void build_host_geom(HostGeom& h, int ks, int nt, const double* knots, int max_k, int kgl,
                            const double* glx, const double* glw)
{
    typedef std::invalid_argument Error;
    if (ks < 2 || ks > kMaxOrder) throw Error("spline order k must be in 2..20");
    if (nt < 2 * ks + 1) throw Error("knot vector too short");
    if (max_k < 0 || max_k > 120) throw Error("max_k must be in 0..120");
    if (kgl < 1 || kgl > 64) throw Error("k_GL must be in 1..64");
    Geom& g = h.g;
    g.ks = ks;
    g.w = ks - 1;
    g.n = nt - ks;
    g.nb = g.n - 2;
    g.cells = g.n - ks + 1;
    g.K1 = max_k + 1;
    g.kgl = kgl;
    h.max_k = max_k;
    if (g.nb < 1) throw Error("no interior B-splines");
    h.knots.assign(knots, knots + nt);
    for (int q = 1; q < nt; ++q)
        if (knots[q] < knots[q - 1]) throw Error("knots must be non-decreasing");
    h.bp.assign(knots + ks - 1, knots + ks - 1 + g.cells + 1);
    for (int v = 0; v < g.cells; ++v)
        if (!(h.bp[v + 1] > h.bp[v])) throw Error("interior knots must be simple (empty cell found)");
    h.glx.assign(glx, glx + kgl);
    h.glw.assign(glw, glw + kgl);
    h.rowoff.assign(g.nb + 2, 0);
    h.pairs.clear();
    for (int a = 1; a <= g.nb; ++a) {
        h.rowoff[a] = (int)h.pairs.size();
        for (int cc = imax(1, a - g.w); cc <= imin(g.nb, a + g.w); ++cc)
            h.pairs.push_back(PairAC{a, cc});
    }
    h.rowoff[g.nb + 1] = (int)h.pairs.size();
    g.P = (int)h.pairs.size();
    g.ldP = ((g.P + 15) / 16) * 16;
    g.t = h.knots.data();
    g.bp = h.bp.data();
    g.glx = h.glx.data();
    g.glw = h.glw.data();
    g.rowoff = h.rowoff.data();
    g.pair = h.pairs.data();
    h.nnz_4d = 0;
    for (const PairAC& q : h.pairs)
        h.nnz_4d += imax(0, pair_hi_cell(g, q.a, q.c) - pair_lo_cell(g, q.a, q.c) + 1);
    h.nnz_6d = 0;  // per cell: (valid local splines)^4
    for (int v = 1; v <= g.cells; ++v) {
        long long nv = 0;
        for (int s = 0; s < ks; ++s) {
            const int b = v + s - 1;
            if (b >= 1 && b <= g.nb) ++nv;
        }
        h.nnz_6d += nv * nv * nv * nv;
    }
}

// dense complex column-major n_b x n_b  ->  band[n][n'-n+w] (complex)
inline void pack_band(const Geom& g, const double* dense, double* band, const char* what)
{
    const int nb = g.nb, w = g.w, bw = 2 * w + 1;
    for (int np = 1; np <= nb; ++np)
        for (int n = 1; n <= nb; ++n) {
            const double re = dense[2 * ((size_t)(n - 1) + (size_t)nb * (np - 1))];
            const double im = dense[2 * ((size_t)(n - 1) + (size_t)nb * (np - 1)) + 1];
            const int d = np - n + w;
            if (d < 0 || d > 2 * w) {
                if (re != 0.0 || im != 0.0)
                    throw std::invalid_argument(std::string(what) +
                                                " has a non-zero entry outside the B-spline band");
                continue;
            }
            band[((size_t)n * bw + d) * 2] = re;
            band[((size_t)n * bw + d) * 2 + 1] = im;
        }
}
inline size_t band_doubles(const Geom& g) { return (size_t)(g.nb + 1) * (2 * g.w + 1) * 2; }

}  // namespace bs2e
