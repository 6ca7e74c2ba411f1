// core.h -- per-element arithmetic and index logic of the two-electron path.
//
// Everything here is __host__ __device__ so that the CUDA kernels (slater.cu,
// rk.cu, block.cu) and the CPU-side logic checker under tests/hostcheck/ run
// the very same statements.  The checker is test infrastructure: the product
// never executes these functions on the host.
//
// Conventions: spline b-index i in 1..n_b (full index i+1),
// cells v in 1..C, ordered band pairs p=(a,c) with |a-c| < ks numbered
// a-major / c ascending, multipole k in 0..max_k (K1 = max_k+1).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define BS2E_HD __host__ __device__ __forceinline__
#else
#define BS2E_HD inline
#endif

namespace bs2e {

constexpr int kMaxOrder = 20;  // reference work arrays: bspline_tools.f90:163

BS2E_HD int imin(int a, int b) { return a < b ? a : b; }
BS2E_HD int imax(int a, int b) { return a > b ? a : b; }
BS2E_HD int iabs(int a) { return a < 0 ? -a : a; }
BS2E_HD int iclamp(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

struct PairAC { int a, c; };

// ---------------------------------------------------------------------------
// geometry of the B-spline basis (device pointers when used in kernels)
// ---------------------------------------------------------------------------
struct Geom {
    int ks;      // spline order (namelist k)
    int w;       // ks-1, half band width
    int n;       // size(knots)-ks
    int nb;      // n-2
    int cells;   // number of knot intervals C
    int K1;      // max_k+1
    int kgl;     // Gauss-Legendre points per cell
    int P;       // number of ordered band pairs
    int ldP;     // padded leading dimension of one R^k plane (multiple of 16)
    const double* t;     // knots, [n+ks]
    const double* bp;    // breakpoints, [cells+1]
    const double* glx;   // GL nodes on [-1,1], [kgl]
    const double* glw;   // GL weights, [kgl]
    const int* rowoff;   // [nb+2]: rowoff[a] = index of pair (a, max(1,a-w))
    const PairAC* pair;  // [P]: pair index -> (a,c)
};

BS2E_HD int pair_clo(const Geom& g, int a) { return imax(1, a - g.w); }
BS2E_HD int pair_index(const Geom& g, int a, int c) { return g.rowoff[a] + c - pair_clo(g, a); }
// cells on which both splines of the pair live (bspline_tools.f90:50-54)
BS2E_HD int pair_lo_cell(const Geom& g, int a, int c) { return imax(1, imax(a, c) - g.ks + 2); }
BS2E_HD int pair_hi_cell(const Geom& g, int a, int c) { return imin(g.cells, imin(a, c) + 1); }

// gfortran's r**k (integer k): square-and-multiply from the low bit
BS2E_HD double powi(double x, int m)
{
    unsigned n = (unsigned)(m < 0 ? -m : m);
    double y = (n & 1u) ? x : 1.0;
    while (n >>= 1) {
        x = x * x;
        if (n & 1u) y *= x;
    }
    return m < 0 ? 1.0 / y : y;
}

// All `order` B-splines of that order that are non-zero on the knot interval
// t(left) <= x < t(left+1) (1-based left), by the Cox-de Boor recurrence.
// out[s] is the spline with full index left-order+1+s.
BS2E_HD void bspline_values_left(const double* t, int order, int left, double x, double* out)
{
    double dl[kMaxOrder], dr[kMaxOrder];
    out[0] = 1.0;
    for (int j = 1; j < order; ++j) {
        dr[j - 1] = t[left + j - 1] - x;
        dl[j - 1] = x - t[left - j];
        double saved = 0.0;
        for (int i = 0; i < j; ++i) {
            double term = out[i] / (dr[i] + dl[j - 1 - i]);
            out[i] = saved + dr[i] * term;
            saved = dl[j - 1 - i] * term;
        }
        out[j] = saved;
    }
}

// All ks B-splines that are non-zero on cell v, evaluated at x.  out[s] is the
// spline with FULL index v+s (b-index v+s-1).  Replaces the reference's ks
// separate de Boor evaluations on unit coefficient vectors
// (bspline_tools.f90:151-224 called from mat_els.f90:415-419,462-476).
BS2E_HD void bspline_values(const double* t, int ks, int v, double x, double* out)
{
    bspline_values_left(t, ks, ks - 1 + v, x, out);
}

// ---------------------------------------------------------------------------
// stage B: one R^k value from the staged cell integrals
//   R^k(ab;cd), p1=(a,c), p2=(b,d)   (sparse_array_tools.f90:472-488)
// mom_rk/mom_rmk : [K1][P][ks]   cell moments, slot = cell - lo_cell(pair)
// pre            : [K1][P][ks+1] pre[t]  = sum_{slot <  t} mom_rk
// sufx           : [K1][P][ks+1] sufx[t] = sum_{slot >= t} mom_rmk
// rd             : [C][K1][ks*ks][ks*ks] same-cell integrals, local slots
// ---------------------------------------------------------------------------
struct CellData {
    const double* mom_rk;
    const double* mom_rmk;
    const double* pre;
    const double* sufx;
    const double* rd;
};

BS2E_HD double rk_offdiag(const Geom& g, const CellData& cd, int k, int p1, int p2,
                          int lo1, int lo2, int hi2)
{
    const int ks = g.ks;
    const double* pre1 = cd.pre + ((size_t)k * g.P + p1) * (ks + 1);
    const double* suf1 = cd.sufx + ((size_t)k * g.P + p1) * (ks + 1);
    const double* rk2 = cd.mom_rk + ((size_t)k * g.P + p2) * ks;
    const double* rmk2 = cd.mom_rmk + ((size_t)k * g.P + p2) * ks;
    double acc = 0.0;
    for (int s = 0; s <= hi2 - lo2; ++s) {
        const int tt = lo2 + s - lo1;  // slot of cell u in pair 1's frame
        acc += rmk2[s] * pre1[iclamp(tt, 0, ks)];
        acc += rk2[s] * suf1[iclamp(tt + 1, 0, ks)];
    }
    return acc;
}

BS2E_HD double rk_diag(const Geom& g, const CellData& cd, int k, PairAC q1, PairAC q2,
                       int vlo, int vhi)
{
    const int ks = g.ks, ks2 = ks * ks;
    double d1 = 0.0, d2 = 0.0;
#if defined(__CUDA_ARCH__)
#pragma unroll 4   // the loads of four cells in flight together (their addresses do not depend on the sums)
#endif
    for (int v = vlo; v <= vhi; ++v) {
        const int l1 = (q1.a + 1 - v) * ks + (q1.c + 1 - v);
        const int l2 = (q2.a + 1 - v) * ks + (q2.c + 1 - v);
        const double* blk = cd.rd + ((size_t)(v - 1) * g.K1 + k) * (size_t)ks2 * ks2;
        d1 += blk[(size_t)l1 * ks2 + l2];  // electron 2 inside electron 1
        d2 += blk[(size_t)l2 * ks2 + l1];  // electron 1 inside electron 2
    }
    return d1 + d2;
}

BS2E_HD double rk_element(const Geom& g, const CellData& cd, int k, int p1, int p2)
{
    const PairAC q1 = g.pair[p1], q2 = g.pair[p2];
    const int lo1 = pair_lo_cell(g, q1.a, q1.c), hi1 = pair_hi_cell(g, q1.a, q1.c);
    const int lo2 = pair_lo_cell(g, q2.a, q2.c), hi2 = pair_hi_cell(g, q2.a, q2.c);
    double val = rk_offdiag(g, cd, k, p1, p2, lo1, lo2, hi2);
    const int vlo = imax(lo1, lo2), vhi = imin(hi1, hi2);
    if (vlo <= vhi) val += rk_diag(g, cd, k, q1, q2, vlo, vhi);
    return val;
}

// The stage-B kernel (rk.cu) writes the plane of multipole k in two roles that share one launch:
//   streaming tiles  kRkRowsS p1 rows x 2*kRkThreads p2 columns: the entries whose pairs live on disjoint cell
//                    ranges (97 % of the tensor) are one product of total moments -- a pure streaming write from
//                    the packed per-pair records RkRow (two 16-byte loads per column, none per entry);
//   band tiles       kRkRowsB p1 rows x kRkThreads p2 columns along the band |a - b| <= w, outside of which no two
//                    cell ranges meet: one thread per column walks the rows of the tile and computes the entries
//                    with overlapping ranges by the general formula (rk_element), its column's moments staying in L1.
// The two roles write disjoint sets of entries, so they need no ordering.
constexpr int kRkThreads = 256;
constexpr int kRkRowsS = 32;   // p1 rows per streaming tile
constexpr int kRkRowsB = 16;   // p1 rows per band tile

// cell range and total moments of a band pair for one multipole (written by stage A: pair_prefix_item)
struct alignas(16) RkRow { int lo, hi; double trk, trmk; double pad; };

BS2E_HD RkRow rk_row_empty()
{
    RkRow r;
    r.lo = 1; r.hi = 0; r.trk = 0.0; r.trmk = 0.0; r.pad = 0.0;
    return r;
}

BS2E_HD RkRow rk_row_data(const Geom& g, const CellData& cd, int k, int p)
{
    RkRow r = rk_row_empty();
    if (p < g.P) {
        const PairAC q = g.pair[p];
        r.lo = pair_lo_cell(g, q.a, q.c);
        r.hi = pair_hi_cell(g, q.a, q.c);
        r.trk = cd.pre[((size_t)k * g.P + p) * (g.ks + 1) + g.ks];
        r.trmk = cd.sufx[((size_t)k * g.P + p) * (g.ks + 1)];
    }
    return r;
}

// the cell ranges of the two pairs meet: the entry needs the general formula
BS2E_HD bool rk_ranges_meet(const RkRow& a, const RkRow& c) { return !(a.hi < c.lo) && !(c.hi < a.lo); }
// disjoint ranges: electron 1 strictly inside electron 2's radius, or the other way round
BS2E_HD double rk_disjoint_value(const RkRow& a, const RkRow& c) { return a.hi < c.lo ? a.trk * c.trmk : a.trmk * c.trk; }

// columns [c0, c1) of the band of the p1 rows [r0, r1): every pair (b, .) with |a - b| <= w for an a of the rows
BS2E_HD void rk_band_columns(const Geom& g, int r0, int r1, int* c0, int* c1)
{
    const int a_min = g.pair[r0].a, a_max = g.pair[r1 - 1].a;
    *c0 = g.rowoff[imax(1, a_min - g.w)];
    *c1 = g.rowoff[imin(g.nb, a_max + g.w) + 1];
}

// streaming role, one thread: columns p2, p2+1 (p2 even) of the rows [r0, r0 + nrows); rows[] holds their records,
// c0 / c1 those of the two columns (rk_row_empty() past the last pair: the padding columns are written as zeros).
BS2E_HD void rk_stream_columns(const Geom& g, double* R, const RkRow* rows, int r0, int nrows, int k, int p2,
                               const RkRow& c0, const RkRow& c1)
{
    const bool ok0 = p2 < g.P, ok1 = p2 + 1 < g.P;
    double* dst = R + ((size_t)k * g.P + r0) * g.ldP + p2;
    for (int row = 0; row < nrows; ++row, dst += g.ldP) {
        const RkRow a = rows[row];
        const bool g0 = ok0 && rk_ranges_meet(a, c0), g1 = ok1 && rk_ranges_meet(a, c1);
        const double o0 = (ok0 && !g0) ? rk_disjoint_value(a, c0) : 0.0;
        const double o1 = (ok1 && !g1) ? rk_disjoint_value(a, c1) : 0.0;
#if defined(__CUDA_ARCH__)
        if (!g0 && !g1) *reinterpret_cast<double2*>(dst) = make_double2(o0, o1);
        else {
            if (!g0) dst[0] = o0;
            if (!g1) dst[1] = o1;
        }
#else
        if (!g0) dst[0] = o0;
        if (!g1) dst[1] = o1;
#endif
    }
}

// rk_element with the per-pair vectors handed in by the caller (the band role of the stage-B kernel keeps them in
// shared memory): pre1 / suf1 [ks+1] of pair p1, rk2 / rmk2 [ks] of pair p2, a / c the pairs' records.  Same
// statements in the same order as rk_offdiag + rk_diag.
BS2E_HD double rk_element_staged(const Geom& g, const CellData& cd, int k, PairAC q1, PairAC q2, const RkRow& a,
                                 const RkRow& c, const double* pre1, const double* suf1, const double* rk2,
                                 const double* rmk2)
{
    const int ks = g.ks;
    double val = 0.0;
    for (int s = 0; s <= c.hi - c.lo; ++s) {
        const int tt = c.lo + s - a.lo;
        val += rmk2[s] * pre1[iclamp(tt, 0, ks)];
        val += rk2[s] * suf1[iclamp(tt + 1, 0, ks)];
    }
    const int vlo = imax(a.lo, c.lo), vhi = imin(a.hi, c.hi);
    if (vlo <= vhi) val += rk_diag(g, cd, k, q1, q2, vlo, vhi);
    return val;
}

// band role with staged vectors, one thread: column p2 of the rows [r0, r0 + nrows); pre1s / suf1s hold the rows'
// vectors ([row][ks+1]), rk2 / rmk2 the column's
BS2E_HD void rk_band_column_staged(const Geom& g, const CellData& cd, double* R, const RkRow* rows, int r0, int nrows,
                                   int k, int p2, const RkRow& c, const double* pre1s, const double* suf1s,
                                   const double* rk2, const double* rmk2)
{
    const PairAC q2 = g.pair[p2];
    for (int row = 0; row < nrows; ++row)
        if (rk_ranges_meet(rows[row], c))
            R[((size_t)k * g.P + r0 + row) * g.ldP + p2] =
                rk_element_staged(g, cd, k, g.pair[r0 + row], q2, rows[row], c, pre1s + (size_t)row * (g.ks + 1),
                                  suf1s + (size_t)row * (g.ks + 1), rk2, rmk2);
}

// band role, one thread: column p2 of the rows [r0, r0 + nrows)
BS2E_HD void rk_band_column(const Geom& g, const CellData& cd, double* R, const RkRow* rows, int r0, int nrows, int k,
                            int p2, const RkRow& c)
{
    for (int row = 0; row < nrows; ++row)
        if (rk_ranges_meet(rows[row], c))
            R[((size_t)k * g.P + r0 + row) * g.ldP + p2] = rk_element(g, cd, k, r0 + row, p2);
}

// ---------------------------------------------------------------------------
// stage C: configuration blocks and the band-partner enumeration
//   (hamiltonian.f90:150-205 pattern; orbital_tools.f90:157-193 ordering)
// ---------------------------------------------------------------------------
struct NcRow {      // one n_c row of an (l_c,l_d) configuration block
    int nd_lo;      // first n_d present (nd_hi < nd_lo: row absent)
    int nd_hi;
    int start;      // 1-based configuration index of (n_c, nd_lo)
    int pad;
};

struct BlockDesc {  // configurations sharing (l_1,l_2), contiguous in the list
    int l1, l2;
    int nc_lo, nc_hi;  // range of n_1 present
};

constexpr unsigned kDirAny = 1u;  // exists k: |ang_k| > 5e-15     (hamiltonian.f90:174)
constexpr unsigned kExAny = 2u;   // same for the exchange ordering (l_d,l_c)

struct KRange { signed char dlo, dhi, xlo, xhi; };  // k ranges (step 2) with non-zero factors

// The planned rows of a block: a union of ascending, disjoint, inclusive 1-based row
// ranges (the whole block, a row range, or the share of one GPU: bs2e_block_plan_ranges).
// Local row = position inside the union; no per-row table is kept.
struct RowRanges {
    int n;           // number of ranges (n == 0: nothing planned)
    const int* lo;   // [n]
    const int* hi;   // [n]
    const int* off;  // [n] local index of row lo[q]
    int lo0, hi0;    // the first range again, by value: a single range needs no table look-up
};

// local index of configuration `row` (1-based), -1 when the row is not planned
BS2E_HD int row_local_of(const RowRanges& rr, int row)
{
    if (rr.n == 1) return (row >= rr.lo0 && row <= rr.hi0) ? row - rr.lo0 : -1;
    int a = 0, n = rr.n;  // largest q with lo[q] <= row
    if (n == 0 || row < rr.lo0) return -1;
    while (n > 1) {
        const int half = n >> 1;
        if (rr.lo[a + half] <= row) a += half;
        n -= half;
    }
    return row <= rr.hi[a] ? rr.off[a] + (row - rr.lo[a]) : -1;
}
// configuration index (1-based) of local row `local`
BS2E_HD int row_of_local(const RowRanges& rr, int local)
{
    if (rr.n == 1) return rr.lo0 + local;
    int a = 0, n = rr.n;
    while (n > 1) {
        const int half = n >> 1;
        if (rr.off[a + half] <= local) a += half;
        n -= half;
    }
    return rr.lo[a] + (local - rr.off[a]);
}

struct Plan {
    int nblk;
    int n_config;
    int full;
    int L;
    int max_nd;                  // largest n_2 of any configuration of the block list
    const BlockDesc* blk;        // [nblk]
    const NcRow* ncrow;          // [nblk][nb+1], index n_c
    const unsigned char* flags;  // [nblk][nblk]
    const KRange* krange;        // [nblk][nblk]
    const double* angD;          // [nblk][nblk][K1]   direct, |.|<5e-16 zeroed
    const double* angX;          // [nblk][nblk][K1]   exchange * (-1)^(lc+ld+L)
    const double* angP;          // [nblk][nblk][2*nkp] the same packed by parity (site_core.h)
    int nkp;
    // per-row tables: only built for the row-wise kernels (fallback fill, dipole blocks);
    // the site kernels derive the rows of a radial site from ncrow
    const unsigned short* row_n1;   // [n_config]
    const unsigned short* row_n2;   // [n_config]
    const unsigned short* row_blk;  // [n_config]
    // the planned rows
    int nrows;
    RowRanges rr;
};

// configuration index (1-based) of (block bi; n_a, n_b), 0 when there is no such configuration
BS2E_HD int config_index(const Geom& g, const Plan& pl, int bi, int na, int nb)
{
    const NcRow r = pl.ncrow[(size_t)bi * (g.nb + 1) + na];
    return (nb >= r.nd_lo && nb <= r.nd_hi) ? r.start + (nb - r.nd_lo) : 0;
}

// union of two closed integer intervals as <= 2 disjoint ascending intervals
struct Union2 { int lo[2], hi[2]; int n; };

BS2E_HD Union2 union2(int a1, int b1, int a2, int b2)
{
    Union2 u;
    u.n = 0;
    u.lo[0] = u.lo[1] = 0;
    u.hi[0] = u.hi[1] = -1;
    const bool e1 = a1 > b1, e2 = a2 > b2;
    if (e1 && e2) return u;
    if (e1) { u.lo[0] = a2; u.hi[0] = b2; u.n = 1; return u; }
    if (e2) { u.lo[0] = a1; u.hi[0] = b1; u.n = 1; return u; }
    if (a2 < a1) { int t = a1; a1 = a2; a2 = t; t = b1; b1 = b2; b2 = t; }
    if (a2 <= b1 + 1) {  // overlapping or adjacent: one interval
        u.lo[0] = a1; u.hi[0] = imax(b1, b2); u.n = 1;
    } else {
        u.lo[0] = a1; u.hi[0] = b1; u.lo[1] = a2; u.hi[1] = b2; u.n = 2;
    }
    return u;
}

BS2E_HD int union2_count(const Union2& u)
{
    int c = 0;
    for (int q = 0; q < u.n; ++q) c += u.hi[q] - u.lo[q] + 1;
    return c;
}

struct RowInfo { int i; int bi; int na, nb; int la, lb; };

BS2E_HD RowInfo row_info(const Plan& pl, int i /*1-based*/)
{
    RowInfo r;
    r.i = i;
    r.bi = pl.row_blk[i - 1];
    r.na = pl.row_n1[i - 1];
    r.nb = pl.row_n2[i - 1];
    r.la = pl.blk[r.bi].l1;
    r.lb = pl.blk[r.bi].l2;
    return r;
}

// per (row, column block): which selection rules can fire
struct Coupling {
    bool dirany, exany, same, samex;
    BS2E_HD bool any() const { return dirany || exany || same; }
};

BS2E_HD Coupling coupling(const Plan& pl, const RowInfo& r, int bj)
{
    Coupling c;
    const unsigned f = pl.flags[(size_t)r.bi * pl.nblk + bj];
    c.dirany = (f & kDirAny) != 0;
    c.exany = (f & kExAny) != 0;
    c.same = (bj == r.bi);                 // l_eq  (l_a=l_c, l_b=l_d)
    c.samex = c.same && (r.la == r.lb);    // l_eq_ex (configs keep l_1 >= l_2)
    return c;
}

// n_c windows of a row inside column block bj (ascending, <= 2 intervals)
BS2E_HD Union2 nc_windows(const Geom& g, const Plan& pl, const RowInfo& r, int bj)
{
    const BlockDesc b = pl.blk[bj];
    int lo = b.nc_lo, hi = b.nc_hi;
    if (!pl.full && bj == r.bi) lo = imax(lo, r.na);  // j >= i
    return union2(imax(lo, r.na - g.w), imin(hi, r.na + g.w),
                  imax(lo, r.nb - g.w), imin(hi, r.nb + g.w));
}

// One (row, column block, n_c) segment: the n_d intervals with direct support
// (D) and with exchange support (X), already clipped to the configurations
// that exist and to j >= i.
struct Segment {
    int dlo, dhi, xlo, xhi;
    int jbase;  // configuration index j = jbase + n_d
};

BS2E_HD Segment segment(const Geom& g, const Plan& pl, const RowInfo& r, int bj, int nc)
{
    Segment s;
    const NcRow row = pl.ncrow[(size_t)bj * (g.nb + 1) + nc];
    int lo = row.nd_lo, hi = row.nd_hi;
    if (!pl.full && bj == r.bi && nc == r.na) lo = imax(lo, r.nb);  // j >= i
    s.jbase = row.start - row.nd_lo;
    if (iabs(r.na - nc) <= g.w) { s.dlo = imax(lo, r.nb - g.w); s.dhi = imin(hi, r.nb + g.w); }
    else { s.dlo = 0; s.dhi = -1; }
    if (iabs(r.nb - nc) <= g.w) { s.xlo = imax(lo, r.na - g.w); s.xhi = imin(hi, r.na + g.w); }
    else { s.xlo = 0; s.xhi = -1; }
    return s;
}

// stored sets of a segment (hamiltonian.f90:188-198)
BS2E_HD Union2 seg_H(const Segment& s, const Coupling& c)
{
    const bool d = c.same || c.dirany, x = c.same || c.exany;
    return union2(d ? s.dlo : 0, d ? s.dhi : -1, x ? s.xlo : 0, x ? s.xhi : -1);
}
BS2E_HD Union2 seg_S(const Segment& s, const Coupling& c)
{
    return union2(c.same ? s.dlo : 0, c.same ? s.dhi : -1, c.samex ? s.xlo : 0, c.samex ? s.xhi : -1);
}

// number of stored H and S entries of one row (replaces hamiltonian.f90:348-416)
BS2E_HD void row_count(const Geom& g, const Plan& pl, int i, long long* nH, long long* nS)
{
    const RowInfo r = row_info(pl, i);
    long long cH = 0, cS = 0;
    for (int bj = (pl.full ? 0 : r.bi); bj < pl.nblk; ++bj) {
        const Coupling c = coupling(pl, r, bj);
        if (!c.any()) continue;
        const Union2 win = nc_windows(g, pl, r, bj);
        for (int q = 0; q < win.n; ++q)
            for (int nc = win.lo[q]; nc <= win.hi[q]; ++nc) {
                const Segment s = segment(g, pl, r, bj, nc);
                cH += union2_count(seg_H(s, c));
                if (c.same) cS += union2_count(seg_S(s, c));
            }
    }
    *nH = cH;
    *nS = cS;
}

// Traversal of the stored H entries of one row in ascending column order, in
// chunks of <= 32 consecutive n_d (one warp step each).  f(bj, nc, seg, cpl,
// base, hi): columns n_d = base .. min(base+31, hi) of segment seg.
template <class F>
BS2E_HD void for_each_chunk(const Geom& g, const Plan& pl, const RowInfo& r, F&& f)
{
    for (int bj = (pl.full ? 0 : r.bi); bj < pl.nblk; ++bj) {
        const Coupling c = coupling(pl, r, bj);
        if (!c.any()) continue;
        const Union2 win = nc_windows(g, pl, r, bj);
        for (int q = 0; q < win.n; ++q)
            for (int nc = win.lo[q]; nc <= win.hi[q]; ++nc) {
                const Segment s = segment(g, pl, r, bj, nc);
                const Union2 uh = seg_H(s, c);
                for (int z = 0; z < uh.n; ++z)
                    for (int base = uh.lo[z]; base <= uh.hi[z]; base += 32)
                        f(bj, nc, s, c, base, uh.hi[z]);
            }
    }
}

// one-particle matrices in band storage: M[l][n][n'-n+w], complex interleaved;
// entries outside the band are exact zeros in the reference's dense arrays
struct OneBody {
    const double* Hb;  // [lmax+1][nb+1][2w+1][2]
    const double* Sb;  // [nb+1][2w+1][2]
};

struct Cplx { double re, im; };
BS2E_HD Cplx cmul(Cplx a, Cplx b) { return Cplx{a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re}; }
BS2E_HD Cplx cadd(Cplx a, Cplx b) { return Cplx{a.re + b.re, a.im + b.im}; }

BS2E_HD Cplx band_S(const Geom& g, const OneBody& ob, int n, int np)
{
    const int d = np - n + g.w;
    if (d < 0 || d > 2 * g.w) return Cplx{0.0, 0.0};
    const double* q = ob.Sb + ((size_t)n * (2 * g.w + 1) + d) * 2;
    return Cplx{q[0], q[1]};
}
BS2E_HD Cplx band_H(const Geom& g, const OneBody& ob, int l, int n, int np)
{
    const int d = np - n + g.w;
    if (d < 0 || d > 2 * g.w) return Cplx{0.0, 0.0};
    const double* q = ob.Hb + (((size_t)l * (g.nb + 1) + n) * (2 * g.w + 1) + d) * 2;
    return Cplx{q[0], q[1]};
}

struct Element { Cplx H, S; bool storeS; };

// k ranges of a pair in packed form: direct terms k = dlo + 2i, i < nkd; exchange likewise
struct PairK { int dlo, nkd, xlo, nkx; };
BS2E_HD PairK pair_k(KRange kr)
{
    PairK p;
    p.dlo = kr.dlo;
    p.nkd = kr.dhi >= kr.dlo ? (kr.dhi - kr.dlo) / 2 + 1 : 0;
    p.xlo = kr.xlo;
    p.nkx = kr.xhi >= kr.xlo ? (kr.xhi - kr.xlo) / 2 + 1 : 0;
    return p;
}

// ---- element formulas -------------------------------------------------------
//   mat_els.f90:552-571 r_12_tens, :608-633 c_mat_neq_tens,
//   :664-678 S_mat_neq, :697-715 H_1p_neq; hamiltonian.f90:183-193

// one-body and overlap part of an entry whose column block equals the row block
// (H_1p_neq, S_mat_neq); lc, ld = l values of the column block (= those of the row)
BS2E_HD void one_body_terms(const Geom& g, const Plan& pl, const OneBody& ob, const RowInfo& r,
                            bool same, bool samex, int lc, int ld, int nc, int nd, Cplx* hout,
                            Cplx* sout)
{
    Cplx h = Cplx{0.0, 0.0}, s = Cplx{0.0, 0.0};
    if (same) {
        const Cplx Sbd = band_S(g, ob, r.nb, nd), Sac = band_S(g, ob, r.na, nc);
        h = cadd(h, cmul(band_H(g, ob, r.la, r.na, nc), Sbd));
        h = cadd(h, cmul(band_H(g, ob, r.lb, r.nb, nd), Sac));
        s = cadd(s, cmul(Sac, Sbd));
    }
    if (samex) {
        const double sgn = ((pl.L + lc + ld) & 1) ? -1.0 : 1.0;
        const Cplx Sbc = band_S(g, ob, r.nb, nc), Sad = band_S(g, ob, r.na, nd);
        Cplx hx = cadd(cmul(band_H(g, ob, r.la, r.na, nd), Sbc),
                       cmul(band_H(g, ob, r.lb, r.nb, nc), Sad));
        h = cadd(h, Cplx{hx.re * sgn, hx.im * sgn});
        Cplx sx2 = cmul(Cplx{sgn * Sad.re, sgn * Sad.im}, Sbc);
        s = cadd(s, sx2);
    }
    *hout = h;
    *sout = s;
}

// value of the (i,j) entry, j = (bj, nc, nd), with R^k gathered from the global tensor R[k][p1][p2] (row kernel); the
// stride between consecutive terms of one parity is two planes
BS2E_HD Element element_value(const Geom& g, const Plan& pl, const OneBody& ob,
                              const double* R, const RowInfo& r, const Coupling& c,
                              int bj, int nc, int nd, bool sup, bool sup_ex)
{
    const size_t plane = (size_t)g.P * g.ldP;
    const size_t cpl = ((size_t)r.bi * pl.nblk + bj);
    const PairK pk = pair_k(pl.krange[cpl]);
    const double* Rd = R;
    const double* Rx = R;
    if (sup) Rd = R + (size_t)pair_index(g, r.na, nc) * g.ldP + pair_index(g, r.nb, nd) + pk.dlo * plane;
    if (sup_ex) Rx = R + (size_t)pair_index(g, r.nb, nc) * g.ldP + pair_index(g, r.na, nd) + pk.xlo * plane;
    const double* ad = pl.angD + cpl * g.K1;
    const double* ax = pl.angX + cpl * g.K1;
    Element e;
    e.H = Cplx{0.0, 0.0};
    e.S = Cplx{0.0, 0.0};
    const bool allowed = (sup && c.dirany) || (sup_ex && c.exany);
    if (allowed) {
        double res = 0.0;
        if (sup) {
            double acc = 0.0;
            for (int i = 0; i < pk.nkd; ++i) acc += Rd[(size_t)(2 * i) * plane] * ad[pk.dlo + 2 * i];
            res += acc;
        }
        if (sup_ex) {
            double acc = 0.0;
            for (int i = 0; i < pk.nkx; ++i) acc += Rx[(size_t)(2 * i) * plane] * ax[pk.xlo + 2 * i];
            res += acc;
        }
        e.H.re = res;
    }
    e.storeS = (sup && c.same) || (sup_ex && c.samex);
    if (e.storeS) {
        const BlockDesc bc = pl.blk[bj];
        Cplx h, s;
        one_body_terms(g, pl, ob, r, c.same, c.samex, bc.l1, bc.l2, nc, nd, &h, &s);
        e.H = cadd(e.H, h);
        e.S = s;
    }
    return e;
}

}  // namespace bs2e
