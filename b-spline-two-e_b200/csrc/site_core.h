// site_core.h -- lane-level bodies of the site-centric stage-C fill kernel
// (block.cu: site_fill_kernel).  __host__ __device__ like core.h so that the
// CPU logic checker (tests/hostcheck) executes the same statements.
//
// A radial SITE is a pair (n_a, n_b) of B-spline indices.  Every row
// (l_a,l_b; n_a,n_b) of a symmetry block that sits on the site couples to the
// same columns (n_c,n_d) up to clipping:
//     direct   window  D : |n_c - n_a| <= w, |n_d - n_b| <= w
//     exchange window  X : |n_c - n_b| <= w, |n_d - n_a| <= w
// (hamiltonian.f90:164-165) and reads the same values of R^k; only the
// angular factors and the (l_c,l_d) clipping differ.  The union of the two
// windows, enumerated in CSR order (n_c ascending, n_d ascending), is the
// CANDIDATE list of the site.  One thread of the site's CTA owns one candidate
// column for the whole site: it keeps the column's R^k values of all
// multipoles (both windows) in registers, and for every column block works out
// once per storage mode whether and where the column is stored; the rows of
// the site then differ only in their angular factors and output base.
#pragma once
#include "core.h"

namespace bs2e {

struct Site {
    int na, nb;
    Union2 ncu;  // n_c values of the candidates (union of the two n_c windows)
    int nnc;     // number of n_c slots
    int cDlo, cDhi, dDlo, dDhi, dw, nD;  // D window: n_c range, n_d range, width, entries
    int cXlo, cXhi, dXlo, dXhi, xw, nX;  // X window
};

// wantX = false: every exchange window of the site is clipped away by the basis
// (n_a - w exceeds the largest n_2 of any configuration), the site is treated as if
// it had the direct window only (half the n_c slots, no exchange values to load)
BS2E_HD bool site_wants_X(const Geom& g, int max_nd, int na) { return imax(1, na - g.w) <= max_nd; }

BS2E_HD Site make_site(const Geom& g, int na, int nb, bool wantX)
{
    Site s;
    s.na = na;
    s.nb = nb;
    s.cDlo = imax(1, na - g.w); s.cDhi = imin(g.nb, na + g.w);
    s.dDlo = imax(1, nb - g.w); s.dDhi = imin(g.nb, nb + g.w);
    if (wantX) {
        s.cXlo = s.dDlo; s.cXhi = s.dDhi;
        s.dXlo = s.cDlo; s.dXhi = s.cDhi;
    } else {
        s.cXlo = 1; s.cXhi = 0;
        s.dXlo = 1; s.dXhi = 0;
    }
    s.dw = s.dDhi - s.dDlo + 1;
    s.xw = s.dXhi - s.dXlo + 1;
    s.nD = (s.cDhi - s.cDlo + 1) * s.dw;
    s.nX = (s.cXhi - s.cXlo + 1) * s.xw;
    s.ncu = union2(s.cDlo, s.cDhi, s.cXlo, s.cXhi);
    s.nnc = union2_count(s.ncu);
    return s;
}

// upper bounds used to size shared memory for a launch
BS2E_HD int site_max_nc(const Geom& g) { return 2 * (2 * g.w + 1); }
BS2E_HD int site_max_slots(const Geom& g) { return 2 * (2 * g.w + 1) * (2 * g.w + 1); }

BS2E_HD int union2_at(const Union2& u, int idx)
{
    const int len0 = u.hi[0] - u.lo[0] + 1;
    return idx < len0 ? u.lo[0] + idx : u.lo[1] + (idx - len0);
}

BS2E_HD int site_nc(const Site& s, int q) { return union2_at(s.ncu, q); }
BS2E_HD bool site_nc_inD(const Site& s, int nc) { return nc >= s.cDlo && nc <= s.cDhi; }
BS2E_HD bool site_nc_inX(const Site& s, int nc) { return nc >= s.cXlo && nc <= s.cXhi; }

// n_d values of the candidates of one n_c
BS2E_HD Union2 site_nd_union(const Site& s, int nc)
{
    const bool d = site_nc_inD(s, nc), x = site_nc_inX(s, nc);
    return union2(d ? s.dDlo : 0, d ? s.dDhi : -1, x ? s.dXlo : 0, x ? s.dXhi : -1);
}

// ---- per (site, column block) tables ---------------------------------------
// Clipping of the two windows by the configurations that exist in column block
// bj depends on the site and on bj only, not on the row: per n_c slot the
// clipped direct / exchange n_d intervals (empty: lo=1, hi=0) and the
// configuration index base j = jbase + n_d.
struct alignas(16) SiteEntry {
    unsigned short dlo, dhi, xlo, xhi;
    int jbase;
    int pad;
};

BS2E_HD SiteEntry site_entry(const Geom& g, const Plan& pl, const Site& s, int bj, int q)
{
    const int nc = site_nc(s, q);
    const NcRow row = pl.ncrow[(size_t)bj * (g.nb + 1) + nc];
    const int lo = row.nd_lo, hi = row.nd_hi;
    int dlo = 1, dhi = 0, xlo = 1, xhi = 0;
    if (hi >= lo) {
        if (site_nc_inD(s, nc)) { dlo = imax(lo, s.dDlo); dhi = imin(hi, s.dDhi); }
        if (site_nc_inX(s, nc)) { xlo = imax(lo, s.dXlo); xhi = imin(hi, s.dXhi); }
        if (dhi < dlo) { dlo = 1; dhi = 0; }
        if (xhi < xlo) { xlo = 1; xhi = 0; }
    }
    SiteEntry e;
    e.dlo = (unsigned short)dlo; e.dhi = (unsigned short)dhi;
    e.xlo = (unsigned short)xlo; e.xhi = (unsigned short)xhi;
    e.jbase = row.start - row.nd_lo;
    e.pad = 0;
    return e;
}

// j >= i inside the diagonal pair (upper triangle, hamiltonian.f90:150 low => i):
// all rows of the site share (n_a,n_b), so the cut is a property of the site
BS2E_HD SiteEntry entry_cut(SiteEntry e, const Site& s, int nc)
{
    if (nc < s.na) { e.dlo = 1; e.dhi = 0; e.xlo = 1; e.xhi = 0; }
    else if (nc == s.na) {
        if ((int)e.dhi >= (int)e.dlo) { e.dlo = (unsigned short)imax(e.dlo, s.nb); if (e.dhi < e.dlo) { e.dlo = 1; e.dhi = 0; } }
        if ((int)e.xhi >= (int)e.xlo) { e.xlo = (unsigned short)imax(e.xlo, s.nb); if (e.xhi < e.xlo) { e.xlo = 1; e.xhi = 0; } }
    }
    return e;
}

// which windows a (row block, column block) pair stores in H (hamiltonian.f90:195-198)
constexpr int kModeD = 0, kModeX = 1, kModeDX = 2, kModeDiag = 3, kModes = 4;

BS2E_HD int pair_mode(const Plan& pl, const RowInfo& r, int bj)
{
    if (!pl.full && bj < r.bi) return -1;
    const Coupling c = coupling(pl, r, bj);
    if (c.same) return kModeDiag;
    if (c.dirany) return c.exany ? kModeDX : kModeD;
    return c.exany ? kModeX : -1;
}
BS2E_HD bool mode_useD(int mode) { return mode != kModeX; }
BS2E_HD bool mode_useX(int mode) { return mode != kModeD; }

BS2E_HD int ilen(int lo, int hi) { return hi >= lo ? hi - lo + 1 : 0; }
// |[lo,hi] ∩ (-inf, x)|
BS2E_HD int below(int lo, int hi, int x) { return ilen(lo, imin(hi, x - 1)); }

// stored entries of the union of (useA ? A : {}) and (useB ? B : {}) that lie below x
BS2E_HD int union_below(bool useA, int alo, int ahi, bool useB, int blo, int bhi, int x)
{
    int c = 0;
    if (useA) c += below(alo, ahi, x);
    if (useB) c += below(blo, bhi, x);
    if (useA && useB) c -= below(imax(alo, blo), imin(ahi, bhi), x);
    return c;
}

BS2E_HD int entry_count(const SiteEntry& e, bool useD, bool useX)
{
    return union_below(useD, e.dlo, e.dhi, useX, e.xlo, e.xhi, 0x7fffffff);
}

// position of n_c value v in the ascending union of the two n_c ranges
BS2E_HD int union_pos(const Site& s, int v)
{
    return below(s.cDlo, s.cDhi, v) + below(s.cXlo, s.cXhi, v) -
           below(imax(s.cDlo, s.cXlo), imin(s.cDhi, s.cXhi), v);
}

// ---- per (row, column block) pair: storage mode, k parity, offset in the row ----
// One 32-bit word per pair of the row group: offset of the pair's first H entry
// inside its row, the storage mode (after effective_mode) and the parity of the
// multipoles of the direct term, pd = (l_a + l_c) & 1 (wigner_tools.f90:131: the
// factor vanishes unless l_a+k+l_c is even).  The exchange parity follows from
// the column block alone: px = pd ^ ((l_c + l_d) & 1).
constexpr unsigned kPmValid = 0x80000000u;
BS2E_HD unsigned pm_pack(int off, int mode, int pd) { return kPmValid | ((unsigned)(mode | (pd << 2)) << 24) | (unsigned)off; }
BS2E_HD bool pm_valid(unsigned v) { return (v & kPmValid) != 0; }
BS2E_HD int pm_off(unsigned v) { return (int)(v & 0xffffffu); }
BS2E_HD int pm_mode(unsigned v) { return (int)((v >> 24) & 3u); }
BS2E_HD int pm_pd(unsigned v) { return (int)((v >> 26) & 1u); }

// rows of a group are addressed by bit masks (<= 32 rows per group)
BS2E_HD int popc32(unsigned m)
{
#if defined(__CUDA_ARCH__)
    return __popc(m);
#else
    return __builtin_popcount(m);
#endif
}
BS2E_HD int lowest_bit(unsigned m)
{
#if defined(__CUDA_ARCH__)
    return __ffs((int)m) - 1;
#else
    return __builtin_ctz(m);
#endif
}

// ---- candidate columns of a site ------------------------------------------------
// The columns (n_c,n_d) a row of the site can couple to, in CSR order (n_c, then n_d
// ascending), independent of the column block: the union of the two windows (list
// DX), or the direct window alone when every exchange window of the site is clipped
// away.  Thread t of the site's CTA OWNS candidate t for the whole site: it loads the
// candidate's R^k values for all multipoles into registers once and then walks every
// (column block, row) pair of the site.
BS2E_HD int site_cand_DX_count(const Site& s, int q) { return union2_count(site_nd_union(s, site_nc(s, q))); }

// largest q in [0,nnc) with prefix[q] <= p; the trip count depends on nnc only
template <class T>
BS2E_HD int prefix_slot(const T* prefix, int nnc, int p)
{
    int lo = 0, n = nnc;  // the slot lies in [lo, lo+n)
    while (n > 1) {
        const int half = n >> 1;
        if ((int)prefix[lo + half] <= p) lo += half;
        n -= half;
    }
    return lo;
}

struct OwnCand {
    int q, nc, nd;
    bool inD, inX;           // member of the direct / exchange window of the site
    int rowD, colD;          // R^k(n_a n_b; n_c n_d) = R[k][rowD][colD]
    int rowX, colX;          // R^k(n_a n_b; n_d n_c) = R[k][rowX][colX]  (= R^k(n_b n_a; n_c n_d))
};

// candidate t of the site; cprefix[q] = number of list-DX entries of the n_c slots
// before q (only read when wantX)
BS2E_HD OwnCand site_own_cand(const Geom& g, const Site& s, const int* cprefix, bool wantX, int t)
{
    OwnCand c;
    if (wantX) {
        c.q = prefix_slot(cprefix, s.nnc, t);
        c.nc = site_nc(s, c.q);
        c.nd = union2_at(site_nd_union(s, c.nc), t - cprefix[c.q]);
    } else {
        const int r = t / s.dw;
        c.nc = s.cDlo + r;
        c.nd = s.dDlo + (t - r * s.dw);
        c.q = union_pos(s, c.nc);
    }
    c.inD = site_nc_inD(s, c.nc) && c.nd >= s.dDlo && c.nd <= s.dDhi;
    c.inX = wantX && site_nc_inX(s, c.nc) && c.nd >= s.dXlo && c.nd <= s.dXhi;
    c.rowD = c.colD = c.rowX = c.colX = 0;
    if (c.inD) { c.rowD = pair_index(g, s.na, c.nc); c.colD = pair_index(g, s.nb, c.nd); }
    if (c.inX) { c.rowX = pair_index(g, s.nb, c.nc); c.colX = pair_index(g, s.na, c.nd); }
    return c;
}
BS2E_HD int site_num_cand(const Site& s, const int* cprefix, bool wantX) { return wantX ? cprefix[s.nnc] : s.nD; }

// What candidate c is inside the column list of (column block, storage mode):
// whether it is stored, through which window(s), at which rank and column index.
struct ModeSlot {
    bool sup, sup_ex;  // stored through the direct / exchange window
    int rank;          // position inside the (column block, mode) list
    int jcol;          // configuration index of the column, 1-based
    SiteEntry e;       // clipped windows of the n_c slot (diagonal cut applied)
};

BS2E_HD ModeSlot site_mode_slot(const Site& s, const OwnCand& c, SiteEntry e, const unsigned short* hpq,
                                int mode, bool cut)
{
    ModeSlot m;
    if (cut) e = entry_cut(e, s, c.nc);
    m.e = e;
    const bool useD = mode_useD(mode), useX = mode_useX(mode);
    m.sup = useD && c.nd >= (int)e.dlo && c.nd <= (int)e.dhi;
    m.sup_ex = useX && c.nd >= (int)e.xlo && c.nd <= (int)e.xhi;
    // only read for stored columns: inside a single window the rank is the distance from its start
    if (!useX) m.rank = (int)hpq[c.q] + (c.nd - (int)e.dlo);
    else if (!useD) m.rank = (int)hpq[c.q] + (c.nd - (int)e.xlo);
    else m.rank = (int)hpq[c.q] + union_below(true, e.dlo, e.dhi, true, e.xlo, e.xhi, c.nd);
    m.jcol = e.jbase + c.nd;
    return m;
}

// ---- angular factors packed for the site kernel ----------------------------------
// angP[(bi*nblk + bj)*2*nkp + l]: l < nkp: direct factor of k = pd + 2l, pd = (la+lc)&1;
// nkp + l: exchange factor (times (-1)^(lc+ld+L)) of k = px + 2l, px = (la+ld)&1; zero
// past max_k and where the reference skips the term (|ang| < 5e-16, mat_els.f90:568).
// nkp is even so that both halves are 16-byte aligned.
BS2E_HD int site_nkp(int kmax) { return (((kmax + 1) / 2) + 1) & ~1; }

// sum over the multipoles of one parity in ascending order (mat_els.f90:566-570);
// R[k] holds all multipoles of the candidate (zeros past max_k)
template <int KMAX>
BS2E_HD double site_dot_par(const double* cf, const double* R, int par)
{
    double acc = 0.0;
    if (par == 0) {
#pragma unroll
        for (int i = 0; 2 * i < KMAX; ++i) acc += cf[i] * R[2 * i];
    } else {
#pragma unroll
        for (int i = 0; 2 * i + 1 < KMAX; ++i) acc += cf[i] * R[2 * i + 1];
    }
    return acc;
}

// ---- one-particle matrices of the site ----------------------------------------------
// The diagonal pair of a row needs H_l(n_a,.), H_l(n_b,.), S(n_a,.), S(n_b,.) inside the
// band only; the CTA copies those 2(l_max+2) band rows to shared memory once:
//   Hs[(l*2 + r)*(2w+1) + d], Ss[r*(2w+1) + d], r = 0: n = n_a, r = 1: n = n_b,
//   d = n' - n + w, complex interleaved.
struct SiteOneBody { const double* Hs; const double* Ss; };
BS2E_HD int site_1p_doubles(const Geom& g, int nl) { return (nl + 1) * 2 * (2 * g.w + 1) * 2; }
// element idx of the copy (idx < site_1p_doubles / 2 complex numbers, H rows first)
BS2E_HD Cplx site_1p_source(const Geom& g, const OneBody& ob, const Site& s, int nl, int idx)
{
    const int bw = 2 * g.w + 1;
    const int row = idx / bw, d = idx - row * bw;
    const int n = (row & 1) ? s.nb : s.na;
    const int np = n + d - g.w;
    if (np < 1 || np > g.nb) return Cplx{0.0, 0.0};
    return row < 2 * nl ? band_H(g, ob, row >> 1, n, np) : band_S(g, ob, n, np);
}
BS2E_HD Cplx site_S(const Geom& g, const SiteOneBody& so, int r, int n, int np)
{
    const int d = np - n + g.w;
    if (d < 0 || d > 2 * g.w) return Cplx{0.0, 0.0};
    const double* q = so.Ss + (r * (2 * g.w + 1) + d) * 2;
    return Cplx{q[0], q[1]};
}
BS2E_HD Cplx site_H(const Geom& g, const SiteOneBody& so, int l, int r, int n, int np)
{
    const int d = np - n + g.w;
    if (d < 0 || d > 2 * g.w) return Cplx{0.0, 0.0};
    const double* q = so.Hs + ((l * 2 + r) * (2 * g.w + 1) + d) * 2;
    return Cplx{q[0], q[1]};
}

// one-body and overlap part of a stored entry of the diagonal pair (column block ==
// row block): adds H_1p to (re,im) and writes the S entry when one is stored.  Same
// statements as one_body_terms (core.h) with the band rows read from the site's copy.
BS2E_HD void site_diag_terms(const Geom& g, const Plan& pl, const SiteOneBody& so, const Site& s, int la, int lb,
                             const OwnCand& c, const ModeSlot& m, bool samex, const unsigned short* spq,
                             long long sbase, double* re, double* im, long long* Sidx, double* Sdat)
{
    const bool storeS = m.sup || (m.sup_ex && samex);
    if (!storeS) return;
    const int nc = c.nc, nd = c.nd;
    Cplx h = Cplx{0.0, 0.0}, sv = Cplx{0.0, 0.0};
    {
        const Cplx Sbd = site_S(g, so, 1, s.nb, nd), Sac = site_S(g, so, 0, s.na, nc);
        h = cadd(h, cmul(site_H(g, so, la, 0, s.na, nc), Sbd));
        h = cadd(h, cmul(site_H(g, so, lb, 1, s.nb, nd), Sac));
        sv = cadd(sv, cmul(Sac, Sbd));
    }
    if (samex) {
        const double sgn = ((pl.L + la + lb) & 1) ? -1.0 : 1.0;
        const Cplx Sbc = site_S(g, so, 1, s.nb, nc), Sad = site_S(g, so, 0, s.na, nd);
        Cplx hx = cadd(cmul(site_H(g, so, la, 0, s.na, nd), Sbc), cmul(site_H(g, so, lb, 1, s.nb, nc), Sad));
        h = cadd(h, Cplx{hx.re * sgn, hx.im * sgn});
        Cplx sx2 = cmul(Cplx{sgn * Sad.re, sgn * Sad.im}, Sbc);
        sv = cadd(sv, sx2);
    }
    *re += h.re;
    *im += h.im;
    const long long pos =
        sbase + spq[c.q] + union_below(true, m.e.dlo, m.e.dhi, samex, m.e.xlo, m.e.xhi, c.nd);
    Sidx[pos] = m.jcol;
    Sdat[2 * pos] = sv.re;
    Sdat[2 * pos + 1] = sv.im;
}

// ---- tensor-core site kernel (site_mma.cu): bit masks over the candidate list ---------------
// The candidates of a site are cut into SEGMENTS of 32; which candidates a (column group,
// storage mode) pair stores is one 32-bit mask per segment plus the number of stored entries
// before the segment, so that a thread finds "stored?" and the rank of its candidate with one
// 8-byte load, an AND and a population count -- for any record, without per-mode arithmetic.
constexpr int kSegCand = 32;
struct alignas(8) MaskWord { unsigned mask, pre; };

// one n_c slot of the candidate list: index of its first candidate and the n_d values it holds
struct CandSlot { int base; Union2 u; };
BS2E_HD CandSlot cand_slot(const Site& s, const int* cprefix, bool wantX, int q)
{
    CandSlot c;
    if (wantX) {
        c.base = cprefix[q];
        c.u = site_nd_union(s, site_nc(s, q));
    } else {
        c.base = q * s.dw;
        c.u = union2(s.dDlo, s.dDhi, 0, -1);
    }
    return c;
}
// candidate index of n_d inside the slot (n_d must belong to the slot's union)
BS2E_HD int cand_index(const CandSlot& c, int nd)
{
    const int len0 = c.u.hi[0] - c.u.lo[0] + 1;
    return c.base + ((c.u.n == 2 && nd > c.u.hi[0]) ? len0 + nd - c.u.lo[1] : nd - c.u.lo[0]);
}
// candidates t0..t1 (inclusive) as (segment, bits) pieces: orfn(segment, bits)
template <class F>
BS2E_HD void mask_range(int t0, int t1, F&& orfn)
{
    for (int seg = t0 / kSegCand; seg <= t1 / kSegCand; ++seg) {
        const int a = imax(t0, seg * kSegCand) - seg * kSegCand, b = imin(t1, seg * kSegCand + kSegCand - 1) - seg * kSegCand;
        const unsigned bits = (b - a == 31) ? 0xffffffffu : (((1u << (b - a + 1)) - 1u) << a);
        orfn(seg, bits);
    }
}
// the direct / exchange intervals of a clipped slot entry as candidate ranges: fd(t0,t1), fx(t0,t1)
template <class FD, class FX>
BS2E_HD void entry_cand_ranges(const CandSlot& c, const SiteEntry& e, FD&& fd, FX&& fx)
{
    if ((int)e.dhi >= (int)e.dlo) fd(cand_index(c, e.dlo), cand_index(c, e.dhi));
    if ((int)e.xhi >= (int)e.xlo) fx(cand_index(c, e.xlo), cand_index(c, e.xhi));
}
BS2E_HD bool mask_has(const MaskWord& w, int bit) { return (w.mask >> bit) & 1u; }
BS2E_HD int mask_rank(const MaskWord& w, int bit) { return (int)w.pre + popc32(w.mask & ((1u << bit) - 1u)); }

// record of the tensor-core kernel: one (row, column group) pair of the site
struct alignas(16) MmaRec {
    long long hpos;       // 0-based position of the pair's first H entry
    int cf;               // index of the pair's packed factors, bi*nblk + bj (-1: padding)
    unsigned short tbl;   // row of the mask table
    unsigned char bj;     // column group (row of the jbase table)
    unsigned char ri;     // row of the group (diagonal pairs)
};
// rows of the mask table: (column group, mode) first, then the H and S masks of the diagonal pair of
// each row of the group, then one all-zero row for padding records
BS2E_HD int mask_modes(bool wantX) { return wantX ? 3 : 1; }
BS2E_HD int mask_rows(int nblk, int G, bool wantX) { return nblk * mask_modes(wantX) + 2 * G + 1; }
BS2E_HD int mask_row_pair(int bj, int mode, bool wantX) { return wantX ? bj * 3 + mode : bj; }
BS2E_HD int mask_row_diagH(int nblk, int ri, bool wantX) { return nblk * mask_modes(wantX) + ri; }
BS2E_HD int mask_row_diagS(int nblk, int G, int ri, bool wantX) { return nblk * mask_modes(wantX) + G + ri; }
BS2E_HD int mask_row_zero(int nblk, int G, bool wantX) { return nblk * mask_modes(wantX) + 2 * G; }

// one-body and overlap part of an entry of the diagonal pair (H_1p_neq, S_mat_neq: mat_els.f90:664-678,
// 697-715), band rows from the site's shared-memory copy; same statements as one_body_terms (core.h)
BS2E_HD void site_onebody_at(const Geom& g, int L, const SiteOneBody& so, int na, int nb, int la, int lb,
                             int nc, int nd, bool samex, Cplx* hout, Cplx* sout)
{
    Cplx h = Cplx{0.0, 0.0}, sv = Cplx{0.0, 0.0};
    {
        const Cplx Sbd = site_S(g, so, 1, nb, nd), Sac = site_S(g, so, 0, na, nc);
        h = cadd(h, cmul(site_H(g, so, la, 0, na, nc), Sbd));
        h = cadd(h, cmul(site_H(g, so, lb, 1, nb, nd), Sac));
        sv = cadd(sv, cmul(Sac, Sbd));
    }
    if (samex) {
        const double sgn = ((L + la + lb) & 1) ? -1.0 : 1.0;
        const Cplx Sbc = site_S(g, so, 1, nb, nc), Sad = site_S(g, so, 0, na, nd);
        Cplx hx = cadd(cmul(site_H(g, so, la, 0, na, nd), Sbc), cmul(site_H(g, so, lb, 1, nb, nc), Sad));
        h = cadd(h, Cplx{hx.re * sgn, hx.im * sgn});
        Cplx sx2 = cmul(Cplx{sgn * Sad.re, sgn * Sad.im}, Sbc);
        sv = cadd(sv, sx2);
    }
    *hout = h;
    *sout = sv;
}
BS2E_HD void site_onebody(const Geom& g, const Plan& pl, const SiteOneBody& so, const Site& s, int la, int lb,
                          int nc, int nd, bool samex, Cplx* hout, Cplx* sout)
{
    site_onebody_at(g, pl.L, so, s.na, s.nb, la, lb, nc, nd, samex, hout, sout);
}

// shared-memory stride (doubles) of a record's factor row: the B fragments of mma.m8n8k4 read
// [record = lane/4][k = lane%4] as 8-byte words, conflict-free when stride mod 16 is 4 or 12
BS2E_HD constexpr int mma_cf_stride(int cfs)
{
    int s = cfs;
    while (s % 16 != 4 && s % 16 != 12) s += 2;
    return s;
}

// storage mode that the pair effectively stores: a D+X pair whose exchange
// (direct) windows are all empty for this site and column block is a pure D (X) pair
BS2E_HD int effective_mode(int mode, int totD, int totX)
{
    if (mode == kModeDX) {
        if (totX == 0) return kModeD;
        if (totD == 0) return kModeX;
    }
    return mode;
}

// number of multipole registers the site kernel is instantiated for (0: none, the
// row kernel is used)
BS2E_HD int site_kmax_for(int K1)
{
    if (K1 <= 7) return 7;
    if (K1 <= 13) return 13;
    if (K1 <= 21) return 21;
    if (K1 <= 31) return 31;
    return 0;
}

}  // namespace bs2e
