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
// CANDIDATE list of the site; slot numbering of the staged R^k values is
// D-window row-major followed by X-window row-major.
#pragma once
#include "core.h"

namespace bs2e {

struct Site {
    int na, nb;
    Union2 ncu;  // n_c values of the candidates (union of the two n_c windows)
    int nnc;     // number of n_c slots
    int cDlo, cDhi, dDlo, dDhi, dw, nD;  // D window: n_c range, n_d range, width, slots
    int cXlo, cXhi, dXlo, dXhi, xw, nX;  // X window
};

BS2E_HD Site make_site(const Geom& g, int na, int nb)
{
    Site s;
    s.na = na;
    s.nb = nb;
    s.cDlo = imax(1, na - g.w); s.cDhi = imin(g.nb, na + g.w);
    s.dDlo = imax(1, nb - g.w); s.dDhi = imin(g.nb, nb + g.w);
    s.cXlo = s.dDlo; s.cXhi = s.dDhi;
    s.dXlo = s.cDlo; s.dXhi = s.cDhi;
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

// slot of a column (n_c,n_d) in the staged D / X window (caller guarantees membership)
BS2E_HD int site_slotD(const Site& s, int nc, int nd) { return (nc - s.cDlo) * s.dw + (nd - s.dDlo); }
BS2E_HD int site_slotX(const Site& s, int nc, int nd) { return s.nD + (nc - s.cXlo) * s.xw + (nd - s.dXlo); }

// address of the R^k(k=0) value that fills a slot; add k*P*ldP for multipole k
BS2E_HD size_t site_slot_source(const Geom& g, const Site& s, int slot)
{
    if (slot < s.nD) {
        const int nc = s.cDlo + slot / s.dw, nd = s.dDlo + slot % s.dw;
        return (size_t)pair_index(g, s.na, nc) * g.ldP + pair_index(g, s.nb, nd);
    }
    const int t = slot - s.nD;
    const int nc = s.cXlo + t / s.xw, nd = s.dXlo + t % s.xw;
    return (size_t)pair_index(g, s.nb, nc) * g.ldP + pair_index(g, s.na, nd);
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

// rho-th (0-based, ascending) stored n_d of the entry
BS2E_HD int entry_nd(const SiteEntry& e, bool useD, bool useX, int rho)
{
    const Union2 u = union2(useD ? (int)e.dlo : 1, useD ? (int)e.dhi : 0,
                            useX ? (int)e.xlo : 1, useX ? (int)e.xhi : 0);
    return union2_at(u, rho);
}

// largest q in [0,nnc) with prefix[q] <= o  (prefix ascending, prefix[0] = 0)
BS2E_HD int prefix_search(const unsigned short* prefix, int nnc, int top, int o)
{
    int q = 0;
    for (int st = top; st > 0; st >>= 1) {
        const int m = q + st;
        if (m < nnc && (int)prefix[m] <= o) q = m;
    }
    return q;
}
BS2E_HD int search_top(int nnc)
{
    int top = 1;
    while (top * 2 < nnc) top *= 2;
    return top;
}

}  // namespace bs2e
