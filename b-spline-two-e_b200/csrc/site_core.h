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

// ---- one stored entry of a (row, column block) pair --------------------------
// Everything that is uniform over the pair.  Pointers into shared memory in the
// kernel, into plain arrays in the CPU checker.
struct PairCtx {
    const SiteEntry* Tb;        // [nnc] clipped windows of the column block
    const unsigned short* hpq;  // [nnc+1] prefix of stored H entries over the n_c slots
    const unsigned short* spq;  // [nnc]   same for S (diagonal pair only)
    const double* Rv;           // staged R^k values, Rv[k*nsmax + slot]
    const double* wa_d;         // packed direct factors, k = pk.dlo + 2i
    const double* wa_x;         // packed exchange factors
    int nsmax;
    PairK pk;
    int bj;
    int win;                    // single-window form: kModeD or kModeX; else kModeDX
    bool diag;                  // column block == row block: one-body terms and S
    bool dirany, exany, samex, cut;
    long long hbase, sbase;     // first H entry of the pair / first S entry of the row
};

// How a pair is walked: with one window only (all entries of the other window are
// clipped away for this site and column block, or the mode stores one window) the
// n_c slots hold <= 2w+1 consecutive n_d; otherwise the union of both windows.
BS2E_HD int pair_window(int mode, int totD, int totX)
{
    if (mode == kModeD || mode == kModeX) return mode;
    if (totX == 0) return kModeD;
    if (totD == 0) return kModeX;
    return kModeDX;
}

// position of n_c value v in the ascending union of the two n_c ranges
BS2E_HD int union_pos(const Site& s, int v)
{
    return below(s.cDlo, s.cDhi, v) + below(s.cXlo, s.cXhi, v) -
           below(imax(s.cDlo, s.cXlo), imin(s.cDhi, s.cXhi), v);
}

// A pair is walked in <= 3 segments of consecutive n_c slots [q0,q1): slots that lie in
// one n_c range only hold one window (win = kModeD / kModeX, <= 2w+1 consecutive n_d);
// slots in both ranges (|n_a - n_b| <= 2w) hold the union (win = kModeDX).
struct Seg { int q0, q1, win; };
struct Segs { Seg a, b, c; };  // ascending n_c; empty segments have q1 <= q0
BS2E_HD Seg segs_at(const Segs& g3, int t) { return t == 0 ? g3.a : (t == 1 ? g3.b : g3.c); }

BS2E_HD Seg seg_of_range(const Site& s, int lo, int hi, int win)
{
    Seg r;
    r.q0 = union_pos(s, lo);
    r.q1 = hi >= lo ? union_pos(s, hi) + 1 : r.q0;
    r.win = win;
    return r;
}

BS2E_HD Segs pair_segments(const Site& s, int win, bool cut)
{
    Segs o;
    o.a = o.b = o.c = Seg{0, 0, kModeD};
    if (win == kModeD) {
        o.a = seg_of_range(s, s.cDlo, s.cDhi, kModeD);
    } else if (win == kModeX) {
        o.a = seg_of_range(s, s.cXlo, s.cXhi, kModeX);
    } else {
        const int ilo = imax(s.cDlo, s.cXlo), ihi = imin(s.cDhi, s.cXhi);
        const bool dlow = s.cDlo <= s.cXlo, dhigh = s.cDhi > s.cXhi;
        if (ihi < ilo) {  // disjoint n_c ranges
            o.a = seg_of_range(s, dlow ? s.cDlo : s.cXlo, dlow ? s.cDhi : s.cXhi, dlow ? kModeD : kModeX);
            o.c = seg_of_range(s, dlow ? s.cXlo : s.cDlo, dlow ? s.cXhi : s.cDhi, dlow ? kModeX : kModeD);
        } else {
            o.a = seg_of_range(s, imin(s.cDlo, s.cXlo), ilo - 1, dlow ? kModeD : kModeX);
            o.b = seg_of_range(s, ilo, ihi, kModeDX);
            o.c = seg_of_range(s, ihi + 1, imax(s.cDhi, s.cXhi), dhigh ? kModeD : kModeX);
        }
    }
    if (cut) {  // slots with n_c < n_a are cut away (j >= i)
        const int qc = union_pos(s, s.na);
        o.a.q0 = imax(o.a.q0, qc);
        o.b.q0 = imax(o.b.q0, qc);
        o.c.q0 = imax(o.c.q0, qc);
    }
    return o;
}

// entry number idx (ascending n_d) of n_c slot q.  WIDE = false: one window (win);
// WIDE = true: union of both windows, cnt = number of entries of the slot.
template <bool WIDE>
BS2E_HD void site_lane(const Geom& g, const Plan& pl, const OneBody& ob, const Site& s,
                       const RowInfo& r, const PairCtx& pc, int win, int q, int idx, int cnt,
                       long long* Hidx, double* Hdat, long long* Sidx, double* Sdat)
{
    SiteEntry e = pc.Tb[q];
    const int nc = site_nc(s, q);
    if (pc.cut) e = entry_cut(e, s, nc);
    int nd;
    if (!WIDE) {
        const int lo = win == kModeD ? (int)e.dlo : (int)e.xlo;
        const int hi = win == kModeD ? (int)e.dhi : (int)e.xhi;
        nd = lo + idx;
        if (nd > hi) return;
    } else {
        if (idx >= cnt) return;
        nd = entry_nd(e, true, true, idx);
    }
    // inside the stored set the clipping is common to both windows
    const bool sup = site_nc_inD(s, nc) && nd >= s.dDlo && nd <= s.dDhi;
    const bool sup_ex = site_nc_inX(s, nc) && nd >= s.dXlo && nd <= s.dXhi;
    const int stride = 2 * pc.nsmax;
    double re = 0.0, im = 0.0;
    const bool allowed = (sup && pc.dirany) || (sup_ex && pc.exany);
    if (allowed) {
        double res = 0.0;
        if (sup) res += k_dot(pc.Rv + pc.pk.dlo * pc.nsmax + site_slotD(s, nc, nd), stride, pc.wa_d, pc.pk.nkd);
        if (sup_ex) res += k_dot(pc.Rv + pc.pk.xlo * pc.nsmax + site_slotX(s, nc, nd), stride, pc.wa_x, pc.pk.nkx);
        re = res;
    }
    const long long j = (long long)e.jbase + nd;
    if (pc.diag) {
        const bool storeS = sup || (sup_ex && pc.samex);
        if (storeS) {
            const BlockDesc bc = pl.blk[pc.bj];
            Cplx h, sv;
            one_body_terms(g, pl, ob, r, true, pc.samex, bc.l1, bc.l2, nc, nd, &h, &sv);
            re += h.re;
            im += h.im;
            const long long pos =
                pc.sbase + pc.spq[q] + union_below(true, e.dlo, e.dhi, pc.samex, e.xlo, e.xhi, nd);
            Sidx[pos] = j;
            Sdat[2 * pos] = sv.re;
            Sdat[2 * pos + 1] = sv.im;
        }
    }
    const long long pos = pc.hbase + pc.hpq[q] + idx;
    Hidx[pos] = j;
#if defined(__CUDA_ARCH__)
    *reinterpret_cast<double2*>(Hdat + 2 * pos) = make_double2(re, im);
#else
    Hdat[2 * pos] = re;
    Hdat[2 * pos + 1] = im;
#endif
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

}  // namespace bs2e
