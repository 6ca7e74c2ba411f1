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
    int cDlo, cDhi, dDlo, dDhi, dw, nD;  // D window: n_c range, n_d range, width, entries
    int cXlo, cXhi, dXlo, dXhi, xw, nX;  // X window
    // staged R^k values: Rv[win*xoff + k*kst + r*cpad + c], r = n_c - first n_c of the
    // window, c = n_d - first n_d (the box a TMA tile load drops into shared memory)
    // The first column of a TMA box must be 16-byte aligned, so a box starts at the even
    // column below the window and shD / shX (0 or 1) is the window's offset inside it.
    int cpad, kst, xoff, shD, shX;
};

// padded column count of a staged window (TMA: inner box extent must be a multiple of 16 B)
BS2E_HD int site_cpad(const Geom& g) { return 2 * g.w + 2; }  // 2w+1 columns + 1 for the alignment shift
BS2E_HD int site_kst(const Geom& g) { return (2 * g.w + 1) * site_cpad(g); }
// doubles per staged window, rounded so that the second window stays 128-byte aligned
BS2E_HD int site_win_doubles(const Geom& g) { return (g.K1 * site_kst(g) + 15) & ~15; }

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
    s.cpad = site_cpad(g);
    s.kst = site_kst(g);
    s.xoff = site_win_doubles(g);
    s.shD = pair_index(g, nb, s.dDlo) & 1;
    s.shX = pair_index(g, na, s.dXlo) & 1;
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
BS2E_HD int site_slotD(const Site& s, int nc, int nd) { return (nc - s.cDlo) * s.cpad + (nd - s.dDlo) + s.shD; }
BS2E_HD int site_slotX(const Site& s, int nc, int nd) { return s.xoff + (nc - s.cXlo) * s.cpad + (nd - s.dXlo) + s.shX; }

// first row / column of R (pair indices) of the staged windows: window D holds
// R^k[pair(n_a,n_c)][pair(n_b,n_d)], window X holds R^k[pair(n_b,n_c)][pair(n_a,n_d)]
// (the exchange integral read through R^k(ab;dc) = R^k(ba;cd))
BS2E_HD int site_rowD(const Geom& g, const Site& s) { return pair_index(g, s.na, s.cDlo); }
BS2E_HD int site_colD(const Geom& g, const Site& s) { return pair_index(g, s.nb, s.dDlo) & ~1; }
BS2E_HD int site_rowX(const Geom& g, const Site& s) { return pair_index(g, s.nb, s.cXlo); }
BS2E_HD int site_colX(const Geom& g, const Site& s) { return pair_index(g, s.na, s.dXlo) & ~1; }

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

// ---- candidate lists of a site ------------------------------------------------
// The columns (n_c,n_d) a row of the site can couple to, in CSR order (n_c, then
// n_d ascending), independent of the column block: list D = direct window, list X =
// exchange window, list DX = union.  One entry: a = q | n_d << 8,
// b = slotD | slotX << 16 (0xffff: the column is not in that window).
struct alignas(8) Cand { unsigned a, b; };
constexpr unsigned kNoSlot = 0xffffu;

BS2E_HD Cand site_cand(const Site& s, int q, int nc, int nd)
{
    Cand c;
    c.a = (unsigned)q | ((unsigned)nd << 8);
    unsigned sd = kNoSlot, sx = kNoSlot;
    if (site_nc_inD(s, nc) && nd >= s.dDlo && nd <= s.dDhi) sd = (unsigned)site_slotD(s, nc, nd);
    if (site_nc_inX(s, nc) && nd >= s.dXlo && nd <= s.dXhi) sx = (unsigned)site_slotX(s, nc, nd);
    c.b = sd | (sx << 16);
    return c;
}
BS2E_HD int cand_q(const Cand& c) { return (int)(c.a & 0xffu); }
BS2E_HD int cand_nd(const Cand& c) { return (int)(c.a >> 8); }
BS2E_HD unsigned cand_slotD(const Cand& c) { return c.b & 0xffffu; }
BS2E_HD unsigned cand_slotX(const Cand& c) { return c.b >> 16; }

// entry t of list D / list X (row-major over the window = slot order)
BS2E_HD Cand site_cand_D(const Site& s, int t)
{
    const int nc = s.cDlo + t / s.dw, nd = s.dDlo + t % s.dw;
    return site_cand(s, union_pos(s, nc), nc, nd);
}
BS2E_HD Cand site_cand_X(const Site& s, int t)
{
    const int nc = s.cXlo + t / s.xw, nd = s.dXlo + t % s.xw;
    return site_cand(s, union_pos(s, nc), nc, nd);
}
// number of entries of n_c slot q in list DX, and its idx-th entry
BS2E_HD int site_cand_DX_count(const Site& s, int q) { return union2_count(site_nd_union(s, site_nc(s, q))); }
BS2E_HD Cand site_cand_DX(const Site& s, int q, int idx)
{
    const int nc = site_nc(s, q);
    return site_cand(s, q, nc, union2_at(site_nd_union(s, nc), idx));
}

// ---- one (row, column block) pair ---------------------------------------------
// Everything that is uniform over the pair.  Pointers into shared memory in the
// kernel, into plain arrays in the CPU checker.
struct PairCtx {
    const SiteEntry* Tb;        // [nnc] clipped windows of the column block
    const unsigned short* hpq;  // [nnc+1] prefix of stored H entries over the n_c slots
    const unsigned short* spq;  // [nnc]   same for S (diagonal pair only)
    const double* Rv;           // staged R^k values, Rv[k*kst + slot]
    const double* wa_d;         // packed direct factors, k = pk.dlo + 2i
    const double* wa_x;         // packed exchange factors
    int kst;
    PairK pk;
    int bj;
    bool diag;                  // column block == row block: one-body terms and S
    bool dirany, exany, samex, cut;
    long long hbase, sbase;     // first H entry of the pair / first S entry of the row
};

// Which list a pair walks: one window when the mode stores one window or when the
// other window is clipped away completely for this site and column block.
BS2E_HD int pair_window(int mode, int totD, int totX)
{
    if (mode == kModeD || mode == kModeX) return mode;
    if (totX == 0) return kModeD;
    if (totD == 0) return kModeX;
    return kModeDX;
}

// One candidate of the pair's list.  LIST = kModeD / kModeX: every stored entry lies
// in that window; LIST = kModeDX: union of both windows.
template <int LIST>
BS2E_HD void site_item(const Geom& g, const Plan& pl, const OneBody& ob, const Site& s,
                       const RowInfo& r, const PairCtx& pc, Cand cd, long long* Hidx, double* Hdat,
                       long long* Sidx, double* Sdat)
{
    const int q = cand_q(cd), nd = cand_nd(cd);
    SiteEntry e = pc.Tb[q];
    if (pc.cut) e = entry_cut(e, s, site_nc(s, q));
    const bool sup = nd >= (int)e.dlo && nd <= (int)e.dhi;
    const bool sup_ex = nd >= (int)e.xlo && nd <= (int)e.xhi;
    int rank;
    if (LIST == kModeD) {
        if (!sup) return;
        rank = nd - (int)e.dlo;
    } else if (LIST == kModeX) {
        if (!sup_ex) return;
        rank = nd - (int)e.xlo;
    } else {
        if (!sup && !sup_ex) return;
        rank = union_below(true, e.dlo, e.dhi, true, e.xlo, e.xhi, nd);
    }
    const int stride = 2 * pc.kst;
    double re = 0.0, im = 0.0;
    const bool allowed = (sup && pc.dirany) || (sup_ex && pc.exany);
    if (allowed) {
        double res = 0.0;
        if (sup) res += k_dot(pc.Rv + pc.pk.dlo * pc.kst + cand_slotD(cd), stride, pc.wa_d, pc.pk.nkd);
        if (sup_ex) res += k_dot(pc.Rv + pc.pk.xlo * pc.kst + cand_slotX(cd), stride, pc.wa_x, pc.pk.nkx);
        re = res;
    }
    const long long j = (long long)e.jbase + nd;
    if (pc.diag) {
        const bool storeS = sup || (sup_ex && pc.samex);
        if (storeS) {
            const BlockDesc bc = pl.blk[pc.bj];
            Cplx h, sv;
            one_body_terms(g, pl, ob, r, true, pc.samex, bc.l1, bc.l2, site_nc(s, q), nd, &h, &sv);
            re += h.re;
            im += h.im;
            const long long pos =
                pc.sbase + pc.spq[q] + union_below(true, e.dlo, e.dhi, pc.samex, e.xlo, e.xhi, nd);
            Sidx[pos] = j;
            Sdat[2 * pos] = sv.re;
            Sdat[2 * pos + 1] = sv.im;
        }
    }
    const long long pos = pc.hbase + pc.hpq[q] + rank;
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
