// slater_core.h -- thread-level bodies of the stage-A kernels (slater.cu).
// Each phase is a __host__ __device__ function of (thread id, thread count)
// so that the CUDA kernels and the CPU logic checker (tests/hostcheck) execute
// the same statements; a phase boundary is a __syncthreads() in the kernel.
#pragma once
#include "core.h"

namespace bs2e {

// ---- cell moments (mat_els.f90:392-439) ------------------------------------
//   r_k[p,v,k]   = scale * sum_q w_q B_i(r_q) B_j(r_q) r_q^k
//   r_m_k[p,v,k] = scale * sum_q w_q B_i(r_q) B_j(r_q) / r_q^(k+1)
// stored as mom[k][p][slot], slot = v - first cell of the pair
struct MomSmem {
    double* B;   // [kgl][ks]   B-spline table of this cell
    double* rp;  // [kgl][K1]   r_q^k
    double* rq;  // [kgl][K1]   r_q^(k+1)
};
BS2E_HD size_t mom_smem_doubles(const Geom& g) { return (size_t)g.kgl * g.ks + 2 * (size_t)g.kgl * g.K1; }
BS2E_HD MomSmem mom_smem_carve(const Geom& g, double* sm)
{
    MomSmem m;
    m.B = sm;
    m.rp = m.B + g.kgl * g.ks;
    m.rq = m.rp + g.kgl * g.K1;
    return m;
}

BS2E_HD void mom_phase_tables(const Geom& g, int v, const MomSmem& m, int tid, int nthr)
{
    const int ks = g.ks, K1 = g.K1;
    const double a = g.bp[v - 1], b = g.bp[v];
    const double scale = 0.5 * (b - a), translate = 0.5 * (b + a);
    for (int q = tid; q < g.kgl; q += nthr) {
        const double r = scale * g.glx[q] + translate;
        double vals[kMaxOrder];
        bspline_values(g.t, ks, v, r, vals);
        for (int s = 0; s < ks; ++s) m.B[q * ks + s] = vals[s];
        for (int k = 0; k < K1; ++k) {
            m.rp[q * K1 + k] = powi(r, k);
            m.rq[q * K1 + k] = powi(r, k + 1);
        }
    }
}

BS2E_HD void mom_phase_integrate(const Geom& g, int v, const MomSmem& m, int tid, int nthr,
                                 double* mom_rk, double* mom_rmk)
{
    const int ks = g.ks, K1 = g.K1, ks2 = ks * ks;
    const double scale = 0.5 * (g.bp[v] - g.bp[v - 1]);
    for (int item = tid; item < ks2 * K1; item += nthr) {
        const int k = item / ks2, ss = item % ks2;
        const int s = ss / ks, s2 = ss % ks;
        const int ia = v + s - 1, ic = v + s2 - 1;  // b-indices of the two splines
        if (ia < 1 || ia > g.nb || ic < 1 || ic > g.nb) continue;
        double acc = 0.0, accm = 0.0;
        for (int q = 0; q < g.kgl; ++q) {
            const double wb = g.glw[q] * m.B[q * ks + s] * m.B[q * ks + s2];
            acc += wb * m.rp[q * K1 + k];
            accm += wb / m.rq[q * K1 + k];
        }
        const int p = pair_index(g, ia, ic);
        const int slot = v - pair_lo_cell(g, ia, ic);
        const size_t o = ((size_t)k * g.P + p) * ks + slot;
        mom_rk[o] = scale * acc;
        mom_rmk[o] = scale * accm;
    }
}

// pre[t]  = sum_{slot <  t} r_k ,  t = 0..ks   (pre[ks]  = total)
// sufx[t] = sum_{slot >= t} r_m_k, t = 0..ks   (sufx[0] = total)
// rkrow[idx]: the pair's cell range and the two totals, packed for the streaming role of stage B (core.h: RkRow)
BS2E_HD void pair_prefix_item(const Geom& g, size_t idx, const double* mom_rk,
                              const double* mom_rmk, double* pre, double* sufx, RkRow* rkrow)
{
    const int ks = g.ks;
    const double* rk = mom_rk + idx * ks;
    const double* rmk = mom_rmk + idx * ks;
    double* pr = pre + idx * (ks + 1);
    double* sf = sufx + idx * (ks + 1);
    double run = 0.0;
    for (int t = 0; t < ks; ++t) { pr[t] = run; run += rk[t]; }
    pr[ks] = run;
    run = 0.0;
    sf[ks] = 0.0;
    for (int t = ks - 1; t >= 0; --t) { run += rmk[t]; sf[t] = run; }
    const PairAC q = g.pair[idx % (size_t)g.P];
    RkRow r;
    r.lo = pair_lo_cell(g, q.a, q.c);
    r.hi = pair_hi_cell(g, q.a, q.c);
    r.trk = pr[ks];
    r.trmk = sf[0];
    r.pad = 0.0;
    rkrow[idx] = r;
}

// ---- same-cell double integral (mat_els.f90:441-491) -----------------------
// r_d_k[(i,i'),(j,j'),v,k] = scale * sum_q w_q B_i B_i'(r_q) / r_q^(k+1) * I_q
//   I_q = scale_q * sum_p w_p B_j B_j'(r_qp) r_qp^k ,  r_qp in [a, r_q]
// (nested Gauss-Legendre of mat_els.f90:460-483).  Per (cell,k) the outer sum
// is a (ks^2 x kgl) x (kgl x ks^2) product of the tables W and I.
// Output rd[v][k][i*ks+i'][j*ks+j'] with local slots i = full index - v.
struct DiagSmem {
    double* Bo;   // [kgl][ks]        outer table
    double* Bi;   // [kgl][kgl][ks]   inner table
    double* ro;   // [kgl]            outer radii
    double* sj;   // [kgl]            inner Jacobians
    double* rin;  // [kgl][kgl]       inner radii
    double* pw;   // [kgl][kgl]       w_p * r_qp^k
    double* Im;   // [kgl][ks2]       I_q[j,j']
    double* Wm;   // [kgl][ks2]       W_q[i,i']
};
BS2E_HD size_t diag_smem_doubles(const Geom& g)
{
    const size_t kgl = g.kgl, ks2 = (size_t)g.ks * g.ks;
    return kgl * g.ks + kgl * kgl * g.ks + 2 * kgl + 2 * kgl * kgl + 2 * kgl * ks2;
}
BS2E_HD DiagSmem diag_smem_carve(const Geom& g, double* sm)
{
    const int kgl = g.kgl, ks = g.ks, ks2 = ks * ks;
    DiagSmem d;
    d.Bo = sm;
    d.Bi = d.Bo + kgl * ks;
    d.ro = d.Bi + kgl * kgl * ks;
    d.sj = d.ro + kgl;
    d.rin = d.sj + kgl;
    d.pw = d.rin + kgl * kgl;
    d.Im = d.pw + kgl * kgl;
    d.Wm = d.Im + kgl * ks2;
    return d;
}

BS2E_HD void diag_phase_tables(const Geom& g, int v, const DiagSmem& d, int tid, int nthr)
{
    const int ks = g.ks, kgl = g.kgl;
    const double a = g.bp[v - 1], b = g.bp[v];
    const double scale = 0.5 * (b - a), translate = 0.5 * (b + a);
    for (int q = tid; q < kgl; q += nthr) {
        const double r = scale * g.glx[q] + translate;
        double vals[kMaxOrder];
        bspline_values(g.t, ks, v, r, vals);
        for (int s = 0; s < ks; ++s) d.Bo[q * ks + s] = vals[s];
        d.ro[q] = r;
        d.sj[q] = 0.5 * (r - a);
    }
    for (int qp = tid; qp < kgl * kgl; qp += nthr) {
        const int q = qp / kgl, p = qp % kgl;
        const double r = scale * g.glx[q] + translate;
        const double r_j = 0.5 * (r - a) * g.glx[p] + 0.5 * (r + a);
        double vals[kMaxOrder];
        bspline_values(g.t, ks, v, r_j, vals);
        for (int s = 0; s < ks; ++s) d.Bi[(size_t)qp * ks + s] = vals[s];
        d.rin[qp] = r_j;
    }
}

BS2E_HD void diag_phase_powers(const Geom& g, int k, const DiagSmem& d, int tid, int nthr)
{
    const int kgl = g.kgl;
    for (int qp = tid; qp < kgl * kgl; qp += nthr) d.pw[qp] = g.glw[qp % kgl] * powi(d.rin[qp], k);
}

BS2E_HD void diag_phase_inner(const Geom& g, int k, const DiagSmem& d, int tid, int nthr)
{
    const int ks = g.ks, kgl = g.kgl, ks2 = ks * ks;
    for (int item = tid; item < kgl * ks2; item += nthr) {
        const int q = item / ks2, jj = item % ks2;
        const int j = jj / ks, j2 = jj % ks;
        const double* bi = d.Bi + (size_t)q * kgl * ks;
        double acc = 0.0;
        for (int p = 0; p < kgl; ++p) acc += d.pw[q * kgl + p] * bi[p * ks + j] * bi[p * ks + j2];
        d.Im[item] = d.sj[q] * acc;
        d.Wm[item] = g.glw[q] * d.Bo[q * ks + j] * d.Bo[q * ks + j2] / powi(d.ro[q], k + 1);
    }
}

BS2E_HD void diag_phase_outer(const Geom& g, int v, int k, const DiagSmem& d, int tid, int nthr,
                              double* rd)
{
    const int ks = g.ks, kgl = g.kgl, ks2 = ks * ks;
    const double scale = 0.5 * (g.bp[v] - g.bp[v - 1]);
    const int tpd = (ks2 + 3) / 4;  // 4x4 register tiles per dimension
    double* out = rd + ((size_t)(v - 1) * g.K1 + k) * (size_t)ks2 * ks2;
    for (int tile = tid; tile < tpd * tpd; tile += nthr) {
        const int i0 = (tile / tpd) * 4, j0 = (tile % tpd) * 4;
        double acc[4][4];
#pragma unroll
        for (int x = 0; x < 4; ++x)
#pragma unroll
            for (int y = 0; y < 4; ++y) acc[x][y] = 0.0;
        for (int q = 0; q < kgl; ++q) {
            double wv[4], iv[4];
#pragma unroll
            for (int x = 0; x < 4; ++x) {
                wv[x] = (i0 + x < ks2) ? d.Wm[q * ks2 + i0 + x] : 0.0;
                iv[x] = (j0 + x < ks2) ? d.Im[q * ks2 + j0 + x] : 0.0;
            }
#pragma unroll
            for (int x = 0; x < 4; ++x)
#pragma unroll
                for (int y = 0; y < 4; ++y) acc[x][y] += wv[x] * iv[y];
        }
#pragma unroll
        for (int x = 0; x < 4; ++x)
#pragma unroll
            for (int y = 0; y < 4; ++y)
                if (i0 + x < ks2 && j0 + y < ks2)
                    out[(size_t)(i0 + x) * ks2 + j0 + y] = scale * acc[x][y];
    }
}

}  // namespace bs2e
