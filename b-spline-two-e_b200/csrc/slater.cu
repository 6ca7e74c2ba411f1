// slater.cu -- stage A: primitive cell integrals of the Slater R^k expansion.
//
// Stands in for setup_Slater_integrals (src/mat_els/mat_els.f90:172-292) and
// its kernels compute_Slater_off_diag (:392-439) / compute_Slater_diag
// (:441-491).  The reference re-evaluates every B-spline at every quadrature
// point for every index tuple and every k; here the B-spline values on the
// Gauss-Legendre nodes of a cell are tabulated once in shared memory and every
// integral of that cell is formed from the table.
//
//   cell_moments_kernel : r_k, r_m_k                (one CTA per cell)
//   pair_prefix_kernel  : running sums over cells used by stage B
//   diag_cells_kernel   : r_d_k, the triangular same-cell double integral
//                         (one CTA per cell and k-slice)
#include <algorithm>
#include <cstring>

#include "ctx.h"
#include "slater_core.h"

namespace bs2e {

__global__ void __launch_bounds__(128)
cell_moments_kernel(Geom g, double* __restrict__ mom_rk, double* __restrict__ mom_rmk)
{
    extern __shared__ double sm[];
    const int v = blockIdx.x + 1;
    const MomSmem m = mom_smem_carve(g, sm);
    mom_phase_tables(g, v, m, threadIdx.x, blockDim.x);
    __syncthreads();
    mom_phase_integrate(g, v, m, threadIdx.x, blockDim.x, mom_rk, mom_rmk);
}

__global__ void pair_prefix_kernel(Geom g, const double* __restrict__ mom_rk,
                                   const double* __restrict__ mom_rmk,
                                   double* __restrict__ pre, double* __restrict__ sufx, RkRow* __restrict__ rkrow)
{
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)g.K1 * g.P) return;
    pair_prefix_item(g, idx, mom_rk, mom_rmk, pre, sufx, rkrow);
}

__global__ void __launch_bounds__(256)
diag_cells_kernel(Geom g, double* __restrict__ rd, int v0)
{
    extern __shared__ double sm[];
    const int v = blockIdx.x + v0;
    const DiagSmem d = diag_smem_carve(g, sm);
    diag_phase_tables(g, v, d, threadIdx.x, blockDim.x);
    __syncthreads();
    for (int k = blockIdx.y; k < g.K1; k += gridDim.y) {
        diag_phase_powers(g, k, d, threadIdx.x, blockDim.x);
        __syncthreads();
        diag_phase_inner(g, k, d, threadIdx.x, blockDim.x);
        __syncthreads();
        diag_phase_outer(g, v, k, d, threadIdx.x, blockDim.x, rd);
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------
void run_slater_cells(bs2e_ctx* c)
{
    const Geom& g = c->dg;
    const size_t nmom = (size_t)g.K1 * g.P * g.ks;
    const size_t npre = (size_t)g.K1 * g.P * (g.ks + 1);
    const size_t ks2 = (size_t)g.ks * g.ks;
    const size_t nrd = (size_t)g.cells * g.K1 * ks2 * ks2;
    if (!c->d_mom_rk) {
        c->d_mom_rk = dev_alloc<double>(nmom);
        c->d_mom_rmk = dev_alloc<double>(nmom);
        c->d_pre = dev_alloc<double>(npre);
        c->d_sufx = dev_alloc<double>(npre);
        c->d_rd = dev_alloc<double>(nrd);
        c->d_rkrow = dev_alloc<RkRow>((size_t)g.K1 * g.P);
    }
    BS2E_CUDA(cudaMemsetAsync(c->d_mom_rk, 0, sizeof(double) * nmom, c->stream));
    BS2E_CUDA(cudaMemsetAsync(c->d_mom_rmk, 0, sizeof(double) * nmom, c->stream));

    {
        const size_t smem = sizeof(double) * mom_smem_doubles(g);
        if (smem > 200 * 1024) throw Error("cell_moments: k_GL*(k+2*(max_k+1)) too large for shared memory");
        BS2E_CUDA(cudaFuncSetAttribute(cell_moments_kernel,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        cell_moments_kernel<<<g.cells, 128, smem, c->stream>>>(g, c->d_mom_rk, c->d_mom_rmk);
        BS2E_LAUNCHED();
    }
    {
        const size_t n = (size_t)g.K1 * g.P;
        pair_prefix_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(
            g, c->d_mom_rk, c->d_mom_rmk, c->d_pre, c->d_sufx, c->d_rkrow);
        BS2E_LAUNCHED();
    }
    {
        const size_t smem = sizeof(double) * diag_smem_doubles(g);
        if (smem > 200 * 1024)
            throw Error("diag_cells: k_GL^2*k B-spline table does not fit in shared memory");
        BS2E_CUDA(cudaFuncSetAttribute(diag_cells_kernel,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        // same-cell integrals of the cells the rows of the slice live on (bs2e_rk_rows; all cells by default):
        // pairs (a, c), a in [a_lo, a_hi], |a - c| <= w  ->  cells a_lo-ks+2 .. a_hi+1 (pair_lo_cell / pair_hi_cell)
        const int v0 = std::max(1, c->slice_lo - g.ks + 2), v1 = std::min(g.cells, c->slice_hi + 1);
        const int ncell = std::max(1, v1 - v0 + 1);
        int ksplit = (16 * 148 + ncell - 1) / ncell;
        ksplit = std::max(1, std::min(ksplit, g.K1));
        diag_cells_kernel<<<dim3(ncell, ksplit), 256, smem, c->stream>>>(g, c->d_rd, v0);
        BS2E_LAUNCHED();
    }
    c->have_cells = true;
    c->cells_lo = c->slice_lo;
    c->cells_hi = c->slice_hi;
}

// ---------------------------------------------------------------------------
// getters in the reference's entry order
// ---------------------------------------------------------------------------
void fetch_r_k(bs2e_ctx* c, double* r_k, double* r_m_k, int64_t* iv, int64_t* ia, int64_t* ja)
{
    if (!c->have_cells) throw Error("bs2e_get_r_k: call bs2e_slater_cells first");
    const Geom& g = c->hg;
    const size_t nmom = (size_t)g.K1 * g.P * g.ks;
    std::vector<double> hrk, hrmk;
    if (r_k) {
        hrk.resize(nmom);
        BS2E_CUDA(cudaMemcpyAsync(hrk.data(), c->d_mom_rk, sizeof(double) * nmom,
                                  cudaMemcpyDeviceToHost, c->stream));
    }
    if (r_m_k) {
        hrmk.resize(nmom);
        BS2E_CUDA(cudaMemcpyAsync(hrmk.data(), c->d_mom_rmk, sizeof(double) * nmom,
                                  cudaMemcpyDeviceToHost, c->stream));
    }
    BS2E_CUDA(cudaStreamSynchronize(c->stream));
    // entry order of setup_Slater_off_diag (mat_els.f90:199-225): j_b, i_b, i_r
    const long long nnz = c->nnz_4d;
    long long ptr = 0;
    for (int j_b = 1; j_b <= g.nb; ++j_b)
        for (int i_b = imax(1, j_b - g.w); i_b <= imin(g.nb, j_b + g.w); ++i_b) {
            const int lo = pair_lo_cell(g, i_b, j_b), hi = pair_hi_cell(g, i_b, j_b);
            const int p = pair_index(g, i_b, j_b);
            for (int i_r = lo; i_r <= hi; ++i_r) {
                for (int k = 0; k < g.K1; ++k) {
                    const size_t o = ((size_t)k * g.P + p) * g.ks + (i_r - lo);
                    if (r_k) r_k[ptr + nnz * k] = hrk[o];
                    if (r_m_k) r_m_k[ptr + nnz * k] = hrmk[o];
                }
                if (iv) iv[ptr] = i_r;
                if (ia) ia[ptr] = i_b;
                if (ja) ja[ptr] = j_b;
                ++ptr;
            }
        }
    if (ptr != nnz) throw Error("bs2e_get_r_k: internal entry count mismatch");
}

void fetch_r_d_k(bs2e_ctx* c, double* r_d_k, int64_t* iv, int64_t* ia, int64_t* ja, int64_t* ipa,
                 int64_t* jpa)
{
    if (!c->have_cells) throw Error("bs2e_get_r_d_k: call bs2e_slater_cells first");
    if (c->cells_lo != 1 || c->cells_hi != c->hg.nb) throw Error("bs2e_get_r_d_k: stage A ran for a row slice only (bs2e_rk_rows)");
    const Geom& g = c->hg;
    const size_t ks2 = (size_t)g.ks * g.ks;
    const size_t nrd = (size_t)g.cells * g.K1 * ks2 * ks2;
    std::vector<double> h;
    if (r_d_k) {
        h.resize(nrd);
        BS2E_CUDA(cudaMemcpyAsync(h.data(), c->d_rd, sizeof(double) * nrd, cudaMemcpyDeviceToHost,
                                  c->stream));
        BS2E_CUDA(cudaStreamSynchronize(c->stream));
    }
    // entry order of setup_Slater_diag (mat_els.f90:247-288): j_b_p, j_b, i_b_p, i_b, i_r
    const long long nnz = c->nnz_6d;
    const int w = g.w, nb = g.nb, ks = g.ks;
    long long ptr = 0;
    for (int j_b_p = 1; j_b_p <= nb; ++j_b_p)
        for (int j_b = imax(1, j_b_p - w); j_b <= imin(nb, j_b_p + w); ++j_b) {
            const int jmin = imin(j_b, j_b_p), jmax = imax(j_b, j_b_p);
            // a common cell needs every spline within ks-1 of every other
            for (int i_b_p = imax(1, jmax - w); i_b_p <= imin(nb, jmin + w); ++i_b_p)
                for (int i_b = imax(1, imax(i_b_p, jmax) - w); i_b <= imin(nb, imin(i_b_p, jmin) + w);
                     ++i_b) {
                    const int fmin = imin(imin(i_b, i_b_p), jmin) + 1;
                    const int fmax = imax(imax(i_b, i_b_p), jmax) + 1;
                    const int lo = imax(1, fmax - ks + 1), hi = imin(g.cells, fmin);
                    for (int i_r = lo; i_r <= hi; ++i_r) {
                        if (r_d_k) {
                            const size_t li = (size_t)(i_b + 1 - i_r) * ks + (i_b_p + 1 - i_r);
                            const size_t lj = (size_t)(j_b + 1 - i_r) * ks + (j_b_p + 1 - i_r);
                            for (int k = 0; k < g.K1; ++k)
                                r_d_k[ptr + nnz * k] =
                                    h[((size_t)(i_r - 1) * g.K1 + k) * ks2 * ks2 + li * ks2 + lj];
                        }
                        if (iv) iv[ptr] = i_r;
                        if (ia) ia[ptr] = i_b;
                        if (ipa) ipa[ptr] = i_b_p;
                        if (ja) ja[ptr] = j_b;
                        if (jpa) jpa[ptr] = j_b_p;
                        ++ptr;
                    }
                }
        }
    if (ptr != nnz) throw Error("bs2e_get_r_d_k: internal entry count mismatch");
}

}  // namespace bs2e
