// rk.cu -- stage B: materialise the radial Slater integrals R^k(ab;cd).
//
// Stands in for compute_R_K_map (src/tools/sparse_array_tools.f90:452-493) and
// the Nd_DOK hash map behind it (:495-555).  The map is replaced by a dense
// tensor R[k][p1][p2] over ordered band pairs p1=(a,c), p2=(b,d) (P x P values
// per multipole, exactly the count_nnz_R_k keys the reference inserts), with
// a padded leading dimension so that rows are 128-byte aligned.
//
// Where the two pairs live on disjoint cell ranges (the vast majority of the
// tensor) the double sum over cells collapses to one product of total
// moments, so the kernel is a streaming write; near the diagonal the sum runs
// over the <= ks common cells using the prefix tables of stage A, and the
// same-cell integrals r_d_k are added.  Bound: HBM write, 8 B per R^k value.
#include "ctx.h"

namespace bs2e {

__global__ void __launch_bounds__(kRkThreads)
rk_build_kernel(Geom g, CellData cd, double* __restrict__ R)
{
    __shared__ RkRow rows[kRkRows];
    __shared__ unsigned short list[kRkRows * kRkThreads * 2];
    __shared__ int count;
    const int bx = blockIdx.x, by = blockIdx.y, k = blockIdx.z, tx = threadIdx.x;
    if (tx < kRkRows) rows[tx] = rk_row_data(g, cd, k, by * kRkRows + tx);
    if (tx == 0) count = 0;
    __syncthreads();
    rk_stream_thread(g, cd, R, rows, bx, by, k, tx,
                     [&](int code) { list[atomicAdd(&count, 1)] = (unsigned short)code; });
    __syncthreads();
    const int n = count;
    for (int i = tx; i < n; i += kRkThreads) rk_general_item(g, cd, R, bx, by, k, list[i]);
}

void run_rk_build(bs2e_ctx* c)
{
    if (!c->have_cells) throw Error("bs2e_rk_build: call bs2e_slater_cells first");
    const Geom& g = c->dg;
    if (!c->d_R) {
        c->d_R = dev_alloc<double>((size_t)g.K1 * g.P * g.ldP);
    }
    dim3 grid((g.ldP / 2 + kRkThreads - 1) / kRkThreads, (g.P + kRkRows - 1) / kRkRows, g.K1);
    rk_build_kernel<<<grid, kRkThreads, 0, c->stream>>>(g, c->cell_data(), c->d_R);
    BS2E_LAUNCHED();
    c->have_R = true;
}

// ---- Nd_DOK%get_val and plane export (host-facing, used by tests) ----------
__global__ void rk_gather_kernel(Geom g, const double* __restrict__ R, long long n_keys,
                                 const long long* __restrict__ keys, double* __restrict__ vals,
                                 int* __restrict__ bad)
{
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_keys * g.K1) return;
    const long long key = idx / g.K1;
    const int k = (int)(idx % g.K1);
    const long long a = keys[4 * key], b = keys[4 * key + 1], cc = keys[4 * key + 2],
                    d = keys[4 * key + 3];
    const bool in = a >= 1 && a <= g.nb && b >= 1 && b <= g.nb && cc >= 1 && cc <= g.nb && d >= 1 &&
                    d <= g.nb && (a - cc <= g.w) && (cc - a <= g.w) && (b - d <= g.w) && (d - b <= g.w);
    if (!in) { *bad = 1; vals[idx] = 0.0; return; }
    const int p1 = pair_index(g, (int)a, (int)cc), p2 = pair_index(g, (int)b, (int)d);
    vals[idx] = R[((size_t)k * g.P + p1) * g.ldP + p2];
}

void fetch_rk_keys(bs2e_ctx* c, long long n_keys, const int64_t* keys, double* vals)
{
    if (!c->have_R) throw Error("bs2e_rk_get: call bs2e_rk_build first");
    if (n_keys <= 0) return;
    const Geom& g = c->dg;
    long long* d_keys = dev_alloc<long long>(4 * n_keys);
    double* d_vals = dev_alloc<double>(n_keys * g.K1);
    int* d_bad = dev_alloc<int>(1);
    int bad = 0;
    try {
        BS2E_CUDA(cudaMemcpyAsync(d_keys, keys, sizeof(long long) * 4 * n_keys,
                                  cudaMemcpyHostToDevice, c->stream));
        BS2E_CUDA(cudaMemsetAsync(d_bad, 0, sizeof(int), c->stream));
        const long long n = n_keys * g.K1;
        rk_gather_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(g, c->d_R, n_keys,
                                                                            d_keys, d_vals, d_bad);
        BS2E_LAUNCHED();
        BS2E_CUDA(cudaMemcpyAsync(vals, d_vals, sizeof(double) * n, cudaMemcpyDeviceToHost,
                                  c->stream));
        BS2E_CUDA(cudaMemcpyAsync(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        BS2E_CUDA(cudaStreamSynchronize(c->stream));
    } catch (...) {
        cudaFree(d_keys); cudaFree(d_vals); cudaFree(d_bad);
        throw;
    }
    cudaFree(d_keys); cudaFree(d_vals); cudaFree(d_bad);
    if (bad) throw Error("bs2e_rk_get: key outside the band structure (no such R^k entry)");
}

void fetch_rk_plane(bs2e_ctx* c, int k, double* out)
{
    if (!c->have_R) throw Error("bs2e_rk_plane: call bs2e_rk_build first");
    const Geom& g = c->dg;
    if (k < 0 || k >= g.K1) throw Error("bs2e_rk_plane: k out of range");
    BS2E_CUDA(cudaMemcpy2DAsync(out, sizeof(double) * g.P,
                                c->d_R + (size_t)k * g.P * g.ldP, sizeof(double) * g.ldP,
                                sizeof(double) * g.P, g.P, cudaMemcpyDeviceToHost, c->stream));
    BS2E_CUDA(cudaStreamSynchronize(c->stream));
}

}  // namespace bs2e
