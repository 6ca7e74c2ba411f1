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
#include <algorithm>

#include "ctx.h"

namespace bs2e {

struct RkGrid {
    int row_lo, row_hi;   // p1 rows to build (bs2e_rk_rows; the whole plane by default)
    int band_x, band_y;   // band tiles: column tiles per row tile, row tiles
    int stream_x;         // streaming tiles per row tile
    int staged;           // band role keeps the per-pair vectors of its tile in shared memory
};

__global__ void __launch_bounds__(kRkThreads)
rk_build_kernel(const __grid_constant__ Geom g, const __grid_constant__ CellData cd, const RkRow* __restrict__ rkrow,
                double* __restrict__ R, const __grid_constant__ RkGrid q)
{
    __shared__ RkRow rows[kRkRowsS];
    extern __shared__ double stage[];   // band role: per-pair vectors of the tile's rows and columns (q.staged)
    const int k = blockIdx.y, tx = threadIdx.x;
    const RkRow* rk = rkrow + (size_t)k * g.P;
    // the band tiles are spread evenly over the launch order, so that every SM holds a mix of the latency-bound
    // band CTAs and the bandwidth-bound streaming CTAs at any time
    const long long nband = (long long)q.band_x * q.band_y, ntot = gridDim.x;
    const long long before = (long long)blockIdx.x * nband / ntot, upto = ((long long)blockIdx.x + 1) * nband / ntot;
    int b;
    if (upto > before) {
        // ---- band role ----
        b = (int)before;
        const int ty = b / q.band_x, bt = b - ty * q.band_x;
        const int r0 = q.row_lo + ty * kRkRowsB, nrows = imin(kRkRowsB, q.row_hi - r0);
        int c0, c1;
        rk_band_columns(g, r0, r0 + nrows, &c0, &c1);
        if (c0 + bt * kRkThreads >= c1) return;
        if (tx < nrows) rows[tx] = rk[r0 + tx];
        const int p2 = c0 + bt * kRkThreads + tx;
        if (!q.staged) {
            __syncthreads();
            if (p2 < c1) rk_band_column(g, cd, R, rows, r0, nrows, k, p2, rk[p2]);
            return;
        }
        // the vectors of the tile's pairs are contiguous in the stage-A tables: two linear copies for the rows,
        // two for the columns (column vectors at an odd stride: thread t reads row t of the table)
        const int ks = g.ks, cstr = 2 * ks + 1;
        double* pre1s = stage;
        double* suf1s = pre1s + kRkRowsB * (ks + 1);
        double* colm = suf1s + kRkRowsB * (ks + 1);
        {
            const size_t rbase = ((size_t)k * g.P + r0) * (ks + 1);
            for (int i = tx; i < nrows * (ks + 1); i += kRkThreads) {
                pre1s[i] = cd.pre[rbase + i];
                suf1s[i] = cd.sufx[rbase + i];
            }
            const int pc0 = c0 + bt * kRkThreads, ncols = imin(kRkThreads, c1 - pc0);
            const size_t cbase = ((size_t)k * g.P + pc0) * ks;
            for (int i = tx; i < ncols * ks; i += kRkThreads) {
                const int col = i / ks, sl = i - col * ks;
                colm[col * cstr + sl] = cd.mom_rk[cbase + i];
                colm[col * cstr + ks + sl] = cd.mom_rmk[cbase + i];
            }
        }
        __syncthreads();
        if (p2 < c1)
            rk_band_column_staged(g, cd, R, rows, r0, nrows, k, p2, rk[p2], pre1s, suf1s, colm + tx * cstr,
                                  colm + tx * cstr + ks);
        return;
    }
    // ---- streaming role ----
    b = (int)(blockIdx.x - before);
    const int ty = b / q.stream_x, bx = b - ty * q.stream_x;
    const int r0 = q.row_lo + ty * kRkRowsS, nrows = imin(kRkRowsS, q.row_hi - r0);
    if (tx < nrows) rows[tx] = rk[r0 + tx];
    const int p2 = (bx * kRkThreads + tx) * 2;
    RkRow c0 = rk_row_empty(), c1 = rk_row_empty();
    if (p2 < g.P) c0 = rk[p2];
    if (p2 + 1 < g.P) c1 = rk[p2 + 1];
    __syncthreads();
    if (p2 < g.ldP) rk_stream_columns(g, R, rows, r0, nrows, k, p2, c0, c1);
}

// widest band of any row tile, in column tiles (host geometry)
static int rk_band_tiles(const Geom& hg, int row_lo, int row_hi)
{
    int widest = 1;
    for (int r0 = row_lo; r0 < row_hi; r0 += kRkRowsB) {
        int c0, c1;
        rk_band_columns(hg, r0, std::min(r0 + kRkRowsB, row_hi), &c0, &c1);
        widest = std::max(widest, (c1 - c0 + kRkThreads - 1) / kRkThreads);
    }
    return widest;
}

void run_rk_build(bs2e_ctx* c)
{
    if (!c->have_cells) throw Error("bs2e_rk_build: call bs2e_slater_cells first");
    if (c->cells_lo > c->slice_lo || c->cells_hi < c->slice_hi)
        throw Error("bs2e_rk_build: bs2e_slater_cells ran for a narrower row slice (bs2e_rk_rows)");
    const Geom& g = c->dg;
    if (!c->d_R) {
        c->d_R = dev_alloc<double>((size_t)g.K1 * g.P * g.ldP);
    }
    // rows of the slice (bs2e_rk_rows; the whole tensor by default): pairs (a, .) with a in [a_lo, a_hi]
    RkGrid q;
    q.row_lo = c->hg.rowoff[c->slice_lo];
    q.row_hi = c->hg.rowoff[c->slice_hi + 1];
    const int nrows = q.row_hi - q.row_lo;
    q.band_y = (nrows + kRkRowsB - 1) / kRkRowsB;
    q.band_x = rk_band_tiles(c->hg, q.row_lo, q.row_hi);
    q.stream_x = (g.ldP / 2 + kRkThreads - 1) / kRkThreads;
    const int stream_y = (nrows + kRkRowsS - 1) / kRkRowsS;
    // shared memory of the band role (every CTA of the launch reserves it): staged while several CTAs fit an SM
    size_t smem = sizeof(double) * ((size_t)2 * kRkRowsB * (g.ks + 1) + (size_t)kRkThreads * (2 * g.ks + 1));
    q.staged = smem <= 44 * 1024;   // spline order <= 9 (five CTAs per SM at order 8)
    if (!q.staged) smem = 0;
    dim3 grid((unsigned)(q.band_x * q.band_y + q.stream_x * stream_y), g.K1);
    rk_build_kernel<<<grid, kRkThreads, smem, c->stream>>>(g, c->cell_data(), c->d_rkrow, c->d_R, q);
    BS2E_LAUNCHED();
    c->have_R = true;
    c->R_lo = c->slice_lo;
    c->R_hi = c->slice_hi;
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
    if (c->R_lo != 1 || c->R_hi != c->hg.nb) throw Error("bs2e_rk_get: only a row slice of the tensor was built (bs2e_rk_rows)");
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
    if (c->R_lo != 1 || c->R_hi != c->hg.nb) throw Error("bs2e_rk_plane: only a row slice of the tensor was built (bs2e_rk_rows)");
    const Geom& g = c->dg;
    if (k < 0 || k >= g.K1) throw Error("bs2e_rk_plane: k out of range");
    BS2E_CUDA(cudaMemcpy2DAsync(out, sizeof(double) * g.P,
                                c->d_R + (size_t)k * g.P * g.ldP, sizeof(double) * g.ldP,
                                sizeof(double) * g.P, g.P, cudaMemcpyDeviceToHost, c->stream));
    BS2E_CUDA(cudaStreamSynchronize(c->stream));
}

}  // namespace bs2e
