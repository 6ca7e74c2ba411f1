// block.cu -- stage C: assembly of one (L, parity) symmetry block of the
// coupled two-electron Hamiltonian H and overlap S as CSR matrices.
//
// Stands in for count_nnz (src/mat_els/hamiltonian.f90:348-416) and
// construct_block_tensor (:106-283) with the element formulas of
// src/mat_els/mat_els.f90:552-571,608-633,664-715.  The reference scans all
// n_config^2 configuration pairs twice and evaluates 3j/6j symbols for every
// pair; here
//   * the angular factors are tabulated once per pair of (l1,l2) blocks on
//     the host (exact arithmetic, wigner.cpp) and uploaded,
//   * the band partners of a row are GENERATED from the block structure of
//     the configuration list (no pair scan): block_count_kernel counts them
//     in closed form, an exclusive scan builds index_ptr, and
//     block_fill_kernel (one warp per row) writes indices and values in
//     ascending column order with coalesced 8/16-byte stores.
// Bound: HBM (24 B written per stored element + R^k gathers).
#include <cub/device/device_scan.cuh>

#include <algorithm>
#include <map>
#include <mutex>
#include <set>
#include <unordered_map>

#include "ctx.h"
#include "plan.h"

namespace bs2e {

// ---------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------
__global__ void block_count_kernel(Geom g, Plan pl, long long row_lo, long long nrows,
                                   long long* __restrict__ cntH, long long* __restrict__ cntS)
{
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx > nrows) return;
    long long h = 0, s = 0;
    if (idx < nrows) row_count(g, pl, (int)(row_lo + idx), &h, &s);
    cntH[idx] = h;  // slot nrows holds 0 so that the scan yields the total
    cntS[idx] = s;
}

__global__ void ptr_one_based_kernel(long long n, long long* __restrict__ a, long long* __restrict__ b)
{
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n) return;
    a[idx] += 1;
    b[idx] += 1;
}

constexpr int kFillWarps = 8;

__global__ void __launch_bounds__(kFillWarps * 32)
block_fill_kernel(Geom g, Plan pl, OneBody ob, const double* __restrict__ R, long long row_lo,
                  long long nrows, const long long* __restrict__ Hptr,
                  const long long* __restrict__ Sptr, long long* __restrict__ Hidx,
                  double2* __restrict__ Hdat, long long* __restrict__ Sidx,
                  double2* __restrict__ Sdat)
{
    const long long wrow = (long long)blockIdx.x * kFillWarps + (threadIdx.x >> 5);
    if (wrow >= nrows) return;
    const int lane = threadIdx.x & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    const RowInfo r = row_info(pl, (int)(row_lo + wrow));
    long long hpos = Hptr[wrow] - 1, spos = Sptr[wrow] - 1;

    for_each_chunk(g, pl, r, [&](int bj, int nc, const Segment& s, const Coupling& c, int base, int hi) {
        const int nd = base + lane;
        const bool act = nd <= hi;
        bool storeS = false;
        Element e;
        if (act) {
            const bool sup = nd >= s.dlo && nd <= s.dhi;
            const bool sup_ex = nd >= s.xlo && nd <= s.xhi;
            e = element_value(g, pl, ob, R, r, c, bj, nc, nd, sup, sup_ex);
            storeS = e.storeS;
            Hidx[hpos + lane] = (long long)s.jbase + nd;
            Hdat[hpos + lane] = make_double2(e.H.re, e.H.im);
        }
        hpos += imin(32, hi - base + 1);
        const unsigned m = __ballot_sync(0xffffffffu, storeS);
        if (storeS) {
            const long long pos = spos + __popc(m & lt_mask);
            Sidx[pos] = (long long)s.jbase + nd;
            Sdat[pos] = make_double2(e.S.re, e.S.im);
        }
        spos += __popc(m);
    });
}

__global__ void checksum_kernel(long long n, const long long* __restrict__ idx,
                                const double* __restrict__ dat, unsigned long long* out)
{
    unsigned long long acc = 0;
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < n;
         q += (long long)gridDim.x * blockDim.x) {
        const unsigned long long w = (unsigned long long)(q + 1);
        acc += w * (unsigned long long)idx[q];
        acc += w * (unsigned long long)__double_as_longlong(dat[2 * q]);
        acc += (w << 1) * (unsigned long long)__double_as_longlong(dat[2 * q + 1]);
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(out, acc);
}

// count pass + exclusive scan -> 1-based index_ptr of the planned rows
void block_count_scan(bs2e_block* b, bool read_totals)
{
    bs2e_ctx* c = b->ctx;
    cudaStream_t st = c->stream;
    const long long nrows = b->row_hi - b->row_lo + 1;
    block_count_kernel<<<(unsigned)((nrows + 1 + 127) / 128), 128, 0, st>>>(
        c->dg, b->dplan, b->row_lo, nrows, b->d_cntH, b->d_cntS);
    BS2E_LAUNCHED();
    size_t tmp = b->scan_tmp_bytes;
    BS2E_CUDA(cub::DeviceScan::ExclusiveSum(b->d_scan_tmp, tmp, b->d_cntH, b->d_Hptr, nrows + 1, st));
    g_launches.fetch_add(2);  // DeviceScanInitKernel + DeviceScanKernel
    BS2E_CUDA(cub::DeviceScan::ExclusiveSum(b->d_scan_tmp, tmp, b->d_cntS, b->d_Sptr, nrows + 1, st));
    g_launches.fetch_add(2);
    ptr_one_based_kernel<<<(unsigned)((nrows + 1 + 255) / 256), 256, 0, st>>>(nrows + 1, b->d_Hptr,
                                                                             b->d_Sptr);
    BS2E_LAUNCHED();
    if (read_totals) {
        long long lastH = 0, lastS = 0;
        BS2E_CUDA(cudaMemcpyAsync(&lastH, b->d_Hptr + nrows, sizeof(long long),
                                  cudaMemcpyDeviceToHost, st));
        BS2E_CUDA(cudaMemcpyAsync(&lastS, b->d_Sptr + nrows, sizeof(long long),
                                  cudaMemcpyDeviceToHost, st));
        BS2E_CUDA(cudaStreamSynchronize(st));
        b->nnzH = lastH - 1;
        b->nnzS = lastS - 1;
    }
}

// ---------------------------------------------------------------------------
// plan: derive the block structure from the configuration list, count, scan
// ---------------------------------------------------------------------------
bs2e_block* block_plan(bs2e_ctx* c, int L, long long n_config, const int64_t* conf_n,
                       const int64_t* conf_l, int full, long long row_lo, long long row_hi)
{
    HostPlan hp;
    try {
        hp = build_host_plan(c->hg, L, n_config, conf_n, conf_l, full, row_lo, row_hi);
    } catch (const std::invalid_argument& e) {
        throw Error(e.what());
    }
    const int nblk = hp.nblk;
    bs2e_block* b = new bs2e_block();
    b->ctx = c;
    b->L = L;
    b->full = full ? 1 : 0;
    b->n_config = n_config;
    b->row_lo = row_lo;
    b->row_hi = row_hi;
    b->lmax = hp.lmax;
    try {
        cudaStream_t st = c->stream;
        b->d_blk = dev_upload(hp.blocks, st);
        b->d_ncrow = dev_upload(hp.ncrow, st);
        b->d_flags = dev_upload(hp.flags, st);
        b->d_krange = dev_upload(hp.krange, st);
        b->d_angD = dev_upload(hp.angD, st);
        b->d_angX = dev_upload(hp.angX, st);
        b->d_row_n1 = dev_upload(hp.row_n1, st);
        b->d_row_n2 = dev_upload(hp.row_n2, st);
        b->d_row_blk = dev_upload(hp.row_blk, st);
        Plan& pl = b->dplan;
        pl.nblk = nblk;
        pl.n_config = (int)n_config;
        pl.full = b->full;
        pl.L = L;
        pl.blk = b->d_blk;
        pl.ncrow = b->d_ncrow;
        pl.flags = b->d_flags;
        pl.krange = b->d_krange;
        pl.angD = b->d_angD;
        pl.angX = b->d_angX;
        pl.row_n1 = b->d_row_n1;
        pl.row_n2 = b->d_row_n2;
        pl.row_blk = b->d_row_blk;

        const long long nrows = row_hi - row_lo + 1;
        b->d_cntH = dev_alloc<long long>(nrows + 1);
        b->d_cntS = dev_alloc<long long>(nrows + 1);
        b->d_Hptr = dev_alloc<long long>(nrows + 1);
        b->d_Sptr = dev_alloc<long long>(nrows + 1);
        size_t tmp = 0;
        BS2E_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp, b->d_cntH, b->d_Hptr, nrows + 1, st));
        b->scan_tmp_bytes = tmp;
        BS2E_CUDA(cudaMalloc(&b->d_scan_tmp, tmp ? tmp : 1));
        block_count_scan(b, true);
    } catch (...) {
        block_free(b);
        throw;
    }
    return b;
}

void block_assemble(bs2e_block* b)
{
    bs2e_ctx* c = b->ctx;
    if (!c->have_R) throw Error("block_assemble: call bs2e_rk_build first");
    if (!c->have_1p) throw Error("block_assemble: call bs2e_set_one_particle first");
    if (b->lmax > c->lmax_1p) throw Error("block_assemble: configuration l exceeds max_l_1p of H_vec");
    if (!b->d_Hidx) {
        b->d_Hidx = dev_alloc<long long>(b->nnzH);
        b->d_Hdat = dev_alloc<double>(2 * (size_t)b->nnzH);
        b->d_Sidx = dev_alloc<long long>(b->nnzS);
        b->d_Sdat = dev_alloc<double>(2 * (size_t)b->nnzS);
    }
    const long long nrows = b->row_hi - b->row_lo + 1;
    block_fill_kernel<<<(unsigned)((nrows + kFillWarps - 1) / kFillWarps), kFillWarps * 32, 0,
                        c->stream>>>(c->dg, b->dplan, c->one_body(), c->d_R, b->row_lo, nrows,
                                     b->d_Hptr, b->d_Sptr, b->d_Hidx,
                                     reinterpret_cast<double2*>(b->d_Hdat), b->d_Sidx,
                                     reinterpret_cast<double2*>(b->d_Sdat));
    BS2E_LAUNCHED();
    b->assembled = true;
}

void block_download(bs2e_block* b, int64_t* H_ptr, int64_t* H_idx, double* H_dat, int64_t* S_ptr,
                    int64_t* S_idx, double* S_dat)
{
    if (!b->assembled) throw Error("block_download: call bs2e_block_assemble first");
    cudaStream_t st = b->ctx->stream;
    const long long nrows = b->row_hi - b->row_lo + 1;
    auto d2h = [&](void* dst, const void* src, size_t bytes) {
        if (dst && bytes) BS2E_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, st));
    };
    d2h(H_ptr, b->d_Hptr, sizeof(long long) * (nrows + 1));
    d2h(S_ptr, b->d_Sptr, sizeof(long long) * (nrows + 1));
    d2h(H_idx, b->d_Hidx, sizeof(long long) * b->nnzH);
    d2h(S_idx, b->d_Sidx, sizeof(long long) * b->nnzS);
    d2h(H_dat, b->d_Hdat, sizeof(double) * 2 * b->nnzH);
    d2h(S_dat, b->d_Sdat, sizeof(double) * 2 * b->nnzS);
    BS2E_CUDA(cudaStreamSynchronize(st));
}

void block_row_counts(bs2e_block* b, int64_t* cH, int64_t* cS)
{
    cudaStream_t st = b->ctx->stream;
    const long long nrows = b->row_hi - b->row_lo + 1;
    if (cH) BS2E_CUDA(cudaMemcpyAsync(cH, b->d_cntH, sizeof(long long) * nrows, cudaMemcpyDeviceToHost, st));
    if (cS) BS2E_CUDA(cudaMemcpyAsync(cS, b->d_cntS, sizeof(long long) * nrows, cudaMemcpyDeviceToHost, st));
    BS2E_CUDA(cudaStreamSynchronize(st));
}

void block_checksum(bs2e_block* b, uint64_t* sH, uint64_t* sS)
{
    if (!b->assembled) throw Error("block_checksum: call bs2e_block_assemble first");
    cudaStream_t st = b->ctx->stream;
    unsigned long long* d = dev_alloc<unsigned long long>(2);
    unsigned long long h[2] = {0, 0};
    try {
        BS2E_CUDA(cudaMemsetAsync(d, 0, 2 * sizeof(unsigned long long), st));
        if (b->nnzH > 0) {
            checksum_kernel<<<148 * 8, 256, 0, st>>>(b->nnzH, b->d_Hidx, b->d_Hdat, d);
            BS2E_LAUNCHED();
        }
        if (b->nnzS > 0) {
            checksum_kernel<<<148 * 8, 256, 0, st>>>(b->nnzS, b->d_Sidx, b->d_Sdat, d + 1);
            BS2E_LAUNCHED();
        }
        BS2E_CUDA(cudaMemcpyAsync(h, d, sizeof(h), cudaMemcpyDeviceToHost, st));
        BS2E_CUDA(cudaStreamSynchronize(st));
    } catch (...) {
        cudaFree(d);
        throw;
    }
    cudaFree(d);
    if (sH) *sH = h[0];
    if (sS) *sS = h[1];
}

void block_free(bs2e_block* b)
{
    if (!b) return;
    cudaFree(b->d_blk); cudaFree(b->d_ncrow); cudaFree(b->d_flags); cudaFree(b->d_krange);
    cudaFree(b->d_angD); cudaFree(b->d_angX);
    cudaFree(b->d_row_n1); cudaFree(b->d_row_n2); cudaFree(b->d_row_blk);
    cudaFree(b->d_cntH); cudaFree(b->d_cntS); cudaFree(b->d_Hptr); cudaFree(b->d_Sptr);
    cudaFree(b->d_Hidx); cudaFree(b->d_Sidx); cudaFree(b->d_Hdat); cudaFree(b->d_Sdat);
    cudaFree(b->d_scan_tmp);
    delete b;
}

}  // namespace bs2e
