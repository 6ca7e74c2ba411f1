// Microbenchmark (profiling aid, not product): write bandwidth of one B200 for the store shapes the
// stage-C fill can use.
//   coalesced   every warp store instruction writes 512 contiguous, aligned bytes (STG.128)
//   runs128     every warp store writes four 128-byte runs that start at odd multiples of 16 bytes
//               (8 lanes per CSR row: the C-fragment layout of the tensor-core kernel)
//   runs128+64  the same plus an 8-byte index store per element (four 64-byte runs at odd multiples of 8)
//   bulk        cp.async.bulk shared -> global of `chunk` bytes at 16-byte aligned, otherwise arbitrary offsets
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__global__ void k_coalesced(double2* out, size_t n_per_cta, int iters)
{
    double2* p = out + (size_t)blockIdx.x * n_per_cta;
    const double2 v = make_double2(threadIdx.x, 0.0);
    for (int it = 0; it < iters; ++it)
        for (size_t i = threadIdx.x; i < n_per_cta; i += blockDim.x) __stcs(p + i, v);
}

// each group of 8 lanes writes rows of `run` elements; row r of the CTA starts at element r*pitch + 1 (odd)
__global__ void k_runs(double2* out, long long* idx, size_t n_per_cta, int with_idx)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int grp = lane >> 3, l8 = lane & 7;
    double2* p = out + (size_t)blockIdx.x * n_per_cta;
    long long* q = idx + (size_t)blockIdx.x * n_per_cta;
    // rows of 225 elements (one CSR row segment of a site), written 8 elements at a time by alternating warps
    const int rowlen = 225;
    const size_t nrows = n_per_cta / rowlen;
    for (size_t r0 = 0; r0 < nrows; r0 += 4) {
        const size_t row = r0 + grp;
        if (row >= nrows) continue;
        for (int seg = warp; seg * 32 < rowlen; seg += nw)
            for (int ct = 0; ct < 4; ++ct) {
                const int e = seg * 32 + ct * 8 + l8;
                if (e < rowlen) {
                    __stcs(p + row * rowlen + e, make_double2(e, 0.0));
                    if (with_idx) __stcs(q + row * rowlen + e, (long long)e);
                }
            }
    }
}

// the row-wise shape: one warp store = 32 consecutive values (512 B) of one CSR row, then their 32 indices (256 B);
// rows of 225 entries back to back (so that runs start at every 16-byte / 8-byte phase), 8 warps take the 8 segments
__global__ void k_rowwise(double2* out, long long* idx, size_t n_per_cta, int with_idx, int rowlen)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    double2* p = out + (size_t)blockIdx.x * n_per_cta;
    long long* q = idx + (size_t)blockIdx.x * n_per_cta;
    const size_t nrows = n_per_cta / rowlen;
    for (size_t row = 0; row < nrows; ++row)
        for (int seg = warp; seg * 32 < rowlen; seg += nw) {
            const int e = seg * 32 + lane;
            if (e < rowlen) {
                __stcs(p + row * rowlen + e, make_double2(e, 0.0));
                if (with_idx) __stcs(q + row * rowlen + e, (long long)e);
            }
        }
}

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__global__ void k_bulk(char* out, size_t bytes_per_cta, int chunk, int shift)
{
    extern __shared__ __align__(128) char sm[];
    for (int i = threadIdx.x; i < 32768 / 8; i += blockDim.x) reinterpret_cast<double*>(sm)[i] = i;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    char* base = out + (size_t)blockIdx.x * bytes_per_cta + shift;
    const size_t nchunks = (bytes_per_cta - 256) / chunk;
    // every thread issues chunks round robin; at most 4 groups in flight per thread
    int inflight = 0;
    for (size_t c = threadIdx.x; c < nchunks; c += blockDim.x) {
        const unsigned src = smem_u32(sm + ((c * 48) & 16383 & ~15));
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(base + c * chunk), "r"(src), "r"(chunk)
                     : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        if (++inflight >= 4) { asm volatile("cp.async.bulk.wait_group.read 3;" ::: "memory"); inflight = 3; }
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// the shape proposed for the fill: every warp owns a 4 KB staging tile (8 rows x 32 values), fills it with
// STS.128, hands each row to the bulk-copy engine (512 B at an odd 16-byte offset) and writes the 8-byte
// indices of the same elements with STG.64 (idx_mode 1) or stages them too and bulk-copies the aligned
// interior (idx_mode 2); idx_mode 0: values only
__global__ void k_mixed(char* out, long long* idx, size_t elems_per_warp, int idx_mode)
{
    extern __shared__ __align__(128) char sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int l4 = lane >> 2, l3 = lane & 3;
    char* tile = sm + warp * 6656;            // 4096 values + 2560 indices (8 rows x 40 x 8)
    const size_t w = (size_t)blockIdx.x * nw + warp;
    double2* vout = reinterpret_cast<double2*>(out) + w * elems_per_warp + 1;   // odd element offset
    long long* iout = idx + w * elems_per_warp + 1;
    const size_t ntiles = (elems_per_warp - 2) / 256;
    for (size_t t = 0; t < ntiles; ++t) {
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncwarp();
#pragma unroll
        for (int e = 0; e < 2; ++e)
#pragma unroll
            for (int ct = 0; ct < 4; ++ct) {
                const int row = 2 * l3 + e, col = ct * 8 + l4;
                *reinterpret_cast<double2*>(tile + row * 512 + col * 16) = make_double2((double)t, 0.0);
                if (idx_mode == 1) __stcs(iout + (t * 8 + row) * 32 + col, (long long)col);
                if (idx_mode == 2) *reinterpret_cast<long long*>(tile + 4096 + row * 320 + 8 + col * 8) = col;
            }
        if (idx_mode == 3) {   // indices row-wise: one 256-byte run per store instruction
#pragma unroll
            for (int row = 0; row < 8; ++row) __stcs(iout + (t * 8 + row) * 32 + lane, (long long)lane);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane < 8) {
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(vout + (t * 8 + lane) * 32),
                         "r"(smem_u32(tile + lane * 512)), "r"(512)
                         : "memory");
        } else if (lane < 16 && idx_mode == 2) {
            const int r = lane - 8;   // interior 30 indices (16-byte aligned both sides), the two ends by STG
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(iout + (t * 8 + r) * 32 + 1),
                         "r"(smem_u32(tile + 4096 + r * 320 + 16)), "r"(240)
                         : "memory");
        } else if (lane < 32 && idx_mode == 2) {
            const int r = (lane - 16) >> 1, end = lane & 1;
            __stcs(iout + (t * 8 + r) * 32 + (end ? 31 : 0), (long long)end);
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

int main()
{
    const size_t total = (size_t)8 << 30;   // 8 GiB of output
    char* buf; long long* idx;
    cudaMalloc(&buf, total + 4096);
    cudaMalloc(&idx, total / 2 + 4096);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms;
    const int ctas = 148 * 8;
    auto report = [&](const char* name, double bytes) {
        cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
        printf("%-44s %8.1f GB/s  (%s)\n", name, bytes / (ms * 1e-3) / 1e9, cudaGetErrorString(cudaGetLastError()));
    };
    cudaFuncSetAttribute(k_mixed, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 6656);
    for (int rep = 0; rep < 2; ++rep) {
        const size_t n_per = total / 16 / ctas;
        cudaEventRecord(e0);
        k_coalesced<<<ctas, 256>>>((double2*)buf, n_per, 1);
        report("coalesced STG.128", (double)n_per * ctas * 16);
        cudaEventRecord(e0);
        k_runs<<<ctas, 256>>>((double2*)buf, idx, n_per, 0);
        report("128-byte runs at odd offsets (values only)", (double)(n_per / 225) * 225 * ctas * 16);
        cudaEventRecord(e0);
        k_runs<<<ctas, 256>>>((double2*)buf, idx, n_per, 1);
        report("128-byte value runs + 64-byte index runs", (double)(n_per / 225) * 225 * ctas * 24);
        for (int rowlen : {225, 256}) {
            for (int wi = 0; wi < 2; ++wi) {
                cudaEventRecord(e0);
                k_rowwise<<<ctas, 256>>>((double2*)buf, idx, n_per, wi, rowlen);
                char name[96];
                snprintf(name, sizeof name, "row-wise 512 B runs%s, rows of %d", wi ? " + 256 B index runs" : "", rowlen);
                report(name, (double)(n_per / rowlen) * rowlen * ctas * (wi ? 24 : 16));
            }
        }
        for (int mode = 0; mode < 4; ++mode) {
            const size_t epw = (total / 16 / (ctas * 8)) & ~(size_t)255;
            cudaEventRecord(e0);
            k_mixed<<<ctas, 256, 8 * 6656>>>(buf, idx, epw, mode);
            const char* nm[4] = {"staged values -> bulk (512 B rows), no indices", "staged values -> bulk + indices by STG.64",
                                 "staged values and indices -> bulk (+2 STG per row)",
                                 "staged values -> bulk + indices row-wise (256 B runs)"};
            report(nm[mode], (double)((epw - 2) / 256) * 256 * ctas * 8 * (mode ? 24 : 16));
        }
        for (int chunk : {512, 3600}) {
            for (int shift : {0, 16}) {
                const size_t bpc = (total / ctas) & ~(size_t)255;
                cudaEventRecord(e0);
                k_bulk<<<ctas, 128, 32768>>>(buf, bpc, chunk, shift);
                char name[96];
                snprintf(name, sizeof name, "bulk S2G, %5d-byte chunks, offset %2d mod 128", chunk, shift);
                report(name, (double)((bpc - 256) / chunk) * chunk * ctas);
            }
        }
    }
    return 0;
}
