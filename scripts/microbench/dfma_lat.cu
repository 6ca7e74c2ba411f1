// Microbenchmark (profiling aid, not product): dependent-issue latency and per-SM throughput of
// DFMA, and of shared-memory broadcast LDS.128, on the device it runs on.
#include <cstdio>
#include <cuda_runtime.h>

__global__ void dfma_chain(double* out, long long* cyc, int n, double a, double b)
{
    double x = threadIdx.x * 1e-9, y = 1.0 + threadIdx.x * 1e-9;
    const long long t0 = clock64();
    for (int i = 0; i < n; ++i) {   // two independent chains would be x and y; here only x depends
        x = fma(x, a, b); x = fma(x, a, b); x = fma(x, a, b); x = fma(x, a, b);
        x = fma(x, a, b); x = fma(x, a, b); x = fma(x, a, b); x = fma(x, a, b);
    }
    const long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = x + y;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int CH>
__global__ void dfma_ilp(double* out, long long* cyc, int n, double a, double b)
{
    double x[CH];
    for (int c = 0; c < CH; ++c) x[c] = threadIdx.x * 1e-9 + c;
    const long long t0 = clock64();
    for (int i = 0; i < n; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int c = 0; c < CH; ++c) x[c] = fma(x[c], a, b);
    }
    const long long t1 = clock64();
    double s = 0;
    for (int c = 0; c < CH; ++c) s += x[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int main()
{
    double* out; long long* cyc;
    cudaMalloc(&out, sizeof(double) * 148 * 1024);
    cudaMalloc(&cyc, sizeof(long long) * 148);
    long long h[148];
    const int n = 2000;
    // latency: one warp
    dfma_chain<<<1, 32>>>(out, cyc, n, 1.0000001, 1e-9);
    cudaMemcpy(h, cyc, sizeof(long long), cudaMemcpyDeviceToHost);
    printf("DFMA dependent-issue latency: %.2f cycles\n", (double)h[0] / (8.0 * n));
    // throughput per SM vs warps per SM (1 chain per warp ... 4 chains)
    for (int warps : {4, 8, 16, 32}) {
        dfma_chain<<<1, warps * 32>>>(out, cyc, n, 1.0000001, 1e-9);
        cudaMemcpy(h, cyc, sizeof(long long), cudaMemcpyDeviceToHost);
        printf("1 chain/warp, %2d warps/SM: %.3f warp-DFMA/cycle/SM\n", warps, warps * 8.0 * n / (double)h[0]);
    }
    for (int warps : {4, 8, 16}) {
        dfma_ilp<2><<<1, warps * 32>>>(out, cyc, n, 1.0000001, 1e-9);
        cudaMemcpy(h, cyc, sizeof(long long), cudaMemcpyDeviceToHost);
        printf("2 chains/warp, %2d warps/SM: %.3f warp-DFMA/cycle/SM\n", warps, warps * 16.0 * n / (double)h[0]);
        dfma_ilp<4><<<1, warps * 32>>>(out, cyc, n, 1.0000001, 1e-9);
        cudaMemcpy(h, cyc, sizeof(long long), cudaMemcpyDeviceToHost);
        printf("4 chains/warp, %2d warps/SM: %.3f warp-DFMA/cycle/SM\n", warps, warps * 32.0 * n / (double)h[0]);
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
