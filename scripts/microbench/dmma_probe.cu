// Microbenchmark (profiling aid, not product): FP64 tensor-core mma.sync.m8n8k4 (SASS DMMA.8x8x4) on the
// device it runs on -- dependent-issue latency, per-SM throughput against warps and independent
// accumulator chains, the fragment layout, and the accumulation order inside one instruction (is the
// result the ascending-k FMA chain, bit for bit?).
#include <cmath>
#include <cstdio>
#include <cstring>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b, double c0, double c1)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%4,%5};"
                 : "=d"(d0), "=d"(d1) : "d"(a), "d"(b), "d"(c0), "d"(c1));
}

template <int CH>
__global__ void dmma_chains(double* out, long long* cyc, int n, double a, double b)
{
    double c0[CH], c1[CH];
    for (int c = 0; c < CH; ++c) { c0[c] = threadIdx.x * 1e-9 + c; c1[c] = 1.0; }
    const long long t0 = clock64();
    for (int i = 0; i < n; ++i) {
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int c = 0; c < CH; ++c) dmma(c0[c], c1[c], a, b, c0[c], c1[c]);
    }
    const long long t1 = clock64();
    double s = 0;
    for (int c = 0; c < CH; ++c) s += c0[c] + c1[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// one 8x8x4 product with A[m][k], B[k][n] given densely; writes C[m][n]
__global__ void dmma_layout(const double* A, const double* B, const double* C, double* D)
{
    const int lane = threadIdx.x;
    const double a = A[(lane >> 2) * 4 + (lane & 3)];        // A[row = lane/4][k = lane%4]
    const double b = B[(lane & 3) * 8 + (lane >> 2)];        // B[k = lane%4][col = lane/4]
    const int r = lane >> 2, c = 2 * (lane & 3);
    double d0, d1;
    dmma(d0, d1, a, b, C[r * 8 + c], C[r * 8 + c + 1]);      // C[row = lane/4][col = 2*(lane%4) + {0,1}]
    D[r * 8 + c] = d0;
    D[r * 8 + c + 1] = d1;
}

int main()
{
    double* out; long long* cyc;
    cudaMalloc(&out, sizeof(double) * 1024);
    cudaMalloc(&cyc, sizeof(long long) * 4);
    long long h = 0;
    const int n = 2000;
    dmma_chains<1><<<1, 32>>>(out, cyc, n, 1.0000001, 1e-9);
    cudaMemcpy(&h, cyc, sizeof(long long), cudaMemcpyDeviceToHost);
    printf("DMMA.8x8x4 dependent-issue latency: %.2f cycles\n", (double)h / (4.0 * n));
    for (int warps : {1, 2, 4, 8, 16, 32}) {
        dmma_chains<1><<<1, warps * 32>>>(out, cyc, n, 1.0000001, 1e-9);
        cudaMemcpy(&h, cyc, sizeof(long long), cudaMemcpyDeviceToHost);
        const double r1 = warps * 4.0 * n / (double)h;
        dmma_chains<2><<<1, warps * 32>>>(out, cyc, n, 1.0000001, 1e-9);
        cudaMemcpy(&h, cyc, sizeof(long long), cudaMemcpyDeviceToHost);
        const double r2 = warps * 8.0 * n / (double)h;
        dmma_chains<4><<<1, warps * 32>>>(out, cyc, n, 1.0000001, 1e-9);
        cudaMemcpy(&h, cyc, sizeof(long long), cudaMemcpyDeviceToHost);
        const double r4 = warps * 16.0 * n / (double)h;
        printf("%2d warps/SM: %.4f / %.4f / %.4f DMMA/cycle/SM with 1 / 2 / 4 chains per warp  (x512 flop: %.1f flop/cycle/SM)\n",
               warps, r1, r2, r4, r4 * 512);
    }
    // layout and accumulation order
    double A[32], B[32], C[64], D[64];
    unsigned long long s = 88172645463325252ull;
    auto rnd = [&] { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return (double)(s >> 11) / 9007199254740992.0 - 0.5; };
    int same_asc = 0, same_desc = 0, same_pair = 0, close = 0, trials = 200;
    double *dA, *dB, *dC, *dD;
    cudaMalloc(&dA, sizeof(A)); cudaMalloc(&dB, sizeof(B)); cudaMalloc(&dC, sizeof(C)); cudaMalloc(&dD, sizeof(D));
    for (int t = 0; t < trials; ++t) {
        for (double& v : A) v = rnd() * std::ldexp(1.0, (int)(rnd() * 40));
        for (double& v : B) v = rnd();
        for (double& v : C) v = rnd() * 1e-3;
        cudaMemcpy(dA, A, sizeof(A), cudaMemcpyHostToDevice);
        cudaMemcpy(dB, B, sizeof(B), cudaMemcpyHostToDevice);
        cudaMemcpy(dC, C, sizeof(C), cudaMemcpyHostToDevice);
        dmma_layout<<<1, 32>>>(dA, dB, dC, dD);
        cudaMemcpy(D, dD, sizeof(D), cudaMemcpyDeviceToHost);
        bool asc = true, desc = true, pair = true, cl = true;
        for (int m = 0; m < 8; ++m)
            for (int nn = 0; nn < 8; ++nn) {
                double x = C[m * 8 + nn], y = C[m * 8 + nn], mag = std::fabs(C[m * 8 + nn]);
                for (int k = 0; k < 4; ++k) x = std::fma(A[m * 4 + k], B[k * 8 + nn], x);
                for (int k = 0; k < 4; ++k) mag += std::fabs(A[m * 4 + k] * B[k * 8 + nn]);
                for (int k = 3; k >= 0; --k) y = std::fma(A[m * 4 + k], B[k * 8 + nn], y);
                const double p = std::fma(A[m * 4 + 1], B[8 + nn], A[m * 4] * B[nn]) +
                                 std::fma(A[m * 4 + 3], B[24 + nn], A[m * 4 + 2] * B[16 + nn]) + C[m * 8 + nn];
                const double got = D[m * 8 + nn];
                asc = asc && got == x;
                desc = desc && got == y;
                pair = pair && got == p;
                cl = cl && std::fabs(got - x) <= 1e-15 * mag;
            }
        same_asc += asc; same_desc += desc; same_pair += pair; close += cl;
    }
    printf("layout check (A row-major lane/4,lane%%4; B lane%%4,lane/4; C lane/4, 2*(lane%%4)+i): %d/%d products agree to rounding\n", close, trials);
    printf("bitwise equal to the ascending-k FMA chain: %d/%d, descending: %d/%d, pairwise: %d/%d\n",
           same_asc, trials, same_desc, trials, same_pair, trials);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
