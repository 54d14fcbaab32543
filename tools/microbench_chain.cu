// Development micro-benchmark: per-step cost of a one-warp chain of dependent 32 x 32 matrix-vector products (the join of the
// time-parallel evaluation): what bounds it -- the DFMA dependency chain, the shuffles, the loads or the store.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void fill(double *M, size_t n) { for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) M[i] = 1e-3 * (double)(i % 97); }
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
// LOADS: 0 none, 1 one step ahead, 2 two steps ahead; ACC: independent partial sums; SHFL: 1 broadcast by shuffle, 0 own value; STORE
template <int LOADS, int ACC, int SHFL, int STORE>
__global__ void chain(const double *M, double *V, int nstep, long long *out) {
    const int lane = threadIdx.x;
    double x = 1.0 + lane, a[32], an[32], an2[32];
#pragma unroll
    for (int k = 0; k < 32; ++k) { an[k] = M[k * 32 + lane]; an2[k] = M[1024 + k * 32 + lane]; }
    const long long t0 = clock64();
    const unsigned long long g0 = gtime();
    for (int s = 0; s < nstep; ++s) {
#pragma unroll
        for (int k = 0; k < 32; ++k) { a[k] = an[k]; if (LOADS == 2) an[k] = an2[k]; }
        if (LOADS && s + LOADS < nstep) {
#pragma unroll
            for (int k = 0; k < 32; ++k) (LOADS == 2 ? an2[k] : an[k]) = M[(size_t)(s + LOADS) * 1024 + k * 32 + lane];
        }
        double acc[ACC];
#pragma unroll
        for (int k = 0; k < ACC; ++k) acc[k] = 0.0;
#pragma unroll
        for (int k = 0; k < 32; ++k) acc[k % ACC] = fma(a[k], SHFL ? __shfl_sync(0xffffffffu, x, k) : x, acc[k % ACC]);
#pragma unroll
        for (int w = ACC / 2; w > 0; w >>= 1) {
#pragma unroll
            for (int k = 0; k < w; ++k) acc[k] += acc[k + w];
        }
        x = acc[0];
        if (STORE) V[(size_t)s * 32 + lane] = x;
    }
    const long long t1 = clock64();
    const unsigned long long g1 = gtime();
    if (lane == 0) { out[0] = t1 - t0; out[1] = (long long)(g1 - g0); }
    V[lane] = x;
}
template <int LOADS, int ACC, int SHFL, int STORE>
void run(const char *what, double *M, double *V, long long *clk, int nstep) {
    long long h[2];
    for (int rep = 0; rep < 2; ++rep) {
        fill<<<148, 256>>>(M, (size_t)nstep * 1024);
        chain<LOADS, ACC, SHFL, STORE><<<1, 32>>>(M, V, nstep, clk);
        cudaMemcpy(h, clk, 16, cudaMemcpyDeviceToHost);
    }
    printf("%-44s %7.1f cycles  %7.1f ns per step\n", what, (double)h[0] / nstep, (double)h[1] / nstep);
}
int main() {
    const int nstep = 512;
    double *M, *V; long long *clk;
    cudaMalloc(&M, (size_t)nstep * 1024 * 8); cudaMalloc(&V, (size_t)nstep * 32 * 8); cudaMalloc(&clk, 16);
    run<0, 2, 1, 0>("no loads, 2 sums, shuffles", M, V, clk, nstep);
    run<0, 8, 1, 0>("no loads, 8 sums, shuffles", M, V, clk, nstep);
    run<0, 32, 1, 0>("no loads, 32 sums, shuffles", M, V, clk, nstep);
    run<0, 8, 0, 0>("no loads, 8 sums, no shuffles", M, V, clk, nstep);
    run<1, 8, 1, 0>("loads 1 ahead, 8 sums, shuffles", M, V, clk, nstep);
    run<2, 8, 1, 0>("loads 2 ahead, 8 sums, shuffles", M, V, clk, nstep);
    run<1, 8, 1, 1>("loads 1 ahead, 8 sums, shuffles, store", M, V, clk, nstep);
    run<2, 8, 1, 1>("loads 2 ahead, 8 sums, shuffles, store", M, V, clk, nstep);
    return 0;
}
