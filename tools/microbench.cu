// Latency/throughput micro-benchmarks used to reason about the trajectory kernel:
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/microbench tools/microbench.cu && /tmp/microbench
#include <cstdio>
#include <cuda_runtime.h>

__global__ void dfma_chain(double *out, long long *cyc, int iters, double a, double b) {
    double x = threadIdx.x * 1e-3;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) x = fma(x, a, b);
    }
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}
template <int ILP>
__global__ void dfma_ilp(double *out, long long *cyc, int iters, double a, double b) {
    double x[ILP];
    for (int k = 0; k < ILP; ++k) x[k] = threadIdx.x * 1e-3 + k;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int k = 0; k < ILP; ++k) x[k] = fma(x[k], a, b);
    }
    long long t1 = clock64();
    double s = 0;
    for (int k = 0; k < ILP; ++k) s += x[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
// DFMA with three distinct, changing register operands (no operand reuse): register-file bandwidth test
template <int ILP>
__global__ void dfma_3op(double *out, long long *cyc, int iters) {
    double x[ILP], y[ILP], z[ILP];
    for (int k = 0; k < ILP; ++k) { x[k] = threadIdx.x * 1e-3 + k; y[k] = 0.999 + 1e-6 * k; z[k] = 1.0 - 1e-7 * k; }
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int k = 0; k < ILP; ++k) {
                x[k] = fma(y[k], z[(k + 1) % ILP], x[k]);
                y[k] = fma(z[k], x[(k + 3) % ILP], y[k]);
                z[k] = fma(x[k], y[(k + 5) % ILP], z[k]);
            }
    }
    long long t1 = clock64();
    double s = 0;
    for (int k = 0; k < ILP; ++k) s += x[k] + y[k] + z[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
__global__ void shfl_chain(double *out, long long *cyc, int iters) {
    double x = threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) x = __shfl_xor_sync(0xffffffffu, x, 1) + 1.0;
    }
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}
__global__ void smem_chain(double *out, long long *cyc, int iters) {
    __shared__ double buf[64];
    double x = threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            buf[(u & 1) * 32 + threadIdx.x] = x;
            __syncwarp();
            x = buf[(u & 1) * 32 + (threadIdx.x ^ 1)] + 1.0;
        }
    }
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}

// Do FP64 MMA (mma.sync m8n8k4) and DFMA share an execution pipe?  MODE 0: DFMA only, 1: DMMA only, 2: both interleaved.
// NF independent DFMA chains and ND independent DMMA accumulators per thread per inner iteration.
template <int MODE, int NF, int ND>
__global__ void dmma_dfma_mix(double *out, long long *cyc, int iters, double a, double b) {
    double x[NF], c[ND][2];
    for (int k = 0; k < NF; ++k) x[k] = threadIdx.x * 1e-3 + k;
    for (int k = 0; k < ND; ++k) { c[k][0] = k; c[k][1] = -k; }
    double fa = 1e-3 * (threadIdx.x % 7), fb = 1e-3 * (threadIdx.x % 5);
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (MODE != 1) {
#pragma unroll
                for (int k = 0; k < NF; ++k) x[k] = fma(x[k], a, b);
            }
            if (MODE != 0) {
#pragma unroll
                for (int k = 0; k < ND; ++k)
                    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                                 : "+d"(c[k][0]), "+d"(c[k][1]) : "d"(fa), "d"(fb));
            }
        }
    }
    long long t1 = clock64();
    double s = 0;
    for (int k = 0; k < NF; ++k) s += x[k];
    for (int k = 0; k < ND; ++k) s += c[k][0] + c[k][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
__global__ void dmma_chain(double *out, long long *cyc, int iters) {
    double c0 = 0, c1 = 1, fa = 1e-3 * (threadIdx.x % 7), fb = 1e-3 * (threadIdx.x % 5);
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(fa), "d"(fb));
    }
    long long t1 = clock64();
    out[threadIdx.x] = c0 + c1;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}

int main() {
    double *out; long long *cyc, h;
    cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 8);
    const int iters = 4096;
    dfma_chain<<<1, 32>>>(out, cyc, iters, 0.999, 1e-9); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("dependent DFMA latency: %.2f cycles\n", (double)h / (iters * 16.0));
    shfl_chain<<<1, 32>>>(out, cyc, iters); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("SHFL(64-bit)+DADD dependent round: %.2f cycles\n", (double)h / (iters * 16.0));
    smem_chain<<<1, 32>>>(out, cyc, iters); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("STS.64 + syncwarp + LDS.64 + DADD dependent round: %.2f cycles\n", (double)h / (iters * 16.0));
#define RUN(ILP, WARPS)                                                                                        \
    dfma_ilp<ILP><<<1, 32 * WARPS>>>(out, cyc, iters, 0.999, 1e-9); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); \
    printf("1 SM, %d warps x ILP %d: %.2f cycles per warp-DFMA per SMSP-warp (%.1f%% of 2-cycle issue)\n", WARPS, ILP,         \
           (double)h / (iters * 8.0 * ILP), 100.0 * 2.0 * ((WARPS + 3) / 4) / ((double)h / (iters * 8.0 * ILP)));
    RUN(1, 4) RUN(2, 4) RUN(4, 4) RUN(8, 4) RUN(1, 8) RUN(2, 8) RUN(4, 8) RUN(1, 16) RUN(2, 16)
#define RUN3(ILP, WARPS)                                                                                       \
    dfma_3op<ILP><<<1, 32 * WARPS>>>(out, cyc, 1024); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);             \
    printf("3-operand DFMA, %d warps x ILP %d: %.2f cycles per warp-DFMA per scheduler\n", WARPS, ILP,               \
           (double)h / (1024 * 4.0 * 3 * ILP) / ((WARPS + 3) / 4));
    RUN3(8, 4) RUN3(8, 8) RUN3(4, 8) RUN3(8, 16)
    dmma_chain<<<1, 32>>>(out, cyc, iters); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("dependent DMMA m8n8k4 latency: %.2f cycles\n", (double)h / (iters * 16.0));
#define MIX(MODE, NF, ND, WARPS)                                                                                 \
    dmma_dfma_mix<MODE, NF, ND><<<1, 32 * WARPS>>>(out, cyc, 1024, 0.999, 1e-9); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); \
    printf("1 SM, %d warps, per inner iteration %d DFMA + %d DMMA (mode %d): %.1f cycles\n", WARPS, MODE == 1 ? 0 : NF, MODE == 0 ? 0 : ND, MODE, \
           (double)h / (1024 * 4.0));
    MIX(0, 8, 2, 8) MIX(1, 8, 2, 8) MIX(2, 8, 2, 8) MIX(0, 8, 4, 8) MIX(1, 8, 4, 8) MIX(2, 8, 4, 8) MIX(1, 8, 1, 8) MIX(2, 8, 1, 8)
    return 0;
}
