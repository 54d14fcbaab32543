// DFMA / DMUL / DADD issue interval per scheduler as a function of the register-operand pattern (operand-reuse cache).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/mb_reuse tools/microbench_reuse.cu && /tmp/mb_reuse
// MODE 0: t[k] = fma(c, x[k], t[k])      one coefficient shared by 8 consecutive FMAs (reuse on one operand)
// MODE 1: t[k] = fma(c[k], x[k], t[k])   three distinct registers per FMA, nothing shared between neighbours
// MODE 2: t[k] = fma(c[k&1], x[k], t[k]) coefficient alternates (no back-to-back sharing)
// MODE 3: t[k] = fma(c[k>>1], x[k], t[k]) coefficient shared by pairs
// MODE 4: y[k] = c * x[k] ; t[k] += y[k]   DMUL (shared c) + DADD
// MODE 6: as MODE 1 with warp-uniform coefficients (the compiler keeps them in uniform registers)
// MODE 5: t[k] = fma(c, x[k], t[k]) with x[k] changing every round (x rotates through registers)
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k(double *out, long long *cyc, int iters, double c0) {
    double t[8], x[8], c[8];
    for (int i = 0; i < 8; ++i) { t[i] = threadIdx.x * 1e-3 + i; x[i] = 1.0 + 1e-6 * (i + threadIdx.x); c[i] = c0 + 1e-9 * (i + (MODE == 6 ? 0 : threadIdx.x)); }
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (MODE == 0) t[i] = fma(c[0], x[i], t[i]);
                if (MODE == 1 || MODE == 6) t[i] = fma(c[i], x[i], t[i]);
                if (MODE == 2) t[i] = fma(c[i & 1], x[i], t[i]);
                if (MODE == 3) t[i] = fma(c[i >> 1], x[i], t[i]);
                if (MODE == 4) { double y = c[0] * x[i]; t[i] += y; }
                if (MODE == 5) t[i] = fma(c[0], x[(i + u) & 7], t[i]);
            }
            if (MODE == 5) {
#pragma unroll
                for (int i = 0; i < 8; ++i) x[i] = fma(c[1], t[i], x[i]);
            }
        }
    }
    long long t1 = clock64();
    double s = 0;
    for (int i = 0; i < 8; ++i) s += t[i] + x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

int main() {
    double *out; long long *cyc, h;
    cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 8);
    const int iters = 2048;
#define RUN(MODE, WARPS, NINST)                                                                                  \
    k<MODE><<<1, 32 * WARPS>>>(out, cyc, iters, 1e-7); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);            \
    printf("mode %d, %2d warps: %.2f cycles per FP64 warp-instruction per scheduler\n", MODE, WARPS, (double)h / (iters * 4.0 * NINST) / ((WARPS + 3) / 4));
    RUN(0, 8, 8) RUN(1, 8, 8) RUN(2, 8, 8) RUN(3, 8, 8) RUN(4, 8, 16) RUN(5, 8, 16) RUN(6, 8, 8)
    RUN(0, 4, 8) RUN(1, 4, 8) RUN(2, 4, 8) RUN(3, 4, 8) RUN(4, 4, 16) RUN(5, 4, 16)
    return 0;
}
