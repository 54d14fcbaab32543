// FP64 roofline denominator: MEASURED_PEAKS.json has no FP64 entry, so the library measures
// the DFMA issue rate itself — 8 independent FMA chains per thread, all SMs full, CUDA-event timed.
#include "jq_common.h"
#include "../../include/juqbox_b200.h"


__global__ void __launch_bounds__(256) jq_dfma_peak_kernel(double *out, int iters, double a, double b) {
    double x0 = threadIdx.x * 1e-3, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
            x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}

// Same, but every DFMA has three distinct, changing register operands (no operand-reuse cache hits): the register
// file then delivers one warp-DFMA per 3 cycles per scheduler instead of 2 (measured on B200), which is the realistic
// ceiling for code whose FMAs are not coefficient-broadcast shaped.
__global__ void __launch_bounds__(256) jq_dfma_3op_kernel(double *out, int iters) {
    double x[8], y[8], z[8];
    for (int k = 0; k < 8; ++k) { x[k] = threadIdx.x * 1e-3 + k; y[k] = 0.999 + 1e-6 * k; z[k] = 1.0 - 1e-7 * k; }
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            x[k] = fma(y[k], z[(k + 1) % 8], x[k]);
            y[k] = fma(z[k], x[(k + 3) % 8], y[k]);
            z[k] = fma(x[k], y[(k + 5) % 8], z[k]);
        }
    }
    double s = 0;
    for (int k = 0; k < 8; ++k) s += x[k] + y[k] + z[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// FP64 tensor-core (DMMA) issue rate: mma.sync m16n8k16 f64, four independent accumulator tiles per warp.  The
// trajectory kernels do not use this pipe (the control operators have <= 2 nonzeros per row, so a dense contraction
// would spend 4.4x (cnot2) to 15x (cnot3) more flops than the structural-nonzero form); the number is reported so
// that the decision can be checked: DMMA peak / dense-to-nnz flop ratio < achieved DFMA rate.
__global__ void __launch_bounds__(256) jq_dmma_peak_kernel(double *out, int iters) {
    double a[8], b[4], c[4][4];
    for (int k = 0; k < 8; ++k) a[k] = 1e-3 * (threadIdx.x % 7) + 1e-4 * k;
    for (int k = 0; k < 4; ++k) b[k] = 1e-3 * (threadIdx.x % 5) - 1e-4 * k;
    for (int t = 0; t < 4; ++t) for (int k = 0; k < 4; ++k) c[t][k] = t + k;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int t = 0; t < 4; ++t)
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, "
                         "{%12,%13,%14,%15}, {%0,%1,%2,%3};"
                         : "+d"(c[t][0]), "+d"(c[t][1]), "+d"(c[t][2]), "+d"(c[t][3])
                         : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                           "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
    }
    double s = 0;
    for (int t = 0; t < 4; ++t) for (int k = 0; k < 4; ++k) s += c[t][k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

static int time_kernel(int device, int which, double *tflops) {
    if (!tflops) return JQ_ERR_ARG;
    if (cudaSetDevice(device) != cudaSuccess) return JQ_ERR_CUDA;
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    const int blocks = sms * 8, threads = 256, iters = which == 0 ? 4096 : 8192;
    double *buf = nullptr;
    if (cudaMalloc(&buf, sizeof(double) * blocks * threads) != cudaSuccess) return JQ_ERR_ALLOC;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    double best = 0.0;
    for (int rep = 0; rep < 6; ++rep) {
        cudaEventRecord(e0);
        if (which == 0) jq_dfma_peak_kernel<<<blocks, threads>>>(buf, iters, 0.999999, 1e-9);
        else if (which == 1) jq_dfma_3op_kernel<<<blocks, threads>>>(buf, iters);
        else jq_dmma_peak_kernel<<<blocks, threads>>>(buf, iters);
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) { cudaFree(buf); return JQ_ERR_CUDA; }
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        // per thread and iteration: 64 / 24 DFMA, or 4 warp-wide m16n8k16 MMAs = 4*16*8*16/32 multiply-adds per lane
        const double fmas = (which == 0 ? 64.0 : which == 1 ? 24.0 : 256.0) * iters * (double)blocks * threads;
        if (rep > 0 && ms > 0.f) best = best > 2.0 * fmas / (ms * 1e9) ? best : 2.0 * fmas / (ms * 1e9);
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(buf);
    *tflops = best;
    return 0;
}

extern "C" int jq_fp64_peak_dmma(int device, double *tflops) { return time_kernel(device, 2, tflops); }

extern "C" int jq_fp64_peak_3op(int device, double *tflops) { return time_kernel(device, 1, tflops); }

extern "C" int jq_fp64_peak(int device, double *tflops) {
    return time_kernel(device, 0, tflops);
}
