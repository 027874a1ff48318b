// Microbenchmark behind DESIGN.md "scan vs block-Toeplitz": what the FP64 TENSOR path (DMMA, mma.sync.m8n8k4.f64)
// delivers on B200 next to the FP64 FMA pipe, alone and in the same instruction stream.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_vs_dfma dmma_vs_dfma.cu && ./dmma_vs_dfma
// A recursive section costs 4-5 DFMA per sample as a scan.  As a block-Toeplitz product over blocks of 8 samples
// (the m8n8k4 shape: T[8x8] X[8 samples x 8 streams] in two k = 4 steps, + the state's contribution and the state
// update) it costs >= 4 DMMA warp instructions of 256 MACs per 64 samples = 16 MAC per sample against 5: the tensor
// formulation only wins if a DMMA warp instruction retires >= 3.2 x the MACs per cycle of the DFMA pipe (32 per warp
// instruction), or if the two pipes run side by side.  This file measures both.
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double &d0, double &d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// NF independent DFMA chains and NM independent DMMA accumulator pairs per thread, per inner step
template <int NF, int NM>
__global__ void mix(double *out, long long *cyc, double a, double b, int iters) {
    double x[NF > 0 ? NF : 1], c0[NM > 0 ? NM : 1], c1[NM > 0 ? NM : 1];
#pragma unroll
    for (int j = 0; j < NF; ++j) x[j] = a + threadIdx.x + j;
#pragma unroll
    for (int j = 0; j < NM; ++j) { c0[j] = j; c1[j] = -j; }
    const double fa = a * 1e-3 + threadIdx.x * 1e-6, fb = b * 1e-3;
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
#pragma unroll
            for (int j = 0; j < NM; ++j) dmma(c0[j], c1[j], fa, fb);
#pragma unroll
            for (int j = 0; j < NF; ++j) x[j] = fma(x[j], b, a);
        }
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int j = 0; j < NF; ++j) s += x[j];
#pragma unroll
    for (int j = 0; j < NM; ++j) s += c0[j] + c1[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int NF, int NM>
static void run(const char *what, double *out, long long *cyc, int sms, double ghz) {
    const int iters = 2048, warps = 16;
    mix<NF, NM><<<sms, 32 * warps>>>(out, cyc, 1.0, 0.999, iters);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    mix<NF, NM><<<sms, 32 * warps>>>(out, cyc, 1.0, 0.999, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    const double n_f = (double)warps * iters * 8 * NF, n_m = (double)warps * iters * 8 * NM;   // warp instructions per SM
    const double macs = (n_f * 32 + n_m * 256) * sms;
    printf("%-34s %6.3f DFMA + %6.3f DMMA warp-inst/cycle/SM   %7.1f MAC/cycle/SM   %6.2f TFLOP/s (events, %d SMs, %.3f ms)\n",
           what, n_f / h, n_m / h, (n_f * 32 + n_m * 256) / h, 2.0 * macs / (ms * 1e-3) / 1e12, sms, ms);
    (void)ghz;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    double *out; long long *cyc;
    cudaMalloc(&out, (size_t)p.multiProcessorCount * 512 * 8); cudaMalloc(&cyc, (size_t)p.multiProcessorCount * 8);
    const int sms = p.multiProcessorCount;
    printf("%s, %d SMs\n", p.name, sms);
    run<4, 0>("DFMA only (ILP 4, 16 warps)", out, cyc, sms, 0);
    run<0, 1>("DMMA only (1 accumulator pair)", out, cyc, sms, 0);
    run<0, 4>("DMMA only (4 accumulator pairs)", out, cyc, sms, 0);
    run<4, 1>("DFMA x4 + DMMA x1 interleaved", out, cyc, sms, 0);
    run<4, 4>("DFMA x4 + DMMA x4 interleaved", out, cyc, sms, 0);
    run<1, 4>("DFMA x1 + DMMA x4 interleaved", out, cyc, sms, 0);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
