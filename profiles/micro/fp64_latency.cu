// Microbenchmark: dependent-issue latency and per-SMSP throughput of FP64 ops on the device (B200).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_latency fp64_latency.cu && ./fp64_latency
#include <cstdio>
#include <cuda_runtime.h>

template <int OP>
__global__ void lat(double *out, long long *cyc, double a, double b, int iters) {
    double x = a + threadIdx.x, y = b;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            if (OP == 0) x = fma(x, y, y);
            if (OP == 1) x = x + y;
            if (OP == 2) x = fmin(x + y, a);
            if (OP == 3) { double up = fmin(x + y, a), dn = fmax(x - b, 0.0); x = (x <= a) ? up : dn; }
            if (OP == 4) x = __shfl_sync(0xffffffffu, x, (k + 1) & 31);
            if (OP == 5) x = (double)(float)x * y;
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = x;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

// throughput: W warps per block, each with ILP independent chains
template <int ILP>
__global__ void thr(double *out, long long *cyc, double a, double b, int iters) {
    double x[ILP];
#pragma unroll
    for (int j = 0; j < ILP; ++j) x[j] = a + threadIdx.x + j;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k)
#pragma unroll
            for (int j = 0; j < ILP; ++j) x[j] = fma(x[j], b, a);
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int j = 0; j < ILP; ++j) s += x[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

int main() {
    double *out; long long *cyc, h;
    cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 8);
    const int iters = 4096;
    const char *names[] = {"DFMA", "DADD", "DADD+fmin", "att_update (2 DADD, fmin, fmax, select)", "SHFL.IDX 64-bit", "F2F f64->f32->f64 + DMUL"};
#define RUN(OP) lat<OP><<<1, 32>>>(out, cyc, 1.0000001, 1e-9, iters); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); \
    printf("latency %-45s %.2f cycles/op\n", names[OP], (double)h / (iters * 16.0));
    RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5)
    for (int warps = 1; warps <= 16; warps *= 2) {
        thr<1><<<1, 32 * warps>>>(out, cyc, 1.0, 0.999, iters); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("throughput 1 SM, %2d warps, ILP1: %.3f DFMA warp-inst/cycle/SM\n", warps, warps * iters * 8.0 / h);
        thr<4><<<1, 32 * warps>>>(out, cyc, 1.0, 0.999, iters); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("throughput 1 SM, %2d warps, ILP4: %.3f DFMA warp-inst/cycle/SM\n", warps, warps * iters * 8.0 * 4 / h);
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
