// Microbenchmark of the compressor attenuation recurrence (cycles per dependent step) on B200.
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ double upd_int(double att, double m, double inc, double dec) {
    const long long ia = __double_as_longlong(att), im = __double_as_longlong(m);
    const double s = att + inc, d = att - dec;
    const long long is = __double_as_longlong(s);
    const double r = (ia > im) ? d : m;
    return (ia <= im && is < im) ? s : r;
}
__device__ __forceinline__ double upd_fp(double att, double m, double inc, double dec) {
    const double up = fmin(att + inc, m), dn = fmax(att - dec, 0.0);
    return (att <= m) ? up : dn;
}
// high-word-only compare: valid when the operands differ in the top 32 bits or we accept a tie-break on the low word
__device__ __forceinline__ double upd_hi(double att, double m, double inc, double dec) {
    const double s = att + inc, d = att - dec;
    const bool p = att > m;                       // DSETP off the critical path? (depends on att)
    const double t = (s < m) ? s : m;
    return p ? d : t;
}

template <int V>
__global__ void reg_loop(double *out, long long *cyc, double m, double inc, double dec, int iters) {
    double att = 0.0;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            if (V == 0) att = upd_int(att, m, inc, dec);
            if (V == 1) att = upd_fp(att, m, inc, dec);
            if (V == 2) att = upd_hi(att, m, inc, dec);
        }
        m += 1e-7;
    }
    long long t1 = clock64();
    out[threadIdx.x] = att;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}

template <int UNROLL>
__global__ void smem_loop(double *out, long long *cyc, int total, int reps) {
    __shared__ double2 q01[256];
    __shared__ double q2[256];
    __shared__ double qa[256];
    for (int i = threadIdx.x; i < 256; i += 32) { q01[i] = make_double2(3.0 + 1e-3 * i, 0.0125); q2[i] = 0.00125; }
    __syncwarp();
    double att = 0.0;
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
#pragma unroll UNROLL
        for (int j = 0; j < total; ++j) {
            const double2 a = q01[j];
            att = upd_int(att, a.x, a.y, q2[j]);
            qa[j] = att;
        }
        __syncwarp();
    }
    long long t1 = clock64();
    out[threadIdx.x] = att + qa[threadIdx.x];
    if (threadIdx.x == 0) *cyc = t1 - t0;
}

int main() {
    double *out; long long *cyc, h;
    cudaMalloc(&out, 4096); cudaMalloc(&cyc, 8);
    const int iters = 2048;
    reg_loop<0><<<1, 32>>>(out, cyc, 6.0, 0.025, 0.0025, iters); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("register loop, integer compares : %.2f cycles/step\n", (double)h / (iters * 16.0));
    reg_loop<1><<<1, 32>>>(out, cyc, 6.0, 0.025, 0.0025, iters); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("register loop, fmin/fmax        : %.2f cycles/step\n", (double)h / (iters * 16.0));
    reg_loop<2><<<1, 32>>>(out, cyc, 6.0, 0.025, 0.0025, iters); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("register loop, DSETP selects    : %.2f cycles/step\n", (double)h / (iters * 16.0));
    smem_loop<1><<<1, 32>>>(out, cyc, 256, 64); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("smem queue loop, unroll 1       : %.2f cycles/step\n", (double)h / (256.0 * 64));
    smem_loop<4><<<1, 32>>>(out, cyc, 256, 64); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("smem queue loop, unroll 4       : %.2f cycles/step\n", (double)h / (256.0 * 64));
    smem_loop<8><<<1, 32>>>(out, cyc, 256, 64); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("smem queue loop, unroll 8       : %.2f cycles/step\n", (double)h / (256.0 * 64));
    smem_loop<16><<<1, 32>>>(out, cyc, 256, 64); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("smem queue loop, unroll 16      : %.2f cycles/step\n", (double)h / (256.0 * 64));
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
