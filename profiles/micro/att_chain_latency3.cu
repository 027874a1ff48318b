// Exact attenuation recurrence from a shared-memory queue: find the cheapest loop shape (B200, one warp).
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ double upd_fp(double att, double m, double inc, double dec, double tau) {
    const bool above = att > m;
    const bool rising = att < tau;
    const double s = att + inc, d = att - dec;
    const double r = above ? d : m;
    return (rising && !above) ? s : r;
}
__device__ __forceinline__ double upd_int(double att, double m, double inc, double dec, double tau) {
    const long long ia = __double_as_longlong(att);
    const bool above = ia > __double_as_longlong(m);
    const bool rising = ia < __double_as_longlong(tau);
    const double s = att + inc, d = att - dec;
    const double r = above ? d : m;
    return (rising && !above) ? s : r;
}

// MODE 0: operands straight from smem each step; 1: ping-pong register blocks of 8; 2: registers only (no smem)
template <int MODE, bool FP, bool ST>
__global__ void loop(double *out, long long *cyc, int total, int reps) {
    __shared__ double2 q01[272], q23[272];   // (m, tau), (inc, dec)
    __shared__ double qa[272];
    for (int i = threadIdx.x; i < 272; i += 32) {
        // alternate regimes like a tracking compressor: m jitters up and down around the attenuation
        const double m = 3.0 + ((i & 1) ? 0.004 : -0.004) + 1e-4 * (i % 7);
        q01[i] = make_double2(m, m - 0.0125);
        q23[i] = make_double2(0.0125, 0.00125);
    }
    __syncwarp();
    double att = 3.0;
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
        if (MODE == 0) {
#pragma unroll 8
            for (int j = 0; j < total; ++j) {
                const double2 a = q01[j], b = q23[j];
                att = FP ? upd_fp(att, a.x, b.x, b.y, a.y) : upd_int(att, a.x, b.x, b.y, a.y);
                if (ST) qa[j] = att;
            }
        } else if (MODE == 1) {
            double2 A0[8], A1[8], B0[8], B1[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) { A0[k] = q01[k]; A1[k] = q23[k]; }
            for (int j0 = 0; j0 < total; j0 += 16) {
#pragma unroll
                for (int k = 0; k < 8; ++k) { B0[k] = q01[j0 + 8 + k]; B1[k] = q23[j0 + 8 + k]; }
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    att = FP ? upd_fp(att, A0[k].x, A1[k].x, A1[k].y, A0[k].y) : upd_int(att, A0[k].x, A1[k].x, A1[k].y, A0[k].y);
                    if (ST) qa[j0 + k] = att;
                }
#pragma unroll
                for (int k = 0; k < 8; ++k) { A0[k] = q01[j0 + 16 + k]; A1[k] = q23[j0 + 16 + k]; }
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    att = FP ? upd_fp(att, B0[k].x, B1[k].x, B1[k].y, B0[k].y) : upd_int(att, B0[k].x, B1[k].x, B1[k].y, B0[k].y);
                    if (ST) qa[j0 + 8 + k] = att;
                }
            }
        } else {
            const double2 a = q01[r & 15], b = q23[r & 15];
#pragma unroll 16
            for (int j = 0; j < total; ++j) {
                const double m = (j & 1) ? a.x : a.x + 0.01;
                att = FP ? upd_fp(att, m, b.x, b.y, m - 0.0125) : upd_int(att, m, b.x, b.y, m - 0.0125);
            }
        }
        __syncwarp();
    }
    long long t1 = clock64();
    out[threadIdx.x] = att + qa[threadIdx.x];
    if (threadIdx.x == 0) *cyc = t1 - t0;
}

#define RUN(MODE, FP, ST, name) loop<MODE, FP, ST><<<1, 32>>>(out, cyc, 256, 64); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); \
    printf("%-60s %.2f cycles/step\n", name, (double)h / (256.0 * 64));
int main() {
    double *out; long long *cyc, h;
    cudaMalloc(&out, 4096); cudaMalloc(&cyc, 8);
    RUN(2, true, false, "registers only, FP compares")
    RUN(2, false, false, "registers only, int compares")
    RUN(0, true, false, "smem operands each step, FP compares")
    RUN(0, true, true, "smem operands each step, FP compares, store")
    RUN(0, false, true, "smem operands each step, int compares, store")
    RUN(1, true, false, "ping-pong register blocks, FP compares")
    RUN(1, true, true, "ping-pong register blocks, FP compares, store")
    RUN(1, false, true, "ping-pong register blocks, int compares, store")
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
