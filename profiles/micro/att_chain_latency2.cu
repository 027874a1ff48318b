// Variants of the k_att_chain block loop, cycles per dependent step on B200 (one warp).
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ double upd_tau(double att, double m, double inc, double dec, double tau) {
    const long long ia = __double_as_longlong(att);
    const bool above = ia > __double_as_longlong(m);
    const bool rising = ia < __double_as_longlong(tau);
    const double s = att + inc, d = att - dec;
    const double r = above ? d : m;
    return (rising && !above) ? s : r;
}
__device__ __forceinline__ double upd_tau_fp(double att, double m, double inc, double dec, double tau) {
    const bool above = att > m;
    const bool rising = att < tau;
    const double s = att + inc, d = att - dec;
    const double r = above ? d : m;
    return (rising && !above) ? s : r;
}
// single add: pick the addend first, then one DADD, then clamp-select with m
__device__ __forceinline__ double upd_one_add(double att, double m, double inc, double dec, double tau) {
    const long long ia = __double_as_longlong(att);
    const bool above = ia > __double_as_longlong(m);
    const bool rising = ia < __double_as_longlong(tau);
    const double delta = above ? -dec : inc;
    const double t = att + delta;
    return (above || rising) ? t : m;
}

template <int V, bool GE, bool ST>
__global__ void loop(double *out, long long *cyc, int total, int reps, int before_grp) {
    __shared__ double2 q01[264], q23[264];
    __shared__ double qa[264];
    for (int i = threadIdx.x; i < 264; i += 32) {
        q01[i] = make_double2(3.0 + 1e-3 * i, 0.0125);
        q23[i] = make_double2(0.00125, 3.0 + 1e-3 * i - 0.0125);
    }
    __syncwarp();
    double att = 0.0, ge = 0.0;
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
        double2 a01[8], a23[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) { a01[k] = q01[k]; a23[k] = q23[k]; }
        for (int j0 = 0; j0 < total; j0 += 8) {
            double2 n01[8], n23[8];
            if (j0 + 8 < total) {
#pragma unroll
                for (int k = 0; k < 8; ++k) { n01[k] = q01[j0 + 8 + k]; n23[k] = q23[j0 + 8 + k]; }
            }
            double res[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                if (V == 0) att = upd_tau(att, a01[k].x, a01[k].y, a23[k].x, a23[k].y);
                if (V == 1) att = upd_tau_fp(att, a01[k].x, a01[k].y, a23[k].x, a23[k].y);
                if (V == 2) att = upd_one_add(att, a01[k].x, a01[k].y, a23[k].x, a23[k].y);
                if (GE) { if (before_grp == j0 + k + 1) ge = att; }
                res[k] = att;
            }
            if (ST) {
#pragma unroll
                for (int k = 0; k < 8; ++k) qa[j0 + k] = res[k];
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) { a01[k] = n01[k]; a23[k] = n23[k]; }
        }
        __syncwarp();
    }
    long long t1 = clock64();
    out[threadIdx.x] = att + ge + qa[threadIdx.x];
    if (threadIdx.x == 0) *cyc = t1 - t0;
}

#define RUN(V, GE, ST, name) loop<V, GE, ST><<<1, 32>>>(out, cyc, 256, 64, 77); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); \
    printf("%-52s %.2f cycles/step\n", name, (double)h / (256.0 * 64));
int main() {
    double *out; long long *cyc, h;
    cudaMalloc(&out, 4096); cudaMalloc(&cyc, 8);
    RUN(0, true, true, "int compares, ge capture, stores (kernel as is)")
    RUN(0, false, true, "int compares, stores")
    RUN(0, false, false, "int compares only")
    RUN(1, false, false, "DSETP compares only")
    RUN(1, true, true, "DSETP compares, ge, stores")
    RUN(2, false, false, "one-add, int compares only")
    RUN(2, true, true, "one-add, int compares, ge, stores")
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
