// Predicated-execution form of the attenuation recurrence vs select form (ping-pong register blocks from smem).
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ double upd_sel(double att, double m, double inc, double dec, double tau) {
    const long long ia = __double_as_longlong(att);
    const bool above = ia > __double_as_longlong(m);
    const bool rising = ia < __double_as_longlong(tau);
    const double s = att + inc, d = att - dec;
    const double r = above ? d : m;
    return (rising && !above) ? s : r;
}
// predicates from the OLD attenuation with FP compares, then three mutually exclusive predicated writes
__device__ __forceinline__ double upd_pred(double att, double m, double inc, double dec, double tau) {
    asm("{\n\t"
        ".reg .pred pa, pr, pc;\n\t"
        "setp.gt.f64 pa, %0, %1;\n\t"          // above
        "setp.lt.f64 pr, %0, %4;\n\t"          // rising (only meaningful when !above)
        "and.pred pr, pr, !pa;\n\t"
        "or.pred pc, pa, pr;\n\t"
        "@pa sub.f64 %0, %0, %3;\n\t"
        "@pr add.f64 %0, %0, %2;\n\t"
        "@!pc mov.f64 %0, %1;\n\t"
        "}"
        : "+d"(att) : "d"(m), "d"(inc), "d"(dec), "d"(tau));
    return att;
}
__device__ __forceinline__ double upd_pred_int(double att, double m, double inc, double dec, double tau) {
    asm("{\n\t"
        ".reg .pred pa, pr, pc;\n\t"
        ".reg .b64 ia, im, it;\n\t"
        "mov.b64 ia, %0; mov.b64 im, %1; mov.b64 it, %4;\n\t"
        "setp.gt.s64 pa, ia, im;\n\t"
        "setp.lt.s64 pr, ia, it;\n\t"
        "and.pred pr, pr, !pa;\n\t"
        "or.pred pc, pa, pr;\n\t"
        "@pa sub.f64 %0, %0, %3;\n\t"
        "@pr add.f64 %0, %0, %2;\n\t"
        "@!pc mov.f64 %0, %1;\n\t"
        "}"
        : "+d"(att) : "d"(m), "d"(inc), "d"(dec), "d"(tau));
    return att;
}

template <int V>
__global__ void loop(double *out, long long *cyc, int total, int reps) {
    __shared__ double2 q01[272], q23[272];   // (m, tau), (inc, dec)
    __shared__ double qa[272];
    for (int i = threadIdx.x; i < 272; i += 32) {
        const double m = 3.0 + ((i & 1) ? 0.004 : -0.004) + 1e-4 * (i % 7);
        q01[i] = make_double2(m, m - 0.0125);
        q23[i] = make_double2(0.0125, 0.00125);
    }
    __syncwarp();
    double att = 3.0;
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
        double2 A0[8], A1[8], B0[8], B1[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) { A0[k] = q01[k]; A1[k] = q23[k]; }
        for (int j0 = 0; j0 < total; j0 += 16) {
#pragma unroll
            for (int k = 0; k < 8; ++k) { B0[k] = q01[j0 + 8 + k]; B1[k] = q23[j0 + 8 + k]; }
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                att = V == 0 ? upd_sel(att, A0[k].x, A1[k].x, A1[k].y, A0[k].y)
                    : V == 1 ? upd_pred(att, A0[k].x, A1[k].x, A1[k].y, A0[k].y)
                             : upd_pred_int(att, A0[k].x, A1[k].x, A1[k].y, A0[k].y);
                qa[j0 + k] = att;
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) { A0[k] = q01[j0 + 16 + k]; A1[k] = q23[j0 + 16 + k]; }
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                att = V == 0 ? upd_sel(att, B0[k].x, B1[k].x, B1[k].y, B0[k].y)
                    : V == 1 ? upd_pred(att, B0[k].x, B1[k].x, B1[k].y, B0[k].y)
                             : upd_pred_int(att, B0[k].x, B1[k].x, B1[k].y, B0[k].y);
                qa[j0 + 8 + k] = att;
            }
        }
        __syncwarp();
    }
    long long t1 = clock64();
    out[threadIdx.x] = att + qa[threadIdx.x];
    if (threadIdx.x == 0) *cyc = t1 - t0;
}

#define RUN(V, name) loop<V><<<1, 32>>>(out, cyc, 256, 64); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); \
    cudaMemcpy(&res, out, 8, cudaMemcpyDeviceToHost); printf("%-50s %.2f cycles/step (result %.17g)\n", name, (double)h / (256.0 * 64), res);
int main() {
    double *out, res; long long *cyc, h;
    cudaMalloc(&out, 4096); cudaMalloc(&cyc, 8);
    RUN(0, "select form, integer compares (current kernel)")
    RUN(1, "predicated writes, FP compares")
    RUN(2, "predicated writes, integer compares")
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
