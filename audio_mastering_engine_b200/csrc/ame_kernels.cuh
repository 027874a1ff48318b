// Device kernels of libame (sm_100a).  See DESIGN.md for the data layout and the per-kernel rooflines.
//
// Time parallelism: every 30 s chunk of the reference restarts all filter state from zero
// (audio_mastering_engine.py:185-199), and inside a chunk the filters are stable LTI systems, so a
// tile of T frames is computed by ONE lane pair (L lane, R lane) that first runs `warm` frames of the
// preceding audio from zero state.  `warm` is chosen on the host so that the state error has decayed
// below 1e-13 of the signal level (FP64 rounding level) when the tile proper starts; the first tile
// of a chunk needs no warm-up and is exact by construction.
//
// All 32 lanes of a warp run the SAME number of 4-frame groups (the warp maximum), so every shuffle is a
// plain full-mask SHFL; lanes past their own range keep computing on stale input and simply do not store.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/ame.h"

namespace ame {

constexpr int kGainTile = 32768;   // frames per CTA in k_apply_gain / k_band_sum
constexpr unsigned kFull = 0xffffffffu;

struct TileJob {           // one lane pair of k_eq / k_band_split
    int64_t chunk_begin;   // absolute frame index (packed buffer) where filter state is reset
    int64_t tile_begin;    // first frame this job writes (absolute, multiple of 4 unless == chunk_begin)
    int64_t tile_end;      // one past the last frame it writes
    int32_t track;
    int32_t variant;       // EQ stage mask (bit s = stage s active)
};

struct ChainJob {          // one warp of k_compress: one band of one chunk of a multiband track
    int64_t mb_begin;      // of the chunk, in the multiband-only packing (bands planes)
    int64_t n;             // frames
    int64_t ck_begin;      // first slot of this chain in the per-group checkpoint array
    int32_t band;
    int32_t table;
    uint32_t thr_i;        // rms > thresh_rms  <=>  rms >= thr_i
    int32_t look;          // look_frames
};

struct KwJob { int32_t track; int32_t sb_begin; int32_t sb_end; int32_t pad; };

struct GainJob { int64_t begin; int64_t end; int32_t track; int32_t pad; };

// indexed by integer rms 0..32768.  tau = the smallest attenuation a with fl(a + inc) >= m, so that
// (att + inc < m) <=> (att < tau) exactly and the branch predicates depend on the OLD attenuation only.
struct AttEntry { double m, inc, dec, tau; };

struct TrackDev {          // device-side per-track bookkeeping
    int64_t sb_offset;     // start of this track's 100 ms energies in the energy array
    int32_t n_sb;          // number of complete 100 ms sub-blocks (halo included)
    int32_t s100;          // frames per 100 ms = (fs + 5) / 10   (ebur128.c)
    int32_t first_block;   // time shards: 400 ms blocks before this one belong to the previous shard / the warm-up
    int32_t pad;
    int64_t n_total;       // halo + span frames
};

__constant__ double c_hist_bounds[1001];
__constant__ double c_hist_energy[1000];

// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 ldg16(const uint4 *p) { return __ldg(p); }

__device__ __forceinline__ double bq_step(const ame_biquad &c, double &z0, double &z1, double x) {
    // DF-II transposed, the form scipy evaluates (lfilter / sosfilt) - FMA-contracted.
    double y = fma(c.b0, x, z0);
    z0 = fma(-c.a1, y, fma(c.b1, x, z1));
    z1 = fma(-c.a2, y, c.b2 * x);
    return y;
}

// exact int16 -> double without the conversion pipe: 2^52 + 2^31 + x as raw bits, minus the bias
__device__ __forceinline__ double i16_to_f64(int x) {
    return __hiloint2double(0x43300000, (int)(0x80000000u ^ (unsigned)x)) - 4503601774854144.0;
}

// np.clip(x,-1,1) * 32767 -> astype(int16)  (truncate toward zero), float64 flavour
// (clip first or clamp the truncated product: same integer - the product is monotone in v, +-32767 are exact, and
// cvt.rzi saturates; the integer clamp avoids two DSETP-based double min/max)
__device__ __forceinline__ int to_pcm_f64(double v) {
    return min(max(__double2int_rz(__dmul_rn(v, 32767.0)), -32767), 32767);
}
__device__ __forceinline__ int to_pcm_f32(float v) {
    v = fminf(fmaxf(v, -1.0f), 1.0f);
    return __float2int_rz(__fmul_rn(v, 32767.0f));
}
__device__ __forceinline__ int sat16(int v) { return v > 32767 ? 32767 : (v < -32768 ? -32768 : v); }
__device__ __forceinline__ uint32_t pack16(int lo, int hi) { return (uint32_t)(uint16_t)lo | ((uint32_t)(uint16_t)hi << 16); }

// Butterworth second-order section, numerator b0 * (1 + 2S z^-1 + z^-2) with S = +1 (zeros at z = -1:
// low-pass type) or S = -1 (zeros at z = +1: high-pass type), a0 = 1.  Every section scipy.signal.butter
// returns for the reference's shelves / band-passes / crossovers has this shape (validated on the host),
// which saves two coefficient registers per section.  DF-II transposed as scipy evaluates it:
//   y = b0 x + z0 ; z0 = (z1 + b1 x) - a1 y ; z1 = b2 x - a2 y        with b1 x = 2S (b0 x) exactly.
template <int S>
__device__ __forceinline__ double bw_step(double b0, double a1, double a2, double &z0, double &z1, double x) {
    const double y = fma(b0, x, z0);
    const double t = b0 * x;
    z0 = fma(-a1, y, fma(2.0 * S, t, z1));
    z1 = fma(-a2, y, t);
    return y;
}
template <int S>   // b0 == 1
__device__ __forceinline__ double bw_step1(double a1, double a2, double &z0, double &z1, double x) {
    const double y = x + z0;
    z0 = fma(-a1, y, fma(2.0 * S, x, z1));
    z1 = fma(-a2, y, x);
    return y;
}

struct PeakCoef { double b0, a1[4], a2[4]; };     // butter(4, bandpass, sos): signs (+,+,-,-), sections 1..3 unit gain
__device__ __forceinline__ void load_peak(PeakCoef &c, const ame_eq_stage &st) {
    c.b0 = st.s[0].b0;
#pragma unroll
    for (int i = 0; i < 4; ++i) { c.a1[i] = st.s[i].a1; c.a2[i] = st.s[i].a2; }
}
__device__ __forceinline__ double peak_step(const PeakCoef &c, double *z, double x) {
    double t = bw_step<1>(c.b0, c.a1[0], c.a2[0], z[0], z[1], x);
    t = bw_step1<1>(c.a1[1], c.a2[1], z[2], z[3], t);
    t = bw_step1<-1>(c.a1[2], c.a2[2], z[4], z[5], t);
    return bw_step1<-1>(c.a1[3], c.a2[3], z[6], z[7], t);
}

// exact int16 -> x / 32768 as double in ONE add: bits of 2^37 + (x + 2^31) * 2^-15, minus 2^37 + 2^16
__device__ __forceinline__ double i16_to_unit(int x) {
    return __hiloint2double(0x42400000, (int)(0x80000000u ^ (unsigned)x)) - 137439019008.0;
}

// ------------------------------------------------------------------------------------------------
// k_eq: int16 in -> [warmth -> int16] -> float32 -> 4-stage EQ in FP64 -> float32 -> [width] -> int16
// ONE THREAD per tile, both channels: the L and R cascades are two independent dependency chains in one
// instruction stream (the kernel is bound by FP64 latency, not by registers), and the cross-channel
// stages (warmth, width, packing) need no shuffles.  MASK = active EQ stages, WARM = warmth on.
// ------------------------------------------------------------------------------------------------
template <int MASK, bool WARM>
__device__ __forceinline__ void eq_tile(const TileJob &job, const ame_track_params *__restrict__ tp,
                                        const double *__restrict__ luts, const int16_t *__restrict__ in,
                                        int16_t *__restrict__ pre) {
    const bool widen = (tp->flags & AME_F_WIDTH) != 0;
    const float wfac = tp->width;
    const double *lut = WARM ? luts + (size_t)tp->warm_lut * 65536 + 32768 : nullptr;
    double wl_b0 = 0, wl_b1 = 0, wl_a1 = 0, wl_gm1 = 0, wh_b0 = 0, wh_b1 = 0, wh_a1 = 0, wh_gm1 = 0;
    if (WARM) {
        wl_b0 = tp->wl_b0; wl_b1 = tp->wl_b1; wl_a1 = tp->wl_a1; wl_gm1 = tp->wl_gm1;
        wh_b0 = tp->wh_b0; wh_b1 = tp->wh_b1; wh_a1 = tp->wh_a1; wh_gm1 = tp->wh_gm1;
    }
    double s0_b0 = 0, s0_a1 = 0, s0_a2 = 0, g0 = 0, gm0 = 0, s3_b0 = 0, s3_a1 = 0, s3_a2 = 0, g3 = 0, gm3 = 0, gm1 = 0, gm2 = 0;
    bool boost0 = false, boost3 = false;
    PeakCoef p1, p2;
    if (MASK & 1) {
        s0_b0 = tp->eq[0].s[0].b0; s0_a1 = tp->eq[0].s[0].a1; s0_a2 = tp->eq[0].s[0].a2;
        g0 = tp->eq[0].g; gm0 = tp->eq[0].gm1; boost0 = tp->eq[0].kind == AME_EQ_SHELF_BOOST;
    }
    if (MASK & 2) { load_peak(p1, tp->eq[1]); gm1 = tp->eq[1].gm1; }
    if (MASK & 4) { load_peak(p2, tp->eq[2]); gm2 = tp->eq[2].gm1; }
    if (MASK & 8) {
        s3_b0 = tp->eq[3].s[0].b0; s3_a1 = tp->eq[3].s[0].a1; s3_a2 = tp->eq[3].s[0].a2;
        g3 = tp->eq[3].g; gm3 = tp->eq[3].gm1; boost3 = tp->eq[3].kind == AME_EQ_SHELF_BOOST;
    }
    double zl[20], zr[20];
#pragma unroll
    for (int i = 0; i < 20; ++i) { zl[i] = 0.0; zr[i] = 0.0; }

    const int64_t warm = (MASK != 0) ? (int64_t)tp->warm_eq : 0;
    int64_t f_lo = job.tile_begin - warm;
    if (f_lo < job.chunk_begin) f_lo = job.chunk_begin;
    const int64_t f_hi = job.tile_end;
    if (f_hi <= f_lo) return;
    const int64_t g0f = f_lo & ~(int64_t)3;               // first 4-aligned group
    const int n_it = (int)((f_hi - g0f + 3) >> 2);

    auto cascade = [&](double v, double *z) -> float {     // one channel through the 4 EQ stages
        if (MASK & 1) {   // apply_shelf_filter 250 Hz low (:283-289)
            const double f = bw_step<1>(s0_b0, s0_a1, s0_a2, z[0], z[1], v);
            v = boost0 ? v + (f - v) * gm0 : v * g0 + (f - v * g0);
        }
        if (MASK & 2) v = v + peak_step(p1, z + 2, v) * gm1;    // apply_peak_filter 1 kHz (:290-298)
        if (MASK & 4) v = v + peak_step(p2, z + 10, v) * gm2;   // apply_peak_filter 4 kHz
        if (MASK & 8) {   // apply_shelf_filter 8 kHz high
            const double f = bw_step<-1>(s3_b0, s3_a1, s3_a2, z[18], z[19], v);
            v = boost3 ? v + (f - v) * gm3 : v * g3 + (f - v * g3);
        }
        return __double2float_rn(v);                       // samples[:, i] = ... into the float32 array (:274)
    };

    // one frame -> packed (L | R << 16) int16 output.  lutL / lutR = tanh table values (fetched a group ahead).
    auto frame = [&](uint32_t w, double lutL, double lutR) -> uint32_t {
        int xl = (int)(int16_t)(w & 0xffffu), xr = (int)(int16_t)(w >> 16);
        if (WARM) {
            // apply_analog_character (:258-266): tanh in float32 (table = the host's own np.tanh, widened
            // exactly to double), then two order-2 "shelves" that lfilter(axis=-1) runs ACROSS the channels:
            //   y0 = b0*L ; y1 = (b1*L - a1*y0) + b0*R ; blend x + (y - x)*(g - 1)       (no FMA contraction)
            double L = lutL, R = lutR;
            double y0 = __dmul_rn(wl_b0, L);
            double y1 = __dadd_rn(__dsub_rn(__dmul_rn(wl_b1, L), __dmul_rn(wl_a1, y0)), __dmul_rn(wl_b0, R));
            const double L1 = __dadd_rn(L, __dmul_rn(__dsub_rn(y0, L), wl_gm1));
            const double R1 = __dadd_rn(R, __dmul_rn(__dsub_rn(y1, R), wl_gm1));
            y0 = __dmul_rn(wh_b0, L1);
            y1 = __dadd_rn(__dsub_rn(__dmul_rn(wh_b1, L1), __dmul_rn(wh_a1, y0)), __dmul_rn(wh_b0, R1));
            L = __dadd_rn(L1, __dmul_rn(__dsub_rn(y0, L1), wh_gm1));
            R = __dadd_rn(R1, __dmul_rn(__dsub_rn(y1, R1), wh_gm1));
            xl = to_pcm_f64(L);                            // float_array_to_audio_segment (:254-257)
            xr = to_pcm_f64(R);
        }
        // audio_segment_to_float_array (:250-253): x / 32768 is exact in float32 and in float64
        float yl, yr;
        if (MASK != 0) {
            yl = cascade(i16_to_unit(xl), zl);
            yr = cascade(i16_to_unit(xr), zr);
        } else {
            yl = __fmul_rn((float)xl, 1.0f / 32768.0f);
            yr = __fmul_rn((float)xr, 1.0f / 32768.0f);
        }
        if (widen) {          // apply_stereo_width (:267-271), float32 arithmetic
            const float mid = __fmul_rn(__fadd_rn(yl, yr), 0.5f);
            const float side = __fmul_rn(__fmul_rn(__fsub_rn(yl, yr), 0.5f), wfac);
            yl = __fadd_rn(mid, side);
            yr = __fsub_rn(mid, side);
        }
        return pack16(to_pcm_f32(yl), to_pcm_f32(yr));     // clip inside to_pcm == np.clip of (:270) then (:255)
    };

    const uint4 *src = reinterpret_cast<const uint4 *>(in) + (g0f >> 2);
    uint4 *dst = reinterpret_cast<uint4 *>(pre) + (g0f >> 2);
    // software pipeline: input words two groups ahead, tanh-table values one group ahead
    uint4 cur = ldg16(src), nxt = make_uint4(0, 0, 0, 0);
    if (g0f + 0 < job.chunk_begin) cur.x = 0;              // frames before the chunk start keep the zero state
    if (g0f + 1 < job.chunk_begin) cur.y = 0;
    if (g0f + 2 < job.chunk_begin) cur.z = 0;
    if (n_it > 1) nxt = ldg16(src + 1);
    double lutL[4] = {0, 0, 0, 0}, lutR[4] = {0, 0, 0, 0};
    auto fetch_lut = [&](const uint4 &q, double *l, double *r) {
        const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            l[k] = __ldg(lut + (int)(int16_t)(w[k] & 0xffffu));
            r[k] = __ldg(lut + (int)(int16_t)(w[k] >> 16));
        }
    };
    if (WARM) fetch_lut(cur, lutL, lutR);
    for (int it = 0; it < n_it; ++it) {
        uint4 nn = nxt;
        if (it + 2 < n_it) nn = ldg16(src + it + 2);
        double nL[4] = {0, 0, 0, 0}, nR[4] = {0, 0, 0, 0};
        if (WARM) fetch_lut(nxt, nL, nR);
        const int64_t g = g0f + 4 * (int64_t)it;
        const uint32_t w[4] = {cur.x, cur.y, cur.z, cur.w};
        uint32_t o[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) o[k] = frame(w[k], lutL[k], lutR[k]);
        if (g >= job.tile_begin && g + 4 <= f_hi) {
            dst[it] = make_uint4(o[0], o[1], o[2], o[3]);
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (g + k >= job.tile_begin && g + k < f_hi) reinterpret_cast<uint32_t *>(dst + it)[k] = o[k];
        }
        cur = nxt; nxt = nn;
#pragma unroll
        for (int k = 0; k < 4; ++k) { lutL[k] = nL[k]; lutR[k] = nR[k]; }
    }
}

__global__ void __launch_bounds__(128, 2)
k_eq(const TileJob *__restrict__ jobs, int n_jobs, const ame_track_params *__restrict__ tracks,
     const double *__restrict__ luts, const int16_t *__restrict__ in, int16_t *__restrict__ pre) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_jobs) return;
    const TileJob job = jobs[j];
    const ame_track_params *tp = tracks + job.track;
    switch (job.variant) {      // bits 0-3: EQ stages, bit 4: warmth
#define AME_EQ_CASE(M) case M: eq_tile<M, false>(job, tp, luts, in, pre); break; \
                       case M + 16: eq_tile<M, true>(job, tp, luts, in, pre); break;
        AME_EQ_CASE(0) AME_EQ_CASE(1) AME_EQ_CASE(2) AME_EQ_CASE(3) AME_EQ_CASE(4) AME_EQ_CASE(5)
        AME_EQ_CASE(6) AME_EQ_CASE(7) AME_EQ_CASE(8) AME_EQ_CASE(9) AME_EQ_CASE(10) AME_EQ_CASE(11)
        AME_EQ_CASE(12) AME_EQ_CASE(13) AME_EQ_CASE(14) AME_EQ_CASE(15)
#undef AME_EQ_CASE
    }
}

// ------------------------------------------------------------------------------------------------
// k_band_split: int16 pre -> Butterworth-4 LP 250 / HP 4k in FP64, mid = x - low - high, each band
// truncated to int16 (apply_multiband_compressor :300-305).  One thread per tile, both channels, same
// warm-up scheme.  bands = 3 planes of mb_frames frames each.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 3)
k_band_split(const TileJob *__restrict__ jobs, int n_jobs, const ame_track_params *__restrict__ tracks,
             const int64_t *__restrict__ mb_delta,   // per track: mb_offset - offset_frames
             const int16_t *__restrict__ pre, int16_t *__restrict__ bands, int64_t mb_frames) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_jobs) return;
    const TileJob job = jobs[j];
    const ame_track_params *tp = tracks + job.track;
    const double lb0 = tp->xlp[0].b0, la10 = tp->xlp[0].a1, la20 = tp->xlp[0].a2, la11 = tp->xlp[1].a1, la21 = tp->xlp[1].a2;
    const double hb0 = tp->xhp[0].b0, ha10 = tp->xhp[0].a1, ha20 = tp->xhp[0].a2, ha11 = tp->xhp[1].a1, ha21 = tp->xhp[1].a2;
    const int64_t delta = mb_delta[job.track];
    double zl[8], zr[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { zl[i] = 0.0; zr[i] = 0.0; }
    int64_t f_lo = job.tile_begin - (int64_t)tp->warm_xover;
    if (f_lo < job.chunk_begin) f_lo = job.chunk_begin;
    const int64_t f_hi = job.tile_end;
    if (f_hi <= f_lo) return;
    const int64_t g0f = f_lo & ~(int64_t)7;                // 8 frames = one 32-byte sector per iteration
    const int n_it = (int)((f_hi - g0f + 7) >> 3);
    uint32_t *b0 = reinterpret_cast<uint32_t *>(bands) + delta;
    uint32_t *b1 = b0 + mb_frames;
    uint32_t *b2 = b1 + mb_frames;

    auto split = [&](int xm, double *z, int &p0, int &p1, int &p2) {
        const double x = i16_to_unit(xm);
        const double lo = bw_step1<1>(la11, la21, z[2], z[3], bw_step<1>(lb0, la10, la20, z[0], z[1], x));
        const double hi = bw_step1<-1>(ha11, ha21, z[6], z[7], bw_step<-1>(hb0, ha10, ha20, z[4], z[5], x));
        const double mid = __dsub_rn(__dsub_rn(x, lo), hi);
        p0 = to_pcm_f64(lo); p1 = to_pcm_f64(mid); p2 = to_pcm_f64(hi);
    };

    const uint4 *src = reinterpret_cast<const uint4 *>(pre) + (g0f >> 2);
    uint4 c0 = ldg16(src), c1 = ldg16(src + 1), n0 = make_uint4(0, 0, 0, 0), n1 = n0;
    {   // keep the zero state until the chunk starts
        uint32_t *cw0 = reinterpret_cast<uint32_t *>(&c0), *cw1 = reinterpret_cast<uint32_t *>(&c1);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (g0f + k < job.chunk_begin) cw0[k] = 0;
            if (g0f + 4 + k < job.chunk_begin) cw1[k] = 0;
        }
    }
    if (n_it > 1) { n0 = ldg16(src + 2); n1 = ldg16(src + 3); }
    for (int it = 0; it < n_it; ++it) {
        uint4 m0 = n0, m1 = n1;
        if (it + 2 < n_it) { m0 = ldg16(src + 2 * it + 4); m1 = ldg16(src + 2 * it + 5); }
        const int64_t g = g0f + 8 * (int64_t)it;
        const uint32_t w[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
        uint32_t o0[8], o1[8], o2[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            int l0, l1, l2, r0, r1, r2;
            split((int)(int16_t)(w[k] & 0xffffu), zl, l0, l1, l2);
            split((int)(int16_t)(w[k] >> 16), zr, r0, r1, r2);
            o0[k] = pack16(l0, r0); o1[k] = pack16(l1, r1); o2[k] = pack16(l2, r2);
        }
        if (g >= job.tile_begin && g + 8 <= f_hi) {
            uint4 *q0 = reinterpret_cast<uint4 *>(b0 + g), *q1 = reinterpret_cast<uint4 *>(b1 + g), *q2 = reinterpret_cast<uint4 *>(b2 + g);
            q0[0] = make_uint4(o0[0], o0[1], o0[2], o0[3]); q0[1] = make_uint4(o0[4], o0[5], o0[6], o0[7]);
            q1[0] = make_uint4(o1[0], o1[1], o1[2], o1[3]); q1[1] = make_uint4(o1[4], o1[5], o1[6], o1[7]);
            q2[0] = make_uint4(o2[0], o2[1], o2[2], o2[3]); q2[1] = make_uint4(o2[4], o2[5], o2[6], o2[7]);
        } else {
#pragma unroll
            for (int k = 0; k < 8; ++k)
                if (g + k >= job.tile_begin && g + k < f_hi) { b0[g + k] = o0[k]; b1[g + k] = o1[k]; b2[g + k] = o2[k]; }
        }
        c0 = n0; c1 = n1; n0 = m0; n1 = m1;
    }
}

// ------------------------------------------------------------------------------------------------
// Multiband compressor = pydub compress_dynamic_range per band (:306-308), split by what is sequential:
//   k_window_flag   (time-parallel)  window rms of the previous look_frames frames; emits the integer rms
//                                    for frames ABOVE threshold and 0 otherwise (2 B per band frame)
//   k_att_chain_spec (speculative)   the attenuation recurrence over the flagged frames of one (chunk, band), cut
//                                    into 64..256 time segments that are walked in parallel from a guessed start
//                                    and repaired until every segment starts from its predecessor's true end
//                                    (exact); emits the attenuation after every flagged frame and entering each
//                                    32-frame group.  Below threshold the reference never releases
//                                    (max_attenuation = 0 => dec = 0), so unflagged frames are no-ops.
//   k_att_chain      (sequential)    the same recurrence by one producer + one consumer warp per chain: the
//                                    fallback for chains whose segments cannot be repaired cheaply, and the
//                                    whole stage with chain_warps = -1.
//   k_compress_apply (time-parallel) attenuation in force at a frame = the value stored for the last flagged frame
//                                    at or before it in its 32-frame group, else the group's entry value;
//                                    gain = 10^(-att/20), audioop.mul, low.overlay(mid).overlay(high) (:309).
// pydub: rms_at(i) = audioop.rms(frames [max(i-look,0), i)) = (unsigned)sqrt(S / n) with S the exact integer
// sum of squares and n = 2 * frames.  rms > thresh  <=>  rms >= thr_i  <=>  S >= thr_i^2 * n (integers), so
// only flagged frames take the square root (S/n is never within 2^-41 of a perfect square unless equal,
// hence the double-precision expression of audioop truncates to the exact integer root).
// ------------------------------------------------------------------------------------------------
constexpr int kWfThreads = 256;
constexpr int kWfTile = 2048;      // frames per CTA of k_window_flag (8 per thread)
constexpr int kSeg = 256;          // frames per k_att_chain iteration / per k_compress_apply warp (8 groups)

struct WfJob { int32_t chain; int32_t pad; int64_t tile_begin; };   // tile_begin relative to the chunk

struct MbChunk {           // one chunk of a multiband track (k_compress_apply)
    int64_t abs_begin;     // absolute frame index in the packed pre buffer
    int64_t mb_begin;      // frame index in the multiband-only packing
    int64_t n;             // frames
    int64_t seg_prefix;    // kSeg-segments in all earlier chunks
    int64_t ck_begin[3];   // first group slot of each band's chain in the checkpoint array
    int32_t track;
    int32_t pad;
};

__device__ __forceinline__ unsigned energy_of(uint32_t w) {          // l^2 + r^2 <= 2^31
    const int l = (int16_t)(w & 0xffffu), r = (int16_t)(w >> 16);
    return (unsigned)(l * l) + (unsigned)(r * r);
}

__device__ __forceinline__ void load8(const uint32_t *__restrict__ p, int64_t idx, int64_t lo, int64_t hi, uint32_t *w) {
    // 8 consecutive frames p[idx .. idx+8) with frames outside [lo, hi) read as zero
    if (idx >= lo && idx + 8 <= hi && ((reinterpret_cast<uintptr_t>(p + idx) & 15) == 0)) {
        const uint4 a = __ldg(reinterpret_cast<const uint4 *>(p + idx));
        const uint4 b = __ldg(reinterpret_cast<const uint4 *>(p + idx) + 1);
        w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w; w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
    } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) w[k] = (idx + k >= lo && idx + k < hi) ? __ldg(p + idx + k) : 0u;
    }
}

// floor(sqrt(s / n)) in exact integers, for s / n < 2^31 and n < 2^11: audioop.rms is (unsigned)sqrt(S / n) in double,
// which truncates to exactly this (header comment above) - a float estimate is within 1 of it, two integer tests
// settle it, and the double division + square root (~55 instructions on the FP64 pipe per flagged frame) go away.
__device__ __forceinline__ unsigned isqrt_ratio(unsigned long long s, unsigned n) {
    unsigned r = (unsigned)__fsqrt_rn(__fdividef((float)s, (float)n));
    if ((unsigned long long)r * r * n > s) --r;
    else if ((unsigned long long)(r + 1) * (r + 1) * n <= s) ++r;
    return r;
}

__global__ void __launch_bounds__(kWfThreads)
k_window_flag(const WfJob *__restrict__ jobs, const ChainJob *__restrict__ chains, const int16_t *__restrict__ bands,
              uint16_t *__restrict__ rms, int64_t mb_frames) {
    __shared__ long long s_scan[kWfThreads / 32];
    __shared__ unsigned long long s_head[kWfThreads / 32];
    const WfJob job = jobs[blockIdx.x];
    const ChainJob cj = chains[job.chain];
    const uint32_t *bp = reinterpret_cast<const uint32_t *>(bands) + (int64_t)cj.band * mb_frames + cj.mb_begin;
    uint16_t *rp = rms + (int64_t)cj.band * mb_frames + cj.mb_begin;
    const int64_t n = cj.n, t0 = job.tile_begin;
    const int look = cj.look;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int64_t i0 = t0 + (int64_t)threadIdx.x * 8;
    // all loads first (the kernel is bound by memory latency): this thread's 8 frames, the 8 frames `look` earlier,
    // and its share of the window entering the tile
    uint32_t w[8], wo[8];
    load8(bp, i0, 0, n, w);
    load8(bp, i0 - look, 0, n, wo);
    // window sum entering the tile: frames [t0 - look, t0)
    unsigned long long head = 0;
    for (int64_t j = t0 - look + threadIdx.x; j < t0; j += kWfThreads)
        if (j >= 0) head += energy_of(__ldg(bp + j));
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) head += __shfl_xor_sync(kFull, head, d);
    if (lane == 0) s_head[wid] = head;
    // D_j = e_j - e_{j-look}; S_i = head + sum_{t0 <= j < i} D_j
    long long pre[8];          // exclusive prefix of D inside the thread
    long long run = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        pre[k] = run;
        run += (long long)energy_of(w[k]) - (long long)energy_of(wo[k]);
    }
    long long incl = run;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const long long v = __shfl_up_sync(kFull, incl, d);
        if (lane >= d) incl += v;
    }
    if (lane == 31) s_scan[wid] = incl;
    __syncthreads();
    long long base = incl - run;
    unsigned long long h = 0;
#pragma unroll
    for (int q = 0; q < kWfThreads / 32; ++q) {
        if (q < wid) base += s_scan[q];
        h += s_head[q];
    }
    base += (long long)h;
    const unsigned long long thr2 = (unsigned long long)cj.thr_i * cj.thr_i;
    const bool never = cj.thr_i > 32768u;          // rms <= 32768 can never exceed it
    uint32_t o[4] = {0, 0, 0, 0};
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int64_t i = i0 + k;
        const long long s = base + pre[k];
        const int64_t nfr = i < look ? i : look;
        unsigned r = 0;
        if (!never && i < n && nfr > 0 && (unsigned long long)s >= thr2 * (unsigned long long)(2 * nfr))
            r = isqrt_ratio((unsigned long long)s, (unsigned)(2 * nfr));
        o[k >> 1] |= (r & 0xffffu) << ((k & 1) * 16);
    }
    if (i0 + 8 <= n && ((reinterpret_cast<uintptr_t>(rp + i0) & 15) == 0)) {
        *reinterpret_cast<uint4 *>(rp + i0) = make_uint4(o[0], o[1], o[2], o[3]);
    } else {
#pragma unroll
        for (int k = 0; k < 8; ++k)
            if (i0 + k < n) rp[i0 + k] = (uint16_t)(o[k >> 1] >> ((k & 1) * 16));
    }
}

// att' = (att <= M) ? min(att + inc, M) : max(att - dec, 0), evaluated as
//   att > M ? att - dec : (att < tau ? att + inc : M)
// * max(., 0) never binds: att - dec > M - M/release_frames >= 0.
// * tau is precomputed per table entry (host), so BOTH predicates are functions of the old attenuation and
//   evaluate in parallel with the two DADDs; the loop-carried path is max(DADD, compare) + select, ~18 cycles
//   on B200 against ~36 for DADD -> fmin (DSETP + FSEL) -> select (profiles/micro/att_chain_latency.cu).
// * all operands are non-negative doubles, for which integer order of the bit patterns == numeric order.
__device__ __forceinline__ double att_update(double att, double m, double inc, double dec, double tau) {
    const long long ia = __double_as_longlong(att);
    const bool above = ia > __double_as_longlong(m);
    const bool rising = ia < __double_as_longlong(tau);
    const double s = att + inc;
    const double d = att - dec;
    const double r = above ? d : m;
    return (rising && !above) ? s : r;
}

constexpr int kChainStages = 2;    // segment buffers in flight between the producer and the consumer warp

__device__ __forceinline__ void named_sync(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }
__device__ __forceinline__ void named_arrive(int id) { asm volatile("bar.arrive %0, 64;" ::"r"(id) : "memory"); }

// k_att_chain: the strictly sequential part.  Two warps per (chunk, band), four chains per CTA:
//   a PRODUCER warp walks the 256-frame segments of the rms plane (lane = 8 consecutive frames), compacts the
//     table entries (M, tau, inc, dec) of the flagged frames into a shared-memory queue in time order, padded to a
//     multiple of 16 with no-op entries, and later scatters the results (attenuation after every flagged frame ->
//     att_f, attenuation entering every 32-frame group -> ckpt);
//   a CONSUMER warp does nothing but the recurrence over those queues: 16 steps per iteration from two
//     ping-pong register blocks so the shared-memory loads of the next 8 steps are in flight while the current 8
//     run - 25 cycles per dependent step on B200 (profiles/micro/att_chain_latency3.cu), and segments without
//     flagged frames cost it one barrier.
// The two warps are decoupled by kChainStages buffers and named barriers (full[s]: producer arrives / consumer
// waits; done[s]: consumer arrives / producer waits).
constexpr int kChainsPerCta = 4;   // consumers = warps 0..3, producers = warps 4..7: one of each per SM sub-partition
struct ChainSmem {                 // per chain
    double2 mt[kChainStages][kSeg + 32];     // (M, tau)   (+16 no-op entries, +8 read-ahead slack)
    double2 id[kChainStages][kSeg + 32];     // (inc, dec)
    double att[kChainStages][kSeg + 16];     // attenuation after each flagged frame
    double att_in[kChainStages];             // attenuation entering the segment
    int total[kChainStages];
    uint16_t off[kChainStages][32];          // producer bookkeeping for the scatter
    uint8_t m8[kChainStages][32];
};

// One chain on two warps (consumer: producer == false).  `sm` is that pair's queue, bar_base its first named barrier.
__device__ __forceinline__ void chain_queue_run(const ChainJob &job, ChainSmem &sm, bool producer, int lane, int bar_base,
                                                const uint16_t *__restrict__ rms, const AttEntry *__restrict__ tables,
                                                double *__restrict__ ckpt, double *__restrict__ att_f, int64_t mb_frames) {
    auto &s_mt = sm.mt; auto &s_id = sm.id; auto &s_att = sm.att; auto &s_att_in = sm.att_in; auto &s_total = sm.total;
    auto &s_off = sm.off; auto &s_m8 = sm.m8;
    const int64_t n = job.n;
    const int64_t n_seg = (n + kSeg - 1) / kSeg;
    const int FULL = bar_base, DONE = FULL + kChainStages;   // 2 * kChainStages named barriers of 64 threads

    if (!producer) {
        // ------------------------------------------------------------------ consumer: the recurrence only
        double att = 0.0;
        for (int64_t seg = 0; seg < n_seg; ++seg) {
            const int st = (int)(seg % kChainStages);
            named_sync(FULL + st);
            const int total = s_total[st];
            if (lane == 0) s_att_in[st] = att;
            // one active lane is enough (the warp would only repeat the same scalar work 32 times); it also keeps the
            // shared-memory traffic of the operand loads at 16 B instead of 32 x 16 B per instruction
            if (total && lane == 0) {
                const double2 *__restrict__ qmt = s_mt[st];
                const double2 *__restrict__ qid = s_id[st];
                double *__restrict__ qa = s_att[st];
                double2 A0[8], A1[8], B0[8], B1[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) { A0[k] = qmt[k]; A1[k] = qid[k]; }
                for (int j0 = 0; j0 < total; j0 += 16) {
#pragma unroll
                    for (int k = 0; k < 8; ++k) { B0[k] = qmt[j0 + 8 + k]; B1[k] = qid[j0 + 8 + k]; }
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        att = att_update(att, A0[k].x, A1[k].x, A1[k].y, A0[k].y);
                        qa[j0 + k] = att;
                    }
#pragma unroll
                    for (int k = 0; k < 8; ++k) { A0[k] = qmt[j0 + 16 + k]; A1[k] = qid[j0 + 16 + k]; }   // read-ahead (may be slack)
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        att = att_update(att, B0[k].x, B1[k].x, B1[k].y, B0[k].y);
                        qa[j0 + 8 + k] = att;
                    }
                }
            }
            att = __shfl_sync(kFull, att, 0);
            named_arrive(DONE + st);
        }
        return;
    }

    // ---------------------------------------------------------------------- producer: gather, compact, scatter
    const AttEntry *tbl = tables + (size_t)job.table * 32769;
    const uint16_t *rp = rms + (int64_t)job.band * mb_frames + job.mb_begin;
    double *af = att_f + (int64_t)job.band * mb_frames + job.mb_begin;
    double *ck = ckpt + job.ck_begin;
    const int64_t n_groups = (n + 31) >> 5;
    const bool vec = (reinterpret_cast<uintptr_t>(rp) & 15) == 0;

    auto load_seg = [&](int64_t seg, uint32_t *h) {     // this lane's 8 rms values as 4 words
        const int64_t i = seg * kSeg + lane * 8;
        h[0] = h[1] = h[2] = h[3] = 0;
        if (seg >= n_seg) return;
        if (vec && i + 8 <= n) {
            const uint4 q = __ldg(reinterpret_cast<const uint4 *>(rp + i));
            h[0] = q.x; h[1] = q.y; h[2] = q.z; h[3] = q.w;
        } else {
#pragma unroll
            for (int k = 0; k < 8; ++k)
                if (i + k < n) h[k >> 1] |= (uint32_t)__ldg(rp + i + k) << ((k & 1) * 16);
        }
    };
    // scatter the results of a finished segment: attenuation after each flagged frame, and entering each group
    auto flush = [&](int64_t seg) {
        const int st = (int)(seg % kChainStages);
        named_sync(DONE + st);
        const int total = s_total[st];
        double ge = s_att_in[st];
        if (total) {
            const double *qa = s_att[st];
            const unsigned m8 = s_m8[st][lane];
            int slot = s_off[st][lane];
            // flagged frames before group g = exclusive count at lane 4g; lane g (< 8) keeps it
            const int before_grp = s_off[st][(lane & 7) * 4];
            if (before_grp) ge = qa[before_grp - 1];
#pragma unroll
            for (int k = 0; k < 8; ++k)
                if (m8 & (1u << k)) af[seg * kSeg + lane * 8 + k] = qa[slot++];
        }
        if (lane < 8 && seg * 8 + lane < n_groups) ck[seg * 8 + lane] = ge;
    };

    uint32_t cur[4], nx1[4], nx2[4];                   // rms words are fetched two segments ahead
    load_seg(0, cur);
    load_seg(1, nx1);
    for (int64_t seg = 0; seg < n_seg; ++seg) {
        const int st = (int)(seg % kChainStages);
        load_seg(seg + 2, nx2);
        if (seg >= kChainStages) flush(seg - kChainStages);    // also frees stage st
        unsigned m8 = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k)
            if ((cur[k >> 1] >> ((k & 1) * 16)) & 0xffffu) m8 |= 1u << k;
        const int cnt = __popc(m8);
        int incl = cnt;                                // inclusive scan of the per-lane counts
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int v = __shfl_up_sync(kFull, incl, d);
            if (lane >= d) incl += v;
        }
        const int total = __shfl_sync(kFull, incl, 31);
        if (total) {
            double2 *qmt = s_mt[st], *qid = s_id[st];
            int slot = incl - cnt;
            s_m8[st][lane] = (uint8_t)m8;
            s_off[st][lane] = (uint16_t)slot;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const unsigned r = (cur[k >> 1] >> ((k & 1) * 16)) & 0xffffu;
                if (r) {
                    const double2 a = __ldg(reinterpret_cast<const double2 *>(tbl + r));       // (M, inc)
                    const double2 b = __ldg(reinterpret_cast<const double2 *>(tbl + r) + 1);   // (dec, tau)
                    qmt[slot] = make_double2(a.x, b.y);
                    qid[slot] = make_double2(a.y, b.x);
                    ++slot;
                }
            }
            // no-op entries: M < 0 => "above", dec = 0 => att unchanged
            if (lane < 16) { qmt[total + lane] = make_double2(-1.0, 0.0); qid[total + lane] = make_double2(0.0, 0.0); }
        }
        if (lane == 0) s_total[st] = total;
        named_arrive(FULL + st);
#pragma unroll
        for (int k = 0; k < 4; ++k) { cur[k] = nx1[k]; nx1[k] = nx2[k]; }
    }
    for (int64_t seg = (n_seg > kChainStages ? n_seg - kChainStages : 0); seg < n_seg; ++seg) flush(seg);
}

__global__ void __launch_bounds__(kChainsPerCta * 64)
k_att_chain(const ChainJob *__restrict__ jobs, int n_jobs, const uint16_t *__restrict__ rms,
            const AttEntry *__restrict__ tables, double *__restrict__ ckpt, double *__restrict__ att_f, int64_t mb_frames,
            const int *__restrict__ only) {      // only != NULL: just the chains k_att_chain_spec left to this kernel
    extern __shared__ __align__(16) unsigned char s_raw[];
    // Warp w of a CTA lands on SM sub-partition w % 4.  With 64-thread CTAs every consumer warp sat on sub-partitions
    // 1 and 3 and, at ~8 chains per SM, four latency-bound recurrences shared one issue port (24 instead of 10
    // cycles/frame for 1152 chains).  Four chains per CTA put one consumer and one producer on every sub-partition.
    const int warp = threadIdx.x >> 5;
    const int chain = warp & (kChainsPerCta - 1);
    const int job_i = blockIdx.x * kChainsPerCta + chain;
    if (job_i >= n_jobs || (only && !only[job_i])) return;   // both warps of that chain leave; barriers are per chain
    chain_queue_run(jobs[job_i], reinterpret_cast<ChainSmem *>(s_raw)[chain], warp >= kChainsPerCta, threadIdx.x & 31,
                    chain * 2 * kChainStages, rms, tables, ckpt, att_f, mb_frames);
}

// k_att_chain_spec: the recurrence, made parallel in time by SPECULATION AND REPAIR - exact, not approximate.
// One CTA per (chunk, band); S = n_lanes of its threads cut the chunk into S contiguous segments and every lane
// walks its own segment (the lanes of a warp run in lockstep, so a step of the 25-cycle dependent chain advances
// 32 segments at once):
//   pass 1   every lane starts from attenuation 0 (lane 0 really does - the reference resets it per chunk);
//   repair   lane t takes the end value lane t-1 produced in the previous pass.  If that differs from the start it
//            used, it walks its segment again carrying BOTH attenuations (old start, new start - two independent
//            chains in one thread, no extra latency) and stops at the first frame where they are bit-equal: from
//            there on the trajectory it stored before is the right one, and so is its old end value.  A lane that
//            never meets publishes a new end value;
//   until no lane's start changed.  By induction every lane has then been walked from the true end of its
//   predecessor, i.e. the stored values are those of the sequential loop.  A chain that does not settle within its
//   repair budget is redone by two warps of the CTA as the producer / consumer pair of k_att_chain (or, from a
//   one-warp CTA, flagged in gave_up[] for the filtered k_att_chain launch that follows).
// Trajectories meet whenever both clamp to the same max_attenuation (att in [tau, M] -> M), which the compressor
// does all the time while it tracks the level: on the bench tracks 0.3k-20k frames after a segment start (two or
// three passes).  A signal that never clamps would degrade to one segment per pass; the budget below cuts that off.
// Unflagged frames carry rms 0 and table entry 0 (M = inc = dec = tau = 0) is a no-op, so the walk needs no branch
// per frame; 8-frame blocks without any flagged frame are skipped.
// Emits the attenuation after every flagged frame (att_f, sparse) and entering every 32-frame group (ckpt).
constexpr int kChainMaxThreads = 256;
constexpr int kWalkBlock = 16;     // frames per walker iteration = one 32-byte rms load

struct ChainCtx {
    const uint16_t *rp;        // rms plane of this chain
    const AttEntry *tbl;
    double *af, *ck;
    int64_t n;
    bool vec;                  // rp is 32-byte aligned
};

struct RmsBlock { uint32_t w[kWalkBlock / 2]; };

__device__ __forceinline__ RmsBlock chain_load_rms(const ChainCtx &c, int64_t i) {   // rms of frames i..i+15, 0 past the end
    RmsBlock q;
#pragma unroll
    for (int k = 0; k < kWalkBlock / 2; ++k) q.w[k] = 0;
    if (i + kWalkBlock <= c.n && c.vec) {
        // volatile: must stay under the alignment test (a plain asm counts as pure and may be executed speculatively)
        asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(q.w[0]), "=r"(q.w[1]), "=r"(q.w[2]), "=r"(q.w[3]), "=r"(q.w[4]), "=r"(q.w[5]), "=r"(q.w[6]), "=r"(q.w[7])
                     : "l"(c.rp + i));
    } else if (i < c.n) {
#pragma unroll
        for (int k = 0; k < kWalkBlock; ++k)
            if (i + k < c.n) q.w[k >> 1] |= (uint32_t)__ldg(c.rp + i + k) << ((k & 1) * 16);
    }
    return q;
}

__device__ __forceinline__ bool chain_any(const RmsBlock &q, int h) {     // a flagged frame in half h of the block?
    return (q.w[4 * h] | q.w[4 * h + 1] | q.w[4 * h + 2] | q.w[4 * h + 3]) != 0;
}

// table entries of the 8 frames of half h of a block: one 32-byte load per flagged frame, the no-op entry otherwise.
// (plain asm, not volatile: the table is constant and tbl + r is always a valid aligned entry, so the compiler may
// schedule - or speculate - the loads as it likes)
__device__ __forceinline__ void chain_entries(AttEntry *e, const AttEntry *tbl, const RmsBlock &q, int h) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const unsigned r = (q.w[4 * h + (k >> 1)] >> ((k & 1) * 16)) & 0xffffu;
        e[k] = AttEntry{0.0, 0.0, 0.0, 0.0};
        if (r) asm("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(e[k].m), "=d"(e[k].inc), "=d"(e[k].dec), "=d"(e[k].tau) : "l"(tbl + r));
    }
}

struct SegStat { double max_m, sum_dec; int n_flag, n_grp; };   // of a segment: largest max_attenuation, total release if always above
                                                                // it, flagged frames, 8-frame groups holding one

// 8 steps of the recurrence over frames i..i+7 (half h of block q); true when DUAL and a == b afterwards.
// The first (single) walk also gathers the segment statistics.
template <bool DUAL>
__device__ __forceinline__ bool chain_step8(const ChainCtx &c, const RmsBlock &q, int h, const AttEntry *e, int64_t i, double &a, double &b,
                                            SegStat &st) {
    if (!chain_any(q, h)) return false;
    if (!DUAL) ++st.n_grp;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        b = att_update(b, e[k].m, e[k].inc, e[k].dec, e[k].tau);
        if (DUAL) a = att_update(a, e[k].m, e[k].inc, e[k].dec, e[k].tau);
        const bool flagged = ((q.w[4 * h + (k >> 1)] >> ((k & 1) * 16)) & 0xffffu) != 0;
        if (!DUAL) { st.max_m = fmax(st.max_m, e[k].m); st.sum_dec += e[k].dec; st.n_flag += flagged; }
        if (flagged) c.af[i + k] = b;
    }
    return DUAL && __double_as_longlong(a) == __double_as_longlong(b);
}

// Walk frames [b0, b1) (b0 a multiple of 32) from attenuation b; with DUAL also from a, returning true at the first
// 8-frame group after which the two are bit-equal.  b holds the attenuation reached.  The rms words run two blocks
// ahead of the recurrence and the table entries one half block ahead (two register blocks, as many loads in flight
// as the 8 dependent steps they hide behind).
template <bool DUAL>
__device__ __forceinline__ bool chain_walk(const ChainCtx &c, int64_t b0, int64_t b1, double a, double &b, SegStat &st) {
    if (b0 >= b1) return false;
    RmsBlock cur = chain_load_rms(c, b0), nxt = chain_load_rms(c, b0 + kWalkBlock);
    AttEntry ea[8], eb[8];
    chain_entries(ea, c.tbl, cur, 0);
    for (int64_t i = b0; i < b1; i += kWalkBlock) {
        const RmsBlock nn = chain_load_rms(c, i + 2 * kWalkBlock);
        if ((i & 31) == 0) c.ck[i >> 5] = b;
        if (!chain_any(cur, 0) && !chain_any(cur, 1)) {      // 16 silent frames: a dozen instructions instead of ~150
            if (chain_any(nxt, 0)) chain_entries(ea, c.tbl, nxt, 0);
            cur = nxt; nxt = nn;
            continue;
        }
        chain_entries(eb, c.tbl, cur, 1);
        if (chain_step8<DUAL>(c, cur, 0, ea, i, a, b, st)) return true;
        chain_entries(ea, c.tbl, nxt, 0);
        if (chain_step8<DUAL>(c, cur, 1, eb, i + 8, a, b, st)) return true;
        cur = nxt; nxt = nn;
    }
    return false;
}

__global__ void __launch_bounds__(kChainMaxThreads)
k_att_chain_spec(const ChainJob *__restrict__ jobs, const uint16_t *__restrict__ rms, const AttEntry *__restrict__ tables,
                 double *__restrict__ ckpt, double *__restrict__ att_f, int64_t mb_frames, int *__restrict__ gave_up,
                 int n_lanes) {            // segments per chain (<= blockDim.x; threads past it only serve the fallback)
    __shared__ double s_end[kChainMaxThreads], s_max_m[kChainMaxThreads], s_dec[kChainMaxThreads];
    __shared__ int s_nflag[kChainMaxThreads], s_ngrp[kChainMaxThreads], s_stuck, s_budget;
    __shared__ short s_prev[kChainMaxThreads];
    const ChainJob job = jobs[blockIdx.x];
    const int S = n_lanes, t = threadIdx.x;     // threads t >= S get an empty segment [n, n)
    ChainCtx c;
    c.rp = rms + (int64_t)job.band * mb_frames + job.mb_begin;
    c.tbl = tables + (size_t)job.table * 32769;
    c.af = att_f + (int64_t)job.band * mb_frames + job.mb_begin;
    c.ck = ckpt + job.ck_begin;
    c.n = job.n;
    c.vec = (reinterpret_cast<uintptr_t>(c.rp) & 31) == 0;
    // whole 32-frame groups per segment: a checkpoint and a 32-byte rms load never straddle two lanes
    const int64_t seg = (((c.n + S - 1) / S) + 31) & ~(int64_t)31;
    const int64_t b0 = min(c.n, (int64_t)t * seg), b1 = min(c.n, b0 + seg);
    // First guess for the attenuation entering the segment: the max_attenuation of the last flagged frame before it
    // (looked for in the 64 frames in front), which is exactly right whenever the compressor was clamped to it there -
    // about every second frame while it tracks a rising level - and costs nothing when it is wrong; else 0.
    double start = 0.0;
    if (t > 0 && t < S && b0 < b1) {
        const int64_t lo = max((int64_t)0, b0 - 64);
        for (int64_t i = b0 - 1; i >= lo; --i) {
            const unsigned r = __ldg(c.rp + i);
            if (r) { start = c.tbl[r].m; break; }
        }
    }
    double end = start;
    SegStat st{0.0, 0.0, 0, 0};
    chain_walk<false>(c, b0, b1, 0.0, end, st);
    // What the first walk tells, before any repair is paid for (thread 0, S <= 256 segments):
    // * budget.  Measured on B200 (profiles/r01e_summary.md): a repair pass costs a lone warp ~120 cycles per frame of
    //   every 8-frame group that holds a flagged frame (the walk skips the others; the slowest lane carries ~1.3x the
    //   mean) plus ~5 per frame of the segment for scanning; the queue kernel, which compacts the flagged frames,
    //   25 cycles per flagged frame plus ~4 per frame.  Speculation may cost 80 % of what the queue kernel needs for the chain: a
    //   band at 12 % flagged frames spread over every group is cheap for the queue and dear to repair, one at 40 %
    //   in bursts the other way round.
    // * forecast.  Follow a lower bound of the TRUE attenuation through the segments: where it enters a segment above
    //   everything that segment can ask for, even after all the release it could get there, the true trajectory never
    //   clamps in it, cannot meet the speculation, and hands the problem to the next lane.  A run of such segments
    //   costs one repair pass each (one loud passage, then a bed just over threshold: the reference releases by
    //   M / release_frames per frame, i.e. hardly at all); a run far longer than the budget is a chain for k_att_chain.
    // Both only decide WHO computes the chain; what is stored is decided by the bit-equality test alone.
    s_end[t] = end; s_max_m[t] = st.max_m; s_dec[t] = st.sum_dec; s_nflag[t] = st.n_flag; s_ngrp[t] = st.n_grp;
    __syncthreads();
    if (t == 0) {
        double lb = s_end[0];
        int run = 0, longest = 0;
        long long flagged = s_nflag[0], groups = s_ngrp[0];
        for (int u = 1; u < S; ++u) {
            flagged += s_nflag[u];
            groups += s_ngrp[u];
            const double lo = lb - 1.000001 * s_dec[u];
            if (lb > 0.0 && lo > s_max_m[u]) { lb = lo; run += s_nflag[u] != 0; longest = max(longest, run); }   // silent segments re-walk for free
            else { lb = s_end[u]; run = 0; }
        }
        const long long budget = 4 * (25 * flagged + 4 * c.n) / (5 * (1250 * groups / S + 5 * seg)) - 1;
        s_budget = (int)max(2LL, min(budget, (long long)S));
        s_stuck = longest + 2 > s_budget;      // a run of L such segments needs about L + 2 repairs
    }
    __syncthreads();
    // a segment without a flagged frame hands its start on unchanged: lane t takes the end of the last lane before it
    // that has one (else a sparse band would pay a pass per silent segment just to pass a number along).  Silent
    // lanes still re-walk when their start changes - a skip per 16 frames - to refresh their checkpoints.
    {
        int u = t - 1;
        while (u >= 0 && s_nflag[u] == 0) --u;
        s_prev[t] = (short)u;
    }
    const int max_repairs = s_budget;
    int repairs = 0, stuck = s_stuck;
    while (!stuck) {
        s_end[t] = end;
        __syncthreads();
        const double from = s_prev[t] >= 0 ? s_end[s_prev[t]] : 0.0;
        const bool redo = t < S && __double_as_longlong(from) != __double_as_longlong(start);
        const int n_redo = __syncthreads_count(redo);      // also orders the reads of s_end before the next round's writes
        if (!n_redo) break;
        if (repairs >= max_repairs || (repairs >= 2 && 2 * n_redo > S)) { stuck = 1; break; }   // over budget / not settling
        ++repairs;
        if (redo) {
            double b = from;
            if (!chain_walk<true>(c, b0, b1, start, b, st)) end = b;
            start = from;
        }
    }
    // A chain that will not settle is strictly sequential: with two warps or more the CTA turns into one producer /
    // consumer pair of the queue kernel on the spot (the other chains of the launch keep repairing meanwhile);
    // a single-warp CTA leaves it to the filtered k_att_chain launch that follows.
    const bool here = stuck && blockDim.x >= 64;
    if (t == 0) gave_up[blockIdx.x] = stuck && !here;
    if (here && t < 64) {
        extern __shared__ __align__(16) unsigned char s_raw[];
        chain_queue_run(job, *reinterpret_cast<ChainSmem *>(s_raw), t >= 32, t & 31, 1, rms, tables, ckpt, att_f, mb_frames);
    }
}

__device__ __forceinline__ int mul_floor(int x, double f) {   // audioop.c fbound()
    double v = __dmul_rn((double)x, f);
    if (v > 32767.0) v = 32767.0;
    else if (v < -32767.0) v = -32768.0;
    return __double2int_rd(v);
}

// k_compress_apply: one warp per 256-frame segment, all three bands.  The attenuation in force at a frame is
// the value k_att_chain stored for the last flagged frame at or before it inside its 32-frame group, or the
// group's entry value; gain = 10^(-att/20) (pydub db_to_float), audioop.mul = floor(clip(x * gain)), skipped
// when att == 0 exactly as pydub does, then low.overlay(mid).overlay(high) = saturating adds (:309).
// pydub db_to_float(-att) = 10 ** (-att / 20); out of line so the 24 call sites of k_compress_apply share one copy
// (the fully inlined kernel thrashed the instruction cache: stall_no_instruction 1.7 per issue)
__device__ __noinline__ double gain_of_att(double att) { return exp10(-att / 20.0); }

__global__ void __launch_bounds__(128)
k_compress_apply(const MbChunk *__restrict__ chunks, int n_chunks, int64_t seg_lo, int64_t seg_hi,
                 const int16_t *__restrict__ bands, const uint16_t *__restrict__ rms,
                 const double *__restrict__ ckpt, const double *__restrict__ att_f, int16_t *__restrict__ pre,
                 int64_t mb_frames) {
    const int64_t seg = seg_lo + (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (seg >= seg_hi) return;
    const int lane = threadIdx.x & 31;
    int lo = 0, hi = n_chunks - 1;                 // last chunk (of this launch's slice) with seg_prefix <= seg
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (chunks[mid].seg_prefix <= seg) lo = mid; else hi = mid - 1;
    }
    const MbChunk ck = chunks[lo];
    const int64_t lseg = seg - ck.seg_prefix;
    const int64_t f0 = lseg * kSeg;
    const int64_t n = ck.n;
    // issue every load of the segment first (24 groups x {rms, band word, entry attenuation}): the kernel is
    // bound by memory latency, not by arithmetic
    unsigned rv[3][8];
    uint32_t wv[3][8];
    double ce[3][8];
#pragma unroll
    for (int b = 0; b < 3; ++b) {
        const uint32_t *bp = reinterpret_cast<const uint32_t *>(bands) + (int64_t)b * mb_frames + ck.mb_begin;
        const uint16_t *rp = rms + (int64_t)b * mb_frames + ck.mb_begin;
        const double *cp = ckpt + ck.ck_begin[b] + lseg * 8;
#pragma unroll
        for (int g = 0; g < 8; ++g) {
            const int64_t i = f0 + g * 32 + lane;
            const bool valid = i < n;
            rv[b][g] = valid ? (unsigned)__ldg(rp + i) : 0u;
            wv[b][g] = valid ? __ldg(bp + i) : 0u;
            ce[b][g] = (f0 + g * 32 < n) ? __ldg(cp + g) : 0.0;
        }
    }
    int accl[8], accr[8];
#pragma unroll
    for (int b = 0; b < 3; ++b) {
        const double *ap = att_f + (int64_t)b * mb_frames + ck.mb_begin;
        double c_att = 0.0, c_fac = 1.0;           // per-lane cache of the last gain computed
#pragma unroll
        for (int g = 0; g < 8; ++g) {
            const int64_t i = f0 + g * 32 + lane;
            const uint32_t w = wv[b][g];
            int l = (int16_t)(w & 0xffffu), r = (int16_t)(w >> 16);
            const unsigned flags = __ballot_sync(kFull, rv[b][g] != 0);
            const unsigned below = flags & (0xffffffffu >> (31 - lane));
            const double mine = below ? __ldg(ap + (i - lane) + (31 - __clz(below))) : ce[b][g];
            if (mine != 0.0) {
                if (mine != c_att) { c_att = mine; c_fac = gain_of_att(mine); }
                l = mul_floor(l, c_fac);
                r = mul_floor(r, c_fac);
            }
            accl[g] = b ? sat16(accl[g] + l) : l;
            accr[g] = b ? sat16(accr[g] + r) : r;
        }
    }
    uint32_t *out = reinterpret_cast<uint32_t *>(pre) + ck.abs_begin;
#pragma unroll
    for (int g = 0; g < 8; ++g) {
        const int64_t i = f0 + g * 32 + lane;
        if (i < n) out[i] = pack16(accl[g], accr[g]);
    }
}

// ------------------------------------------------------------------------------------------------
// k_kweight_energy: s16 -> x/32768 -> BS.1770 pre-filter + RLB (2 biquads, FP64) -> sum of squares
// per 100 ms sub-block (ebur128.c filter + gating-block sums).  K-filter state runs through the
// whole track (the reference measures the concatenated file), so warm-up may cross chunk joins.
// One thread per tile of sub-blocks, both channels (two independent chains).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
k_kweight_energy(const KwJob *__restrict__ jobs, int n_jobs, const ame_track_params *__restrict__ tracks,
                 const TrackDev *__restrict__ tdev, const int16_t *__restrict__ pre,
                 double *__restrict__ energy, int *__restrict__ peak) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_jobs) return;
    const KwJob job = jobs[j];
    const ame_track_params *tp = tracks + job.track;
    const TrackDev td = tdev[job.track];
    const ame_biquad k0 = tp->kw[0], k1 = tp->kw[1];
    const int64_t s100 = td.s100;
    const int64_t base = tp->offset_frames;                      // multiple of 8
    const int64_t t_begin = (int64_t)job.sb_begin * s100;        // relative to the track
    const int64_t t_end = (int64_t)job.sb_end * s100;
    int64_t f_lo = t_begin - (int64_t)tp->warm_kw;
    if (f_lo < 0) f_lo = 0;
    f_lo &= ~(int64_t)7;                                         // extra warm-up frames are harmless
    double zl[4] = {0, 0, 0, 0}, zr[4] = {0, 0, 0, 0}, accl = 0, accr = 0;
    int pk = 0;
    int64_t next_end = t_begin + s100;
    int sb = job.sb_begin;
    const uint4 *src = reinterpret_cast<const uint4 *>(reinterpret_cast<const uint32_t *>(pre) + base + f_lo);
    const int n_it = (int)((t_end - f_lo + 7) >> 3);             // 8 frames = one 32-byte sector per iteration
    uint4 c0 = ldg16(src), c1 = ldg16(src + 1), n0 = c0, n1 = c1;
    if (n_it > 1) { n0 = ldg16(src + 2); n1 = ldg16(src + 3); }
    for (int it = 0; it < n_it; ++it) {
        uint4 m0 = n0, m1 = n1;
        if (it + 2 < n_it) { m0 = ldg16(src + 2 * it + 4); m1 = ldg16(src + 2 * it + 5); }
        const int64_t g = f_lo + 8 * (int64_t)it;
        const uint32_t w[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int64_t f = g + k;
            const int xl = (int)(int16_t)(w[k] & 0xffffu), xr = (int)(int16_t)(w[k] >> 16);
            const double yl = bq_step(k1, zl[2], zl[3], bq_step(k0, zl[0], zl[1], i16_to_unit(xl)));
            const double yr = bq_step(k1, zr[2], zr[3], bq_step(k0, zr[0], zr[1], i16_to_unit(xr)));
            if (f >= t_begin && f < t_end) {
                accl = fma(yl, yl, accl);
                accr = fma(yr, yr, accr);
                pk = max(pk, max(abs(xl), abs(xr)));
                if (f + 1 == next_end) {
                    energy[td.sb_offset + sb] = accl + accr;     // ebur128: per-channel sums, then added
                    accl = 0; accr = 0; ++sb; next_end += s100;
                }
            }
        }
        c0 = n0; c1 = n1; n0 = m0; n1 = m1;
    }
    atomicMax(peak + job.track, pk);
}

// tail frames beyond the last complete sub-block still count for the sample peak
__global__ void k_tail_peak(const ame_track_params *__restrict__ tracks, const TrackDev *__restrict__ tdev,
                            int track_lo, int track_hi, const int16_t *__restrict__ pre, int *__restrict__ peak) {
    const int t = track_lo + blockIdx.x;
    if (t >= track_hi) return;
    const int64_t begin = (int64_t)tdev[t].n_sb * tdev[t].s100;
    const int16_t *p = pre + 2 * tracks[t].offset_frames;
    int pk = 0;
    for (int64_t i = 2 * begin + threadIdx.x; i < 2 * tdev[t].n_total; i += blockDim.x) pk = max(pk, abs((int)p[i]));
    atomicMax(peak + t, pk);
}

__device__ __forceinline__ int hist_index(double e) {   // ebur128.c find_histogram_index
    int lo = 0, hi = 1000;
    do {
        int mid = (lo + hi) >> 1;
        if (e >= c_hist_bounds[mid]) lo = mid; else hi = mid;
    } while (hi - lo != 1);
    return lo;
}

// k_block_hist: 400 ms blocks every 100 ms -> 1000-bin histogram (absolute gate = bin floor, -70 LUFS), and the
// short-term histogram behind loudness range: 3 s windows, the first ending at 3 s, then one per second
// (ebur128.c: short_term_frame_counter reaches 30 sub-blocks, is reset to 20).
__global__ void __launch_bounds__(256)
k_block_hist(const TrackDev *__restrict__ tdev, int track_lo, const double *__restrict__ energy, long long *__restrict__ hist,
             int *__restrict__ hist_st) {
    __shared__ unsigned s_hist[1000];
    __shared__ unsigned s_st[1000];
    const int t = track_lo + blockIdx.x;
    for (int i = threadIdx.x; i < 1000; i += blockDim.x) { s_hist[i] = 0; s_st[i] = 0; }
    __syncthreads();
    const TrackDev td = tdev[t];
    const double *e = energy + td.sb_offset;
    const double denom = (double)(4 * (int64_t)td.s100);
    for (int j = td.first_block + threadIdx.x; j + 3 < td.n_sb; j += blockDim.x) {
        const double s = (((e[j] + e[j + 1]) + e[j + 2]) + e[j + 3]) / denom;
        if (s >= c_hist_bounds[0]) atomicAdd(&s_hist[hist_index(s)], 1u);
    }
    const double denom_st = (double)(30 * (int64_t)td.s100);
    for (int k = threadIdx.x; 10 * k + 29 < td.n_sb; k += blockDim.x) {
        double s = 0.0;
        for (int j = 0; j < 30; ++j) s += e[10 * k + j];
        s /= denom_st;
        if (s >= c_hist_bounds[0]) atomicAdd(&s_st[hist_index(s)], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 1000; i += blockDim.x) {
        hist[(int64_t)t * 1000 + i] = (long long)s_hist[i];
        hist_st[(int64_t)t * 1000 + i] = (int)s_st[i];
    }
}

// k_finalize: ebur128_gated_loudness + loudness range + the linear-mode gain of af_loudnorm, one thread per track
__global__ void k_finalize(const ame_track_params *__restrict__ tracks, int track_lo, int track_hi,
                           const long long *__restrict__ hist, const int *__restrict__ hist_st, const int *__restrict__ peak,
                           ame_track_result *__restrict__ res) {
    const int t = track_lo + blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= track_hi) return;
    const long long *h = hist + (int64_t)t * 1000;
    ame_track_result r;
    r.input_i = -INFINITY; r.measured_i_2dp = -INFINITY; r.gain = 1.0; r.rel_threshold = 0.0;
    r.n_blocks = 0; r.normalized = 0; r.sample_peak = peak[t];
    r.input_lra = 0.0; r.input_thresh = -70.0;
    double rel = 0.0; long long count = 0;
    for (int j = 0; j < 1000; ++j) { rel += (double)h[j] * c_hist_energy[j]; count += h[j]; }
    r.n_blocks = count;
    if (count > 0) {
        rel /= (double)count;
        rel *= 0.1;                                  // RELATIVE_GATE_FACTOR = 10^(-10/10)
        r.rel_threshold = rel;
        r.input_thresh = 10.0 * log10(rel) - 0.691;  // ff_ebur128_relative_threshold
        int start;
        if (rel < c_hist_bounds[0]) start = 0;
        else { start = hist_index(rel); if (rel > c_hist_energy[start]) ++start; }
        double gated = 0.0; long long above = 0;
        for (int j = start; j < 1000; ++j) { gated += (double)h[j] * c_hist_energy[j]; above += h[j]; }
        if (above > 0) {
            gated /= (double)above;
            r.input_i = 10.0 * log10(gated) - 0.691;
            if (tracks[t].flags & AME_F_NORMALIZE) {
                r.measured_i_2dp = rint(r.input_i * 100.0) / 100.0;   // the '%.2f' string of pass 1
                r.gain = pow(10.0, (tracks[t].target_lufs - r.measured_i_2dp) / 20.0);
                r.normalized = 1;
            }
        }
    }
    // ff_ebur128_loudness_range: short-term blocks above (mean power - 20 dB), 10th .. 95th percentile
    {
        const int *hs = hist_st + (int64_t)t * 1000;
        long long n = 0; double power = 0.0;
        for (int j = 0; j < 1000; ++j) { n += hs[j]; power += (double)hs[j] * c_hist_energy[j]; }
        if (n > 0) {
            power /= (double)n;
            const double integ = 0.01 * power;       // MINUS_20DB
            int idx;
            if (integ < c_hist_bounds[0]) idx = 0;
            else { idx = hist_index(integ); if (integ > c_hist_energy[idx]) ++idx; }
            long long m = 0;
            for (int j = idx; j < 1000; ++j) m += hs[j];
            if (m > 0) {
                const long long p_lo = (long long)((double)(m - 1) * 0.1 + 0.5), p_hi = (long long)((double)(m - 1) * 0.95 + 0.5);
                long long acc = 0; int j = idx;
                while (acc <= p_lo) acc += hs[j++];
                const double l_en = c_hist_energy[j - 1];
                while (acc <= p_hi) acc += hs[j++];
                const double h_en = c_hist_energy[j - 1];
                r.input_lra = (10.0 * log10(h_en) - 0.691) - (10.0 * log10(l_en) - 0.691);
            }
        }
    }
    res[t] = r;
}

// k_apply_gain: loudnorm linear mode: s16 -> x/32768 -> * gain -> lrint(x * 32768) clipped to s16
__global__ void __launch_bounds__(256)
k_apply_gain(const GainJob *__restrict__ jobs, const ame_track_result *__restrict__ res,
             const int16_t *__restrict__ pre, int16_t *__restrict__ out) {
    const GainJob job = jobs[blockIdx.x];
    const ame_track_result r = res[job.track];
    const uint4 *src = reinterpret_cast<const uint4 *>(pre);
    uint4 *dst = reinterpret_cast<uint4 *>(out);
    const int64_t v0 = job.begin >> 2, v1 = (job.end + 3) >> 2;   // 4 frames per uint4; tiles are 4-aligned
    if (!r.normalized) {
        for (int64_t v = v0 + threadIdx.x; v < v1; v += blockDim.x) dst[v] = __ldg(src + v);
        return;
    }
    const double g = r.gain;
    for (int64_t v = v0 + threadIdx.x; v < v1; v += blockDim.x) {
        const uint4 q = __ldg(src + v);
        uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int l = (int16_t)(w[k] & 0xffffu), rr = (int16_t)(w[k] >> 16);
            int ol = __double2int_rn(__dmul_rn(__dmul_rn((double)l * (1.0 / 32768.0), g), 32768.0));
            int orr = __double2int_rn(__dmul_rn(__dmul_rn((double)rr * (1.0 / 32768.0), g), 32768.0));
            w[k] = (uint32_t)(uint16_t)sat16(ol) | ((uint32_t)(uint16_t)sat16(orr) << 16);
        }
        dst[v] = make_uint4(w[0], w[1], w[2], w[3]);
    }
}

}  // namespace ame
