// Device kernels of libame (sm_100a).  See DESIGN.md for the data layout and the per-kernel rooflines.
//
// Time parallelism: every 30 s chunk of the reference restarts all filter state from zero
// (audio_mastering_engine.py:185-199), and inside a chunk the filters are stable LTI systems, so a
// tile of T frames is computed by ONE lane pair (L lane, R lane) that first runs `warm` frames of the
// preceding audio from zero state.  `warm` is chosen on the host so that the state error has decayed
// below 1e-13 of the signal level (FP64 rounding level) when the tile proper starts; the first tile
// of a chunk needs no warm-up and is exact by construction.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/ame.h"

namespace ame {

constexpr int kCK = 256;           // compressor checkpoint spacing (frames)
constexpr int kRmsTile = 2048;     // frames per CTA in k_window_rms
constexpr int kRmsThreads = 256;
constexpr int kGainTile = 32768;   // frames per CTA in k_apply_gain

struct TileJob {           // one lane pair of k_eq / k_band_split
    int64_t chunk_begin;   // absolute frame index (packed buffer) where filter state is reset
    int64_t tile_begin;    // first frame this job writes (absolute, multiple of 4 unless == chunk_begin)
    int64_t tile_end;      // one past the last frame it writes
    int32_t track;
    int32_t variant;       // EQ stage mask (bit s = stage s active)
};

struct MbChunk {           // one chunk of a multiband track
    int64_t abs_begin;     // absolute frame index in the packed in/pre buffers
    int64_t mb_begin;      // frame index in the multiband-only packing (bands / rms planes)
    int64_t n;             // frames
    int64_t seg_prefix;    // number of kCK segments in all earlier chunks
    int32_t track;
    int32_t pad;
};

struct RmsJob { int32_t chunk; int32_t band; int64_t tile_begin; };   // tile_begin relative to chunk

struct KwJob { int32_t track; int32_t sb_begin; int32_t sb_end; int32_t pad; };

struct GainJob { int64_t begin; int64_t end; int32_t track; int32_t pad; };

struct AttEntry { double m, inc, dec, pad; };   // indexed by integer rms 0..32768

struct TrackDev {          // device-side per-track bookkeeping
    int64_t sb_offset;     // start of this track's 100 ms energies in the energy array
    int32_t n_sb;          // number of complete 100 ms sub-blocks
    int32_t s100;          // frames per 100 ms = (fs + 5) / 10   (ebur128.c)
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

// np.clip(x,-1,1) * 32767 -> astype(int16)  (truncate toward zero), float64 flavour
__device__ __forceinline__ int to_pcm_f64(double v) {
    v = fmin(fmax(v, -1.0), 1.0);
    return __double2int_rz(__dmul_rn(v, 32767.0));
}
__device__ __forceinline__ int to_pcm_f32(float v) {
    v = fminf(fmaxf(v, -1.0f), 1.0f);
    return __float2int_rz(__fmul_rn(v, 32767.0f));
}

struct EqCoef {
    ame_biquad s0;               // low shelf
    ame_biquad p1[4];            // 1 kHz peak
    ame_biquad p2[4];            // 4 kHz peak
    ame_biquad s3;               // high shelf
    double g0, gm0, gm1, gm2, g3, gm3;
    int kind0, kind3;
};

// ------------------------------------------------------------------------------------------------
// k_eq: int16 in -> [warmth -> int16] -> float32 -> 4-stage EQ in FP64 -> float32 -> [width] -> int16
// One lane per channel, lanes (2j, 2j+1) = (L, R) of job j.
// ------------------------------------------------------------------------------------------------
template <int MASK>
__device__ __forceinline__ void eq_tile(const TileJob &job, const ame_track_params *__restrict__ tp,
                                        const float *__restrict__ luts, const int16_t *__restrict__ in,
                                        int16_t *__restrict__ pre, int ch, unsigned pmask) {
    const unsigned flags = tp->flags;
    const bool warmth = (flags & AME_F_WARMTH) != 0;
    const bool widen = (flags & AME_F_WIDTH) != 0;
    const float *lut = warmth ? luts + (size_t)tp->warm_lut * 65536 + 32768 : nullptr;
    const double wl_b0 = tp->wl_b0, wl_b1 = tp->wl_b1, wl_a1 = tp->wl_a1, wl_gm1 = tp->wl_gm1;
    const double wh_b0 = tp->wh_b0, wh_b1 = tp->wh_b1, wh_a1 = tp->wh_a1, wh_gm1 = tp->wh_gm1;
    const float wfac = tp->width;

    EqCoef c;
    if (MASK & 1) { c.s0 = tp->eq[0].s[0]; c.g0 = tp->eq[0].g; c.gm0 = tp->eq[0].gm1; c.kind0 = tp->eq[0].kind; }
    if (MASK & 2) {
#pragma unroll
        for (int i = 0; i < 4; ++i) c.p1[i] = tp->eq[1].s[i];
        c.gm1 = tp->eq[1].gm1;
    }
    if (MASK & 4) {
#pragma unroll
        for (int i = 0; i < 4; ++i) c.p2[i] = tp->eq[2].s[i];
        c.gm2 = tp->eq[2].gm1;
    }
    if (MASK & 8) { c.s3 = tp->eq[3].s[0]; c.g3 = tp->eq[3].g; c.gm3 = tp->eq[3].gm1; c.kind3 = tp->eq[3].kind; }

    double z[20];
#pragma unroll
    for (int i = 0; i < 20; ++i) z[i] = 0.0;

    const int64_t warm = (MASK != 0) ? (int64_t)tp->warm_eq : 0;
    int64_t g_lo = job.tile_begin - warm;
    if (g_lo < job.chunk_begin) g_lo = job.chunk_begin;
    const int64_t g_hi = job.tile_end;

    auto frame = [&](uint32_t w) -> int {
        int xl = (int)(int16_t)(w & 0xffffu), xr = (int)(int16_t)(w >> 16);
        int xm = ch ? xr : xl;
        float xf;
        if (warmth) {
            // apply_analog_character (:258-266): tanh in float32 (table = the host's own np.tanh),
            // then two order-2 "shelves" that lfilter(axis=-1) runs ACROSS the two channels.
            double L = (double)lut[xl];
            double mine = ch ? (double)lut[xr] : L;
            // 120 Hz low: y0 = b0*L ; y1 = (b1*L - a1*y0) + b0*R ; blend x + (y-x)*(g-1)
            double t0 = __dmul_rn(wl_b0, mine);
            double y0L = __dmul_rn(wl_b0, L);
            double zz = __dsub_rn(__dmul_rn(wl_b1, L), __dmul_rn(wl_a1, y0L));
            double y = ch ? __dadd_rn(zz, t0) : t0;
            double L1 = __dadd_rn(L, __dmul_rn(__dsub_rn(y0L, L), wl_gm1));
            mine = __dadd_rn(mine, __dmul_rn(__dsub_rn(y, mine), wl_gm1));
            // 12 kHz high, same structure on the blended values
            t0 = __dmul_rn(wh_b0, mine);
            y0L = __dmul_rn(wh_b0, L1);
            zz = __dsub_rn(__dmul_rn(wh_b1, L1), __dmul_rn(wh_a1, y0L));
            y = ch ? __dadd_rn(zz, t0) : t0;
            mine = __dadd_rn(mine, __dmul_rn(__dsub_rn(y, mine), wh_gm1));
            xm = to_pcm_f64(mine);               // float_array_to_audio_segment (:254-257)
        }
        xf = __fmul_rn((float)xm, 1.0f / 32768.0f);   // audio_segment_to_float_array (:250-253)
        float yf = xf;
        if (MASK != 0) {
            double v = (double)xf;
            if (MASK & 1) {   // apply_shelf_filter 250 Hz low (:283-289)
                double f = bq_step(c.s0, z[0], z[1], v);
                v = (c.kind0 == AME_EQ_SHELF_BOOST) ? v + (f - v) * c.gm0 : v * c.g0 + (f - v * c.g0);
            }
            if (MASK & 2) {   // apply_peak_filter 1 kHz (:290-298)
                double t = v;
#pragma unroll
                for (int i = 0; i < 4; ++i) t = bq_step(c.p1[i], z[2 + 2 * i], z[3 + 2 * i], t);
                v = v + t * c.gm1;
            }
            if (MASK & 4) {   // apply_peak_filter 4 kHz
                double t = v;
#pragma unroll
                for (int i = 0; i < 4; ++i) t = bq_step(c.p2[i], z[10 + 2 * i], z[11 + 2 * i], t);
                v = v + t * c.gm2;
            }
            if (MASK & 8) {   // apply_shelf_filter 8 kHz high
                double f = bq_step(c.s3, z[18], z[19], v);
                v = (c.kind3 == AME_EQ_SHELF_BOOST) ? v + (f - v) * c.gm3 : v * c.g3 + (f - v * c.g3);
            }
            yf = __double2float_rn(v);            // samples[:, i] = ... into the float32 array (:274)
        }
        if (widen) {          // apply_stereo_width (:267-271), float32 arithmetic
            float other = __shfl_xor_sync(pmask, yf, 1);
            float l = ch ? other : yf, r = ch ? yf : other;
            float mid = __fmul_rn(__fadd_rn(l, r), 0.5f);
            float side = __fmul_rn(__fmul_rn(__fsub_rn(l, r), 0.5f), wfac);
            yf = ch ? __fsub_rn(mid, side) : __fadd_rn(mid, side);
        }
        return to_pcm_f32(yf);                    // clip inside to_pcm == np.clip of (:270) then (:255)
    };

    int64_t g = g_lo & ~(int64_t)3;
    const uint4 *src = reinterpret_cast<const uint4 *>(in) + (g >> 2);
    uint4 *dst = reinterpret_cast<uint4 *>(pre) + (g >> 2);
    uint4 cur = make_uint4(0, 0, 0, 0);
    if (g < g_hi) cur = ldg16(src);
    for (; g < g_hi; g += 4, ++src, ++dst) {
        uint4 nxt = make_uint4(0, 0, 0, 0);
        if (g + 4 < g_hi) nxt = ldg16(src + 1);
        uint32_t w[4] = {cur.x, cur.y, cur.z, cur.w};
        uint32_t o[4];
        if (g >= g_lo && g + 4 <= g_hi) {         // full group: straight-line code
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                int mine = frame(w[k]);
                int other = __shfl_xor_sync(pmask, mine, 1);
                o[k] = ch ? ((uint32_t)(uint16_t)other | ((uint32_t)(uint16_t)mine << 16))
                          : ((uint32_t)(uint16_t)mine | ((uint32_t)(uint16_t)other << 16));
            }
            if (g >= job.tile_begin) {
                if (ch == 0) *dst = make_uint4(o[0], o[1], o[2], o[3]);
            } else if (g + 4 > job.tile_begin) {
                for (int k = 0; k < 4; ++k)
                    if (g + k >= job.tile_begin && ch == 0) reinterpret_cast<uint32_t *>(dst)[k] = o[k];
            }
        } else {                                  // ragged head / tail of the chunk
            for (int k = 0; k < 4; ++k) {
                int64_t f = g + k;
                if (f >= g_lo && f < g_hi) {
                    int mine = frame(w[k]);
                    int other = __shfl_xor_sync(pmask, mine, 1);
                    uint32_t word = ch ? ((uint32_t)(uint16_t)other | ((uint32_t)(uint16_t)mine << 16))
                                       : ((uint32_t)(uint16_t)mine | ((uint32_t)(uint16_t)other << 16));
                    if (f >= job.tile_begin && ch == 0) reinterpret_cast<uint32_t *>(dst)[k] = word;
                }
            }
        }
        cur = nxt;
    }
}

__global__ void __launch_bounds__(128, 2)
k_eq(const TileJob *__restrict__ jobs, int n_jobs, const ame_track_params *__restrict__ tracks,
     const float *__restrict__ luts, const int16_t *__restrict__ in, int16_t *__restrict__ pre) {
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    const int pair = tid >> 1;
    if (pair >= n_jobs) return;                   // both lanes of a pair leave together
    const int ch = tid & 1;
    const unsigned pmask = 3u << ((threadIdx.x & 31) & ~1);
    const TileJob job = jobs[pair];
    const ame_track_params *tp = tracks + job.track;
    switch (job.variant) {
#define AME_EQ_CASE(M) case M: eq_tile<M>(job, tp, luts, in, pre, ch, pmask); break;
        AME_EQ_CASE(0) AME_EQ_CASE(1) AME_EQ_CASE(2) AME_EQ_CASE(3) AME_EQ_CASE(4) AME_EQ_CASE(5)
        AME_EQ_CASE(6) AME_EQ_CASE(7) AME_EQ_CASE(8) AME_EQ_CASE(9) AME_EQ_CASE(10) AME_EQ_CASE(11)
        AME_EQ_CASE(12) AME_EQ_CASE(13) AME_EQ_CASE(14) AME_EQ_CASE(15)
#undef AME_EQ_CASE
    }
}

// ------------------------------------------------------------------------------------------------
// k_band_split: int16 pre -> Butterworth-4 LP 250 / HP 4k in FP64, mid = x - low - high, each band
// truncated to int16 (apply_multiband_compressor :300-305).  Same lane-pair / warm-up scheme.
// bands = 3 planes of mb_frames frames each.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 4)
k_band_split(const TileJob *__restrict__ jobs, int n_jobs, const ame_track_params *__restrict__ tracks,
             const int64_t *__restrict__ mb_delta,   // per track: mb_offset - offset_frames
             const int16_t *__restrict__ pre, int16_t *__restrict__ bands, int64_t mb_frames) {
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    const int pair = tid >> 1;
    if (pair >= n_jobs) return;
    const int ch = tid & 1;
    const unsigned pmask = 3u << ((threadIdx.x & 31) & ~1);
    const TileJob job = jobs[pair];
    const ame_track_params *tp = tracks + job.track;
    const ame_biquad lp0 = tp->xlp[0], lp1 = tp->xlp[1], hp0 = tp->xhp[0], hp1 = tp->xhp[1];
    const int64_t delta = mb_delta[job.track];
    double z[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) z[i] = 0.0;
    int64_t g_lo = job.tile_begin - (int64_t)tp->warm_xover;
    if (g_lo < job.chunk_begin) g_lo = job.chunk_begin;
    const int64_t g_hi = job.tile_end;
    uint32_t *b0 = reinterpret_cast<uint32_t *>(bands);
    uint32_t *b1 = b0 + mb_frames;
    uint32_t *b2 = b1 + mb_frames;

    int64_t g = g_lo & ~(int64_t)3;
    const uint4 *src = reinterpret_cast<const uint4 *>(pre) + (g >> 2);
    uint4 cur = make_uint4(0, 0, 0, 0);
    if (g < g_hi) cur = ldg16(src);
    for (; g < g_hi; g += 4, ++src) {
        uint4 nxt = make_uint4(0, 0, 0, 0);
        if (g + 4 < g_hi) nxt = ldg16(src + 1);
        uint32_t w[4] = {cur.x, cur.y, cur.z, cur.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int64_t f = g + k;
            if (f >= g_lo && f < g_hi) {
                int xm = ch ? (int)(int16_t)(w[k] >> 16) : (int)(int16_t)(w[k] & 0xffffu);
                double x = (double)__fmul_rn((float)xm, 1.0f / 32768.0f);
                double lo = bq_step(lp1, z[2], z[3], bq_step(lp0, z[0], z[1], x));
                double hi = bq_step(hp1, z[6], z[7], bq_step(hp0, z[4], z[5], x));
                double mid = __dsub_rn(__dsub_rn(x, lo), hi);
                int p0 = to_pcm_f64(lo), p1 = to_pcm_f64(mid), p2 = to_pcm_f64(hi);
                int q0 = __shfl_xor_sync(pmask, p0, 1);
                int q1 = __shfl_xor_sync(pmask, p1, 1);
                int q2 = __shfl_xor_sync(pmask, p2, 1);
                if (f >= job.tile_begin) {
                    const int64_t m = f + delta;
                    if (ch == 0) {
                        b0[m] = (uint32_t)(uint16_t)p0 | ((uint32_t)(uint16_t)q0 << 16);
                        b2[m] = (uint32_t)(uint16_t)p2 | ((uint32_t)(uint16_t)q2 << 16);
                    } else {
                        b1[m] = (uint32_t)(uint16_t)q1 | ((uint32_t)(uint16_t)p1 << 16);
                    }
                }
            }
        }
        cur = nxt;
    }
}

// ------------------------------------------------------------------------------------------------
// k_window_rms: audioop.rms over the previous look_frames frames (both channels), per band.
// pydub rms_at(i) = seg.get_sample_slice(i - look, i).rms ; audioop.rms = (unsigned)sqrt(sum/n).
// One CTA per (chunk, band, tile of kRmsTile frames): exclusive prefix sums of per-frame energies in
// shared memory (exact uint64), window sum = P[i] - P[i-look].
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kRmsThreads)
k_window_rms(const RmsJob *__restrict__ jobs, const MbChunk *__restrict__ chunks,
             const ame_track_params *__restrict__ tracks, const int16_t *__restrict__ bands,
             uint16_t *__restrict__ rms, int64_t mb_frames, int max_look) {
    extern __shared__ unsigned long long s_pref[];     // [max_look + kRmsTile] exclusive prefix sums
    __shared__ unsigned long long s_warp[kRmsThreads / 32];
    const RmsJob job = jobs[blockIdx.x];
    const MbChunk ck = chunks[job.chunk];
    const int look = tracks[ck.track].comp[job.band].look_frames;
    const uint32_t *bp = reinterpret_cast<const uint32_t *>(bands) + (int64_t)job.band * mb_frames + ck.mb_begin;
    uint16_t *rp = rms + (int64_t)job.band * mb_frames + ck.mb_begin;
    const int64_t t0 = job.tile_begin;
    const int64_t t1 = min(t0 + (int64_t)kRmsTile, ck.n);
    const int total = (int)(t1 - t0) + look;           // elements e[0..total): frames t0-look .. t1-1
    // per-thread contiguous run
    const int per = (look + kRmsTile + kRmsThreads - 1) / kRmsThreads;
    const int j0 = threadIdx.x * per;
    unsigned long long run = 0;
    for (int j = j0; j < j0 + per && j < total; ++j) {
        int64_t f = t0 - look + j;
        unsigned long long e = 0;
        if (f >= 0) {
            uint32_t w = __ldg(bp + f);
            long long l = (int16_t)(w & 0xffffu), r = (int16_t)(w >> 16);
            e = (unsigned long long)(l * l + r * r);
        }
        s_pref[j] = run;                               // exclusive within the run
        run += e;
    }
    // block exclusive scan of the run totals
    unsigned long long incl = run;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        unsigned long long v = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += v;
    }
    if (lane == 31) s_warp[wid] = incl;
    __syncthreads();
    unsigned long long base = 0;
    for (int w = 0; w < wid; ++w) base += s_warp[w];
    base += incl - run;
    for (int j = j0; j < j0 + per && j < total; ++j) s_pref[j] += base;
    __syncthreads();
    for (int i = threadIdx.x; i < (int)(t1 - t0); i += kRmsThreads) {
        const int64_t f = t0 + i;                      // frame within chunk
        const int j = i + look;                        // index of frame f in e[]
        const unsigned long long s = s_pref[j] - s_pref[j - look];
        const int64_t nfr = f < look ? f : look;
        unsigned r = 0;
        if (nfr > 0) r = (unsigned)__double2uint_rz(__dsqrt_rn(__ddiv_rn((double)s, (double)(2 * nfr))));
        rp[f] = (uint16_t)r;
    }
}

// ------------------------------------------------------------------------------------------------
// compressor attenuation recurrence (pydub compress_dynamic_range loop body):
//   if rms > thresh and att <= M: att = min(att + inc, M) else att = max(att - dec, 0)
// with M, inc, dec functions of the integer rms (host-built table, same libm as CPython).
// Below threshold M = 0 => dec = 0 => att is frozen (the reference's never-release quirk).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double att_step(double att, unsigned r, unsigned thr_i, const AttEntry *__restrict__ tbl) {
    if (r >= thr_i) {
        const double2 a = __ldg(reinterpret_cast<const double2 *>(tbl + r));
        const double2 b = __ldg(reinterpret_cast<const double2 *>(tbl + r) + 1);
        const double up = fmin(att + a.y, a.x);
        const double dn = fmax(att - b.x, 0.0);
        att = (att <= a.x) ? up : dn;
    }
    return att;
}

struct ChainJob {
    int64_t mb_begin;      // of the chunk, in the mb packing
    int64_t n;             // frames
    int64_t ck_begin;      // first checkpoint slot of this chunk (seg_prefix)
    int32_t band;
    int32_t table;
    uint32_t thr_i;        // rms > thresh_rms  <=>  rms >= thr_i
    int32_t pad;
};

// k_att_chain: the strictly sequential part.  One lane per (chunk, band); stores the attenuation
// entering every kCK-frame segment so k_compress_apply can redo the segments in parallel.
__global__ void __launch_bounds__(32)
k_att_chain(const ChainJob *__restrict__ jobs, int n_jobs, const uint16_t *__restrict__ rms,
            const AttEntry *__restrict__ tables, double *__restrict__ ckpt, int64_t mb_frames, int64_t n_seg_total) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_jobs) return;
    const ChainJob job = jobs[j];
    const AttEntry *tbl = tables + (size_t)job.table * 32769;
    const uint16_t *rp = rms + (int64_t)job.band * mb_frames + job.mb_begin;
    double *ck = ckpt + (int64_t)job.band * n_seg_total + job.ck_begin;
    const unsigned thr = job.thr_i;
    double att = 0.0;
    // head: align to 8 frames (16 bytes) in the rms plane
    int64_t i = 0;
    const int64_t mis = (8 - ((job.mb_begin) & 7)) & 7;
    const int64_t head = mis < job.n ? mis : job.n;
    for (; i < head; ++i) {
        if ((i & (kCK - 1)) == 0) ck[i / kCK] = att;
        att = att_step(att, rp[i], thr, tbl);
    }
    const uint4 *vp = reinterpret_cast<const uint4 *>(rp + i);
    uint4 cur = make_uint4(0, 0, 0, 0);
    if (i + 8 <= job.n) cur = __ldg(vp);
    for (; i + 8 <= job.n; i += 8) {
        ++vp;
        uint4 nxt = make_uint4(0, 0, 0, 0);
        if (i + 16 <= job.n) nxt = __ldg(vp);
        const uint32_t w[4] = {cur.x, cur.y, cur.z, cur.w};
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (((i + k) & (kCK - 1)) == 0) ck[(i + k) / kCK] = att;
            const unsigned r = (w[k >> 1] >> ((k & 1) * 16)) & 0xffffu;
            att = att_step(att, r, thr, tbl);
        }
        cur = nxt;
    }
    for (; i < job.n; ++i) {
        if ((i & (kCK - 1)) == 0) ck[i / kCK] = att;
        att = att_step(att, rp[i], thr, tbl);
    }
}

// audioop.mul: floor(clip(x * f)) with fbound's "val < minval + 1 -> minval" rule
__device__ __forceinline__ int mul_floor(int x, double f) {
    double v = __dmul_rn((double)x, f);
    if (v > 32767.0) v = 32767.0;
    else if (v < -32767.0) v = -32768.0;
    return __double2int_rd(v);
}
__device__ __forceinline__ int sat16(int v) { return v > 32767 ? 32767 : (v < -32768 ? -32768 : v); }

// k_compress_apply: one thread per kCK-frame segment, all three bands: replay the recurrence from
// the checkpoint (bit-identical arithmetic), gain = 10^(-att/20), audioop.mul, then
// low.overlay(mid).overlay(high) = saturating adds (:309).
__global__ void __launch_bounds__(128)
k_compress_apply(const MbChunk *__restrict__ chunks, int n_chunks, int64_t n_seg_total,
                 const ame_track_params *__restrict__ tracks, const int16_t *__restrict__ bands,
                 const uint16_t *__restrict__ rms, const AttEntry *__restrict__ tables,
                 const double *__restrict__ ckpt, int16_t *__restrict__ pre, int64_t mb_frames) {
    const int64_t seg = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (seg >= n_seg_total) return;
    int lo = 0, hi = n_chunks - 1;                 // last chunk with seg_prefix <= seg
    while (lo < hi) {
        int mid = (lo + hi + 1) >> 1;
        if (chunks[mid].seg_prefix <= seg) lo = mid; else hi = mid - 1;
    }
    const MbChunk ck = chunks[lo];
    const ame_track_params *tp = tracks + ck.track;
    const int64_t f0 = (seg - ck.seg_prefix) * kCK;
    const int64_t f1 = min(f0 + (int64_t)kCK, ck.n);
    double att[3], fac[3];
    const AttEntry *tbl[3];
    unsigned thr[3];
    const uint32_t *bp[3];
    const uint16_t *rp[3];
#pragma unroll
    for (int b = 0; b < 3; ++b) {
        att[b] = ckpt[(int64_t)b * n_seg_total + seg];
        fac[b] = (att[b] != 0.0) ? exp10(-att[b] / 20.0) : 1.0;
        tbl[b] = tables + (size_t)tp->comp[b].table * 32769;
        const double t = tp->comp[b].thresh_rms;
        thr[b] = (t >= 65535.0) ? 0x7fffffffu : (unsigned)floor(t) + 1u;
        bp[b] = reinterpret_cast<const uint32_t *>(bands) + (int64_t)b * mb_frames + ck.mb_begin;
        rp[b] = rms + (int64_t)b * mb_frames + ck.mb_begin;
    }
    uint32_t *out = reinterpret_cast<uint32_t *>(pre) + ck.abs_begin;
    for (int64_t f = f0; f < f1; ++f) {
        int accl = 0, accr = 0;
#pragma unroll
        for (int b = 0; b < 3; ++b) {
            const double a = att_step(att[b], rp[b][f], thr[b], tbl[b]);
            if (a != att[b]) { att[b] = a; fac[b] = (a != 0.0) ? exp10(-a / 20.0) : 1.0; }
            const uint32_t w = __ldg(bp[b] + f);
            int l = (int16_t)(w & 0xffffu), r = (int16_t)(w >> 16);
            if (a != 0.0) { l = mul_floor(l, fac[b]); r = mul_floor(r, fac[b]); }
            accl = b ? sat16(accl + l) : l;
            accr = b ? sat16(accr + r) : r;
        }
        out[f] = (uint32_t)(uint16_t)accl | ((uint32_t)(uint16_t)accr << 16);
    }
}

// ------------------------------------------------------------------------------------------------
// k_kweight_energy: s16 -> x/32768 -> BS.1770 pre-filter + RLB (2 biquads, FP64) -> sum of squares
// per 100 ms sub-block (ebur128.c filter + gating-block sums).  K-filter state runs through the
// whole track (the reference measures the concatenated file), so warm-up may cross chunk joins.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
k_kweight_energy(const KwJob *__restrict__ jobs, int n_jobs, const ame_track_params *__restrict__ tracks,
                 const TrackDev *__restrict__ tdev, const int16_t *__restrict__ pre,
                 double *__restrict__ energy, int *__restrict__ peak) {
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    const int pair = tid >> 1;
    if (pair >= n_jobs) return;
    const int ch = tid & 1;
    const unsigned pmask = 3u << ((threadIdx.x & 31) & ~1);
    const KwJob job = jobs[pair];
    const ame_track_params *tp = tracks + job.track;
    const TrackDev td = tdev[job.track];
    const ame_biquad k0 = tp->kw[0], k1 = tp->kw[1];
    const int64_t s100 = td.s100;
    const int64_t base = tp->offset_frames;
    const int64_t t_begin = (int64_t)job.sb_begin * s100;      // relative to track
    const int64_t t_end = (int64_t)job.sb_end * s100;
    int64_t f_lo = t_begin - (int64_t)tp->warm_kw;
    if (f_lo < 0) f_lo = 0;
    double z0 = 0, z1 = 0, z2 = 0, z3 = 0, acc = 0;
    int pk = 0;
    int64_t next_end = t_begin + s100;
    int sb = job.sb_begin;
    const uint32_t *src = reinterpret_cast<const uint32_t *>(pre) + base;
    for (int64_t f = f_lo; f < t_end; ++f) {
        const uint32_t w = __ldg(src + f);
        const int xm = ch ? (int)(int16_t)(w >> 16) : (int)(int16_t)(w & 0xffffu);
        const double x = (double)xm * (1.0 / 32768.0);
        const double y = bq_step(k1, z2, z3, bq_step(k0, z0, z1, x));
        if (f >= t_begin) {
            acc = fma(y, y, acc);
            pk = max(pk, abs(xm));
            if (f + 1 == next_end) {
                const double other = __shfl_xor_sync(pmask, acc, 1);
                if (ch == 0) energy[td.sb_offset + sb] = acc + other;
                acc = 0; ++sb; next_end += s100;
            }
        }
    }
    atomicMax(peak + job.track, pk);
}

// tail frames beyond the last complete sub-block still count for the sample peak
__global__ void k_tail_peak(const ame_track_params *__restrict__ tracks, const TrackDev *__restrict__ tdev,
                            int n_tracks, const int16_t *__restrict__ pre, int *__restrict__ peak) {
    const int t = blockIdx.x;
    if (t >= n_tracks) return;
    const int64_t begin = (int64_t)tdev[t].n_sb * tdev[t].s100;
    const int16_t *p = pre + 2 * tracks[t].offset_frames;
    int pk = 0;
    for (int64_t i = 2 * begin + threadIdx.x; i < 2 * tracks[t].n_frames; i += blockDim.x) pk = max(pk, abs((int)p[i]));
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

// k_block_hist: 400 ms blocks every 100 ms -> 1000-bin histogram (absolute gate = bin floor, -70 LUFS)
__global__ void __launch_bounds__(256)
k_block_hist(const TrackDev *__restrict__ tdev, const double *__restrict__ energy, long long *__restrict__ hist) {
    __shared__ unsigned s_hist[1000];
    const int t = blockIdx.x;
    for (int i = threadIdx.x; i < 1000; i += blockDim.x) s_hist[i] = 0;
    __syncthreads();
    const TrackDev td = tdev[t];
    const double *e = energy + td.sb_offset;
    const double denom = (double)(4 * (int64_t)td.s100);
    for (int j = threadIdx.x; j + 3 < td.n_sb; j += blockDim.x) {
        const double s = (((e[j] + e[j + 1]) + e[j + 2]) + e[j + 3]) / denom;
        if (s >= c_hist_bounds[0]) atomicAdd(&s_hist[hist_index(s)], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 1000; i += blockDim.x) hist[(int64_t)t * 1000 + i] = (long long)s_hist[i];
}

// k_finalize: ebur128_gated_loudness + the linear-mode gain of af_loudnorm, one thread per track
__global__ void k_finalize(const ame_track_params *__restrict__ tracks, int n_tracks,
                           const long long *__restrict__ hist, const int *__restrict__ peak,
                           ame_track_result *__restrict__ res) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tracks) return;
    const long long *h = hist + (int64_t)t * 1000;
    ame_track_result r;
    r.input_i = -INFINITY; r.measured_i_2dp = -INFINITY; r.gain = 1.0; r.rel_threshold = 0.0;
    r.n_blocks = 0; r.normalized = 0; r.sample_peak = peak[t];
    double rel = 0.0; long long count = 0;
    for (int j = 0; j < 1000; ++j) { rel += (double)h[j] * c_hist_energy[j]; count += h[j]; }
    r.n_blocks = count;
    if (count > 0) {
        rel /= (double)count;
        rel *= 0.1;                                  // RELATIVE_GATE_FACTOR = 10^(-10/10)
        r.rel_threshold = rel;
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
