/* C twin of the per-frame loops in oracle/chain.py.  TEST INFRASTRUCTURE ONLY (see oracle/README.md).
 *
 * ame_oracle_compress() restates pydub 0.25.1 effects.compress_dynamic_range (called from the
 * reference at audio_mastering_engine.py:306-308) together with the CPython audioop routines it
 * delegates to (audioop.rms, audioop.mul).  It must stay bit-identical to
 * chain.compress_dynamic_range_py (tests/test_oracle_cport.py); build with -ffp-contract=off.
 *
 * ame_oracle_kfilter_df2() is ebur128.c's literal 4th-order direct-form-II K-weighting loop, used to
 * check that the lfilter-based restatement in chain.k_weighted() agrees with the literal form.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

/* audioop.c fbound(): clamp, then round toward minus infinity */
static int fbound(double val, double minval, double maxval) {
    if (val > maxval) val = maxval;
    else if (val < minval + 1.0) val = minval;
    val = floor(val);
    return (int)val;
}

/* in/out: interleaved stereo int16, n_frames frames.  att_out (optional): attenuation per frame. */
int ame_oracle_compress(const int16_t *in, int16_t *out, int64_t n_frames, double fs,
                        double threshold, double ratio, double attack_ms, double release_ms,
                        double *att_out) {
    const double thresh_rms = 32768.0 * pow(10.0, threshold / 20.0);
    const double attack_frames = attack_ms * (fs / 1000.0);
    const double release_frames = release_ms * (fs / 1000.0);
    const int64_t look = (int64_t)attack_frames;
    const double ln10 = log(10.0);
    double att = 0.0;
    int64_t sumsq = 0; /* exact: <= 2*960*2^30 */
    for (int64_t i = 0; i < n_frames; ++i) {
        /* window = frames [max(i-look,0), i) */
        int64_t lo = i - look; if (lo < 0) lo = 0;
        int64_t nsamp = 2 * (i - lo);
        unsigned int rms = 0;
        if (nsamp > 0) rms = (unsigned int)sqrt((double)sumsq / (double)nsamp);
        double over = 0.0;
        if (rms != 0) {
            double r = (double)rms / thresh_rms;
            if (r != 0.0) {
                double db = 20 * (log(r) / ln10); /* math.log(r, 10) */
                over = db > 0 ? db : 0.0;
            }
        }
        double max_att = (1 - (1.0 / ratio)) * over;
        double inc = max_att / attack_frames;
        double dec = max_att / release_frames;
        if ((double)rms > thresh_rms && att <= max_att) {
            att += inc;
            if (max_att < att) att = max_att;
        } else {
            att -= dec;
            if (att < 0) att = 0; /* python max(att, 0) */
        }
        int l = in[2 * i], r_ = in[2 * i + 1];
        if (att != 0.0) {
            double f = pow(10.0, -att / 20);
            l = fbound((double)l * f, -32768.0, 32767.0);
            r_ = fbound((double)r_ * f, -32768.0, 32767.0);
        }
        out[2 * i] = (int16_t)l; out[2 * i + 1] = (int16_t)r_;
        if (att_out) att_out[i] = att;
        /* slide the window: add frame i, drop frame i-look */
        sumsq += (int64_t)in[2 * i] * in[2 * i] + (int64_t)in[2 * i + 1] * in[2 * i + 1];
        if (i - look >= 0) {
            int64_t d = i - look;
            sumsq -= (int64_t)in[2 * d] * in[2 * d] + (int64_t)in[2 * d + 1] * in[2 * d + 1];
        }
    }
    return 0;
}

/* Integer rms per frame exactly as audioop.rms over frames [max(i-look,0), i). */
int ame_oracle_window_rms(const int16_t *in, uint16_t *rms_out, int64_t n_frames, int64_t look) {
    int64_t sumsq = 0;
    for (int64_t i = 0; i < n_frames; ++i) {
        int64_t lo = i - look; if (lo < 0) lo = 0;
        int64_t nsamp = 2 * (i - lo);
        unsigned int rms = 0;
        if (nsamp > 0) rms = (unsigned int)sqrt((double)sumsq / (double)nsamp);
        rms_out[i] = (uint16_t)(rms > 65535 ? 65535 : rms);
        sumsq += (int64_t)in[2 * i] * in[2 * i] + (int64_t)in[2 * i + 1] * in[2 * i + 1];
        if (i - look >= 0) {
            int64_t d = i - look;
            sumsq -= (int64_t)in[2 * d] * in[2 * d] + (int64_t)in[2 * d + 1] * in[2 * d + 1];
        }
    }
    return 0;
}

/* ebur128.c EBUR128_FILTER: direct form II, v[0..4] per channel; out = K-weighted doubles. */
int ame_oracle_kfilter_df2(const int16_t *in, double *out, int64_t n_frames, const double *b,
                           const double *a) {
    double v[2][5] = {{0}};
    for (int c = 0; c < 2; ++c) {
        for (int64_t i = 0; i < n_frames; ++i) {
            v[c][0] = (double)in[2 * i + c] / 32768.0 - a[1] * v[c][1] - a[2] * v[c][2] -
                      a[3] * v[c][3] - a[4] * v[c][4];
            out[2 * i + c] = b[0] * v[c][0] + b[1] * v[c][1] + b[2] * v[c][2] + b[3] * v[c][3] +
                             b[4] * v[c][4];
            v[c][4] = v[c][3]; v[c][3] = v[c][2]; v[c][2] = v[c][1]; v[c][1] = v[c][0];
        }
    }
    return 0;
}

/* ame_oracle_alimiter() restates FFmpeg libavfilter/af_alimiter.c filter_frame() as the reference calls it at
 * audio_mastering_engine.py:223 (level_in = level_out = 1, auto level on, asc off, latency off) together with the
 * s16 -> dbl (x / 32768) and dbl -> s16 (clip(lrint(y * 32768))) conversions ffmpeg inserts around it.
 * Must stay bit-identical to oracle/limiter.py alimiter_py (tests/test_oracle_limiter.py). */
int ame_oracle_alimiter(const int16_t *in, int16_t *out, int64_t n_frames, double fs, double limit,
                        double attack_ms, double release_ms, double *att_out) {
    const int channels = 2;
    const double attack = attack_ms / 1000.0, release = release_ms / 1000.0;
    int buffer_size = (int)(fs * attack * channels);
    buffer_size -= buffer_size % channels;
    if (buffer_size < channels) return -1;
    double *buffer = (double *)calloc((size_t)buffer_size, sizeof(double));
    double *nextdelta = (double *)calloc((size_t)buffer_size, sizeof(double));
    int *nextpos = (int *)malloc((size_t)buffer_size * sizeof(int));
    if (!buffer || !nextdelta || !nextpos) { free(buffer); free(nextdelta); free(nextpos); return -2; }
    for (int i = 0; i < buffer_size; ++i) nextpos[i] = -1;
    double att = 1.0, delta = 0.0;
    int pos = 0, nextiter = 0, nextlen = 0;
    const double level = 1.0 / limit, level_out = 1.0, level_in = 1.0;
    for (int64_t n = 0; n < n_frames; ++n) {
        double peak = 0;
        for (int c = 0; c < channels; ++c) {
            double sample = ((double)in[n * channels + c] * (1.0 / 32768.0)) * level_in;
            buffer[pos + c] = sample;
            if (fabs(sample) > peak) peak = fabs(sample);
        }
        if (peak > limit) {
            double patt = limit / peak < 1. ? limit / peak : 1.;
            double rdelta = (1.0 - patt) / (fs * release);
            double d = (limit / peak - att) / buffer_size * channels;
            int found = 0, i;
            if (d < delta) {
                delta = d;
                nextpos[0] = pos;
                nextpos[1 % buffer_size] = -1;
                nextdelta[0] = rdelta;
                nextlen = 1;
                nextiter = 0;
            } else {
                for (i = nextiter; i < nextiter + nextlen; i++) {
                    int j = i % buffer_size;
                    double ppeak = 0, pdelta;
                    for (int c = 0; c < channels; c++)
                        if (fabs(buffer[nextpos[j] + c]) > ppeak) ppeak = fabs(buffer[nextpos[j] + c]);
                    pdelta = (limit / peak - limit / ppeak) / (((buffer_size - nextpos[j] + pos) % buffer_size) / channels);
                    if (pdelta < nextdelta[j]) {
                        nextdelta[j] = pdelta;
                        found = 1;
                        break;
                    }
                }
                if (found) {
                    nextlen = i - nextiter + 1;
                    nextpos[(nextiter + nextlen) % buffer_size] = pos;
                    nextdelta[(nextiter + nextlen) % buffer_size] = rdelta;
                    nextpos[(nextiter + nextlen + 1) % buffer_size] = -1;
                    nextlen++;
                }
            }
        }
        const int b0 = (pos + channels) % buffer_size;
        peak = 0;
        for (int c = 0; c < channels; c++)
            if (fabs(buffer[b0 + c]) > peak) peak = fabs(buffer[b0 + c]);
        att += delta;
        double o[2];
        for (int c = 0; c < channels; c++) o[c] = buffer[b0 + c] * att;
        if (b0 == nextpos[nextiter]) {
            delta = nextdelta[nextiter];
            att = limit / peak;
            nextlen -= 1;
            nextpos[nextiter] = -1;
            nextiter = (nextiter + 1) % buffer_size;
        }
        if (att > 1.) { att = 1.; delta = 0.; nextiter = 0; nextlen = 0; nextpos[0] = -1; }
        if (att <= 0.) { att = 0.0000000000001; delta = (1.0 - att) / (fs * release); }
        if (att != 1. && (1. - att) < 0.0000000000001) att = 1.;
        if (delta != 0. && fabs(delta) < 0.00000000000001) delta = 0.;
        for (int c = 0; c < channels; c++) {
            double v = o[c];
            if (v < -limit) v = -limit; else if (v > limit) v = limit;
            v = v * level * level_out;
            double r = rint(v * 32768.0);
            if (r > 32767.0) r = 32767.0; else if (r < -32768.0) r = -32768.0;
            out[n * channels + c] = (int16_t)r;
        }
        if (att_out) att_out[n] = att;
        pos = (pos + channels) % buffer_size;
    }
    free(buffer); free(nextdelta); free(nextpos);
    return 0;
}
