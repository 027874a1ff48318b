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
