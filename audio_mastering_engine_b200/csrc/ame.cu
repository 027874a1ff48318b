// libame host side: plan construction (job tables, workspace) and the C ABI of include/ame.h.
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <tuple>
#include <vector>

#include "ame_kernels.cuh"

using namespace ame;

namespace {

thread_local std::string g_err;

int fail(int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CU(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return fail(AME_E_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

inline int64_t align_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }

template <class T>
int upload(T **dptr, const std::vector<T> &v) {
    *dptr = nullptr;
    if (v.empty()) return AME_OK;
    cudaError_t e = cudaMalloc((void **)dptr, v.size() * sizeof(T));
    if (e != cudaSuccess) return fail(AME_E_NOMEM, "cudaMalloc(%zu) failed: %s", v.size() * sizeof(T), cudaGetErrorString(e));
    CU(cudaMemcpy(*dptr, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    return AME_OK;
}

constexpr int kTargetPairs = 148 * 8 * 16;   // lane pairs that fill 148 SMs at 8 warps each

}  // namespace

struct ame_plan {
    int device = 0;
    int n_tracks = 0;
    std::vector<ame_track_params> tracks;
    std::vector<int64_t> mb_offset;       // per track, -1 if not multiband
    std::vector<TrackDev> tdev;
    int64_t total_frames = 0;             // padded
    int64_t mb_frames = 0;                // padded
    int64_t n_sb_total = 0;
    int max_look = 0;
    int n_eq_jobs = 0, n_split_jobs = 0, n_chain_jobs = 0, n_wf_jobs = 0, n_mb_chunks = 0, n_kw_jobs = 0, n_gain_jobs = 0;
    int eq_slots = 0, split_slots = 0;
    int64_t n_seg_total = 0, n_group_total = 0;
    int eq_tile = 0, split_tile = 0, kw_tile_sb = 0;
    size_t ws_bytes = 0;
    int64_t launches = 0;
    bool any_normalize = false;
    // device
    ame_track_params *d_tracks = nullptr;
    TrackDev *d_tdev = nullptr;
    int64_t *d_mb_delta = nullptr;
    TileJob *d_eq_jobs = nullptr, *d_split_jobs = nullptr;
    ChainJob *d_chain_jobs = nullptr;
    WfJob *d_wf_jobs = nullptr;
    MbChunk *d_mb_chunks = nullptr;
    KwJob *d_kw_jobs = nullptr;
    GainJob *d_gain_jobs = nullptr;
    AttEntry *d_tables = nullptr;
    double *d_luts = nullptr;
    int n_luts = 0;
    int16_t *d_pre = nullptr, *d_bands = nullptr, *d_in = nullptr, *d_out = nullptr;
    uint16_t *d_rms = nullptr;
    double *d_ckpt = nullptr, *d_attf = nullptr, *d_energy = nullptr;
    long long *d_hist = nullptr;
    int *d_peak = nullptr;
    ame_track_result *d_results = nullptr;
    cudaStream_t io_stream = nullptr;
    // optional per-kernel CUDA-event timing (ame_plan_set_timing)
    bool timing = false;
    int t_step = -1;
    std::vector<cudaEvent_t> t_ev;        // [kMaxTimedSteps][AME_N_KERNELS][2]
    std::vector<char> t_used;             // [kMaxTimedSteps][AME_N_KERNELS]
};

constexpr int kMaxTimedSteps = 64;
static const char *const kKernelNames[AME_N_KERNELS] = {"k_eq", "k_band_split", "k_window_flag", "k_att_chain",
    "k_compress_apply", "k_kweight_energy", "k_tail_peak", "k_block_hist", "k_finalize", "k_apply_gain"};
enum { S_EQ = 0, S_SPLIT, S_FLAG, S_CHAIN, S_APPLY, S_KW, S_TAIL, S_HIST, S_FIN, S_GAIN };

static inline void t_begin(ame_plan *p, int slot, cudaStream_t s) {
    if (p->timing && p->t_step >= 0 && p->t_step < kMaxTimedSteps)
        cudaEventRecord(p->t_ev[((size_t)p->t_step * AME_N_KERNELS + slot) * 2], s);
}
static inline void t_end(ame_plan *p, int slot, cudaStream_t s) {
    if (p->timing && p->t_step >= 0 && p->t_step < kMaxTimedSteps) {
        cudaEventRecord(p->t_ev[((size_t)p->t_step * AME_N_KERNELS + slot) * 2 + 1], s);
        p->t_used[(size_t)p->t_step * AME_N_KERNELS + slot] = 1;
    }
}

namespace {

int dmalloc(ame_plan *p, void **ptr, size_t bytes) {
    *ptr = nullptr;
    if (bytes == 0) return AME_OK;
    cudaError_t e = cudaMalloc(ptr, bytes);
    if (e != cudaSuccess) return fail(AME_E_NOMEM, "cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
    p->ws_bytes += bytes;
    return AME_OK;
}

// split chunk [cb, ce) into ceil(n / T) near-equal tiles whose interior boundaries are multiples of 8
void tile_jobs(std::vector<TileJob> &out, int track, int variant, int64_t cb, int64_t ce, int64_t T) {
    const int64_t n = ce - cb;
    if (n <= 0) return;
    const int64_t k = (n + T - 1) / T;
    const int64_t t = align_up((n + k - 1) / k, 8);
    int64_t b = cb;
    while (b < ce) {
        int64_t e = (b + t) & ~(int64_t)7;
        if (e <= b) e = b + t;
        if (e > ce) e = ce;
        out.push_back(TileJob{cb, b, e, track, variant});
        b = e;
    }
}

// pad a track's job list with empty jobs to a whole number of warps (16 lane pairs)
void pad_jobs(std::vector<TileJob> &out, size_t track_first) {
    if (out.size() == track_first) return;
    TileJob d = out.back();
    d.tile_begin = d.tile_end;
    while ((out.size() - track_first) % 16) out.push_back(d);
}

// smallest tile (multiple of 8, >= min_tile) whose job count, padded per track, fits `slots` lane pairs
int64_t pick_tile(const std::vector<std::vector<int64_t>> &chunks_per_track, int64_t slots, int64_t min_tile) {
    auto count = [&](int64_t T) {
        int64_t jobs = 0;
        for (const auto &cs : chunks_per_track) {
            int64_t j = 0;
            for (int64_t n : cs) j += (n + T - 1) / T;
            jobs += j;
        }
        return jobs;
    };
    int64_t lo = min_tile, hi = min_tile;
    for (const auto &cs : chunks_per_track)
        for (int64_t n : cs) hi = std::max(hi, align_up(n, 8));
    if (count(lo) <= slots) return lo;
    while (lo < hi) {                    // count() is non-increasing in T
        const int64_t mid = align_up((lo + hi) / 2, 8);
        if (mid >= hi) break;
        if (count(mid) <= slots) hi = mid; else lo = mid + 8;
    }
    return hi;
}

int validate(const ame_track_params &t, int idx) {
    if (t.n_frames < 0 || t.offset_frames < 0 || (t.offset_frames & 7))
        return fail(AME_E_INVALID, "track %d: offset_frames must be a non-negative multiple of 8", idx);
    if (t.halo_frames < 0 || (t.halo_frames & 7) || (t.sample_rate > 0 && t.halo_frames % ((t.sample_rate + 5) / 10)))
        return fail(AME_E_INVALID, "track %d: halo_frames must be a multiple of 8 and of the 100 ms sub-block", idx);
    if (t.sample_rate < 8000 || t.sample_rate > 384000)
        return fail(AME_E_INVALID, "track %d: unsupported sample rate %d", idx, t.sample_rate);
    if (t.warm_eq < 0 || t.warm_xover < 0 || t.warm_kw < 0)
        return fail(AME_E_INVALID, "track %d: negative warm-up", idx);
    for (int s = 0; s < 4; ++s) {
        const int k = t.eq[s].kind;
        const bool shelf = (s == 0 || s == 3);
        if (k == AME_EQ_BYPASS) continue;
        if (shelf ? (k != AME_EQ_SHELF_BOOST && k != AME_EQ_SHELF_CUT) : (k != AME_EQ_PEAK))
            return fail(AME_E_INVALID, "track %d: eq stage %d has kind %d", idx, s, k);
    }
    // the kernels rely on the Butterworth numerator shape b0 * (1 +- 2 z^-1 + z^-2) that scipy.signal.butter
    // produces for every filter of the reference (sign pattern and unit-gain sections are fixed by design)
    auto shape = [&](const ame_biquad &q, double sgn, bool unit) {
        return q.b1 == sgn * 2.0 * q.b0 && q.b2 == q.b0 && (!unit || q.b0 == 1.0);
    };
    bool ok = true;
    if (t.eq[0].kind != AME_EQ_BYPASS) ok = ok && shape(t.eq[0].s[0], 1, false);
    if (t.eq[3].kind != AME_EQ_BYPASS) ok = ok && shape(t.eq[3].s[0], -1, false);
    for (int s = 1; s <= 2; ++s)
        if (t.eq[s].kind != AME_EQ_BYPASS)
            ok = ok && shape(t.eq[s].s[0], 1, false) && shape(t.eq[s].s[1], 1, true) && shape(t.eq[s].s[2], -1, true) &&
                 shape(t.eq[s].s[3], -1, true);
    if (t.flags & AME_F_MULTIBAND)
        ok = ok && shape(t.xlp[0], 1, false) && shape(t.xlp[1], 1, true) && shape(t.xhp[0], -1, false) && shape(t.xhp[1], -1, true);
    if (!ok) return fail(AME_E_UNSUPPORTED, "track %d: a filter section is not of Butterworth shape b0*(1 +- 2z^-1 + z^-2)", idx);
    if (t.flags & AME_F_MULTIBAND)
        for (int b = 0; b < 3; ++b) {
            const ame_comp_band &c = t.comp[b];
            if (!(c.thresh_rms >= 0) || !(c.attack_frames > 0) || !(c.release_frames > 0) || c.look_frames < 0 ||
                c.look_frames > 4096)
                return fail(AME_E_INVALID, "track %d: bad compressor band %d", idx, b);
        }
    return AME_OK;
}

}  // namespace

extern "C" {

int ame_abi_version(void) { return AME_ABI_VERSION; }
size_t ame_sizeof_track_params(void) { return sizeof(ame_track_params); }
size_t ame_sizeof_track_result(void) { return sizeof(ame_track_result); }
const char *ame_last_error(void) { return g_err.c_str(); }

int ame_device_count(int *count) {
    if (!count) return fail(AME_E_INVALID, "count is NULL");
    *count = 0;
    CU(cudaGetDeviceCount(count));
    return AME_OK;
}

void ame_plan_destroy(ame_plan *p) {
    if (!p) return;
    cudaSetDevice(p->device);
    void *ptrs[] = {p->d_tracks, p->d_tdev, p->d_mb_delta, p->d_eq_jobs, p->d_split_jobs, p->d_wf_jobs, p->d_mb_chunks,
                    p->d_chain_jobs, p->d_kw_jobs, p->d_gain_jobs, p->d_tables, p->d_luts, p->d_pre, p->d_bands,
                    p->d_in, p->d_out, p->d_rms, p->d_ckpt, p->d_attf, p->d_energy, p->d_hist, p->d_peak, p->d_results};
    for (void *q : ptrs)
        if (q) cudaFree(q);
    if (p->io_stream) cudaStreamDestroy(p->io_stream);
    for (cudaEvent_t e : p->t_ev) cudaEventDestroy(e);
    delete p;
}

int ame_plan_create(int device, const ame_track_params *tracks, int32_t n_tracks, const ame_plan_options *opt,
                    ame_plan **out) {
    if (!out) return fail(AME_E_INVALID, "plan out-pointer is NULL");
    *out = nullptr;
    if (!tracks || n_tracks <= 0) return fail(AME_E_INVALID, "no tracks");
    int ndev = 0;
    CU(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return fail(AME_E_CUDA, "CUDA device %d not available (%d devices)", device, ndev);
    CU(cudaSetDevice(device));
    ame_plan_options o{};
    if (opt) o = *opt;

    ame_plan *p = new ame_plan();
    p->device = device;
    p->n_tracks = n_tracks;
    p->tracks.assign(tracks, tracks + n_tracks);
    int rc = AME_OK;
    auto bail = [&](int code) { ame_plan_destroy(p); return code; };

    // ---- layout -------------------------------------------------------------------------------
    int64_t sum_frames = 0, sum_mb = 0;
    std::vector<std::pair<int64_t, int64_t>> spans;
    p->mb_offset.assign(n_tracks, -1);
    p->tdev.resize(n_tracks);
    for (int t = 0; t < n_tracks; ++t) {
        ame_track_params &tp = p->tracks[t];
        if ((rc = validate(tp, t)) != AME_OK) return bail(rc);
        const int64_t n_total = tp.halo_frames + tp.n_frames;
        spans.emplace_back(tp.offset_frames, tp.offset_frames + n_total);
        p->total_frames = std::max(p->total_frames, align_up(tp.offset_frames + n_total, 8));
        sum_frames += tp.n_frames;
        if (tp.flags & AME_F_MULTIBAND) {
            p->mb_offset[t] = p->mb_frames;
            p->mb_frames += align_up(n_total, 8);
            sum_mb += tp.n_frames;
            for (int b = 0; b < 3; ++b) p->max_look = std::max(p->max_look, tp.comp[b].look_frames);
        }
        if (tp.flags & AME_F_NORMALIZE) p->any_normalize = true;
        const int s100 = (tp.sample_rate + 5) / 10;
        p->tdev[t].s100 = s100;
        p->tdev[t].n_sb = (int)(n_total / s100);
        p->tdev[t].n_total = n_total;
        p->tdev[t].pad = 0;
        // blocks whose last sub-block lies in the halo were counted by the previous shard
        p->tdev[t].first_block = tp.halo_frames ? std::max<int>(0, (int)(tp.halo_frames / s100) - 3) : 0;
        p->tdev[t].sb_offset = p->n_sb_total;
        p->n_sb_total += p->tdev[t].n_sb;
    }
    std::sort(spans.begin(), spans.end());
    for (size_t i = 1; i < spans.size(); ++i)
        if (spans[i].first < align_up(spans[i - 1].second, 8)) return bail(fail(AME_E_INVALID, "tracks overlap in the packed buffer"));
    if (p->total_frames == 0) p->total_frames = 8;

    // ---- chunk geometry -----------------------------------------------------------------------
    std::vector<std::vector<int64_t>> chunks_all(n_tracks), chunks_mb;
    for (int t = 0; t < n_tracks; ++t) {
        const ame_track_params &tp = p->tracks[t];
        const int64_t cf = tp.chunk_frames > 0 ? tp.chunk_frames : std::max<int64_t>(tp.n_frames, 1);
        for (int64_t c0 = 0; c0 < tp.n_frames; c0 += cf) chunks_all[t].push_back(std::min(cf, tp.n_frames - c0));
        if (tp.flags & AME_F_MULTIBAND) chunks_mb.push_back(chunks_all[t]);
    }

    // ---- tile sizes: ONE wave of resident lane pairs (no tail wave), else the minimum tile ---------
    int n_sm = 148, occ_eq = 2, occ_split = 4;
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, device);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_eq, k_eq, 128, 0);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_split, k_band_split, 128, 0);
    p->eq_slots = n_sm * std::max(occ_eq, 1) * 128;      // one thread per tile
    p->split_slots = n_sm * std::max(occ_split, 1) * 128;
    constexpr int64_t kMinTile = 512;
    p->eq_tile = o.eq_tile_frames > 0 ? (int)align_up(o.eq_tile_frames, 8) : (int)pick_tile(chunks_all, p->eq_slots, kMinTile);
    p->split_tile = o.xover_tile_frames > 0 ? (int)align_up(o.xover_tile_frames, 8)
                                            : (int)pick_tile(chunks_mb, p->split_slots, kMinTile);
    if (o.kw_tile_subblocks > 0) {
        p->kw_tile_sb = o.kw_tile_subblocks;
    } else {
        // a tile of sub-blocks costs (tile + warm-up) frames: keep the warm-up below ~15 % when the batch is big
        // enough to still give every SM ~512 threads, else shrink the tile towards one sub-block
        int max_warm = 0, min_s100 = 1 << 30;
        for (int t = 0; t < n_tracks; ++t) { max_warm = std::max(max_warm, p->tracks[t].warm_kw); min_s100 = std::min(min_s100, p->tdev[t].s100); }
        const int64_t want = std::max<int64_t>(1, ((int64_t)max_warm * 6 + min_s100 - 1) / min_s100);
        const int64_t fill = std::max<int64_t>(1, p->n_sb_total / ((int64_t)n_sm * 512));
        p->kw_tile_sb = (int)std::min(want, fill);
    }

    // ---- job tables ---------------------------------------------------------------------------
    std::vector<TileJob> eq_jobs, split_jobs;
    std::vector<ChainJob> chain_jobs;
    std::vector<WfJob> wf_jobs;
    std::vector<MbChunk> mb_chunks;
    std::vector<KwJob> kw_jobs;
    std::vector<GainJob> gain_jobs;
    std::vector<int64_t> mb_delta(n_tracks, 0);
    std::vector<AttEntry> tables;
    std::map<std::tuple<double, double, double, double>, int> table_index;

    for (int t = 0; t < n_tracks; ++t) {
        ame_track_params &tp = p->tracks[t];
        int variant = 0;
        for (int s = 0; s < 4; ++s)
            if (tp.eq[s].kind != AME_EQ_BYPASS) variant |= 1 << s;
        if (tp.flags & AME_F_WARMTH) variant |= 16;
        const bool mb = (tp.flags & AME_F_MULTIBAND) != 0;
        if (mb) {
            mb_delta[t] = p->mb_offset[t] - tp.offset_frames;
            for (int b = 0; b < 3; ++b) {
                ame_comp_band &c = tp.comp[b];
                auto key = std::make_tuple(c.thresh_rms, c.coef, c.attack_frames, c.release_frames);
                auto it = table_index.find(key);
                if (it == table_index.end()) {
                    const int idx = (int)table_index.size();
                    table_index[key] = idx;
                    // pydub: (1 - 1/ratio) * max(20 * math.log(rms / thresh_rms, 10), 0); CPython's
                    // two-argument log is log(x) / log(base) on the platform libm.
                    const double ln10 = std::log(10.0);
                    tables.resize(tables.size() + 32769);
                    AttEntry *e = tables.data() + (size_t)idx * 32769;
                    for (int r = 0; r <= 32768; ++r) {
                        double over = 0.0;
                        if (r != 0) {
                            const double ratio = (double)r / c.thresh_rms;
                            if (ratio != 0.0) {
                                const double db = 20 * (std::log(ratio) / ln10);
                                over = db > 0 ? db : 0.0;
                            }
                        }
                        const double m = c.coef * over;
                        const double inc = m / c.attack_frames;
                        // tau: smallest a >= 0 with fl(a + inc) >= m (fl(a + inc) is monotone in a)
                        double tau = m - inc;
                        if (!(tau > 0)) tau = 0.0;
                        while (tau > 0 && tau + inc >= m) tau = std::nextafter(tau, -1.0);
                        while (tau + inc < m) tau = std::nextafter(tau, 1e300);
                        e[r] = AttEntry{m, inc, m / c.release_frames, tau};
                    }
                    c.table = idx;
                } else {
                    c.table = it->second;
                }
            }
        }
        int64_t c0 = 0;
        for (int64_t cn : chunks_all[t]) {
            const int64_t cb = tp.offset_frames + tp.halo_frames + c0, ce = cb + cn;
            tile_jobs(eq_jobs, t, variant, cb, ce, p->eq_tile);
            if (mb) {
                tile_jobs(split_jobs, t, 0, cb, ce, p->split_tile);
                MbChunk ck{cb, p->mb_offset[t] + tp.halo_frames + c0, cn, p->n_seg_total, {0, 0, 0}, t, 0};
                for (int b = 0; b < 3; ++b) {
                    const double thr = tp.comp[b].thresh_rms;
                    const uint32_t thr_i = thr >= 65535.0 ? 0x7fffffffu : (uint32_t)std::floor(thr) + 1u;
                    ck.ck_begin[b] = p->n_group_total;
                    chain_jobs.push_back(ChainJob{ck.mb_begin, cn, p->n_group_total, b, tp.comp[b].table, thr_i, tp.comp[b].look_frames});
                    p->n_group_total += (cn + 31) / 32;
                }
                p->n_seg_total += (cn + kSeg - 1) / kSeg;
                mb_chunks.push_back(ck);
            }
            c0 += cn;
        }
        for (int sb = 0; sb < p->tdev[t].n_sb; sb += p->kw_tile_sb)
            kw_jobs.push_back(KwJob{t, sb, std::min(sb + p->kw_tile_sb, p->tdev[t].n_sb), 0});
        for (int64_t b = 0; b < tp.n_frames; b += kGainTile)
            gain_jobs.push_back(GainJob{tp.offset_frames + tp.halo_frames + b,
                                        tp.offset_frames + tp.halo_frames + std::min<int64_t>(b + kGainTile, tp.n_frames), t, 0});
    }
    // longest chains first: the sequential compressor kernel is bounded by its slowest warp
    std::stable_sort(chain_jobs.begin(), chain_jobs.end(), [](const ChainJob &a, const ChainJob &b) { return a.n > b.n; });
    for (int c = 0; c < (int)chain_jobs.size(); ++c)
        for (int64_t tb = 0; tb < chain_jobs[c].n; tb += kWfTile) wf_jobs.push_back(WfJob{c, 0, tb});

    p->n_eq_jobs = (int)eq_jobs.size();
    p->n_split_jobs = (int)split_jobs.size();
    p->n_chain_jobs = (int)chain_jobs.size();
    p->n_wf_jobs = (int)wf_jobs.size();
    p->n_mb_chunks = (int)mb_chunks.size();
    p->n_kw_jobs = (int)kw_jobs.size();
    p->n_gain_jobs = (int)gain_jobs.size();

    // ---- device state -------------------------------------------------------------------------
    if ((rc = upload(&p->d_tracks, p->tracks)) || (rc = upload(&p->d_tdev, p->tdev)) || (rc = upload(&p->d_mb_delta, mb_delta)) ||
        (rc = upload(&p->d_eq_jobs, eq_jobs)) || (rc = upload(&p->d_split_jobs, split_jobs)) ||
        (rc = upload(&p->d_chain_jobs, chain_jobs)) || (rc = upload(&p->d_wf_jobs, wf_jobs)) || (rc = upload(&p->d_mb_chunks, mb_chunks)) ||
        (rc = upload(&p->d_kw_jobs, kw_jobs)) || (rc = upload(&p->d_gain_jobs, gain_jobs)) || (rc = upload(&p->d_tables, tables)))
        return bail(rc);
    const size_t fb = (size_t)p->total_frames * 4;
    if ((rc = dmalloc(p, (void **)&p->d_pre, fb)) || (rc = dmalloc(p, (void **)&p->d_bands, (size_t)p->mb_frames * 4 * 3)) ||
        (rc = dmalloc(p, (void **)&p->d_rms, (size_t)p->mb_frames * 2 * 3)) ||
        (rc = dmalloc(p, (void **)&p->d_ckpt, (size_t)p->n_group_total * 8)) ||
        (rc = dmalloc(p, (void **)&p->d_attf, (size_t)p->mb_frames * 8 * 3)) ||
        (rc = dmalloc(p, (void **)&p->d_energy, (size_t)std::max<int64_t>(p->n_sb_total, 1) * 8)) ||
        (rc = dmalloc(p, (void **)&p->d_hist, (size_t)n_tracks * 1000 * 8)) ||
        (rc = dmalloc(p, (void **)&p->d_peak, (size_t)n_tracks * 4)) ||
        (rc = dmalloc(p, (void **)&p->d_results, (size_t)n_tracks * sizeof(ame_track_result))))
        return bail(rc);
    if (o.host_io) {
        if ((rc = dmalloc(p, (void **)&p->d_in, fb)) || (rc = dmalloc(p, (void **)&p->d_out, fb))) return bail(rc);
        if (cudaStreamCreateWithFlags(&p->io_stream, cudaStreamNonBlocking) != cudaSuccess)
            return bail(fail(AME_E_CUDA, "cudaStreamCreate failed"));
    }
    if (cudaMemset(p->d_pre, 0, fb) != cudaSuccess) return bail(fail(AME_E_CUDA, "cudaMemset failed"));
    if (p->mb_frames && cudaMemset(p->d_bands, 0, (size_t)p->mb_frames * 12) != cudaSuccess) return bail(fail(AME_E_CUDA, "cudaMemset failed"));

    // ebur128.c histogram tables (same libm calls as the C library)
    {
        static double bounds[1001], energies[1000];
        bounds[0] = std::pow(10.0, (-70.0 + 0.691) / 10.0);
        for (int i = 0; i < 1000; ++i) energies[i] = std::pow(10.0, ((double)i / 10.0 - 69.95 + 0.691) / 10.0);
        for (int i = 1; i < 1001; ++i) bounds[i] = std::pow(10.0, ((double)i / 10.0 - 70.0 + 0.691) / 10.0);
        if (cudaMemcpyToSymbol(c_hist_bounds, bounds, sizeof bounds) != cudaSuccess ||
            cudaMemcpyToSymbol(c_hist_energy, energies, sizeof energies) != cudaSuccess)
            return bail(fail(AME_E_CUDA, "cudaMemcpyToSymbol failed: %s", cudaGetErrorString(cudaGetLastError())));
    }
    *out = p;
    return AME_OK;
}

int ame_plan_set_warm_luts(ame_plan *p, const float *luts, int32_t n_luts) {
    if (!p || !luts || n_luts <= 0) return fail(AME_E_INVALID, "bad warm lut arguments");
    CU(cudaSetDevice(p->device));
    if (p->d_luts) { cudaFree(p->d_luts); p->d_luts = nullptr; }
    // widened exactly to double on the host so the kernel needs no float->double conversion per sample
    std::vector<double> wide((size_t)n_luts * 65536);
    for (size_t i = 0; i < wide.size(); ++i) wide[i] = (double)luts[i];
    int rc = dmalloc(p, (void **)&p->d_luts, wide.size() * sizeof(double));
    if (rc) return rc;
    CU(cudaMemcpy(p->d_luts, wide.data(), wide.size() * sizeof(double), cudaMemcpyHostToDevice));
    p->n_luts = n_luts;
    return AME_OK;
}

int64_t ame_plan_total_frames(const ame_plan *p) { return p ? p->total_frames : 0; }
size_t ame_plan_workspace_bytes(const ame_plan *p) { return p ? p->ws_bytes : 0; }
int64_t ame_plan_launch_count(const ame_plan *p) { return p ? p->launches : 0; }
const int16_t *ame_plan_tap_pre(const ame_plan *p) { return p->d_pre; }
const int16_t *ame_plan_tap_bands(const ame_plan *p) { return p->d_bands; }
const double *ame_plan_tap_subblock_energy(const ame_plan *p) { return p->d_energy; }
int64_t ame_plan_mb_frames(const ame_plan *p) { return p->mb_frames; }
int64_t ame_plan_mb_offset(const ame_plan *p, int32_t t) { return (t < 0 || t >= p->n_tracks) ? -1 : p->mb_offset[t]; }
int64_t ame_plan_subblock_offset(const ame_plan *p, int32_t t) { return (t < 0 || t >= p->n_tracks) ? -1 : p->tdev[t].sb_offset; }

int ame_plan_read_device(ame_plan *p, void *h_dst, const void *d_src, size_t bytes) {
    if (!p || !h_dst || !d_src) return fail(AME_E_INVALID, "NULL argument");
    CU(cudaSetDevice(p->device));
    CU(cudaDeviceSynchronize());
    CU(cudaMemcpy(h_dst, d_src, bytes, cudaMemcpyDeviceToHost));
    return AME_OK;
}

#define LAUNCH_CHECK(p)                                                                            \
    do {                                                                                           \
        cudaError_t e_ = cudaGetLastError();                                                       \
        if (e_ != cudaSuccess) return fail(AME_E_CUDA, "kernel launch failed: %s (%s:%d)", cudaGetErrorString(e_), __FILE__, __LINE__); \
        ++(p)->launches;                                                                           \
    } while (0)

static int check_warmth(const ame_plan *p) {
    for (int t = 0; t < p->n_tracks; ++t)
        if (p->tracks[t].flags & AME_F_WARMTH) {
            const int l = p->tracks[t].warm_lut;
            if (l < 0 || l >= p->n_luts) return fail(AME_E_INVALID, "track %d needs warmth table %d but %d are set", t, l, p->n_luts);
        }
    return AME_OK;
}

int ame_stage_eq(ame_plan *p, const int16_t *d_in, int16_t *d_pre, void *stream) {
    if (!p || !d_in || !d_pre) return fail(AME_E_INVALID, "NULL argument");
    CU(cudaSetDevice(p->device));
    int rc = check_warmth(p);
    if (rc) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    if (p->n_eq_jobs) {
        const int threads = 128, blocks = (p->n_eq_jobs + threads - 1) / threads;
        t_begin(p, S_EQ, s);
        k_eq<<<blocks, threads, 0, s>>>(p->d_eq_jobs, p->n_eq_jobs, p->d_tracks, p->d_luts, d_in, d_pre);
        LAUNCH_CHECK(p);
        t_end(p, S_EQ, s);
    }
    return AME_OK;
}

int ame_stage_band_split(ame_plan *p, const int16_t *d_pre, int16_t *d_bands, void *stream) {
    if (!p || !d_pre) return fail(AME_E_INVALID, "NULL argument");
    CU(cudaSetDevice(p->device));
    cudaStream_t s = (cudaStream_t)stream;
    if (p->n_split_jobs) {
        if (!d_bands) return fail(AME_E_INVALID, "NULL bands buffer");
        const int threads = 128, blocks = (p->n_split_jobs + threads - 1) / threads;
        t_begin(p, S_SPLIT, s);
        k_band_split<<<blocks, threads, 0, s>>>(p->d_split_jobs, p->n_split_jobs, p->d_tracks, p->d_mb_delta, d_pre, d_bands, p->mb_frames);
        LAUNCH_CHECK(p);
        t_end(p, S_SPLIT, s);
    }
    return AME_OK;
}

int ame_stage_compress(ame_plan *p, int16_t *d_bands, int16_t *d_pre, void *stream) {
    if (!p || !d_pre) return fail(AME_E_INVALID, "NULL argument");
    CU(cudaSetDevice(p->device));
    cudaStream_t s = (cudaStream_t)stream;
    if (!p->n_chain_jobs) return AME_OK;
    if (!d_bands) return fail(AME_E_INVALID, "NULL bands buffer");
    t_begin(p, S_FLAG, s);
    k_window_flag<<<p->n_wf_jobs, kWfThreads, 0, s>>>(p->d_wf_jobs, p->d_chain_jobs, d_bands, p->d_rms, p->mb_frames);
    LAUNCH_CHECK(p);
    t_end(p, S_FLAG, s);
    t_begin(p, S_CHAIN, s);
    k_att_chain<<<(p->n_chain_jobs + kChainWarps - 1) / kChainWarps, kChainWarps * 32, 0, s>>>(
        p->d_chain_jobs, p->n_chain_jobs, p->d_rms, p->d_tables, p->d_ckpt, p->d_attf, p->mb_frames);
    LAUNCH_CHECK(p);
    t_end(p, S_CHAIN, s);
    t_begin(p, S_APPLY, s);
    k_compress_apply<<<(unsigned)((p->n_seg_total + 3) / 4), 128, 0, s>>>(p->d_mb_chunks, p->n_mb_chunks, p->n_seg_total, d_bands,
                                                                         p->d_rms, p->d_ckpt, p->d_attf, d_pre, p->mb_frames);
    LAUNCH_CHECK(p);
    t_end(p, S_APPLY, s);
    return AME_OK;
}

int ame_stage_loudness_hist(ame_plan *p, const int16_t *d_pre, int64_t *d_hist, void *stream) {
    if (!p || !d_pre || !d_hist) return fail(AME_E_INVALID, "NULL argument");
    CU(cudaSetDevice(p->device));
    cudaStream_t s = (cudaStream_t)stream;
    CU(cudaMemsetAsync(p->d_peak, 0, (size_t)p->n_tracks * 4, s));
    if (p->n_kw_jobs) {
        const int threads = 128, blocks = (p->n_kw_jobs + threads - 1) / threads;
        t_begin(p, S_KW, s);
        k_kweight_energy<<<blocks, threads, 0, s>>>(p->d_kw_jobs, p->n_kw_jobs, p->d_tracks, p->d_tdev, d_pre, p->d_energy, p->d_peak);
        LAUNCH_CHECK(p);
        t_end(p, S_KW, s);
    }
    t_begin(p, S_TAIL, s);
        k_tail_peak<<<p->n_tracks, 128, 0, s>>>(p->d_tracks, p->d_tdev, p->n_tracks, d_pre, p->d_peak);
    LAUNCH_CHECK(p);
        t_end(p, S_TAIL, s);
    t_begin(p, S_HIST, s);
        k_block_hist<<<p->n_tracks, 256, 0, s>>>(p->d_tdev, p->d_energy, (long long *)d_hist);
    LAUNCH_CHECK(p);
        t_end(p, S_HIST, s);
    return AME_OK;
}

int ame_stage_apply_gain(ame_plan *p, const int16_t *d_pre, const int64_t *d_hist, int16_t *d_out,
                         ame_track_result *results, void *stream) {
    if (!p || !d_pre || !d_hist || !d_out) return fail(AME_E_INVALID, "NULL argument");
    CU(cudaSetDevice(p->device));
    cudaStream_t s = (cudaStream_t)stream;
    t_begin(p, S_FIN, s);
        k_finalize<<<(p->n_tracks + 63) / 64, 64, 0, s>>>(p->d_tracks, p->n_tracks, (const long long *)d_hist, p->d_peak, p->d_results);
    LAUNCH_CHECK(p);
        t_end(p, S_FIN, s);
    if (p->n_gain_jobs) {
        t_begin(p, S_GAIN, s);
        k_apply_gain<<<p->n_gain_jobs, 256, 0, s>>>(p->d_gain_jobs, p->d_results, d_pre, d_out);
        LAUNCH_CHECK(p);
        t_end(p, S_GAIN, s);
    }
    if (results) {
        CU(cudaMemcpyAsync(results, p->d_results, (size_t)p->n_tracks * sizeof(ame_track_result), cudaMemcpyDeviceToHost, s));
        CU(cudaStreamSynchronize(s));
    }
    return AME_OK;
}

int ame_plan_set_timing(ame_plan *p, int enable) {
    if (!p) return fail(AME_E_INVALID, "NULL plan");
    CU(cudaSetDevice(p->device));
    if (enable && p->t_ev.empty()) {
        p->t_ev.resize((size_t)kMaxTimedSteps * AME_N_KERNELS * 2);
        for (auto &e : p->t_ev) CU(cudaEventCreate(&e));
    }
    p->t_used.assign((size_t)kMaxTimedSteps * AME_N_KERNELS, 0);
    p->t_step = -1;
    p->timing = enable != 0;
    return AME_OK;
}

int ame_plan_kernel_times(ame_plan *p, double *ms_sum, int64_t *launches, int *n_steps) {
    if (!p || !ms_sum || !launches) return fail(AME_E_INVALID, "NULL argument");
    CU(cudaSetDevice(p->device));
    CU(cudaDeviceSynchronize());
    const int steps = std::min(p->t_step + 1, kMaxTimedSteps);
    for (int k = 0; k < AME_N_KERNELS; ++k) { ms_sum[k] = 0; launches[k] = 0; }
    for (int st = 0; st < steps; ++st)
        for (int k = 0; k < AME_N_KERNELS; ++k)
            if (!p->t_used.empty() && p->t_used[(size_t)st * AME_N_KERNELS + k]) {
                float ms = 0;
                CU(cudaEventElapsedTime(&ms, p->t_ev[((size_t)st * AME_N_KERNELS + k) * 2], p->t_ev[((size_t)st * AME_N_KERNELS + k) * 2 + 1]));
                ms_sum[k] += ms;
                ++launches[k];
            }
    if (n_steps) *n_steps = steps;
    return AME_OK;
}

const char *ame_kernel_name(int slot) { return (slot >= 0 && slot < AME_N_KERNELS) ? kKernelNames[slot] : ""; }

int ame_measure_device(ame_plan *p, const int16_t *d_in, int64_t *d_hist, void *stream) {
    if (!p) return fail(AME_E_INVALID, "NULL plan");
    p->launches = 0;
    if (p->timing) ++p->t_step;
    int rc;
    if ((rc = ame_stage_eq(p, d_in, p->d_pre, stream))) return rc;
    if ((rc = ame_stage_band_split(p, p->d_pre, p->d_bands, stream))) return rc;
    if ((rc = ame_stage_compress(p, p->d_bands, p->d_pre, stream))) return rc;
    return ame_stage_loudness_hist(p, p->d_pre, d_hist, stream);
}

int ame_normalize_device(ame_plan *p, const int64_t *d_hist, int16_t *d_out, ame_track_result *results, void *stream) {
    if (!p) return fail(AME_E_INVALID, "NULL plan");
    return ame_stage_apply_gain(p, p->d_pre, d_hist, d_out, results, stream);
}

int ame_master_device(ame_plan *p, const int16_t *d_in, int16_t *d_out, ame_track_result *results, void *stream) {
    if (!p) return fail(AME_E_INVALID, "NULL plan");
    int rc = ame_measure_device(p, d_in, (int64_t *)p->d_hist, stream);
    if (rc) return rc;
    return ame_normalize_device(p, (const int64_t *)p->d_hist, d_out, results, stream);
}

int ame_master_host(ame_plan *p, const int16_t *h_in, int16_t *h_out, ame_track_result *results) {
    if (!p || !h_in || !h_out) return fail(AME_E_INVALID, "NULL argument");
    if (!p->d_in || !p->d_out) return fail(AME_E_INVALID, "plan was created without host_io");
    CU(cudaSetDevice(p->device));
    const size_t fb = (size_t)p->total_frames * 4;
    CU(cudaMemcpyAsync(p->d_in, h_in, fb, cudaMemcpyHostToDevice, p->io_stream));
    int rc = ame_master_device(p, p->d_in, p->d_out, nullptr, p->io_stream);
    if (rc) return rc;
    CU(cudaMemcpyAsync(h_out, p->d_out, fb, cudaMemcpyDeviceToHost, p->io_stream));
    if (results)
        CU(cudaMemcpyAsync(results, p->d_results, (size_t)p->n_tracks * sizeof(ame_track_result), cudaMemcpyDeviceToHost, p->io_stream));
    CU(cudaStreamSynchronize(p->io_stream));
    return AME_OK;
}

}  // extern "C"
