// libame host side: plan construction (job tables, workspace, waves) and the C ABI of include/ame.h.
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <dlfcn.h>
#include <map>
#include <mutex>
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

// Every entry point works on its plan's device and leaves the caller's current device as it found it (a process may
// drive plans on several GPUs, and torch allocates on the current device).
struct DeviceGuard {
    int prev = -1;
    cudaError_t err;
    explicit DeviceGuard(int dev) {
        err = cudaGetDevice(&prev);
        if (err == cudaSuccess && prev != dev) err = cudaSetDevice(dev);
    }
    ~DeviceGuard() {
        int cur = -1;
        if (prev >= 0 && cudaGetDevice(&cur) == cudaSuccess && cur != prev) cudaSetDevice(prev);
    }
};
#define GUARD(dev)                                                                                 \
    DeviceGuard guard_(dev);                                                                       \
    if (guard_.err != cudaSuccess) return fail(AME_E_CUDA, "cannot select CUDA device %d: %s", (int)(dev), cudaGetErrorString(guard_.err))

// Device memory comes from one stream-ordered pool per device whose release threshold is lifted: the workspace of a
// destroyed plan stays in the pool and the next plan gets it back in microseconds.  (cudaMalloc + cudaFree of the
// ~25 buffers of a plan cost 50 - 150 ms, several times what mastering a 3 min track takes; profiles/r01e_summary.md.)
// Allocation and release are ordered on the legacy default stream, which every entry point that allocates
// synchronises before it returns; ame_release_cached_memory() hands the pool back to the driver.
constexpr int kMaxDevices = 64;
std::mutex g_pool_mutex;
cudaMemPool_t g_pools[kMaxDevices] = {};
bool g_pool_tried[kMaxDevices] = {};

cudaMemPool_t device_pool() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return nullptr;
    std::lock_guard<std::mutex> lock(g_pool_mutex);
    if (!g_pool_tried[dev]) {
        g_pool_tried[dev] = true;
        int ok = 0;
        cudaDeviceGetAttribute(&ok, cudaDevAttrMemoryPoolsSupported, dev);
        if (ok) {
            cudaMemPoolProps props = {};
            props.allocType = cudaMemAllocationTypePinned;
            props.handleTypes = cudaMemHandleTypeNone;
            props.location.type = cudaMemLocationTypeDevice;
            props.location.id = dev;
            cudaMemPool_t pool = nullptr;
            unsigned long long keep = ~0ull;
            if (cudaMemPoolCreate(&pool, &props) == cudaSuccess &&
                cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep) == cudaSuccess)
                g_pools[dev] = pool;
            else
                cudaGetLastError();
        }
    }
    return g_pools[dev];
}

cudaError_t dev_alloc(void **ptr, size_t bytes) {
    if (cudaMemPool_t pool = device_pool()) return cudaMallocFromPoolAsync(ptr, bytes, pool, 0);
    return cudaMalloc(ptr, bytes);
}

void dev_free(void *ptr) {
    if (!ptr) return;
    if (device_pool()) cudaFreeAsync(ptr, 0); else cudaFree(ptr);
}

template <class T>
int upload(T **dptr, const std::vector<T> &v) {
    *dptr = nullptr;
    if (v.empty()) return AME_OK;
    cudaError_t e = dev_alloc((void **)dptr, v.size() * sizeof(T));
    if (e != cudaSuccess) return fail(AME_E_NOMEM, "device allocation of %zu bytes failed: %s", v.size() * sizeof(T), cudaGetErrorString(e));
    CU(cudaMemcpy(*dptr, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    return AME_OK;
}

constexpr int kLimMaxTile = 1 << 20;     // longest limiter tile (look-ahead + release of the slowest setting must fit one)
constexpr int kMaxTimedSteps = 8;
constexpr int kMaxTimedWaves = 128;
constexpr int kTimedSlots = kMaxTimedWaves * AME_N_KERNELS;   // per step
const char *const kKernelNames[AME_N_KERNELS] = {"k_eq", "k_band_split", "k_window_flag", "k_att_chain",
    "k_compress_apply", "k_kweight_energy", "k_tail_peak", "k_block_hist", "k_finalize", "k_apply_gain", "k_limiter", "k_true_peak"};
enum { S_EQ = 0, S_SPLIT, S_FLAG, S_CHAIN, S_APPLY, S_KW, S_TAIL, S_HIST, S_FIN, S_GAIN, S_LIM, S_TP };

// A wave = a contiguous range of tracks whose jobs are contiguous in every job table.  The device path
// launches each kernel ONCE over all waves; the host path (ame_master_host) launches wave by wave so the
// H2D copy of wave w+1 and the D2H copy of wave w-1 overlap the kernels of wave w.
struct Wave {
    int track_lo = 0, track_hi = 0;
    int64_t frame_lo = 0, frame_hi = 0;            // packed-buffer range (multiples of 8 frames)
    int eq_lo = 0, eq_n = 0, split_lo = 0, split_n = 0, chain_lo = 0, chain_n = 0, wf_lo = 0, wf_n = 0;
    int chunk_lo = 0, chunk_n = 0, kw_lo = 0, kw_n = 0, gain_lo = 0, gain_n = 0, lim_lo = 0, lim_n = 0;
    int64_t seg_lo = 0, seg_hi = 0;
    int slot = 0;                                  // workspace slot (and stream) this wave runs in
    bool limiter = false;                          // a track of the wave has the limiter stage
    bool true_peak = false;                        // a track of the wave wants its oversampled peak
    int64_t mb_frames = 0, n_groups = 0;           // of the wave's own multiband packing
    bool xover_uni = false, kw_uni = false;        // all multiband tracks share the crossover / all k_kweight_energy tracks the K filter
    XoverCfg xover{};
    KwCfg kw{};
};

// Per-wave workspace.  A plan has n_slots of them; wave w runs in slot w % n_slots on that slot's stream, so a slot is
// recycled in stream order and a plan over many waves needs the intermediate buffers of only n_slots waves.
struct Slot {
    int16_t *pre = nullptr;      // pre-normalisation int16 of the wave (the signal the reference materialises at the concat)
    int16_t *bands = nullptr;    // 3 planes of mb_frames frames
    uint16_t *rms = nullptr;     // 3 planes: integer window rms of flagged frames, 0 elsewhere
    uint16_t *list = nullptr;    // 3 planes: per chain the rms values of its flagged frames, dense
    int *tile_cnt = nullptr;     // flagged frames per k_window_flag tile
    int *n_flagged = nullptr;    // flagged frames per chain
    GrpRec *grp = nullptr;       // per 32-frame group of every chain
    double *att = nullptr;       // 3 planes: attenuation after every flagged frame, dense per chain
    int16_t *norm = nullptr;     // the normalised signal of tracks with the limiter stage (k_apply_gain -> k_limiter)
    long long *lim_last = nullptr;   // per k_apply_gain tile: last frame over the limiter's limit, or -1
    LimStore lim_in{}, lim_out{};    // per tile: the limiter state it started from / ended in
    int *lim_need = nullptr;         // per tile: its start is not its predecessor's end (yet)
    cudaStream_t stream = nullptr;
};

}  // namespace

struct ame_plan {
    int device = 0;
    int n_tracks = 0;
    std::vector<ame_track_params> tracks;
    std::vector<int64_t> mb_offset;       // per track, -1 if not multiband
    std::vector<TrackDev> tdev;
    std::vector<Wave> waves;
    Wave all;                             // the union of all waves
    int64_t total_frames = 0;             // padded
    int64_t mb_frames = 0;                // padded; the largest wave's multiband packing = plane stride of every slot
    int64_t slot_frames = 0, slot_groups = 0;
    int slot_tiles = 0, slot_chains = 0;
    int64_t n_sb_total = 0;
    int eq_tile = 0, split_tile = 0, kw_tile_sb = 0;
    int n_sm = 148, chain_warps = 0;      // see chain_lanes()
    int precision = 0;                    // 1 = the FP32 EQ experiment
    size_t ws_bytes = 0;
    int64_t launches = 0;
    // device
    ame_track_params *d_tracks = nullptr;
    TrackDev *d_tdev = nullptr;
    int64_t *d_mb_delta = nullptr;
    TileJob *d_eq_jobs = nullptr, *d_split_jobs = nullptr;
    ChainJob *d_chain_jobs = nullptr;
    WfJob *d_wf_jobs = nullptr;
    MbChunk *d_mb_chunks = nullptr;
    KwJob *d_kw_jobs = nullptr;
    GainJob *d_gain_jobs = nullptr, *d_lim_jobs = nullptr;   // limiter tiles: the same (begin, end, track) records, shorter tiles
    AttEntry *d_tables = nullptr;
    double *d_luts = nullptr;
    int n_luts = 0;
    int16_t *d_in = nullptr, *d_out = nullptr;
    std::vector<Slot> slots;
    double *d_energy = nullptr;
    int *d_chain_stats = nullptr;          // per chain: {flagged steps, passes} of the last k_att_chain
    long long *d_hist = nullptr;
    int *d_hist_st = nullptr;               // short-term (3 s) histogram per track, for loudness range
    int *d_peak = nullptr;
    unsigned *d_tp = nullptr;               // per track: float bits of the oversampled peak (k_true_peak)
    bool any_limiter = false, any_tp = false;
    int lim_keep = 2;                       // queue entries a recorded limiter state holds: look-ahead frames + 2
    int *d_lim_stats = nullptr;             // open tiles after the limiter's round 0 / 1 / 2, accumulated
    int slot_lim_jobs = 0;
    ame_track_result *d_results = nullptr;
    cudaStream_t s_in = nullptr, s_out = nullptr;                       // host path: copy-in / copy-out
    std::vector<cudaEvent_t> ev_in, ev_run;                             // per wave
    // optional per-kernel CUDA-event timing (ame_plan_set_timing)
    bool timing = false;
    int t_step = -1, t_wave = 0;
    std::vector<cudaEvent_t> t_ev;        // [kMaxTimedSteps][kMaxTimedWaves][AME_N_KERNELS][2]
    std::vector<char> t_used;             // [kMaxTimedSteps][kMaxTimedWaves][AME_N_KERNELS]
    std::vector<cudaEvent_t> tl_ev;       // host-path timeline of the last timed ame_master_host: start, then per wave
    int tl_waves = 0;                     //   {H2D done, first kernel may start, kernels done, D2H done}
};

namespace {

inline bool t_on(const ame_plan *p) {
    return p->timing && p->t_step >= 0 && p->t_step < kMaxTimedSteps && p->t_wave < kMaxTimedWaves;
}
inline size_t t_slot(const ame_plan *p, int kernel) {
    return ((size_t)p->t_step * kMaxTimedWaves + p->t_wave) * AME_N_KERNELS + kernel;
}
inline void t_begin(ame_plan *p, int kernel, cudaStream_t s) {
    if (t_on(p)) cudaEventRecord(p->t_ev[t_slot(p, kernel) * 2], s);
}
inline void t_end(ame_plan *p, int kernel, cudaStream_t s) {
    if (t_on(p)) {
        cudaEventRecord(p->t_ev[t_slot(p, kernel) * 2 + 1], s);
        p->t_used[t_slot(p, kernel)] = 1;
    }
}

int dmalloc(ame_plan *p, void **ptr, size_t bytes) {
    *ptr = nullptr;
    if (bytes == 0) return AME_OK;
    cudaError_t e = dev_alloc(ptr, bytes);
    if (e != cudaSuccess) return fail(AME_E_NOMEM, "device allocation of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
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

// smallest tile (multiple of 8, >= min_tile) whose job count fits `slots` threads
int64_t pick_tile(const std::vector<int64_t> &chunks, int64_t slots, int64_t min_tile) {
    auto count = [&](int64_t T) {
        int64_t jobs = 0;
        for (int64_t n : chunks) jobs += (n + T - 1) / T;
        return jobs;
    };
    int64_t lo = min_tile, hi = min_tile;
    for (int64_t n : chunks) hi = std::max(hi, align_up(n, 8));
    if (count(lo) <= slots) return lo;
    while (lo < hi) {                    // count() is non-increasing in T
        const int64_t mid = align_up((lo + hi) / 2, 8);
        if (mid >= hi) break;
        if (count(mid) <= slots) hi = mid; else lo = mid + 8;
    }
    return hi;
}

// k_eq tiles are cost-balanced: a thread's work is about (tile + warm-up) * cost-per-frame, and the cost per
// frame differs a lot between tracks (no EQ at all ... 4 stages + warmth).  Give every track the tile length
// that makes (T + warm) * cost equal to one common budget, the smallest budget whose job count fits `slots`.
double eq_cost_per_frame(const ame_track_params &t) {
    // relative cost of one frame in each k_eq variant, fitted to launches of 96 identical tracks per variant
    // (profiles/r02/eq_variant_cost.txt, warm-up overlap taken out): 4 stages = 128, 2 shelves + 1 peak = 91, one peak 57,
    // two shelves 58, no EQ 47 (bound by its loads, not by arithmetic); warmth + 26 (+ 20 without EQ); width is free
    double c = 24.0;                                    // unpack, convert, pack, store
    if (t.eq[0].kind != AME_EQ_BYPASS) c += 16.0;       // one section + blend, two channels
    if (t.eq[1].kind != AME_EQ_BYPASS) c += 36.0;       // four sections + blend, two channels
    if (t.eq[2].kind != AME_EQ_BYPASS) c += 36.0;
    if (t.eq[3].kind != AME_EQ_BYPASS) c += 16.0;
    const bool warm = (t.flags & AME_F_WARMTH) != 0;
    if (c == 24.0) return warm ? 67.0 : 47.0;
    return c + (warm ? 26.0 : 0.0);                     // two table look-ups + the 2x2 mix in explicit FP64 ops
}

// k_eq runs one thread per tile and switches on the track's variant, so a warp that mixes VARIANTS would run
// both one after the other; tracks of the same variant can share a warp (coefficients are per-thread registers).
// Every track gets a number of jobs (threads) in proportion to its cost (frames x cost per frame); the jobs of a
// wave are then laid out variant by variant and only the end of each variant's run is padded to a whole warp.
// (Until late in round 2 every track got whole WARPS: with ~9 warps per track the rounding left the cheapest
// tracks of a 128-track launch with 14 % more work per thread than the mean, and the launch ended with them.)
int eq_variant_of(const ame_track_params &t) {
    int variant = 0;
    for (int s = 0; s < 4; ++s)
        if (t.eq[s].kind != AME_EQ_BYPASS) variant |= 1 << s;
    if (t.flags & AME_F_WARMTH) variant |= 16;
    return variant;
}

inline int eq_block() {          // 8 warps in ONE CTA per SM (experiments: AME_EQ_BLOCK=128 = two 4-warp CTAs)
    static const int b = [] { const char *e = std::getenv("AME_EQ_BLOCK"); return (e && std::atoi(e) == 128) ? 128 : 256; }();
    return b;
}

// Threads handed out = the resident set minus the padding at the end of every variant's run (a warp).  Padding the runs to
// whole 8-warp CTAs (no CTA would mix two variants) was measured too: 9.84 ms against 9.69 ms on the 128-track launch -
// the ~4 % of idle threads cost more than the ~12 mixed CTAs.
std::vector<int64_t> eq_jobs_per_track(const ame_track_params *tracks, const std::vector<std::vector<int64_t>> &chunks,
                                       const std::vector<double> &cost, int t_lo, int t_hi, int64_t slots, int64_t min_tile) {
    const int n = (int)chunks.size();
    std::vector<int64_t> jobs(n, 0), cap(n, 0), floor_jobs(n, 0);
    std::vector<double> weight(n, 0.0);
    double wsum = 0;
    unsigned variants = 0;
    int n_variants = 0;
    for (int t = t_lo; t < t_hi; ++t) {
        int64_t frames = 0, max_jobs = 0;
        for (int64_t c : chunks[t]) { frames += c; max_jobs += std::max<int64_t>(1, c / min_tile); }
        weight[t] = (double)frames * cost[t];
        cap[t] = std::max<int64_t>(1, max_jobs);
        floor_jobs[t] = std::max<int64_t>(1, (int64_t)chunks[t].size());      // tile_jobs makes one job per chunk at least
        wsum += weight[t];
        const unsigned bit = 1u << eq_variant_of(tracks[t]);
        if (!(variants & bit)) { variants |= bit; ++n_variants; }
    }
    const int64_t avail = std::max<int64_t>(32 * (int64_t)(t_hi - t_lo), slots) - 31 * (int64_t)n_variants;
    std::vector<std::pair<double, int>> rem;
    int64_t used = 0;
    for (int t = t_lo; t < t_hi; ++t) {
        const double share = wsum > 0 ? (double)avail * weight[t] / wsum : 32.0;
        jobs[t] = std::min(cap[t], std::max(floor_jobs[t], (int64_t)share));
        used += jobs[t];
        rem.emplace_back(share - (double)jobs[t], t);
    }
    std::sort(rem.begin(), rem.end(), [](const std::pair<double, int> &a, const std::pair<double, int> &b) { return a.first > b.first; });
    for (size_t i = 0; i < rem.size() && used < avail; ++i)
        if (jobs[rem[i].second] < cap[rem[i].second]) { ++jobs[rem[i].second]; ++used; }
    return jobs;
}

int validate(const ame_track_params &t, int idx) {
    if (t.n_frames < 0 || t.offset_frames < 0 || (t.offset_frames & 7))
        return fail(AME_E_INVALID, "track %d: offset_frames must be a non-negative multiple of 8", idx);
    if (t.sample_rate < 8000 || t.sample_rate > 384000)
        return fail(AME_E_INVALID, "track %d: unsupported sample rate %d", idx, t.sample_rate);
    if (t.halo_frames < 0 || (t.halo_frames & 7) || t.halo_frames % ((t.sample_rate + 5) / 10))
        return fail(AME_E_INVALID, "track %d: halo_frames must be a multiple of 8 and of the 100 ms sub-block", idx);
    if (t.warm_eq < 0 || t.warm_xover < 0 || t.warm_kw < 0)
        return fail(AME_E_INVALID, "track %d: negative warm-up", idx);
    if (t.flags & AME_F_LIMITER) {
        if (t.halo_frames)
            return fail(AME_E_UNSUPPORTED, "track %d: the limiter is one sequential state machine over the whole track and cannot "
                                           "run on a time shard; limit the gathered output instead", idx);
        if (t.lim_frames < 1 || t.lim_frames > kLimQueue - 24 || !(t.lim_limit > 0.0 && t.lim_limit <= 1.0) || !(t.lim_fs_release > 0.0) ||
            t.lim_release_frames < 1 || t.lim_frames + t.lim_release_frames + 8 > kLimMaxTile || t.lim_thr_i < 1)
            return fail(AME_E_INVALID, "track %d: bad limiter parameters (look-ahead %d frames, release %d frames)", idx, t.lim_frames,
                        t.lim_release_frames);
    }
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

// pydub: (1 - 1/ratio) * max(20 * math.log(rms / thresh_rms, 10), 0); CPython's two-argument log is
// log(x) / log(base) on the platform libm.  tau: smallest a >= 0 with fl(a + inc) >= m.
void build_att_table(AttEntry *e, const ame_comp_band &c) {
    const double ln10 = std::log(10.0);
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
        double tau = m - inc;
        if (!(tau > 0)) tau = 0.0;
        while (tau > 0 && tau + inc >= m) tau = std::nextafter(tau, -1.0);
        while (tau + inc < m) tau = std::nextafter(tau, 1e300);
        e[r] = AttEntry{m, inc, m / c.release_frames, tau};
    }
}

#define LAUNCH_CHECK(p)                                                                            \
    do {                                                                                           \
        cudaError_t e_ = cudaGetLastError();                                                       \
        if (e_ != cudaSuccess) return fail(AME_E_CUDA, "kernel launch failed: %s (%s:%d)", cudaGetErrorString(e_), __FILE__, __LINE__); \
        ++(p)->launches;                                                                           \
    } while (0)

int check_warmth(const ame_plan *p) {
    for (int t = 0; t < p->n_tracks; ++t)
        if (p->tracks[t].flags & AME_F_WARMTH) {
            const int l = p->tracks[t].warm_lut;
            if (l < 0 || l >= p->n_luts) return fail(AME_E_INVALID, "track %d needs warmth table %d but %d are set", t, l, p->n_luts);
        }
    return AME_OK;
}

// ---- stage launches over one wave ------------------------------------------------------------------
// `pre` is addressed with ABSOLUTE packed-buffer frame indices everywhere: for a slot the pointer handed in is
// slot.pre - 2 * wave.frame_lo (never dereferenced outside the wave's own range).
struct Bufs {
    const int16_t *in = nullptr;
    int16_t *pre = nullptr, *bands = nullptr, *out = nullptr;
    uint16_t *rms = nullptr, *list = nullptr;
    int *tile_cnt = nullptr, *n_flagged = nullptr;
    GrpRec *grp = nullptr;
    double *att = nullptr;
    int64_t *hist = nullptr;
    int16_t *norm = nullptr;
    long long *lim_last = nullptr;
    LimStore lim_in{}, lim_out{};
    int *lim_need = nullptr;
};

Bufs slot_bufs(ame_plan *p, const Wave &w, const int16_t *d_in, int16_t *d_out) {
    const Slot &sl = p->slots[w.slot];
    Bufs b;
    b.in = d_in; b.out = d_out;
    b.pre = sl.pre - 2 * w.frame_lo;
    b.bands = sl.bands; b.rms = sl.rms; b.list = sl.list; b.tile_cnt = sl.tile_cnt; b.n_flagged = sl.n_flagged;
    b.grp = sl.grp; b.att = sl.att;
    b.norm = sl.norm ? sl.norm - 2 * w.frame_lo : nullptr;
    b.lim_last = sl.lim_last; b.lim_in = sl.lim_in; b.lim_out = sl.lim_out; b.lim_need = sl.lim_need;
    b.hist = (int64_t *)p->d_hist;
    return b;
}

int run_eq(ame_plan *p, const Wave &w, const int16_t *d_in, int16_t *d_pre, cudaStream_t s) {
    if (!w.eq_n) return AME_OK;
    t_begin(p, S_EQ, s);
    if (p->precision == 1)
        k_eq_f32<<<(w.eq_n + 127) / 128, 128, 0, s>>>(p->d_eq_jobs + w.eq_lo, w.eq_n, p->d_tracks, p->d_luts, d_in, d_pre);
    else
        k_eq<<<(w.eq_n + eq_block() - 1) / eq_block(), eq_block(), 0, s>>>(p->d_eq_jobs + w.eq_lo, w.eq_n, p->d_tracks, p->d_luts, d_in, d_pre);
    LAUNCH_CHECK(p);
    t_end(p, S_EQ, s);
    return AME_OK;
}

int run_split(ame_plan *p, const Wave &w, const int16_t *d_pre, int16_t *d_bands, cudaStream_t s) {
    if (!w.split_n) return AME_OK;
    t_begin(p, S_SPLIT, s);
    if (w.xover_uni)
        k_band_split<true><<<(w.split_n + 127) / 128, 128, 0, s>>>(w.xover, p->d_split_jobs + w.split_lo, w.split_n, p->d_tracks, p->d_mb_delta,
                                                                   d_pre, d_bands, p->mb_frames);
    else
        k_band_split<false><<<(w.split_n + 127) / 128, 128, 0, s>>>(w.xover, p->d_split_jobs + w.split_lo, w.split_n, p->d_tracks, p->d_mb_delta,
                                                                    d_pre, d_bands, p->mb_frames);
    LAUNCH_CHECK(p);
    t_end(p, S_SPLIT, s);
    return AME_OK;
}

// Lanes (time segments) per chain of k_att_chain.  More lanes = shorter segments = cheaper passes but more of them once
// a segment is shorter than the distance after which trajectories meet.  Alone on the GPU a chain is fastest with 8
// warps (few chains) or 4; in a batch, where launches of several waves share the SMs, a 216-register lane is worth
// more as room for the other kernels than as a shorter segment: 2 warps (1024 tracks: 223-225 ms per step against
// 228 with 4 warps and 239 with 6, profiles/r02/scheduling_experiments.txt).
int chain_lanes(const ame_plan *p, int n_chains) {
    if (p->chain_warps < 0) return 1;                       // ONE lane per chain: the sequential loop
    if (p->chain_warps > 0) return 32 * p->chain_warps;
    if (n_chains * 2 <= p->n_sm) return 256;
    return (n_chains >= p->n_sm && p->slots.size() > 1) ? 64 : 128;
}

int run_compress(ame_plan *p, const Wave &w, const Bufs &b, cudaStream_t s) {
    if (!w.chain_n) return AME_OK;
    t_begin(p, S_FLAG, s);
    int *tile_base = b.tile_cnt + std::max(p->slot_tiles, 1);
    k_window_flag<<<w.wf_n, kWfThreads, 0, s>>>(p->d_wf_jobs + w.wf_lo, b.bands, b.rms, b.tile_cnt, b.grp);
    LAUNCH_CHECK(p);
    k_tile_prefix<<<w.chain_n, kWfThreads, 0, s>>>(p->d_chain_jobs + w.chain_lo, b.tile_cnt, tile_base, b.n_flagged);
    LAUNCH_CHECK(p);
    k_compact<<<(w.wf_n + kWfThreads / 32 - 1) / (kWfThreads / 32), kWfThreads, 0, s>>>(p->d_wf_jobs + w.wf_lo, w.wf_n, b.rms, b.tile_cnt,
                                                                                          tile_base, b.list, b.grp);
    LAUNCH_CHECK(p);
    t_end(p, S_FLAG, s);
    t_begin(p, S_CHAIN, s);
    const int lanes = chain_lanes(p, w.chain_n);
    k_att_chain<<<w.chain_n, std::max(lanes, 32), 0, s>>>(p->d_chain_jobs + w.chain_lo, b.list, b.n_flagged, p->d_tables, b.att, p->mb_frames,
                                            lanes, p->d_chain_stats + 2 * (size_t)w.chain_lo);
    LAUNCH_CHECK(p);
    t_end(p, S_CHAIN, s);
    t_begin(p, S_APPLY, s);
    k_compress_apply<<<(unsigned)((w.seg_hi - w.seg_lo + 3) / 4), 128, 0, s>>>(p->d_mb_chunks + w.chunk_lo, w.chunk_n, w.seg_lo, w.seg_hi,
                                                                                b.bands, b.grp, b.att, b.pre, p->mb_frames);
    LAUNCH_CHECK(p);
    t_end(p, S_APPLY, s);
    return AME_OK;
}

int run_hist(ame_plan *p, const Wave &w, const int16_t *d_pre, int64_t *d_hist, cudaStream_t s) {
    const int nt = w.track_hi - w.track_lo;
    if (nt <= 0) return AME_OK;
    CU(cudaMemsetAsync(p->d_peak + w.track_lo, 0, (size_t)nt * 4, s));
    if (w.kw_n) {
        t_begin(p, S_KW, s);
        if (w.kw_uni)
            k_kweight_energy<true><<<(w.kw_n + 127) / 128, 128, 0, s>>>(w.kw, p->d_kw_jobs + w.kw_lo, w.kw_n, p->d_tracks, p->d_tdev, d_pre,
                                                                        p->d_energy, p->d_peak);
        else
            k_kweight_energy<false><<<(w.kw_n + 127) / 128, 128, 0, s>>>(w.kw, p->d_kw_jobs + w.kw_lo, w.kw_n, p->d_tracks, p->d_tdev, d_pre,
                                                                         p->d_energy, p->d_peak);
        LAUNCH_CHECK(p);
        t_end(p, S_KW, s);
    }
    t_begin(p, S_TAIL, s);
    k_tail_peak<<<nt, 128, 0, s>>>(p->d_tracks, p->d_tdev, w.track_lo, w.track_hi, d_pre, p->d_peak);
    LAUNCH_CHECK(p);
    t_end(p, S_TAIL, s);
    CU(cudaMemsetAsync(p->d_tp + w.track_lo, 0, (size_t)nt * 4, s));
    if (w.true_peak && w.gain_n) {
        t_begin(p, S_TP, s);
        k_true_peak<<<w.gain_n, 256, 0, s>>>(p->d_gain_jobs + w.gain_lo, p->d_tracks, d_pre, p->d_tp);
        LAUNCH_CHECK(p);
        t_end(p, S_TP, s);
    }
    t_begin(p, S_HIST, s);
    k_block_hist<<<nt, 256, 0, s>>>(p->d_tdev, w.track_lo, p->d_energy, (long long *)d_hist, p->d_hist_st);
    LAUNCH_CHECK(p);
    t_end(p, S_HIST, s);
    return AME_OK;
}

// static gain (+ the limiter's input and per-tile "last frame over the limit") and the limiter; needs d_results
int run_apply(ame_plan *p, const Wave &w, const int16_t *d_pre, int16_t *d_out, const Bufs &b, cudaStream_t s) {
    int16_t *d_norm = b.norm;
    long long *lim_last = b.lim_last;
    if (!w.gain_n) return AME_OK;
    t_begin(p, S_GAIN, s);
    if (w.limiter) CU(cudaMemsetAsync(lim_last, 0xff, (size_t)w.lim_n * sizeof(long long), s));      // -1: no frame over the limit
    k_apply_gain<<<w.gain_n, 256, 0, s>>>(p->d_gain_jobs + w.gain_lo, p->d_results, p->d_tracks, p->d_tdev, d_pre, d_out, d_norm, lim_last);
    LAUNCH_CHECK(p);
    t_end(p, S_GAIN, s);
    if (w.limiter) {
        t_begin(p, S_LIM, s);
        const GainJob *gj = p->d_lim_jobs + w.lim_lo;
        constexpr int kLimRounds = 2;                 // repair rounds before the sequential fallback
        int cap = 64;
        while (cap < p->lim_keep + 2) cap *= 2;       // queue capacity in shared memory: a power of two
        const size_t smem = lim_smem_bytes(cap);
        for (int round = 0; round <= kLimRounds; ++round) {
            if (round == 0) {                         // the elementwise tiles, 256 threads each
                k_limiter<<<w.lim_n, 256, smem, s>>>(gj, w.lim_n, lim_last, p->d_tracks, d_norm, d_out, b.lim_in, b.lim_out, b.lim_need, 0, cap);
                LAUNCH_CHECK(p);
            }
            k_limiter<<<w.lim_n, 32, smem, s>>>(gj, w.lim_n, lim_last, p->d_tracks, d_norm, d_out, b.lim_in, b.lim_out, b.lim_need, round, cap);
            LAUNCH_CHECK(p);
            k_lim_verify<<<(w.lim_n + 3) / 4, 128, 0, s>>>(gj, w.lim_n, p->d_tracks, b.lim_in, b.lim_out, b.lim_need,
                                                             p->d_lim_stats + round);
            LAUNCH_CHECK(p);
        }
        k_lim_fallback<<<w.lim_n, 32, smem, s>>>(gj, w.lim_n, p->d_tracks, d_norm, d_out, b.lim_in, b.lim_out, b.lim_need, cap);
        LAUNCH_CHECK(p);
        t_end(p, S_LIM, s);
    }
    return AME_OK;
}

int run_gain(ame_plan *p, const Wave &w, const int16_t *d_pre, const int64_t *d_hist, int16_t *d_out, const Bufs &b, cudaStream_t s) {
    const int nt = w.track_hi - w.track_lo;
    if (nt <= 0) return AME_OK;
    t_begin(p, S_FIN, s);
    k_finalize<<<nt, 128, 0, s>>>(p->d_tracks, w.track_lo, w.track_hi, (const long long *)d_hist, p->d_hist_st, p->d_peak,
                                             p->d_tp, p->d_results);
    LAUNCH_CHECK(p);
    t_end(p, S_FIN, s);
    return run_apply(p, w, d_pre, d_out, b, s);
}

int run_measure(ame_plan *p, const Wave &w, const Bufs &b, cudaStream_t s) {
    int rc;
    if ((rc = run_eq(p, w, b.in, b.pre, s))) return rc;
    if ((rc = run_split(p, w, b.pre, b.bands, s))) return rc;
    if ((rc = run_compress(p, w, b, s))) return rc;
    return run_hist(p, w, b.pre, b.hist, s);
}

int run_chain_of_stages(ame_plan *p, const Wave &w, const Bufs &b, cudaStream_t s) {
    int rc = run_measure(p, w, b, s);
    if (rc) return rc;
    return run_gain(p, w, b.pre, b.hist, b.out, b, s);
}

// after a failure in the middle of a pipelined call nothing may still be reading the caller's buffers when we return
int drain(ame_plan *p, int rc) {
    (void)p;
    if (rc != AME_OK) cudaDeviceSynchronize();
    return rc;
}

}  // namespace

extern "C" {

int ame_abi_version(void) { return AME_ABI_VERSION; }
size_t ame_sizeof_track_params(void) { return sizeof(ame_track_params); }
size_t ame_sizeof_track_result(void) { return sizeof(ame_track_result); }
size_t ame_sizeof_plan_options(void) { return sizeof(ame_plan_options); }
const char *ame_last_error(void) { return g_err.c_str(); }
const char *ame_kernel_name(int slot) { return (slot >= 0 && slot < AME_N_KERNELS) ? kKernelNames[slot] : ""; }

int ame_device_count(int *count) {
    if (!count) return fail(AME_E_INVALID, "count is NULL");
    *count = 0;
    CU(cudaGetDeviceCount(count));
    return AME_OK;
}

void ame_plan_destroy(ame_plan *p) {
    if (!p) return;
    DeviceGuard guard(p->device);
    void *ptrs[] = {p->d_tracks, p->d_tdev, p->d_mb_delta, p->d_eq_jobs, p->d_split_jobs, p->d_wf_jobs, p->d_mb_chunks,
                    p->d_chain_jobs, p->d_kw_jobs, p->d_gain_jobs, p->d_lim_jobs, p->d_tables, p->d_luts,
                    p->d_in, p->d_out, p->d_energy, p->d_hist, p->d_hist_st, p->d_peak, p->d_tp, p->d_lim_stats, p->d_results,
                    p->d_chain_stats};
    cudaDeviceSynchronize();            // nothing of this plan may still be running when its memory goes back to the pool
    for (void *q : ptrs) dev_free(q);
    for (Slot &sl : p->slots) {
        for (void *q : {(void *)sl.pre, (void *)sl.bands, (void *)sl.rms, (void *)sl.list, (void *)sl.tile_cnt, (void *)sl.n_flagged,
                        (void *)sl.grp, (void *)sl.att, (void *)sl.norm, (void *)sl.lim_last, (void *)sl.lim_in.st, (void *)sl.lim_in.qframe,
                        (void *)sl.lim_in.qdelta, (void *)sl.lim_out.st, (void *)sl.lim_out.qframe, (void *)sl.lim_out.qdelta,
                        (void *)sl.lim_need})
            dev_free(q);
        if (sl.stream) cudaStreamDestroy(sl.stream);
    }
    for (cudaStream_t s : {p->s_in, p->s_out})
        if (s) cudaStreamDestroy(s);
    for (cudaEvent_t e : p->ev_in) cudaEventDestroy(e);
    for (cudaEvent_t e : p->ev_run) cudaEventDestroy(e);
    for (cudaEvent_t e : p->t_ev) cudaEventDestroy(e);
    for (cudaEvent_t e : p->tl_ev) cudaEventDestroy(e);
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
    GUARD(device);
    ame_plan_options o{};
    if (opt) o = *opt;

    ame_plan *p = new ame_plan();
    p->device = device;
    p->n_tracks = n_tracks;
    p->tracks.assign(tracks, tracks + n_tracks);
    int rc = AME_OK;
    auto bail = [&](int code) { ame_plan_destroy(p); return code; };

    // ---- layout -------------------------------------------------------------------------------
    std::vector<std::pair<int64_t, int64_t>> spans;
    p->mb_offset.assign(n_tracks, -1);
    p->tdev.resize(n_tracks);
    int64_t sum_frames = 0;
    for (int t = 0; t < n_tracks; ++t) {
        ame_track_params &tp = p->tracks[t];
        if ((rc = validate(tp, t)) != AME_OK) return bail(rc);
        const int64_t n_total = tp.halo_frames + tp.n_frames;
        if (t && tp.offset_frames < p->tracks[t - 1].offset_frames)
            return bail(fail(AME_E_INVALID, "tracks must be ordered by offset_frames"));
        spans.emplace_back(tp.offset_frames, tp.offset_frames + n_total);
        p->total_frames = std::max(p->total_frames, align_up(tp.offset_frames + n_total, 8));
        sum_frames += tp.n_frames;
        const int s100 = (tp.sample_rate + 5) / 10;
        p->tdev[t].s100 = s100;
        p->tdev[t].n_sb = (int)(n_total / s100);
        p->tdev[t].n_total = n_total;
        // blocks whose last sub-block lies in the halo were counted by the previous shard
        p->tdev[t].first_block = tp.halo_frames ? std::max<int>(0, (int)(tp.halo_frames / s100) - 3) : 0;
        p->tdev[t].sb_offset = p->n_sb_total;
        p->n_sb_total += p->tdev[t].n_sb;
        p->tdev[t].pad = 0;
        p->tdev[t].lim_tile0 = 0;
        p->tdev[t].lim_shift = 15;
        p->any_limiter = p->any_limiter || (tp.flags & AME_F_LIMITER);
        if (tp.flags & AME_F_LIMITER) p->lim_keep = std::max(p->lim_keep, tp.lim_frames + 2);
        p->any_tp = p->any_tp || (tp.flags & AME_F_TRUE_PEAK);
    }
    for (size_t i = 1; i < spans.size(); ++i)
        if (spans[i].first < align_up(spans[i - 1].second, 8)) return bail(fail(AME_E_INVALID, "tracks overlap in the packed buffer"));
    if (p->total_frames == 0) p->total_frames = 8;

    // ---- waves: contiguous track ranges with about equal frames ----------------------------------
    int n_waves = o.n_waves > 0 ? o.n_waves : 1;
    n_waves = std::max(1, std::min(n_waves, n_tracks));
    {
        // With many waves the FIRST wave is a single track: in the host path nothing can be copied back before the
        // first wave has been copied in and mastered, so a small first wave starts the D2H stream early.
        p->waves.resize(n_waves);
        const bool small_first = o.host_io && n_waves >= 4 && n_tracks >= 2 * n_waves;
        int t = 0;
        int64_t acc = 0;
        const int64_t first_frames = small_first ? p->tracks[0].n_frames : 0;
        for (int w = 0; w < n_waves; ++w) {
            Wave &wv = p->waves[w];
            wv.track_lo = t;
            const int must_leave = n_waves - 1 - w;                 // at least one track for every later wave
            if (small_first && w == 0) {
                acc += p->tracks[t++].n_frames;
            } else {
                const int64_t share = small_first ? w : w + 1;
                const int64_t parts = small_first ? n_waves - 1 : n_waves;
                const int64_t target = first_frames + (sum_frames - first_frames) * share / parts;
                while (t < n_tracks - must_leave && (t == wv.track_lo || acc + p->tracks[t].n_frames / 2 <= target)) acc += p->tracks[t++].n_frames;
            }
            if (w == n_waves - 1) t = n_tracks;
            wv.track_hi = t;
            wv.frame_lo = p->tracks[wv.track_lo].offset_frames;
            wv.frame_hi = (wv.track_hi < n_tracks) ? p->tracks[wv.track_hi].offset_frames : p->total_frames;
        }
    }
    // Slots: a wave's intermediate buffers (pre-normalisation signal, bands, rms, attenuations) live in slot
    // w % n_slots and are recycled in stream order, so a batch far larger than the workspace can be planned.
    int n_slots = o.n_slots > 0 ? o.n_slots : std::min(n_waves, 4);
    n_slots = std::max(1, std::min(n_slots, n_waves));
    for (int w = 0; w < n_waves; ++w) {
        Wave &wv = p->waves[w];
        wv.slot = w % n_slots;
        int64_t mb = 0;
        for (int t = wv.track_lo; t < wv.track_hi; ++t)
            if (p->tracks[t].flags & AME_F_MULTIBAND) {
                p->mb_offset[t] = mb;                               // within the wave's own multiband packing
                mb += align_up(p->tdev[t].n_total, 8);
            }
        wv.mb_frames = mb;
        p->mb_frames = std::max(p->mb_frames, mb);
        p->slot_frames = std::max(p->slot_frames, wv.frame_hi - wv.frame_lo);
    }

    // ---- chunk geometry, tile sizes -----------------------------------------------------------------
    // Tiles are sized so that ONE launch fills its share of the machine with one wave of resident threads (no tail
    // wave).  Waves in different slots run concurrently, so a launch gets 1 / min(n_slots, 4) of the resident threads.
    std::vector<std::vector<int64_t>> chunks_all(n_tracks);
    for (int t = 0; t < n_tracks; ++t) {
        const ame_track_params &tp = p->tracks[t];
        const int64_t cf = tp.chunk_frames > 0 ? tp.chunk_frames : std::max<int64_t>(tp.n_frames, 1);
        for (int64_t c0 = 0; c0 < tp.n_frames; c0 += cf) chunks_all[t].push_back(std::min(cf, tp.n_frames - c0));
    }
    int n_sm = 148, occ_eq = 2, occ_split = 3;
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, device);
    p->n_sm = n_sm;
    p->chain_warps = std::max(-1, std::min(o.chain_warps, kChainMaxThreads / 32));
    p->precision = o.precision == 1 ? 1 : 0;
    if (p->precision == 1) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_eq, k_eq_f32, 128, 0);
    else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_eq, k_eq, 128, 0);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_split, k_band_split<true>, 128, 0);
    if (const char *e = std::getenv("AME_EQ_CTAS_PER_SM")) occ_eq = std::max(1, std::atoi(e));       // experiments
    if (const char *e = std::getenv("AME_SPLIT_CTAS_PER_SM")) occ_split = std::max(1, std::atoi(e));
    const int share = std::min(n_slots, 4);
    const int64_t eq_slots = (int64_t)n_sm * std::max(occ_eq, 1) * 128 / share;         // one thread per tile
    const int64_t split_slots = (int64_t)n_sm * std::max(occ_split, 1) * 128 / share;
    constexpr int64_t kMinTile = 512;
    int64_t split_tile = kMinTile;
    std::vector<double> eq_cost(n_tracks);
    std::vector<int64_t> eq_want(n_tracks, 32);
    int max_warm_kw = 0, min_s100 = 1 << 30;
    for (int t = 0; t < n_tracks; ++t) {
        eq_cost[t] = eq_cost_per_frame(p->tracks[t]);
        max_warm_kw = std::max(max_warm_kw, p->tracks[t].warm_kw);
        min_s100 = std::min(min_s100, p->tdev[t].s100);
    }
    const int64_t n_sb_kw = p->n_sb_total;
    // One warp per scheduler already gives these FP64-bound kernels ~77 % of the throughput of two (k_eq alone: 13.7 vs
    // 10.6 ms), and every tile pays its warm-up once: when a wave is so small that a full set of resident threads would
    // make the tiles shorter than ~1.2 x the warm-up (one long 192 kHz track), half the threads do the job sooner -
    // (2T + W) / (2T * 0.77) < (T + W) / T  <=>  T < 1.17 W.
    auto fewer_threads = [](int64_t slots, double frames, double warm_weighted) {
        const double tile = frames / (double)slots, warm = frames > 0 ? warm_weighted / frames : 0.0;
        if (std::getenv("AME_FULL_THREADS")) return slots;      // experiments
        return (tile < 1.17 * warm && slots >= 512) ? slots / 2 : slots;
    };
    for (const Wave &wv : p->waves) {
        std::vector<int64_t> cm;
        double eq_frames = 0, eq_warm = 0, mb_fr = 0, mb_warm = 0;
        for (int t = wv.track_lo; t < wv.track_hi; ++t) {
            const ame_track_params &tp = p->tracks[t];
            eq_frames += (double)tp.n_frames;
            eq_warm += (double)tp.n_frames * tp.warm_eq;
            if (tp.flags & AME_F_MULTIBAND) {
                cm.insert(cm.end(), chunks_all[t].begin(), chunks_all[t].end());
                mb_fr += (double)tp.n_frames;
                mb_warm += (double)tp.n_frames * tp.warm_xover;
            }
        }
        split_tile = std::max(split_tile, pick_tile(cm, fewer_threads(split_slots, mb_fr, mb_warm), kMinTile));
        const std::vector<int64_t> tw = eq_jobs_per_track(p->tracks.data(), chunks_all, eq_cost, wv.track_lo, wv.track_hi,
                                                          fewer_threads(eq_slots, eq_frames, eq_warm), kMinTile);
        for (int t = wv.track_lo; t < wv.track_hi; ++t) eq_want[t] = tw[t];
    }
    p->eq_tile = 0;
    p->split_tile = o.xover_tile_frames > 0 ? (int)align_up(o.xover_tile_frames, 8) : (int)split_tile;
    if (o.kw_tile_subblocks > 0) {
        p->kw_tile_sb = o.kw_tile_subblocks;
    } else if (n_sb_kw > 0) {
        // a tile of sub-blocks costs (tile + warm-up) frames: keep the warm-up below ~15 % when the batch is big
        // enough to still give every SM ~512 threads per launch, else shrink the tile towards one sub-block
        const int64_t want = std::max<int64_t>(1, ((int64_t)max_warm_kw * 6 + min_s100 - 1) / min_s100);
        const int64_t fill = std::max<int64_t>(1, n_sb_kw / n_waves * share / ((int64_t)n_sm * 512));
        p->kw_tile_sb = (int)std::min(want, fill);
    } else {
        p->kw_tile_sb = 1;
    }

    // ---- job tables (every table is ordered by track, hence contiguous per wave) ---------------------
    std::vector<TileJob> eq_jobs, split_jobs;
    std::vector<ChainJob> chain_jobs;
    std::vector<WfJob> wf_jobs;
    std::vector<MbChunk> mb_chunks;
    std::vector<KwJob> kw_jobs;
    std::vector<GainJob> gain_jobs, lim_jobs;
    std::vector<int64_t> mb_delta(n_tracks, 0);
    std::vector<AttEntry> tables;
    std::map<std::tuple<double, double, double, double>, int> table_index;
    int64_t n_seg_total = 0;

    for (Wave &wv : p->waves) {
        wv.eq_lo = (int)eq_jobs.size(); wv.split_lo = (int)split_jobs.size(); wv.chain_lo = (int)chain_jobs.size();
        wv.chunk_lo = (int)mb_chunks.size(); wv.kw_lo = (int)kw_jobs.size(); wv.gain_lo = (int)gain_jobs.size();
        wv.lim_lo = (int)lim_jobs.size();
        wv.seg_lo = n_seg_total;
        int64_t n_groups = 0;                          // group records of this wave (slot-local indices)
        std::map<int, std::vector<TileJob>> eq_bucket;   // k_eq jobs of this wave by variant (a warp runs ONE variant)
        for (int t = wv.track_lo; t < wv.track_hi; ++t) {
            ame_track_params &tp = p->tracks[t];
            const int variant = eq_variant_of(tp);
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
                        tables.resize(tables.size() + 32769);
                        build_att_table(tables.data() + (size_t)idx * 32769, c);
                        c.table = idx;
                    } else {
                        c.table = it->second;
                    }
                }
            }
            // k_eq jobs of this track: eq_want[t] jobs spread over the chunks by length (or the requested tile), collected
            // per variant and laid out at the end of the wave
            {
                std::vector<TileJob> &bucket = eq_bucket[variant];
                int64_t c0e = 0, jobs_before = 0;
                int64_t nf = 0;
                for (int64_t cn : chunks_all[t]) nf += cn;
                nf = std::max<int64_t>(nf, 1);
                const int64_t want = eq_want[t];
                for (int64_t cn : chunks_all[t]) {
                    const int64_t cb = tp.offset_frames + tp.halo_frames + c0e, ce = cb + cn;
                    int64_t T;
                    if (o.eq_tile_frames > 0) {
                        T = align_up(o.eq_tile_frames, 8);
                    } else {
                        // cumulative rounding: the chunks together get exactly `want` jobs (never more)
                        const int64_t upto = (int64_t)((__int128)want * (c0e + cn) / nf);
                        const int64_t jc = std::max<int64_t>(1, upto - jobs_before);
                        jobs_before = upto;
                        T = std::max<int64_t>(kMinTile, align_up((cn + jc - 1) / jc, 8));
                    }
                    tile_jobs(bucket, t, variant, cb, ce, T);
                    p->eq_tile = std::max<int>(p->eq_tile, (int)std::min<int64_t>(T, INT32_MAX));
                    c0e += cn;
                }
            }
            int64_t c0 = 0;
            for (int64_t cn : chunks_all[t]) {
                const int64_t cb = tp.offset_frames + tp.halo_frames + c0, ce = cb + cn;
                if (mb) {
                    tile_jobs(split_jobs, t, 0, cb, ce, p->split_tile);
                    MbChunk ck{cb, p->mb_offset[t] + tp.halo_frames + c0, cn, n_seg_total, {0, 0, 0}, t, 0};
                    for (int b = 0; b < 3; ++b) {
                        const double thr = tp.comp[b].thresh_rms;
                        const uint32_t thr_i = thr >= 65535.0 ? 0x7fffffffu : (uint32_t)std::floor(thr) + 1u;
                        ck.grp_begin[b] = n_groups;
                        chain_jobs.push_back(ChainJob{ck.mb_begin, cn, n_groups, b, tp.comp[b].table, thr_i, tp.comp[b].look_frames, 0, 0});
                        n_groups += (cn + 31) / 32;
                    }
                    n_seg_total += (cn + kSeg - 1) / kSeg;
                    mb_chunks.push_back(ck);
                }
                c0 += cn;
            }
            for (int sb = 0; sb < p->tdev[t].n_sb; sb += p->kw_tile_sb)
                kw_jobs.push_back(KwJob{t, sb, std::min(sb + p->kw_tile_sb, p->tdev[t].n_sb), 0});
            for (int64_t b = 0; b < tp.n_frames; b += kGainTile)
                gain_jobs.push_back(GainJob{tp.offset_frames + tp.halo_frames + b,
                                            tp.offset_frames + tp.halo_frames + std::min<int64_t>(b + kGainTile, tp.n_frames), t, 0});
            if (tp.flags & AME_F_LIMITER) {
                // limiter tiles: a power of two >= G = look-ahead + release + 8 (a tile without an over-limit frame then
                // guarantees the initial state at its end), as short as that allows: more tiles = more warps in flight
                const int64_t G = (int64_t)tp.lim_frames + tp.lim_release_frames + 8;
                int shift = 13;
                while ((1LL << shift) < G) ++shift;
                p->tdev[t].lim_shift = shift;
                p->tdev[t].lim_tile0 = (int)lim_jobs.size() - wv.lim_lo;
                for (int64_t b = 0; b < tp.n_frames; b += 1LL << shift)
                    lim_jobs.push_back(GainJob{tp.offset_frames + tp.halo_frames + b,
                                               tp.offset_frames + tp.halo_frames + std::min<int64_t>(b + (1LL << shift), tp.n_frames), t, 0});
            }
        }
        {
            bool have_x = false, have_k = false;
            wv.xover_uni = wv.kw_uni = true;
            for (int t = wv.track_lo; t < wv.track_hi; ++t) {
                const ame_track_params &tp = p->tracks[t];
                if (tp.flags & AME_F_MULTIBAND) {
                    const XoverCfg x{tp.xlp[0].b0, tp.xlp[0].a1, tp.xlp[0].a2, tp.xlp[1].a1, tp.xlp[1].a2,
                                     tp.xhp[0].b0, tp.xhp[0].a1, tp.xhp[0].a2, tp.xhp[1].a1, tp.xhp[1].a2};
                    if (!have_x) { wv.xover = x; have_x = true; }
                    else if (std::memcmp(&x, &wv.xover, sizeof x)) wv.xover_uni = false;
                }
                {
                    const KwCfg k{tp.kw[0].b0, tp.kw[0].b1, tp.kw[0].b2, tp.kw[0].a1, tp.kw[0].a2, tp.kw[1].a1, tp.kw[1].a2};
                    const bool rlb = tp.kw[1].b0 == 1.0 && tp.kw[1].b1 == -2.0 && tp.kw[1].b2 == 1.0;
                    if (!have_k) { wv.kw = k; have_k = true; }
                    else if (std::memcmp(&k, &wv.kw, sizeof k)) wv.kw_uni = false;
                    if (!rlb) wv.kw_uni = false;
                }
            }
        }
        wv.n_groups = n_groups;
        p->slot_groups = std::max(p->slot_groups, n_groups);
        // longest chains first inside the wave: the launch ends with its slowest CTA
        std::stable_sort(chain_jobs.begin() + wv.chain_lo, chain_jobs.end(), [](const ChainJob &a, const ChainJob &b) { return a.n > b.n; });
        wv.wf_lo = (int)wf_jobs.size();
        for (int c = wv.chain_lo; c < (int)chain_jobs.size(); ++c) {
            chain_jobs[c].tile0 = (int)wf_jobs.size() - wv.wf_lo;
            const ChainJob &cj = chain_jobs[c];
            for (int64_t tb = 0; tb < cj.n; tb += kWfTile)
                wf_jobs.push_back(WfJob{tb, (int64_t)cj.band * p->mb_frames + cj.mb_begin, cj.n, cj.grp_begin, c, cj.look, cj.thr_i, cj.tile0});
        }
        {   // every variant's run is padded to whole warps (a warp runs ONE variant) and the runs follow one another; with one
            // 8-warp CTA per SM nearly every SM then executes ONE variant - each variant is its own ~19 KB unrolled loop in
            // the instruction cache (two 4-warp CTAs of different variants per SM: 10.5 ms on the 128-track launch, one 8-warp
            // CTA: 9.7 ms).  Interleaving the variants' warps in proportion (every SM gets the wave's mix of FP64-heavy and
            // load-bound warps; AME_EQ_INTERLEAVE) gave 18.0 ms: four loops per CTA (profiles/r02/summary.md).
            std::vector<std::pair<double, std::pair<int, int>>> order;     // (position in [0, 1), (variant, warp of the variant))
            for (auto &kv : eq_bucket) {
                std::vector<TileJob> &b = kv.second;
                if (b.empty()) continue;
                TileJob d = b.back();
                d.tile_begin = d.tile_end;
                while (b.size() % 32) b.push_back(d);
                const int nw = (int)(b.size() / 32);
                for (int w = 0; w < nw; ++w) order.push_back({(w + 0.5) / nw, {kv.first, w}});
            }
            if (std::getenv("AME_EQ_INTERLEAVE")) std::stable_sort(order.begin(), order.end());   // experiment (see above): lost
            for (const auto &o2 : order) {
                const std::vector<TileJob> &b = eq_bucket[o2.second.first];
                eq_jobs.insert(eq_jobs.end(), b.begin() + (size_t)o2.second.second * 32, b.begin() + (size_t)o2.second.second * 32 + 32);
            }
        }
        wv.eq_n = (int)eq_jobs.size() - wv.eq_lo; wv.split_n = (int)split_jobs.size() - wv.split_lo;
        wv.chain_n = (int)chain_jobs.size() - wv.chain_lo; wv.wf_n = (int)wf_jobs.size() - wv.wf_lo;
        wv.chunk_n = (int)mb_chunks.size() - wv.chunk_lo; wv.kw_n = (int)kw_jobs.size() - wv.kw_lo;
        wv.gain_n = (int)gain_jobs.size() - wv.gain_lo; wv.seg_hi = n_seg_total;
        wv.lim_n = (int)lim_jobs.size() - wv.lim_lo;
        for (int t = wv.track_lo; t < wv.track_hi; ++t) {
            wv.limiter = wv.limiter || (p->tracks[t].flags & AME_F_LIMITER);
            wv.true_peak = wv.true_peak || (p->tracks[t].flags & AME_F_TRUE_PEAK);
        }
        p->slot_lim_jobs = std::max(p->slot_lim_jobs, wv.lim_n);
        p->slot_tiles = std::max(p->slot_tiles, wv.wf_n);
        p->slot_chains = std::max(p->slot_chains, wv.chain_n);
    }

    // ---- device state -------------------------------------------------------------------------
    if ((rc = upload(&p->d_tracks, p->tracks)) || (rc = upload(&p->d_tdev, p->tdev)) || (rc = upload(&p->d_mb_delta, mb_delta)) ||
        (rc = upload(&p->d_eq_jobs, eq_jobs)) || (rc = upload(&p->d_split_jobs, split_jobs)) ||
        (rc = upload(&p->d_chain_jobs, chain_jobs)) || (rc = upload(&p->d_wf_jobs, wf_jobs)) || (rc = upload(&p->d_mb_chunks, mb_chunks)) ||
        (rc = upload(&p->d_kw_jobs, kw_jobs)) || (rc = upload(&p->d_gain_jobs, gain_jobs)) || (rc = upload(&p->d_lim_jobs, lim_jobs)) || (rc = upload(&p->d_tables, tables)))
        return bail(rc);
    const size_t fb = (size_t)p->total_frames * 4;
    p->slots.resize(n_slots);
    for (Slot &sl : p->slots) {
        if ((rc = dmalloc(p, (void **)&sl.pre, (size_t)p->slot_frames * 4)) ||
            (rc = dmalloc(p, (void **)&sl.bands, (size_t)p->mb_frames * 4 * 3)) ||
            (rc = dmalloc(p, (void **)&sl.rms, (size_t)p->mb_frames * 2 * 3)) ||
            (rc = dmalloc(p, (void **)&sl.list, (size_t)p->mb_frames * 2 * 3)) ||
            (rc = dmalloc(p, (void **)&sl.tile_cnt, (size_t)std::max(p->slot_tiles, 1) * 2 * sizeof(int))) ||     // counts, then ranks
            (rc = dmalloc(p, (void **)&sl.n_flagged, (size_t)std::max(p->slot_chains, 1) * sizeof(int))) ||
            (rc = dmalloc(p, (void **)&sl.grp, (size_t)p->slot_groups * sizeof(GrpRec))) ||
            (rc = dmalloc(p, (void **)&sl.att, (size_t)p->mb_frames * 8 * 3)))
            return bail(rc);
        if (p->any_limiter &&
            ((rc = dmalloc(p, (void **)&sl.norm, (size_t)p->slot_frames * 4)) ||
             (rc = dmalloc(p, (void **)&sl.lim_last, (size_t)std::max(p->slot_lim_jobs, 1) * sizeof(long long))) ||
             (rc = dmalloc(p, (void **)&sl.lim_in.st, (size_t)std::max(p->slot_lim_jobs, 1) * sizeof(LimState))) ||
             (rc = dmalloc(p, (void **)&sl.lim_out.st, (size_t)std::max(p->slot_lim_jobs, 1) * sizeof(LimState))) ||
             (rc = dmalloc(p, (void **)&sl.lim_in.qframe, (size_t)std::max(p->slot_lim_jobs, 1) * p->lim_keep * sizeof(int))) ||
             (rc = dmalloc(p, (void **)&sl.lim_out.qframe, (size_t)std::max(p->slot_lim_jobs, 1) * p->lim_keep * sizeof(int))) ||
             (rc = dmalloc(p, (void **)&sl.lim_in.qdelta, (size_t)std::max(p->slot_lim_jobs, 1) * p->lim_keep * sizeof(double))) ||
             (rc = dmalloc(p, (void **)&sl.lim_out.qdelta, (size_t)std::max(p->slot_lim_jobs, 1) * p->lim_keep * sizeof(double))) ||
             (rc = dmalloc(p, (void **)&sl.lim_need, (size_t)std::max(p->slot_lim_jobs, 1) * sizeof(int)))))
            return bail(rc);
        sl.lim_in.keep = sl.lim_out.keep = p->lim_keep;
        // the filters read a few frames past a track's end (whole 16 / 32-byte groups, never stored): keep them defined
        if (cudaMemsetAsync(sl.pre, 0, (size_t)p->slot_frames * 4, 0) != cudaSuccess ||
            (p->mb_frames && cudaMemsetAsync(sl.bands, 0, (size_t)p->mb_frames * 12, 0) != cudaSuccess))
            return bail(fail(AME_E_CUDA, "cudaMemset failed"));
    }
    if ((rc = dmalloc(p, (void **)&p->d_chain_stats, std::max<size_t>(chain_jobs.size(), 1) * 2 * sizeof(int))) ||
        (rc = dmalloc(p, (void **)&p->d_energy, (size_t)std::max<int64_t>(p->n_sb_total, 1) * 8)) ||
        (rc = dmalloc(p, (void **)&p->d_hist, (size_t)n_tracks * 1000 * 8)) ||
        (rc = dmalloc(p, (void **)&p->d_hist_st, (size_t)n_tracks * 1000 * 4)) ||
        (rc = dmalloc(p, (void **)&p->d_peak, (size_t)n_tracks * 4)) ||
        (rc = dmalloc(p, (void **)&p->d_tp, (size_t)n_tracks * 4)) ||
        (rc = dmalloc(p, (void **)&p->d_lim_stats, 4 * sizeof(int))) ||
        (rc = dmalloc(p, (void **)&p->d_results, (size_t)n_tracks * sizeof(ame_track_result))))
        return bail(rc);
    if (cudaMemsetAsync(p->d_lim_stats, 0, 4 * sizeof(int), 0) != cudaSuccess) return bail(fail(AME_E_CUDA, "cudaMemset failed"));
    if (cudaMemsetAsync(p->d_chain_stats, 0, std::max<size_t>(chain_jobs.size(), 1) * 2 * sizeof(int), 0) != cudaSuccess)
        return bail(fail(AME_E_CUDA, "cudaMemset failed"));
    if (o.host_io) {
        if ((rc = dmalloc(p, (void **)&p->d_in, fb)) || (rc = dmalloc(p, (void **)&p->d_out, fb))) return bail(rc);
        // pool memory is recycled: the padding between tracks is copied back to the caller with the wave, so it must
        // not carry another plan's audio
        if (cudaMemsetAsync(p->d_in, 0, fb, 0) != cudaSuccess || cudaMemsetAsync(p->d_out, 0, fb, 0) != cudaSuccess)
            return bail(fail(AME_E_CUDA, "cudaMemset failed"));
        if (cudaStreamCreateWithFlags(&p->s_in, cudaStreamNonBlocking) != cudaSuccess ||
            cudaStreamCreateWithFlags(&p->s_out, cudaStreamNonBlocking) != cudaSuccess)
            return bail(fail(AME_E_CUDA, "cudaStreamCreate failed"));
    }
    if (o.host_io || n_waves > 1) {
        p->ev_in.resize(n_waves);
        p->ev_run.resize(n_waves);
        for (Slot &sl : p->slots)
            if (cudaStreamCreateWithFlags(&sl.stream, cudaStreamNonBlocking) != cudaSuccess)
                return bail(fail(AME_E_CUDA, "cudaStreamCreate failed"));
        for (int w = 0; w < n_waves; ++w)
            if (cudaEventCreateWithFlags(&p->ev_in[w], cudaEventDisableTiming) != cudaSuccess ||
                cudaEventCreateWithFlags(&p->ev_run[w], cudaEventDisableTiming) != cudaSuccess)
                return bail(fail(AME_E_CUDA, "cudaEventCreate failed"));
    }

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
    // allocations and clears are ordered on the default stream; the plan runs on others
    if (cudaError_t e = cudaStreamSynchronize(0); e != cudaSuccess)
        return bail(fail(AME_E_CUDA, "cudaStreamSynchronize failed: %s", cudaGetErrorString(e)));
    *out = p;
    return AME_OK;
}

int ame_release_cached_memory(int device) {
    GUARD(device);
    CU(cudaDeviceSynchronize());
    if (cudaMemPool_t pool = device_pool()) CU(cudaMemPoolTrimTo(pool, 0));
    return AME_OK;
}

int ame_plan_set_warm_luts(ame_plan *p, const float *luts, int32_t n_luts) {
    if (!p || !luts || n_luts <= 0) return fail(AME_E_INVALID, "bad warm lut arguments");
    GUARD(p->device);
    if (p->d_luts) {
        CU(cudaDeviceSynchronize());
        dev_free(p->d_luts);
        p->d_luts = nullptr;
        p->ws_bytes -= (size_t)p->n_luts * 65536 * sizeof(double);
        p->n_luts = 0;
    }
    // widened exactly to double on the host so the kernel needs no float->double conversion per sample
    std::vector<double> wide((size_t)n_luts * 65536);
    for (size_t i = 0; i < wide.size(); ++i) wide[i] = (double)luts[i];
    int rc = dmalloc(p, (void **)&p->d_luts, wide.size() * sizeof(double));
    if (rc) return rc;
    CU(cudaMemcpy(p->d_luts, wide.data(), wide.size() * sizeof(double), cudaMemcpyHostToDevice));
    CU(cudaStreamSynchronize(0));
    p->n_luts = n_luts;
    return AME_OK;
}

int64_t ame_plan_total_frames(const ame_plan *p) { return p ? p->total_frames : 0; }
size_t ame_plan_workspace_bytes(const ame_plan *p) { return p ? p->ws_bytes : 0; }
int64_t ame_plan_launch_count(const ame_plan *p) { return p ? p->launches : 0; }
int32_t ame_plan_wave_count(const ame_plan *p) { return p ? (int32_t)p->waves.size() : 0; }
int32_t ame_plan_slot_count(const ame_plan *p) { return p ? (int32_t)p->slots.size() : 0; }
const int16_t *ame_plan_tap_pre(const ame_plan *p) { return (p && p->waves.size() == 1) ? p->slots[0].pre : nullptr; }
const int16_t *ame_plan_tap_bands(const ame_plan *p) { return (p && p->waves.size() == 1) ? p->slots[0].bands : nullptr; }
const double *ame_plan_tap_subblock_energy(const ame_plan *p) { return p ? p->d_energy : nullptr; }
int64_t ame_plan_mb_frames(const ame_plan *p) { return p ? p->mb_frames : 0; }
int64_t ame_plan_mb_offset(const ame_plan *p, int32_t t) { return (!p || t < 0 || t >= p->n_tracks) ? -1 : p->mb_offset[t]; }
int64_t ame_plan_subblock_offset(const ame_plan *p, int32_t t) { return (!p || t < 0 || t >= p->n_tracks) ? -1 : p->tdev[t].sb_offset; }

int ame_plan_read_device(ame_plan *p, void *h_dst, const void *d_src, size_t bytes) {
    if (!p || !h_dst || !d_src) return fail(AME_E_INVALID, "NULL argument");
    GUARD(p->device);
    CU(cudaDeviceSynchronize());
    CU(cudaMemcpy(h_dst, d_src, bytes, cudaMemcpyDeviceToHost));
    return AME_OK;
}

int ame_plan_chain_stats(ame_plan *p, int64_t *n_chains, int64_t *steps, int64_t *max_steps, int64_t *passes, int32_t *max_passes) {
    if (!p) return fail(AME_E_INVALID, "NULL plan");
    GUARD(p->device);
    CU(cudaDeviceSynchronize());
    int64_t nc = 0;
    for (const Wave &w : p->waves) nc += w.chain_n;
    std::vector<int> h((size_t)std::max<int64_t>(nc, 1) * 2, 0);
    if (nc) CU(cudaMemcpy(h.data(), p->d_chain_stats, (size_t)nc * 2 * sizeof(int), cudaMemcpyDeviceToHost));
    int64_t st = 0, mx = 0, ps = 0;
    int mp = 0;
    for (int64_t c = 0; c < nc; ++c) {
        st += h[2 * c]; mx = std::max<int64_t>(mx, h[2 * c]);
        ps += h[2 * c + 1]; mp = std::max(mp, h[2 * c + 1]);
    }
    if (n_chains) *n_chains = nc;
    if (steps) *steps = st;
    if (max_steps) *max_steps = mx;
    if (passes) *passes = ps;
    if (max_passes) *max_passes = mp;
    return AME_OK;
}

int ame_plan_limiter_stats(ame_plan *p, int64_t *open_tiles) {
    if (!p || !open_tiles) return fail(AME_E_INVALID, "NULL argument");
    GUARD(p->device);
    CU(cudaDeviceSynchronize());
    int h[4] = {0, 0, 0, 0};
    CU(cudaMemcpy(h, p->d_lim_stats, sizeof h, cudaMemcpyDeviceToHost));
    CU(cudaMemset(p->d_lim_stats, 0, sizeof h));
    for (int i = 0; i < 3; ++i) open_tiles[i] = h[i];
    return AME_OK;
}

int ame_plan_set_timing(ame_plan *p, int enable) {
    if (!p) return fail(AME_E_INVALID, "NULL plan");
    GUARD(p->device);
    if (enable && p->t_ev.empty()) {
        p->t_ev.resize((size_t)kMaxTimedSteps * kTimedSlots * 2);
        for (auto &e : p->t_ev) CU(cudaEventCreate(&e));
    }
    p->t_used.assign((size_t)kMaxTimedSteps * kTimedSlots, 0);
    p->t_step = -1;
    p->timing = enable != 0;
    return AME_OK;
}

int ame_plan_kernel_times(ame_plan *p, double *ms_sum, int64_t *launches, int *n_steps) {
    if (!p || !ms_sum || !launches) return fail(AME_E_INVALID, "NULL argument");
    GUARD(p->device);
    CU(cudaDeviceSynchronize());
    const int steps = std::min(p->t_step + 1, kMaxTimedSteps);
    for (int k = 0; k < AME_N_KERNELS; ++k) { ms_sum[k] = 0; launches[k] = 0; }
    for (size_t i = 0; i < (size_t)steps * kTimedSlots && !p->t_used.empty(); ++i)
        if (p->t_used[i]) {
            float ms = 0;
            CU(cudaEventElapsedTime(&ms, p->t_ev[i * 2], p->t_ev[i * 2 + 1]));
            ms_sum[i % AME_N_KERNELS] += ms;
            ++launches[i % AME_N_KERNELS];
        }
    if (n_steps) *n_steps = steps;
    return AME_OK;
}

int ame_plan_kernel_timeline(ame_plan *p, int step, float *ms, int max_waves) {
    if (!p || !ms) return fail(AME_E_INVALID, "NULL argument");
    GUARD(p->device);
    CU(cudaDeviceSynchronize());
    if (p->t_used.empty() || step < 0 || step > p->t_step || step >= kMaxTimedSteps) return fail(AME_E_INVALID, "step %d was not timed", step);
    const size_t base = (size_t)step * kMaxTimedWaves * AME_N_KERNELS;
    const int n = std::min(std::min((int)p->waves.size(), kMaxTimedWaves), max_waves);
    cudaEvent_t t0 = nullptr;                              // the first kernel of wave 0 opens the step
    for (int k = 0; k < AME_N_KERNELS && !t0; ++k)
        if (p->t_used[base + k]) t0 = p->t_ev[(base + k) * 2];
    if (!t0) return fail(AME_E_INVALID, "step %d was not timed", step);
    for (int w = 0; w < n; ++w)
        for (int k = 0; k < AME_N_KERNELS; ++k) {
            const size_t i = base + (size_t)w * AME_N_KERNELS + k;
            float *o = ms + ((size_t)w * AME_N_KERNELS + k) * 2;
            o[0] = o[1] = NAN;
            if (!p->t_used[i]) continue;
            CU(cudaEventElapsedTime(&o[0], t0, p->t_ev[i * 2]));
            CU(cudaEventElapsedTime(&o[1], t0, p->t_ev[i * 2 + 1]));
        }
    return n;
}

int ame_plan_wave_timeline(ame_plan *p, float *ms, int max_waves) {
    if (!p || !ms) return fail(AME_E_INVALID, "NULL argument");
    GUARD(p->device);
    CU(cudaDeviceSynchronize());
    const int n = std::min(p->tl_waves, max_waves);
    for (int w = 0; w < n; ++w)
        for (int k = 0; k < 4; ++k) CU(cudaEventElapsedTime(&ms[4 * w + k], p->tl_ev[0], p->tl_ev[1 + 4 * w + k]));
    return n;
}

// ---- stage entry points (one launch over the whole batch; plans with ONE wave) ----------------------
#define SINGLE_WAVE(p)                                                                             \
    if ((p)->waves.size() != 1) return fail(AME_E_INVALID, "the stage entry points need a plan with one wave (n_waves <= 1)")

int ame_stage_eq(ame_plan *p, const int16_t *d_in, int16_t *d_pre, void *stream) {
    if (!p || !d_in || !d_pre) return fail(AME_E_INVALID, "NULL argument");
    SINGLE_WAVE(p);
    GUARD(p->device);
    int rc = check_warmth(p);
    if (rc) return rc;
    return run_eq(p, p->waves[0], d_in, d_pre, (cudaStream_t)stream);
}

int ame_stage_band_split(ame_plan *p, const int16_t *d_pre, int16_t *d_bands, void *stream) {
    if (!p || !d_pre) return fail(AME_E_INVALID, "NULL argument");
    SINGLE_WAVE(p);
    GUARD(p->device);
    if (p->waves[0].split_n && !d_bands) return fail(AME_E_INVALID, "NULL bands buffer");
    return run_split(p, p->waves[0], d_pre, d_bands, (cudaStream_t)stream);
}

int ame_stage_compress(ame_plan *p, int16_t *d_bands, int16_t *d_pre, void *stream) {
    if (!p || !d_pre) return fail(AME_E_INVALID, "NULL argument");
    SINGLE_WAVE(p);
    GUARD(p->device);
    if (p->waves[0].chain_n && !d_bands) return fail(AME_E_INVALID, "NULL bands buffer");
    Bufs b = slot_bufs(p, p->waves[0], nullptr, nullptr);
    b.bands = d_bands;
    b.pre = d_pre;
    return run_compress(p, p->waves[0], b, (cudaStream_t)stream);
}

int ame_stage_loudness_hist(ame_plan *p, const int16_t *d_pre, int64_t *d_hist, void *stream) {
    if (!p || !d_pre || !d_hist) return fail(AME_E_INVALID, "NULL argument");
    SINGLE_WAVE(p);
    GUARD(p->device);
    return run_hist(p, p->waves[0], d_pre, d_hist, (cudaStream_t)stream);
}

int ame_stage_apply_gain(ame_plan *p, const int16_t *d_pre, const int64_t *d_hist, int16_t *d_out,
                         ame_track_result *results, void *stream) {
    if (!p || !d_pre || !d_hist || !d_out) return fail(AME_E_INVALID, "NULL argument");
    SINGLE_WAVE(p);
    GUARD(p->device);
    cudaStream_t s = (cudaStream_t)stream;
    const Bufs sb = slot_bufs(p, p->waves[0], nullptr, d_out);
    int rc = run_gain(p, p->waves[0], d_pre, d_hist, d_out, sb, s);
    if (rc) return rc;
    if (results) {
        CU(cudaMemcpyAsync(results, p->d_results, (size_t)p->n_tracks * sizeof(ame_track_result), cudaMemcpyDeviceToHost, s));
        CU(cudaStreamSynchronize(s));
    }
    return AME_OK;
}

// ffmpeg alimiter alone (:223) on an int16 signal that is already normalised: the results are cleared, so that
// k_apply_gain passes the samples through unchanged on its way to the limiter's input buffer
int ame_stage_limiter(ame_plan *p, const int16_t *d_norm, int16_t *d_out, void *stream) {
    if (!p || !d_norm || !d_out) return fail(AME_E_INVALID, "NULL argument");
    SINGLE_WAVE(p);
    GUARD(p->device);
    for (int t = 0; t < p->n_tracks; ++t)
        if (!(p->tracks[t].flags & AME_F_LIMITER)) return fail(AME_E_INVALID, "track %d has no limiter stage (AME_F_LIMITER)", t);
    cudaStream_t s = (cudaStream_t)stream;
    CU(cudaMemsetAsync(p->d_results, 0, (size_t)p->n_tracks * sizeof(ame_track_result), s));
    const Bufs sb = slot_bufs(p, p->waves[0], nullptr, d_out);
    return run_apply(p, p->waves[0], d_norm, d_out, sb, s);
}

// two-phase form: every wave must keep its pre-normalisation signal between the two calls (n_slots == n_waves)
int ame_measure_device(ame_plan *p, const int16_t *d_in, int64_t *d_hist, void *stream) {
    if (!p || !d_in || !d_hist) return fail(AME_E_INVALID, "NULL argument");
    if (p->slots.size() != p->waves.size()) return fail(AME_E_INVALID, "the two-phase calls need n_slots == n_waves");
    GUARD(p->device);
    int rc = check_warmth(p);
    if (rc) return rc;
    p->launches = 0;
    if (p->timing) ++p->t_step;
    for (size_t w = 0; w < p->waves.size(); ++w) {
        p->t_wave = (int)std::min<size_t>(w, kMaxTimedWaves - 1);
        Bufs b = slot_bufs(p, p->waves[w], d_in, nullptr);
        b.hist = d_hist;
        if ((rc = run_measure(p, p->waves[w], b, (cudaStream_t)stream))) return rc;
    }
    return AME_OK;
}

int ame_normalize_device(ame_plan *p, const int64_t *d_hist, int16_t *d_out, ame_track_result *results, void *stream) {
    if (!p || !d_hist || !d_out) return fail(AME_E_INVALID, "NULL argument");
    if (p->slots.size() != p->waves.size()) return fail(AME_E_INVALID, "the two-phase calls need n_slots == n_waves");
    GUARD(p->device);
    cudaStream_t s = (cudaStream_t)stream;
    int rc;
    for (size_t w = 0; w < p->waves.size(); ++w) {
        p->t_wave = (int)std::min<size_t>(w, kMaxTimedWaves - 1);
        const Bufs b = slot_bufs(p, p->waves[w], nullptr, d_out);
        if ((rc = run_gain(p, p->waves[w], b.pre, d_hist, d_out, b, s))) return rc;
    }
    if (results) {
        CU(cudaMemcpyAsync(results, p->d_results, (size_t)p->n_tracks * sizeof(ame_track_result), cudaMemcpyDeviceToHost, s));
        CU(cudaStreamSynchronize(s));
    }
    return AME_OK;
}

// ---- by-time sharding without Python: the two messages of the path over NCCL -------------------------------------
// libame does not link NCCL: the entry points are resolved at run time from the libnccl the host process already
// uses (dlopen with RTLD_NOLOAD first), so the library still loads where there is no NCCL at all.
namespace {
typedef int (*nccl_allreduce_t)(const void *, void *, size_t, int, int, void *, cudaStream_t);
typedef int (*nccl_sendrecv_t)(void *, size_t, int, int, void *, cudaStream_t);
typedef int (*nccl_group_t)(void);
typedef const char *(*nccl_errstr_t)(int);
struct NcclApi {
    nccl_allreduce_t all_reduce = nullptr;
    nccl_sendrecv_t send = nullptr, recv = nullptr;
    nccl_group_t group_start = nullptr, group_end = nullptr;
    nccl_errstr_t err = nullptr;
    bool tried = false, ok = false;
};
NcclApi g_nccl;
std::mutex g_nccl_mutex;
constexpr int kNcclInt8 = 0, kNcclInt64 = 4, kNcclSum = 0;     // ncclDataType_t / ncclRedOp_t values (nccl.h, stable since 2.0)

const NcclApi *nccl_api() {
    std::lock_guard<std::mutex> lock(g_nccl_mutex);
    if (!g_nccl.tried) {
        g_nccl.tried = true;
        void *h = nullptr;
        for (const char *name : {"libnccl.so.2", "libnccl.so"}) {
            h = dlopen(name, RTLD_NOW | RTLD_NOLOAD);
            if (!h) h = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (h) break;
        }
        if (h) {
            g_nccl.all_reduce = (nccl_allreduce_t)dlsym(h, "ncclAllReduce");
            g_nccl.send = (nccl_sendrecv_t)dlsym(h, "ncclSend");
            g_nccl.recv = (nccl_sendrecv_t)dlsym(h, "ncclRecv");
            g_nccl.group_start = (nccl_group_t)dlsym(h, "ncclGroupStart");
            g_nccl.group_end = (nccl_group_t)dlsym(h, "ncclGroupEnd");
            g_nccl.err = (nccl_errstr_t)dlsym(h, "ncclGetErrorString");
            g_nccl.ok = g_nccl.all_reduce && g_nccl.send && g_nccl.recv && g_nccl.group_start && g_nccl.group_end;
        }
    }
    return g_nccl.ok ? &g_nccl : nullptr;
}
#define NCCL(api, call)                                                                            \
    do {                                                                                           \
        int r_ = (call);                                                                           \
        if (r_ != 0) return fail(AME_E_CUDA, "%s failed: %s", #call, (api)->err ? (api)->err(r_) : "NCCL error"); \
    } while (0)
}  // namespace

int ame_hist_allreduce(ame_plan *p, int64_t *d_hist, void *nccl_comm, void *stream) {
    if (!p || !d_hist || !nccl_comm) return fail(AME_E_INVALID, "NULL argument");
    const NcclApi *api = nccl_api();
    if (!api) return fail(AME_E_UNSUPPORTED, "libnccl.so.2 not found in this process");
    GUARD(p->device);
    NCCL(api, api->all_reduce(d_hist, d_hist, (size_t)p->n_tracks * 1000, kNcclInt64, kNcclSum, nccl_comm, (cudaStream_t)stream));
    return AME_OK;
}

int ame_shard_halo_exchange(ame_plan *p, int16_t *d_pre, void *nccl_comm, int prev_rank, int next_rank, int64_t send_frames,
                            void *stream) {
    if (!p || !d_pre || !nccl_comm) return fail(AME_E_INVALID, "NULL argument");
    if (p->n_tracks != 1) return fail(AME_E_INVALID, "a time shard is a plan over one track");
    const ame_track_params &tp = p->tracks[0];
    const int64_t have = tp.halo_frames + tp.n_frames;
    if (next_rank >= 0 && (send_frames <= 0 || send_frames > have))
        return fail(AME_E_UNSUPPORTED, "the shard holds %lld frames, fewer than the %lld the next shard's halo needs", (long long)have,
                    (long long)send_frames);
    const NcclApi *api = nccl_api();
    if (!api) return fail(AME_E_UNSUPPORTED, "libnccl.so.2 not found in this process");
    GUARD(p->device);
    int16_t *base = d_pre + 2 * tp.offset_frames;
    NCCL(api, api->group_start());
    if (next_rank >= 0)
        NCCL(api, api->send(base + 2 * (have - send_frames), (size_t)send_frames * 4, kNcclInt8, next_rank, nccl_comm, (cudaStream_t)stream));
    if (prev_rank >= 0 && tp.halo_frames > 0)
        NCCL(api, api->recv(base, (size_t)tp.halo_frames * 4, kNcclInt8, prev_rank, nccl_comm, (cudaStream_t)stream));
    NCCL(api, api->group_end());
    return AME_OK;
}

int ame_master_device(ame_plan *p, const int16_t *d_in, int16_t *d_out, ame_track_result *results, void *stream) {
    if (!p) return fail(AME_E_INVALID, "NULL plan");
    if (!d_in || !d_out) return fail(AME_E_INVALID, "NULL argument");
    GUARD(p->device);
    int rc = check_warmth(p);
    if (rc) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    const int nw = (int)p->waves.size();
    p->launches = 0;
    if (p->timing) ++p->t_step;
    if (nw <= 1) {
        p->t_wave = 0;
        if ((rc = run_chain_of_stages(p, p->waves[0], slot_bufs(p, p->waves[0], d_in, d_out), s))) return rc;
    } else {
        // several waves: fork the caller's stream into the slots' streams and join again.  Waves of different slots
        // overlap (the latency-bound compressor kernels of one under the FP64-bound filters of another); waves of
        // one slot follow each other in stream order, which is what recycles the slot.
        CU(cudaEventRecord(p->ev_in[0], s));
        const int ns = (int)p->slots.size();
        for (int w = 0; w < nw; ++w) {
            cudaStream_t ws = p->slots[p->waves[w].slot].stream;
            if (w < ns) CU(cudaStreamWaitEvent(ws, p->ev_in[0], 0));
            p->t_wave = std::min(w, kMaxTimedWaves - 1);
            if ((rc = run_chain_of_stages(p, p->waves[w], slot_bufs(p, p->waves[w], d_in, d_out), ws))) return drain(p, rc);
        }
        for (int k = 0; k < ns; ++k) {
            CU(cudaEventRecord(p->ev_run[k], p->slots[k].stream));
            CU(cudaStreamWaitEvent(s, p->ev_run[k], 0));
        }
    }
    if (results) {
        CU(cudaMemcpyAsync(results, p->d_results, (size_t)p->n_tracks * sizeof(ame_track_result), cudaMemcpyDeviceToHost, s));
        CU(cudaStreamSynchronize(s));
    }
    return AME_OK;
}

// Host buffers: wave w's H2D copy (stream s_in), kernels (its slot's stream) and D2H copy (s_out) are chained by
// events, so the copy engines (PCIe is full duplex) and the SMs work on different waves at the same time.  Pinned host
// memory makes the copies truly asynchronous; pageable memory still works, the copies then serialise.
static int master_host_impl(ame_plan *p, const int16_t *h_in, int16_t *h_out, ame_track_result *results) {
    int rc = check_warmth(p);
    if (rc) return rc;
    p->launches = 0;
    if (p->timing) ++p->t_step;
    const int nw = (int)p->waves.size();
    const bool tl = p->timing;
    if (tl) {
        while ((int)p->tl_ev.size() < 1 + 4 * nw) {
            cudaEvent_t e;
            CU(cudaEventCreate(&e));
            p->tl_ev.push_back(e);
        }
        p->tl_waves = nw;
        CU(cudaEventRecord(p->tl_ev[0], p->s_in));
    }
    for (int w = 0; w < nw; ++w) {
        const Wave &wv = p->waves[w];
        const size_t off = (size_t)wv.frame_lo * 4, bytes = (size_t)(wv.frame_hi - wv.frame_lo) * 4;
        CU(cudaMemcpyAsync((char *)p->d_in + off, (const char *)h_in + off, bytes, cudaMemcpyHostToDevice, p->s_in));
        CU(cudaEventRecord(p->ev_in[w], p->s_in));
        if (tl) CU(cudaEventRecord(p->tl_ev[1 + 4 * w], p->s_in));
    }
    for (int w = 0; w < nw; ++w) {
        cudaStream_t ws = p->slots[p->waves[w].slot].stream;
        CU(cudaStreamWaitEvent(ws, p->ev_in[w], 0));
        if (tl) CU(cudaEventRecord(p->tl_ev[2 + 4 * w], ws));
        p->t_wave = std::min(w, kMaxTimedWaves - 1);
        if ((rc = run_chain_of_stages(p, p->waves[w], slot_bufs(p, p->waves[w], p->d_in, p->d_out), ws))) return rc;
        CU(cudaEventRecord(p->ev_run[w], ws));
        if (tl) CU(cudaEventRecord(p->tl_ev[3 + 4 * w], ws));
    }
    for (int w = 0; w < nw; ++w) {
        const Wave &wv = p->waves[w];
        const size_t off = (size_t)wv.frame_lo * 4, bytes = (size_t)(wv.frame_hi - wv.frame_lo) * 4;
        CU(cudaStreamWaitEvent(p->s_out, p->ev_run[w], 0));
        CU(cudaMemcpyAsync((char *)h_out + off, (const char *)p->d_out + off, bytes, cudaMemcpyDeviceToHost, p->s_out));
        if (tl) CU(cudaEventRecord(p->tl_ev[4 + 4 * w], p->s_out));
    }
    if (results)
        CU(cudaMemcpyAsync(results, p->d_results, (size_t)p->n_tracks * sizeof(ame_track_result), cudaMemcpyDeviceToHost, p->s_out));
    CU(cudaStreamSynchronize(p->s_out));
    return AME_OK;
}

int ame_master_host(ame_plan *p, const int16_t *h_in, int16_t *h_out, ame_track_result *results) {
    if (!p || !h_in || !h_out) return fail(AME_E_INVALID, "NULL argument");
    if (!p->d_in || !p->d_out) return fail(AME_E_INVALID, "plan was created without host_io");
    GUARD(p->device);
    // whatever fails in the middle: no copy or kernel may still touch h_in / h_out once we have returned
    return drain(p, master_host_impl(p, h_in, h_out, results));
}

}  // extern "C"
