/* One rank of the by-time path (BASELINE config C3) from a C host: a long track split over N GPUs, one process and
 * one ncclComm_t per GPU.  Every call below is declared in include/ame.h; the host designs the filters and fills
 * ame_track_params exactly as audio_mastering_engine_b200/design.py does (scipy.signal.butter, as the reference at
 * audio_mastering_engine.py:285,296,301,302) - that part is elided here.
 *
 *   cc -I include -c examples/time_shard_nccl.c          (tests/test_abi_and_host.py compiles it)
 */
#include <stddef.h>
#include <stdint.h>
#include "ame.h"

/* d_in: this rank's span of the input track, stored at frames [halo, halo + n) of a buffer of total frames;
 * d_pre, d_out: buffers of the same size; d_bands: 3 * ame_plan_mb_frames(plan) frames; d_hist: int64[1000]. */
int master_my_span(int device, const ame_track_params *shard,   /* .halo_frames > 0 on every rank but the first */
                   const float *warm_luts, int n_luts, void *nccl_comm, int rank, int world, int64_t halo_of_next,
                   const int16_t *d_in, int16_t *d_pre, int16_t *d_bands, int16_t *d_out, int64_t *d_hist,
                   ame_track_result *result, void *stream) {
    ame_plan *plan = NULL;
    int rc = ame_plan_create(device, shard, 1, NULL, &plan);
    if (rc) return rc;
    if (n_luts) rc = ame_plan_set_warm_luts(plan, warm_luts, n_luts);
    /* chunks are independent up to the pre-normalisation signal (the reference restarts every filter per 30 s chunk) */
    if (!rc) rc = ame_stage_eq(plan, d_in, d_pre, stream);
    if (!rc) rc = ame_stage_band_split(plan, d_pre, d_bands, stream);
    if (!rc) rc = ame_stage_compress(plan, d_bands, d_pre, stream);
    /* loudness is global: hand the tail to the next rank (K-filter warm-up + the straddling 400 ms blocks) ... */
    if (!rc) rc = ame_shard_halo_exchange(plan, d_pre, nccl_comm, rank > 0 ? rank - 1 : -1, rank + 1 < world ? rank + 1 : -1,
                                          halo_of_next, stream);
    /* ... histogram the blocks that end in this span, and sum the histograms over the ranks */
    if (!rc) rc = ame_stage_loudness_hist(plan, d_pre, d_hist, stream);
    if (!rc) rc = ame_hist_allreduce(plan, d_hist, nccl_comm, stream);
    /* every rank now derives the same integrated loudness and gain */
    if (!rc) rc = ame_stage_apply_gain(plan, d_pre, d_hist, d_out, result, stream);
    ame_plan_destroy(plan);
    return rc;
}
