// api_batch.cu — the scene loop (BASELINE config 5): whole scenes through the fused pipelines with the next scene's upload
// overlapped with the current scene's kernels and read-back. Reference: process_directory_to_path (api/mod.rs:474-536).
#include <algorithm>
#include <cstring>

#include "ctx.h"

namespace sarpro {

int ensure_upload_stream(sarpro_ctx* ctx) {
    if (!ctx->stream_up) CU(cudaStreamCreateWithFlags(&ctx->stream_up, cudaStreamNonBlocking));
    for (auto& ev : ctx->ev_up)
        if (!ev) CU(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    for (auto& ev : ctx->ev_chunk)
        if (!ev) CU(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    return 0;
}

namespace {

inline size_t sample_bytes(const sarpro_band& b) { return b.dtype == SARPRO_DT_U16 ? 2 : 4; }
inline bool scene_skipped(const sarpro_scene& s) { return !s.b1.data || !s.b2.data || s.b1.rows * s.b1.cols == 0; }

// Queues the upload of scene s into staging slot `slot` on the copy stream; dev[] are the bands the pipeline will read.
int stage_scene(sarpro_ctx* ctx, const sarpro_scene& s, int slot, sarpro_band dev[2], uint64_t* h2d) {
    const sarpro_band* in[2] = {&s.b1, &s.b2};
    for (int k = 0; k < 2; ++k) {
        dev[k] = *in[k];
        if (in[k]->location != SARPRO_LOC_HOST) continue;
        const size_t bytes = (size_t)in[k]->rows * in[k]->cols * sample_bytes(*in[k]);
        RC(reserve(ctx, ctx->batch_stage[slot][k], std::max<size_t>(bytes, 16)));
        CU(cudaMemcpyAsync(ctx->batch_stage[slot][k].p, in[k]->data, bytes, cudaMemcpyHostToDevice, ctx->stream_up));
        *h2d += bytes;
        dev[k].data = ctx->batch_stage[slot][k].p;
        dev[k].location = SARPRO_LOC_DEVICE;
    }
    CU(cudaEventRecord(ctx->ev_up[slot], ctx->stream_up));
    return 0;
}

} // namespace
} // namespace sarpro

using namespace sarpro;

extern "C" int sarpro_pipeline_batch(sarpro_ctx* ctx, const sarpro_scene* scenes, size_t n, int kind, int bit_depth, int strategy,
                                     int mode, int has_target, size_t target, int pad, int tamed_band_step, int continue_on_error,
                                     sarpro_image* outs, sarpro_stats* stats, int* statuses, sarpro_batch_report* report) {
    RC(begin_call(ctx));
    if (report) std::memset(report, 0, sizeof(*report));
    if (n && (!scenes || !outs)) return fail(ctx, SARPRO_ERR_INVALID_ARGUMENT, "NULL argument");
    if (kind != SARPRO_BATCH_MULTIBAND && kind != SARPRO_BATCH_SYNRGB) return fail(ctx, SARPRO_ERR_INVALID_ARGUMENT, "unknown batch kind %d", kind);
    RC(check_enums(ctx, -2, strategy, kind == SARPRO_BATCH_MULTIBAND ? bit_depth : -2));
    RC(ensure_upload_stream(ctx));
    const double t0 = host_ms();
    sarpro_timing acc;
    std::memset(&acc, 0, sizeof(acc));
    uint64_t h2d = 0;
    sarpro_batch_report rep = {0, 0, 0};
    std::string first_err;
    int first_rc = 0;

    // next scene that is not skipped, from k on
    auto next_live = [&](size_t k) {
        while (k < n && scene_skipped(scenes[k])) ++k;
        return k;
    };
    sarpro_band dev[2][2];
    int staged_rc[2] = {0, 0};
    size_t cur = next_live(0);
    int slot = 0;
    if (cur < n) staged_rc[slot] = stage_scene(ctx, scenes[cur], slot, dev[slot], &h2d);
    for (size_t k = 0; k < n; ++k)
        if (scene_skipped(scenes[k])) {
            rep.skipped++;
            if (statuses) statuses[k] = SARPRO_OK;
        }
    while (cur < n) {
        // the staging slot of the scene after this one was last read by the scene before this one, whose call has returned
        // (every pipeline call ends with its stream synchronised): its upload can go out now and runs beside this scene's work
        const size_t nxt = next_live(cur + 1);
        if (nxt < n) staged_rc[slot ^ 1] = stage_scene(ctx, scenes[nxt], slot ^ 1, dev[slot ^ 1], &h2d);
        int rc = staged_rc[slot];
        if (!rc) {
            cudaError_t e = cudaStreamWaitEvent(ctx->stream, ctx->ev_up[slot], 0);
            if (e != cudaSuccess) rc = fail(ctx, SARPRO_ERR_CUDA, "cudaStreamWaitEvent: %s", cudaGetErrorString(e));
        }
        if (!rc) {
            if (kind == SARPRO_BATCH_SYNRGB)
                rc = sarpro_pipeline_synrgb(ctx, &dev[slot][0], &dev[slot][1], strategy, mode, has_target, target, pad, tamed_band_step,
                                            &outs[cur], stats ? &stats[2 * cur] : nullptr);
            else
                rc = sarpro_pipeline_multiband_tiff(ctx, &dev[slot][0], &dev[slot][1], bit_depth, strategy, has_target, target, pad,
                                                    &outs[2 * cur], &outs[2 * cur + 1], stats ? &stats[2 * cur] : nullptr);
        }
        if (statuses) statuses[cur] = rc;
        if (rc) {
            rep.errors++;
            if (!first_rc) { first_rc = rc; first_err = ctx->err; }
            if (!continue_on_error) break;
        } else {
            rep.processed++;
            const sarpro_timing& t = ctx->timing;
            acc.total_ms += t.total_ms;
            acc.kernel_launches += t.kernel_launches;
            acc.host_syncs += t.host_syncs;
            acc.d2h_bytes += t.d2h_bytes;
            for (int i = 0; i < 8; ++i) { acc.stage_ms[i] += t.stage_ms[i]; acc.stage_launches[i] += t.stage_launches[i]; }
        }
        cur = nxt;
        slot ^= 1;
    }
    // an upload that was queued for a scene that will not run must not outlive the call
    cudaStreamSynchronize(ctx->stream_up);
    acc.h2d_bytes = h2d;
    acc.kernel_ms = acc.total_ms;                // device time of the pipelines (copies run beside them)
    acc.total_ms = (float)(host_ms() - t0);      // the whole loop as the caller sees it
    ctx->timing = acc;
    if (report) *report = rep;
    if (first_rc && !continue_on_error) {
        ctx->err = first_err;
        return first_rc;
    }
    if (first_rc) ctx->err = first_err; // kept for the caller's log; the call itself succeeded
    return SARPRO_OK;
}
