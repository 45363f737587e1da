// api_f32.cu — rasters whose samples are not u16-valued (polarization ratios, calibrated inputs).
#include <cstring>

#include "ctx.h"

namespace sarpro {

int f32_general_single(sarpro_ctx* ctx, int slot, const float* a_dev, const float* b_dev, int op, uint64_t rows,
                       uint64_t cols, int bit_depth, int strategy, PlanKind kind, bool has_target, size_t target,
                       bool pad, void* canvas_dev, sarpro_stats* stats) {
    (void)slot; (void)a_dev; (void)b_dev; (void)op; (void)rows; (void)cols; (void)bit_depth; (void)strategy; (void)kind;
    (void)has_target; (void)target; (void)pad; (void)canvas_dev; (void)stats;
    return fail(ctx, SARPRO_ERR_INTERNAL, "general f32 path not built in this revision");
}

} // namespace sarpro

using namespace sarpro;

extern "C" int sarpro_process_scalar_data_inplace(sarpro_ctx* ctx, const float* v, size_t rows, size_t cols, double* db,
                                                  uint8_t* valid_mask) {
    RC(begin_call(ctx));
    const size_t n = rows * cols;
    if (n == 0) return end_call(ctx);
    if (!v || !db || !valid_mask) return fail(ctx, SARPRO_ERR_INVALID_ARGUMENT, "NULL argument");
    BandWs& w = ctx->band[0];
    RC(reserve(ctx, w.f32a, n * 4));
    RC(reserve(ctx, w.full, n * 8));
    RC(reserve(ctx, w.small, n));
    CU(cudaMemcpyAsync(w.f32a.p, v, n * 4, cudaMemcpyHostToDevice, ctx->stream));
    KL(launch_db_mask((const float*)w.f32a.p, n, (double*)w.full.p, (uint8_t*)w.small.p, ctx->sm_count, ctx->stream));
    CU(cudaMemcpyAsync(db, w.full.p, n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(valid_mask, w.small.p, n, cudaMemcpyDeviceToHost, ctx->stream));
    return end_call(ctx);
}
