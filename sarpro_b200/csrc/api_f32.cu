// api_f32.cu — rasters whose samples are not u16-valued (polarization ratios, calibrated inputs):
// scan -> [host: stat edges] -> 4096-bin histogram -> [host: stats, window, level edges] -> quantise.
// CLAHE reuses the DN machinery through a u16 key plane (key = CLAHE bin + 1, 0 = invalid).
#include <cmath>
#include <cstring>
#include <vector>

#include "ctx.h"

namespace sarpro {

// launchers of kernels_f32.cu
cudaError_t launch_f32_scan(const void* a, const void* b, int a_u16, int b_u16, int op, uint64_t n, float valid_thresh,
                            F32Scan* out, int sm_count, cudaStream_t stream);
cudaError_t launch_f32_hist4096(const void* a, const void* b, int a_u16, int b_u16, int op, uint64_t n, float valid_thresh,
                                float min_db, float inv_span4096, const float* edges4096, unsigned long long* hist4096,
                                double* sums, int sm_count, cudaStream_t stream);
cudaError_t launch_f32_quantize(const void* a, const void* b, int a_u16, int b_u16, int op, uint64_t n, float valid_thresh,
                                float low_db, float high_db, float gamma, const float* level_edges, uint32_t n_levels,
                                const uint8_t* remap, int key_plane, uint8_t* out_u8, uint16_t* out_u16, int sm_count,
                                cudaStream_t stream);

namespace {
inline float float_of_bits(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }
} // namespace

// a_dev / b_dev: f32 rasters, or u16 DN rasters when a_u16 / b_u16 is set (the loaders convert: the f32 the reference would
// have read from the same TIFF, gdal.rs:123). When the context is in a sharded call (ctx->shard_reduce), the raster is this
// rank's row band of a scene: the scan and the stat histogram are merged over the ranks (integers and bit patterns: the merged
// values are those of the whole scene) before every rank derives the same window redundantly.
int f32_general_single(sarpro_ctx* ctx, int slot, const void* a_dev, const void* b_dev, int a_u16, int b_u16, int op, uint64_t rows,
                       uint64_t cols, int bit_depth, int strategy, PlanKind kind, const OutGeom& g, void* canvas,
                       sarpro_stats* stats_out) {
    BandWs& w = ctx->band[slot];
    const uint64_t n = rows * cols;
    const bool out8 = kind != PlanKind::Autoscale || bit_depth == SARPRO_U8;
    const size_t esz = out8 ? 1 : 2;
    const size_t n_out = g.oc * g.orr;
    sarpro_stats st;
    std::memset(&st, 0, sizeof(st));

    // ---- pass 1: min / max / count -------------------------------------------------------------
    RC(reserve(ctx, w.f32scan, 4096 * 8 + 64));
    F32Scan* scan_dev = (F32Scan*)w.f32scan.p;
    unsigned long long* hist_dev = (unsigned long long*)((char*)w.f32scan.p + 64);
    double* sums_dev = (double*)((char*)w.f32scan.p + 32);
    F32Scan init{0xffffffffu, 0u, 0ull};
    CU(cudaMemsetAsync(w.f32scan.p, 0, 4096 * 8 + 64, ctx->stream));
    CU(cudaMemcpyAsync(scan_dev, &init, sizeof(init), cudaMemcpyHostToDevice, ctx->stream));
    KS(SARPRO_STAGE_HIST, launch_f32_scan(a_dev, b_dev, a_u16, b_u16, op, n, ctx->valid_thresh, scan_dev, ctx->sm_count, ctx->stream));
    if (ctx->shard_reduce) RC(comm_reduce_f32_scan(ctx, scan_dev));
    F32Scan scan;
    CU(cudaMemcpyAsync(&scan, scan_dev, sizeof(scan), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->timing.host_syncs++;
    if (scan.valid_count == 0) { // autoscale.rs:376-378, 466-468, 716-718
        CU(cudaMemsetAsync(canvas, 0, std::max<size_t>(n_out * esz, 1), ctx->stream));
        if (stats_out) *stats_out = st;
        return 0;
    }
    const float min_v = float_of_bits(scan.min_key), max_v = float_of_bits(scan.max_key);
    const double min_db = db_of_sample(min_v), max_db = db_of_sample(max_v);

    // ---- pass 2: stat histogram -----------------------------------------------------------------
    std::vector<uint64_t> h4096(kStatBins, 0);
    double mean_db = min_db, std_db = 0.0;
    const bool degenerate = std::fabs(max_db - min_db) < 2.220446049250313e-16;
    if (std::isinf(max_db)) {
        // An infinite sample makes span = inf in the reference: every pixel lands in stat bin 0, every
        // percentile is NaN (min + 0*inf) and every quantised sample becomes NaN -> 0 (autoscale.rs:105-134, 441-442).
        h4096[0] = scan.valid_count;
        stats_from_stat_histogram(h4096.data(), scan.valid_count, min_db, max_db, max_db, 0.0, &st);
        choose_window(strategy, kind, &st);
        if (stats_out) *stats_out = st;
        CU(cudaMemsetAsync(canvas, 0, std::max<size_t>(n_out * esz, 1), ctx->stream));
        return 0;
    }
    if (!degenerate) {
        std::vector<float> edges;
        build_stat_edges(min_v, max_v, &edges);
        RC(reserve(ctx, w.edges, (size_t)65536 * 4 + 1024));
        CU(cudaMemcpyAsync(w.edges.p, edges.data(), kStatBins * 4, cudaMemcpyHostToDevice, ctx->stream));
        KS(SARPRO_STAGE_HIST, launch_f32_hist4096(a_dev, b_dev, a_u16, b_u16, op, n, ctx->valid_thresh, (float)min_db,
                                                  (float)(4096.0 / (max_db - min_db)), (const float*)w.edges.p, hist_dev,
                                                  sums_dev, ctx->sm_count, ctx->stream));
        if (ctx->shard_reduce) RC(comm_reduce_f32_hist(ctx, hist_dev, sums_dev));
        double sums[2];
        CU(cudaMemcpyAsync(h4096.data(), hist_dev, kStatBins * 8, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaMemcpyAsync(sums, sums_dev, 16, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        ctx->timing.host_syncs++;
        // mean / std from fp32 logs accumulated in f64 relative to (float)min_db: ~1e-6 dB accurate; they feed
        // log lines and the Adaptive branch test only (autoscale.rs:503)
        const double cnt = (double)scan.valid_count;
        const double m1 = sums[0] / cnt;
        mean_db = (double)(float)min_db + m1;
        std_db = std::sqrt(std::fmax(sums[1] / cnt - m1 * m1, 0.0));
    }
    stats_from_stat_histogram(h4096.data(), scan.valid_count, min_db, max_db, mean_db, std_db, &st);
    choose_window(strategy, kind, &st);
    if (stats_out) *stats_out = st;
    const double low = st.low_clip, high = st.high_clip, gamma = st.gamma;

    // ---- pass 3 -----------------------------------------------------------------------------------
    RC(reserve(ctx, w.edges, (size_t)65536 * 4 + 1024));
    std::vector<float> ledges;
    if (kind == PlanKind::Autoscale && strategy == SARPRO_STRATEGY_CLAHE) {
        // key plane: CLAHE bin + 1 (0 = invalid), then the DN machinery with lut[key] = key - 1
        build_level_edges(LevelKind::ClaheBin, low, high, 1.0, 255, min_v, max_v, &ledges, nullptr, nullptr);
        CU(cudaMemcpyAsync(w.edges.p, ledges.data(), ledges.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
        RC(reserve(ctx, w.dn, n * 2));
        KS(SARPRO_STAGE_CONVERT, launch_f32_quantize(a_dev, b_dev, a_u16, b_u16, op, n, ctx->valid_thresh, (float)low, (float)high, 1.0f,
                                                     (const float*)w.edges.p, 255, nullptr, 1, nullptr, (uint16_t*)w.dn.p,
                                                     ctx->sm_count, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream)); // ledges is a host temporary
        std::vector<uint16_t> lut(kDnBins, 0);
        for (int k = 1; k <= 256; ++k) lut[k] = (uint16_t)(k - 1);
        BandJob j;
        j.dn = (const uint16_t*)w.dn.p;
        j.rows = rows;
        j.cols = cols;
        j.strategy = strategy;
        j.bit_depth = bit_depth;
        j.kind = kind;
        return dn_band_with_preset_lut(ctx, slot, j, lut.data(), 256, g, canvas);
    }
    uint32_t lvl_min = 0, lvl_max = 0;
    const uint32_t n_levels = out8 ? 255u : 65535u;
    build_level_edges(kind == PlanKind::Autoscale ? LevelKind::Quantize : LevelKind::TamedLinearU8, low, high, gamma, n_levels,
                      min_v, max_v, &ledges, &lvl_min, &lvl_max);
    CU(cudaMemcpyAsync(w.edges.p, ledges.data(), ledges.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
    const uint8_t* remap_dev = nullptr;
    if (kind == PlanKind::Autoscale && out8) {
        // scale_u16_to_u8 over all pixels: levels are monotone in the sample, so the extrema are the levels of
        // the smallest / largest valid sample, plus 0 when any pixel is invalid (autoscale.rs:348-364, 669-670)
        uint32_t mn = lvl_min, mx = lvl_max;
        if (scan.valid_count < (ctx->shard_reduce ? ctx->shard_scene_px : n)) mn = 0;
        RC(reserve(ctx, w.remap, 256));
        make_u16_to_u8_remap((uint16_t)mn, (uint16_t)mx, 256, ctx->h_remap + 256 * slot);
        CU(cudaMemcpyAsync(w.remap.p, ctx->h_remap + 256 * slot, 256, cudaMemcpyHostToDevice, ctx->stream));
        remap_dev = (const uint8_t*)w.remap.p;
    }
    void* plane = canvas;
    if (g.resize || g.pad) {
        RC(reserve(ctx, w.full, n * esz));
        plane = w.full.p;
    }
    KS(SARPRO_STAGE_APPLY, launch_f32_quantize(a_dev, b_dev, a_u16, b_u16, op, n, ctx->valid_thresh, (float)low, (float)high, (float)gamma,
                                               (const float*)w.edges.p, n_levels, remap_dev, 0, out8 ? (uint8_t*)plane : nullptr,
                                               out8 ? nullptr : (uint16_t*)plane, ctx->sm_count, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream)); // ledges is a host temporary
    if (!g.resize && !g.pad) return 0;
    if (g.pad) CU(cudaMemsetAsync(canvas, 0, n_out * esz, ctx->stream));
    unsigned char* region = (unsigned char*)canvas + (g.pad_top * g.oc + g.pad_left) * esz;
    if (!g.resize) {
        CU(cudaMemcpy2DAsync(region, g.oc * esz, plane, cols * esz, cols * esz, rows, cudaMemcpyDeviceToDevice, ctx->stream));
        return 0;
    }
    if (g.rc == 0 || g.rr == 0) return 0;
    const int pix16 = out8 ? 0 : 1;
    AxisPlan *ah, *av;
    RC(get_axis(ctx, (uint32_t)cols, (uint32_t)g.rc, pix16, true, HSRC_IMAGE, &ah));
    RC(get_axis(ctx, (uint32_t)rows, (uint32_t)g.rr, pix16, false, 0, &av));
    RC(reserve(ctx, w.temp, rows * g.rc * esz));
    HResizeArgs a{};
    a.src = plane;
    a.src_rows = (uint32_t)rows;
    a.src_cols = (uint32_t)cols;
    a.row0 = 0;
    a.n_rows = (uint32_t)rows;
    a.temp = w.temp.p;
    a.ax = ah->dev();
    RC(run_hpass_generic(ctx, a, HSRC_IMAGE, pix16, ah));
    KS(SARPRO_STAGE_VRESIZE, launch_vresize(w.temp.p, 0, (uint32_t)g.rc, av->dev(), 0, (uint32_t)g.rr, region, (uint32_t)g.oc, 0,
                                            pix16, ctx->stream));
    return 0;
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
