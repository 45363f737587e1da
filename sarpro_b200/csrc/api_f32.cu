// api_f32.cu — rasters whose samples are not u16-valued (polarization ratios, calibrated inputs):
// scan -> [host: stat edges] -> 4096-bin histogram -> [host: stats, window, level edges] -> quantise.
// CLAHE reuses the DN machinery through a u16 key plane (key = CLAHE bin + 1, 0 = invalid).
#include <cmath>
#include <cstring>
#include <limits>
#include <vector>

#include "ctx.h"

namespace sarpro {

// launchers of kernels_f32.cu (nops = 1 or 2 operations over the same operand pair)
cudaError_t launch_f32_scan(const void* a, const void* b, int a_u16, int b_u16, int op0, int op1, int nops, uint64_t n,
                            float valid_thresh, F32Scan* out, int sm_count, cudaStream_t stream);
cudaError_t launch_f32_hist4096(const void* a, const void* b, int a_u16, int b_u16, int op0, int op1, int nops, uint64_t n,
                                float valid_thresh, const float* min_db, const float* inv_span4096, const float* const* edges4096,
                                unsigned long long* const* hist4096, double* const* sums, const F32GuardHost* guards, int sm_count,
                                cudaStream_t stream);
cudaError_t launch_f32_quantize(const void* a, const void* b, int a_u16, int b_u16, int op0, int op1, int nops, uint64_t n,
                                float valid_thresh, const float* low_db, const float* high_db, const float* gamma,
                                const float* const* level_edges, uint32_t n_levels, const uint8_t* const* remap, int key_plane,
                                void* const* out, int out_u8, const F32GuardHost* guards, int sm_count, cudaStream_t stream);

namespace {
inline float float_of_bits(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }
constexpr size_t kScanOff = 0, kSumsOff = 64, kHistOff = 128; // joint workspace of a call: F32Scan[2] | double[2][2] | u64[2][4096]
const F32GuardHost kNoGuard = {0, 0.f, 0.f, 1.f};
} // namespace

// a_dev / b_dev: f32 rasters, or u16 DN rasters when a_u16 / b_u16 is set (the loaders convert: the f32 the reference would
// have read from the same TIFF, gdal.rs:123). nops = 1: one band (op < 0: the raster itself). nops = 2: two polarization
// operations over the same pair, each its own output band (the two calls the reference makes, sentinel1.rs:1497-1579 ->
// pipeline.rs:42-66, with the operands read once per pass); needs full-resolution outputs and a strategy other than CLAHE.
// When the context is in a sharded call (ctx->shard_reduce), the raster is this rank's row band of a scene: the scan and the
// stat histogram are merged over the ranks (integers and bit patterns: the merged values are those of the whole scene)
// before every rank derives the same window redundantly.
int f32_general(sarpro_ctx* ctx, int nops, const int* slots, const void* a_dev, const void* b_dev, int a_u16, int b_u16, const int* ops,
                uint64_t rows, uint64_t cols, int bit_depth, int strategy, PlanKind kind, const OutGeom& g, void* const* canvases,
                sarpro_stats* stats_out) {
    const uint64_t n = rows * cols;
    const bool out8 = kind != PlanKind::Autoscale || bit_depth == SARPRO_U8;
    const size_t esz = out8 ? 1 : 2;
    const size_t n_out = g.oc * g.orr;
    const bool clahe = kind == PlanKind::Autoscale && strategy == SARPRO_STRATEGY_CLAHE;
    if (nops < 1 || nops > 2 || (nops == 2 && (clahe || g.resize || g.pad)))
        return fail(ctx, SARPRO_ERR_INVALID_ARGUMENT, "the two-operation general path writes full-resolution bands of a non-CLAHE strategy");
    const int op0 = ops[0], op1 = nops == 2 ? ops[1] : -1;
    BandWs& w0 = ctx->band[slots[0]];
    sarpro_stats st[2];
    std::memset(st, 0, sizeof(st));

    // ---- pass 1: min / max / count -------------------------------------------------------------
    RC(reserve(ctx, w0.f32scan, kHistOff + 2 * 4096 * 8));
    char* ws = (char*)w0.f32scan.p;
    F32Scan* scan_dev = (F32Scan*)(ws + kScanOff);
    F32Scan init[2] = {{0xffffffffu, 0u, 0ull}, {0xffffffffu, 0u, 0ull}};
    CU(cudaMemsetAsync(ws, 0, kHistOff + 2 * 4096 * 8, ctx->stream));
    CU(cudaMemcpyAsync(scan_dev, init, sizeof(init), cudaMemcpyHostToDevice, ctx->stream));
    KS(SARPRO_STAGE_HIST, launch_f32_scan(a_dev, b_dev, a_u16, b_u16, op0, op1, nops, n, ctx->valid_thresh, scan_dev, ctx->sm_count, ctx->stream));
    if (ctx->shard_reduce) RC(comm_reduce_f32_scan(ctx, scan_dev, nops));
    F32Scan scan[2];
    CU(cudaMemcpyAsync(scan, scan_dev, sizeof(scan), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->timing.host_syncs++;

    // per operation: 0 = regular, 1 = no valid sample (autoscale.rs:376-378, 466-468, 716-718), 2 = an infinite sample, 3 = all equal
    int state[2] = {0, 0};
    float min_v[2], max_v[2];
    double min_db[2], max_db[2], mean_db[2], std_db[2];
    std::vector<uint64_t> h4096[2];
    std::vector<float> edges[2];
    bool need_hist = false;
    for (int o = 0; o < nops; ++o) {
        h4096[o].assign(kStatBins, 0);
        if (scan[o].valid_count == 0) { state[o] = 1; continue; }
        min_v[o] = float_of_bits(scan[o].min_key);
        max_v[o] = float_of_bits(scan[o].max_key);
        min_db[o] = db_of_sample(min_v[o]);
        max_db[o] = db_of_sample(max_v[o]);
        mean_db[o] = min_db[o];
        std_db[o] = 0.0;
        if (std::isinf(max_db[o])) state[o] = 2;
        else if (std::fabs(max_db[o] - min_db[o]) < 2.220446049250313e-16) state[o] = 3;
        else need_hist = true;
    }

    // ---- pass 2: stat histogram -----------------------------------------------------------------
    unsigned long long* hist_dev[2] = {(unsigned long long*)(ws + kHistOff), (unsigned long long*)(ws + kHistOff) + 4096};
    double* sums_dev[2] = {(double*)(ws + kSumsOff), (double*)(ws + kSumsOff) + 2};
    if (need_hist) {
        // an operation that needs no histogram (states 1-3) still rides along in a two-operation launch: its bins are ignored
        float fmin[2] = {0.f, 0.f}, finv[2] = {1.f, 1.f};
        F32GuardHost hg[2] = {{0, 0.f, 0.f, 1.f}, {0, 0.f, 0.f, 1.f}};
        const float* edges_dev[2];
        for (int o = 0; o < nops; ++o) {
            BandWs& w = ctx->band[slots[o]];
            RC(reserve(ctx, w.edges, (size_t)65536 * 4 + 1024));
            edges_dev[o] = (const float*)w.edges.p;
            if (state[o] == 0) {
                build_stat_edges(min_v[o], max_v[o], &edges[o]);
                fmin[o] = (float)min_db[o];
                finv[o] = (float)(4096.0 / (max_db[o] - min_db[o]));
                f32_guard(!ctx->f32_no_guard, min_db[o], max_db[o] - min_db[o], kStatBins, min_v[o], max_v[o], &hg[o].e0, &hg[o].f0, &hg[o].scale,
                          &hg[o].guard);
            } else {
                edges[o].assign(kStatBins, std::numeric_limits<float>::infinity()); // everything in bin 0
            }
            CU(cudaMemcpyAsync(w.edges.p, edges[o].data(), kStatBins * 4, cudaMemcpyHostToDevice, ctx->stream));
        }
        KS(SARPRO_STAGE_HIST, launch_f32_hist4096(a_dev, b_dev, a_u16, b_u16, op0, op1, nops, n, ctx->valid_thresh, fmin, finv, edges_dev,
                                                  hist_dev, sums_dev, hg, ctx->sm_count, ctx->stream));
        if (ctx->shard_reduce) RC(comm_reduce_f32_hist(ctx, hist_dev[0], sums_dev[0], nops));
        double sums[4];
        for (int o = 0; o < nops; ++o)
            if (state[o] == 0) CU(cudaMemcpyAsync(h4096[o].data(), hist_dev[o], kStatBins * 8, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaMemcpyAsync(sums, sums_dev[0], sizeof(sums), cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        ctx->timing.host_syncs++;
        for (int o = 0; o < nops; ++o) {
            if (state[o] != 0) continue;
            // mean / std from fp32 logs accumulated in f64 relative to (float)min_db: ~1e-6 dB accurate; they feed
            // log lines and the Adaptive branch test only (autoscale.rs:503)
            const double cnt = (double)scan[o].valid_count;
            const double m1 = sums[2 * o] / cnt;
            mean_db[o] = (double)(float)min_db[o] + m1;
            std_db[o] = std::sqrt(std::fmax(sums[2 * o + 1] / cnt - m1 * m1, 0.0));
        }
    }
    for (int o = 0; o < nops; ++o) {
        if (state[o] == 1) continue;
        if (state[o] == 2) {
            // An infinite sample makes span = inf in the reference: every pixel lands in stat bin 0, every
            // percentile is NaN (min + 0*inf) and every quantised sample becomes NaN -> 0 (autoscale.rs:105-134, 441-442).
            h4096[o][0] = scan[o].valid_count;
            stats_from_stat_histogram(h4096[o].data(), scan[o].valid_count, min_db[o], max_db[o], max_db[o], 0.0, &st[o]);
        } else {
            stats_from_stat_histogram(h4096[o].data(), scan[o].valid_count, min_db[o], max_db[o], mean_db[o], std_db[o], &st[o]);
        }
        choose_window(strategy, kind, &st[o]);
    }
    if (stats_out)
        for (int o = 0; o < nops; ++o) stats_out[o] = st[o];
    // operations whose result is all zeros
    bool any_regular = false;
    for (int o = 0; o < nops; ++o) {
        if (state[o] == 1 || state[o] == 2) CU(cudaMemsetAsync(canvases[o], 0, std::max<size_t>(n_out * esz, 1), ctx->stream));
        else any_regular = true;
    }
    if (!any_regular) return 0;

    // ---- pass 3 -----------------------------------------------------------------------------------
    if (clahe) { // nops == 1
        BandWs& w = w0;
        const double low = st[0].low_clip, high = st[0].high_clip;
        RC(reserve(ctx, w.edges, (size_t)65536 * 4 + 1024));
        // key plane: CLAHE bin + 1 (0 = invalid), then the DN machinery with lut[key] = key - 1
        std::vector<float> ledges;
        build_level_edges(LevelKind::ClaheBin, low, high, 1.0, 255, min_v[0], max_v[0], &ledges, nullptr, nullptr);
        CU(cudaMemcpyAsync(w.edges.p, ledges.data(), ledges.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
        RC(reserve(ctx, w.dn, n * 2));
        const float lo_f = (float)low, hi_f = (float)high, g1 = 1.0f;
        const float* le = (const float*)w.edges.p;
        void* key_out = w.dn.p;
        KS(SARPRO_STAGE_CONVERT, launch_f32_quantize(a_dev, b_dev, a_u16, b_u16, op0, -1, 1, n, ctx->valid_thresh, &lo_f, &hi_f, &g1, &le, 255,
                                                     nullptr, 1, &key_out, 0, &kNoGuard, ctx->sm_count, ctx->stream));
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
        return dn_band_with_preset_lut(ctx, slots[0], j, lut.data(), 256, g, canvases[0]);
    }
    const uint32_t n_levels = out8 ? 255u : 65535u;
    float lo_f[2] = {0.f, 0.f}, hi_f[2] = {1.f, 1.f}, gm_f[2] = {1.f, 1.f};
    const float* le[2] = {nullptr, nullptr};
    const uint8_t* remap_dev[2] = {nullptr, nullptr};
    void* planes[2] = {nullptr, nullptr};
    F32GuardHost qg[2] = {kNoGuard, kNoGuard};
    std::vector<float> ledges[2];
    for (int o = 0; o < nops; ++o) {
        BandWs& w = ctx->band[slots[o]];
        RC(reserve(ctx, w.edges, (size_t)65536 * 4 + 1024));
        le[o] = (const float*)w.edges.p;
        planes[o] = canvases[o];
        if (state[o] == 1 || state[o] == 2) {
            // rides along in a two-operation launch (its canvas was cleared above and is written again with zeros): every level
            // threshold at +inf
            ledges[o].assign((size_t)n_levels + 1, std::numeric_limits<float>::infinity());
            CU(cudaMemcpyAsync(w.edges.p, ledges[o].data(), ledges[o].size() * 4, cudaMemcpyHostToDevice, ctx->stream));
            continue;
        }
        const double low = st[o].low_clip, high = st[o].high_clip, gamma = st[o].gamma;
        uint32_t lvl_min = 0, lvl_max = 0;
        build_level_edges(kind == PlanKind::Autoscale ? LevelKind::Quantize : LevelKind::TamedLinearU8, low, high, gamma, n_levels,
                          min_v[o], max_v[o], &ledges[o], &lvl_min, &lvl_max);
        CU(cudaMemcpyAsync(w.edges.p, ledges[o].data(), ledges[o].size() * 4, cudaMemcpyHostToDevice, ctx->stream));
        lo_f[o] = (float)low; hi_f[o] = (float)high; gm_f[o] = (float)gamma;
        // levels are linear in dB when gamma == 1 (autoscale.rs:440-442, 649-651 with range = max(high - low, 1)): most samples
        // get their level from the guarded direct evaluation and never touch the threshold table (not when the range was
        // raised to 1 dB: samples above high_clip then stop short of the top level)
        f32_guard(gamma == 1.0 && high - low >= 1.0 && !ctx->f32_no_guard, low, high - low, n_levels, min_v[o], max_v[o], &qg[o].e0, &qg[o].f0,
                  &qg[o].scale, &qg[o].guard);
        if (kind == PlanKind::Autoscale && out8) {
            // scale_u16_to_u8 over all pixels: levels are monotone in the sample, so the extrema are the levels of
            // the smallest / largest valid sample, plus 0 when any pixel is invalid (autoscale.rs:348-364, 669-670)
            uint32_t mn = lvl_min, mx = lvl_max;
            if (scan[o].valid_count < (ctx->shard_reduce ? ctx->shard_scene_px : n)) mn = 0;
            RC(reserve(ctx, w.remap, 256));
            make_u16_to_u8_remap((uint16_t)mn, (uint16_t)mx, 256, ctx->h_remap + 256 * slots[o]);
            CU(cudaMemcpyAsync(w.remap.p, ctx->h_remap + 256 * slots[o], 256, cudaMemcpyHostToDevice, ctx->stream));
            remap_dev[o] = (const uint8_t*)w.remap.p;
        }
        if (g.resize || g.pad) {
            RC(reserve(ctx, w.full, n * esz));
            planes[o] = w.full.p;
        }
    }
    KS(SARPRO_STAGE_APPLY, launch_f32_quantize(a_dev, b_dev, a_u16, b_u16, op0, op1, nops, n, ctx->valid_thresh, lo_f, hi_f, gm_f, le, n_levels,
                                               remap_dev, 0, planes, out8 ? 1 : 0, qg, ctx->sm_count, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream)); // ledges are host temporaries
    if (!g.resize && !g.pad) return 0;
    // nops == 1 from here on
    BandWs& w = w0;
    void* canvas = canvases[0];
    void* plane = planes[0];
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

int f32_general_single(sarpro_ctx* ctx, int slot, const void* a_dev, const void* b_dev, int a_u16, int b_u16, int op, uint64_t rows,
                       uint64_t cols, int bit_depth, int strategy, PlanKind kind, const OutGeom& g, void* canvas,
                       sarpro_stats* stats_out) {
    return f32_general(ctx, 1, &slot, a_dev, b_dev, a_u16, b_u16, &op, rows, cols, bit_depth, strategy, kind, g, &canvas, stats_out);
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
