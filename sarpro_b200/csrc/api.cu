// api.cu — sarpro_ctx and the C ABI of include/sarpro_gpu.h.
//
// Orchestration of the device passes in the order of the reference's callers
// (save.rs:49-65,119-134,199-316,317-368; api/mod.rs:84-369). Data layout in HBM per band:
//   dn        [rows][cols] u16      the raster (uploaded, or the caller's device pointer)
//   tile_hist [tiles][65536] u32    per-CLAHE-tile DN counts (tiles = 64 for CLAHE, else 1)
//   total     [65536] u32           DN histogram -> host planner (256 KB D2H)
//   lut       [65536] u16           DN -> final sample or CLAHE bin (128 KB H2D)
//   tile256 / cdf / cdf32           CLAHE per-tile 256-bin histograms and CDFs
//   temp      [rows][out_cols]      horizontally resized samples
//   small     [out_rows][out_cols]  resized (+ padded) band
// There is no CPU fallback: every entry point needs a live context on a CUDA device.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <chrono>
#include <cstring>
#include <map>
#include <string>
#include <thread>
#include <vector>

#include "ctx.h"

using namespace sarpro;

namespace {
thread_local std::string g_create_error;
} // namespace

namespace sarpro {

int fail(sarpro_ctx* c, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (c) c->err = buf;
    else g_create_error = buf;
    return code;
}

int reserve(sarpro_ctx* ctx, DevBuf& b, size_t bytes) {
    if (bytes <= b.cap) return 0;
    if (b.p) CU(cudaFree(b.p));
    b.p = nullptr;
    b.cap = 0;
    size_t want = (bytes + 255) & ~size_t(255);
    CU(cudaMalloc(&b.p, want));
    b.cap = want;
    return 0;
}
void release(DevBuf& b) {
    if (b.p) cudaFree(b.p);
    b.p = nullptr;
    b.cap = 0;
}

// smallest f32 whose dB exceeds -50 (pipeline.rs:19-22), found with the host libm
float compute_valid_thresh() {
    auto valid = [](float v) { return 10.0 * std::log10(std::fmax((double)v, 1e-10)) > -50.0; };
    uint32_t lo = 0, hi;
    float one = 1.0f;
    std::memcpy(&hi, &one, 4); // valid(1.0) is true, valid(0) false
    while (hi - lo > 1) {
        const uint32_t mid = lo + (hi - lo) / 2;
        float f;
        std::memcpy(&f, &mid, 4);
        if (valid(f)) hi = mid; else lo = mid;
    }
    float f;
    std::memcpy(&f, &hi, 4);
    return f;
}

int upload_rgb_luts(sarpro_ctx* ctx) {
    std::vector<uint8_t> all((size_t)kSynRgbSets * kSynRgbSetBytes, 0);
    auto build = [&](uint32_t set) {
        SynRgbLut l;
        if (set == kSynRgbDefaultSet) build_synrgb_default_lut(&l);
        else build_synrgb_suppressed_lut((int)set, &l);
        uint8_t* dst = all.data() + (size_t)set * kSynRgbSetBytes;
        std::memcpy(dst, l.r, 256);
        std::memcpy(dst + 256, l.g, 256);
        std::memcpy(dst + 512, l.b.data(), 65536);
    };
    std::vector<std::thread> th;
    const unsigned nt = std::max(1u, std::min(8u, std::thread::hardware_concurrency()));
    for (unsigned t = 0; t < nt; ++t)
        th.emplace_back([&, t] { for (uint32_t s = t; s < kSynRgbSets; s += nt) build(s); });
    for (auto& t : th) t.join();
    RC(reserve(ctx, ctx->rgb_luts, all.size()));
    CU(cudaMemcpyAsync(ctx->rgb_luts.p, all.data(), all.size(), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return 0;
}

// ---- histogram work units ---------------------------------------------------------------------
// local rows [0, rows) are scene rows [row_off, row_off+rows); tiles follow the scene geometry.
// own0/own1: local rows that are counted (a sharded rank also holds halo rows it must not count).
int prepare_units(sarpro_ctx* ctx, uint64_t rows, uint64_t cols, bool clahe, uint64_t scene_rows, uint64_t row_off,
                  uint64_t own0, uint64_t own1, uint64_t pitch) {
    // pitch >= cols: the per-column CLAHE tables cover the padding columns of a re-pitched raster as well (geometry from cols)
    const uint64_t tcols = std::max(cols, pitch);
    std::vector<HistUnit> units;
    const uint64_t target_px = 192 * 1024;
    if (!clahe) {
        // column strips of <= 4096 samples so that a unit spans >= 48 rows (one row per warp and iteration)
        const uint64_t n_strips = std::max<uint64_t>(1, (cols + 4095) / 4096);
        const uint64_t sw = (((cols + n_strips - 1) / n_strips) + 7) & ~uint64_t(7);
        for (uint64_t c0 = 0; c0 < cols; c0 += sw) {
            const uint64_t c1 = std::min(cols, c0 + sw);
            const uint64_t chunk = std::max<uint64_t>(128, (target_px / (c1 - c0) + 64) / 128 * 128);
            for (uint64_t r = own0; r < own1; r += chunk)
                units.push_back(HistUnit{(uint32_t)r, (uint32_t)std::min(own1, r + chunk), (uint32_t)c0, (uint32_t)c1, 0u, 0u});
        }
        ctx->n_tiles = 1;
    } else {
        const ClaheGeom g = clahe_geometry(scene_rows, cols);
        for (uint64_t ty = 0; ty < kClaheTiles; ++ty) {
            const uint64_t gr0 = ty * g.tile_h, gr1 = std::min((ty + 1) * g.tile_h, scene_rows);
            // intersect with the local band
            const uint64_t a = std::max(gr0, row_off + own0), b = std::min(gr1, row_off + own1);
            if (a >= b) continue;
            for (uint64_t tx = 0; tx < kClaheTiles; ++tx) {
                const uint64_t c0 = tx * g.tile_w, c1 = std::min((tx + 1) * g.tile_w, cols);
                if (c0 >= c1) continue;
                // whole multiples of 128 rows: a warp of the pass-A kernel keeps 4 rows in flight, 32 warps per CTA
                const uint64_t chunk = std::max<uint64_t>(128, (target_px / (c1 - c0) + 64) / 128 * 128);
                for (uint64_t r = a; r < b; r += chunk)
                    units.push_back(HistUnit{(uint32_t)(r - row_off), (uint32_t)(std::min(b, r + chunk) - row_off),
                                             (uint32_t)c0, (uint32_t)c1, (uint32_t)(ty * kClaheTiles + tx), 0u});
            }
        }
        ctx->n_tiles = kClaheTiles * kClaheTiles;
    }
    ctx->n_units = (uint32_t)units.size();
    ctx->units_r1.clear();
    if (!units.empty()) {
        RC(reserve(ctx, ctx->units, units.size() * sizeof(HistUnit)));
        CU(cudaMemcpyAsync(ctx->units.p, units.data(), units.size() * sizeof(HistUnit), cudaMemcpyHostToDevice, ctx->stream));
        // the same units ordered by their last row, for rasters that arrive in row chunks (streamed uploads)
        std::vector<HistUnit> by_row(units);
        std::stable_sort(by_row.begin(), by_row.end(), [](const HistUnit& x, const HistUnit& y) { return x.r1 < y.r1; });
        RC(reserve(ctx, ctx->units_by_row, by_row.size() * sizeof(HistUnit)));
        CU(cudaMemcpyAsync(ctx->units_by_row.p, by_row.data(), by_row.size() * sizeof(HistUnit), cudaMemcpyHostToDevice, ctx->stream));
        for (const HistUnit& u : by_row) ctx->units_r1.push_back(u.r1);
        CU(cudaStreamSynchronize(ctx->stream)); // `units` / `by_row` are pageable temporaries
    }
    if (clahe) {
        const ClaheGeom g = clahe_geometry(scene_rows, cols);
        uint64_t px[kClaheTiles * kClaheTiles];
        for (uint64_t ty = 0; ty < kClaheTiles; ++ty)
            for (uint64_t tx = 0; tx < kClaheTiles; ++tx) {
                const uint64_t r0 = ty * g.tile_h, r1 = std::min((ty + 1) * g.tile_h, scene_rows);
                const uint64_t c0 = tx * g.tile_w, c1 = std::min((tx + 1) * g.tile_w, cols);
                // tiles beyond the raster hold no pixel and are never sampled (autoscale.rs:308-318)
                px[ty * kClaheTiles + tx] = (r1 > r0 ? r1 - r0 : 0) * (c1 > c0 ? c1 - c0 : 0);
            }
        RC(reserve(ctx, ctx->tile_px, sizeof(px)));
        CU(cudaMemcpyAsync(ctx->tile_px.p, px, sizeof(px), cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        // bilinear geometry per column / row
        RC(reserve(ctx, ctx->col_dx, tcols * 8));
        RC(reserve(ctx, ctx->col_omdx, tcols * 8));
        RC(reserve(ctx, ctx->col_t, tcols * 2));
        RC(reserve(ctx, ctx->row_dy, rows * 8));
        RC(reserve(ctx, ctx->row_omdy, rows * 8));
        RC(reserve(ctx, ctx->row_t, rows * 2));
        RC(reserve(ctx, ctx->col_m, tcols * 4));
        RC(reserve(ctx, ctx->row_sat, rows * 4)); // sat, then sat1
        ctx->clahe_tile_w = g.tile_w;
        ctx->clahe_tile_h = g.tile_h;
        KL(launch_clahe_axis((uint32_t)tcols, 0, (uint32_t)g.tile_w, kClaheTiles, (double*)ctx->col_dx.p,
                             (double*)ctx->col_omdx.p, (uint16_t*)ctx->col_t.p, (int32_t*)ctx->col_m.p, nullptr, ctx->stream));
        KL(launch_clahe_axis((uint32_t)rows, (uint32_t)row_off, (uint32_t)g.tile_h, kClaheTiles, (double*)ctx->row_dy.p,
                             (double*)ctx->row_omdy.p, (uint16_t*)ctx->row_t.p, nullptr, (uint16_t*)ctx->row_sat.p, ctx->stream,
                             (uint16_t*)ctx->row_sat.p + rows));
        ctx->clahe_rows = rows;
    }
    return 0;
}

ClaheDev clahe_dev(sarpro_ctx* ctx, int b) {
    ClaheDev cl;
    cl.cdf = (const double*)ctx->band[b].cdf.p;
    cl.cdf32 = (const float*)ctx->band[b].cdf32.p;
    cl.col_dx = (const double*)ctx->col_dx.p;
    cl.col_omdx = (const double*)ctx->col_omdx.p;
    cl.col_t = (const uint16_t*)ctx->col_t.p;
    cl.row_dy = (const double*)ctx->row_dy.p;
    cl.row_omdy = (const double*)ctx->row_omdy.p;
    cl.row_t = (const uint16_t*)ctx->row_t.p;
    cl.col_m = (const int32_t*)ctx->col_m.p;
    cl.row_sat = (const uint16_t*)ctx->row_sat.p;
    cl.row_sat1 = (const uint16_t*)ctx->row_sat.p + ctx->clahe_rows;
    cl.inv2tw = ctx->clahe_tile_w ? (float)(1.0 / (2.0 * (double)ctx->clahe_tile_w)) : 0.f;
    cl.tile_w = (uint32_t)ctx->clahe_tile_w;
    cl.tiles_x = kClaheTiles;
    return cl;
}

// ---- Lanczos axis plans ---------------------------------------------------------------------------
int upload_vec(sarpro_ctx* ctx, DevBuf& b, const void* src, size_t bytes) {
    RC(reserve(ctx, b, std::max<size_t>(bytes, 16)));
    if (bytes) CU(cudaMemcpyAsync(b.p, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return 0;
}

// Drops every cached axis plan. Only called between calls (begin_call): inside a call the pipeline holds raw AxisPlan
// pointers (the horizontal plan while it looks up the vertical one), so nothing may be evicted there.
static void drop_axis_plans(sarpro_ctx* ctx) {
    for (auto& kv : ctx->axes) delete kv.second; // ~AxisPlan releases the device tables
    ctx->axes.clear();
    for (auto& w : ctx->band) { w.pc_axis_id = 0; w.pc_n_ctas = 0; }
}

int get_axis(sarpro_ctx* ctx, uint32_t in, uint32_t out, bool wide, bool horiz, int src_kind, AxisPlan** res, uint32_t strip_nt,
             uint32_t hmma_in) {
    if (!hmma_in || !horiz) hmma_in = in;
    const AxisKey key{in, out, wide ? 1 : 0, horiz ? 1 : 0, horiz ? src_kind : 0, horiz ? strip_nt : 0u, horiz ? hmma_in : 0u};
    auto it = ctx->axes.find(key);
    if (it != ctx->axes.end()) { *res = it->second; return 0; }
    AxisPlan* ap = new AxisPlan();
    ap->id = ctx->next_axis_id++;
    build_lanczos3_axis(in, out, wide, &ap->h);
    const ResampleAxis& h = ap->h;
    std::vector<uint32_t> packed;
    if (horiz && !wide) {
        ap->pairs = ((h.window + 3 + 3) / 4) * 2; // taps = 2*pairs, multiple of 4, >= window + 3
        packed.assign((size_t)out * ap->pairs, 0);
        for (uint32_t ox = 0; ox < out; ++ox) {
            const uint32_t shift = h.start[ox] & 3u;
            for (uint32_t k = 0; k < h.size[ox]; ++k) {
                const uint32_t j = k + shift;
                const uint32_t c = (uint32_t)(uint16_t)(int16_t)h.coef[(size_t)ox * h.window + k];
                packed[(size_t)ox * ap->pairs + (j >> 1)] |= (j & 1) ? (c << 16) : c;
            }
        }
    }
    HMmaPlanHost mp; // (outlives the uploads below: the stream is synchronised before it goes out of scope)
    std::vector<HStrip> strips;
    int rc = upload_vec(ctx, ap->start, h.start.data(), h.start.size() * 4);
    if (!rc) rc = upload_vec(ctx, ap->size, h.size.data(), h.size.size() * 4);
    if (!rc) rc = upload_vec(ctx, ap->coef, h.coef.data(), h.coef.size() * 4);
    if (!rc) rc = upload_vec(ctx, ap->packed, packed.data(), packed.size() * 4);
    if (!rc && horiz) {
        cudaError_t e = hresize_build_strips(h.start.data(), h.size.data(), out, in, h.window, ap->pairs, wide ? 1 : 0,
                                             src_kind, &ap->oxb, &strips, &ap->rbw, &ap->smem);
        if (e != cudaSuccess) rc = fail(ctx, SARPRO_ERR_INVALID_ARGUMENT, "resize %u -> %u needs more shared memory than an SM has", in, out);
        else {
            ap->n_strips = (uint32_t)strips.size();
            rc = upload_vec(ctx, ap->strips, strips.data(), strips.size() * sizeof(HStrip));
            ap->has_strips = true;
        }
        if (!rc && !wide && src_kind != HSRC_IMAGE) {
            const uint32_t tile_w = (in + kClaheTiles - 1) / kClaheTiles;
            // (taps come from the true width `in`; the walk covers hmma_in >= in columns, the extra ones carry no tap)
            if (hmma_build_plan(h.start.data(), h.size.data(), h.coef.data(), h.window, out, hmma_in,
                                src_kind == HSRC_DN_CLAHE ? tile_w : 0u, &mp, strip_nt)) {
                rc = upload_vec(ctx, ap->m_btab, mp.btab.data(), mp.btab.size() * sizeof(uint4));
                if (!rc) rc = upload_vec(ctx, ap->m_ntile, mp.ntile.data(), mp.ntile.size() * sizeof(int4));
                if (!rc) rc = upload_vec(ctx, ap->m_strips, mp.strips.data(), mp.strips.size() * sizeof(uint4));
                ap->m_weights_h = mp.weights;
                ap->m_b_bytes = mp.b_bytes;
                ap->mma = rc == 0;
            }
        }
    }
    if (!rc) {
        cudaError_t e = cudaStreamSynchronize(ctx->stream); // the host vectors are temporaries
        if (e != cudaSuccess) rc = fail(ctx, SARPRO_ERR_CUDA, "CUDA error %s", cudaGetErrorName(e));
    }
    if (rc) { delete ap; return rc; }
    ctx->axes[key] = ap;
    *res = ap;
    return 0;
}

// Strip length (n-tiles of 8 output columns) of the tensor-core pass B for a raster of `rows` rows. A warp walks one 16-row
// group along one strip at a time, so the number of such walks per warp sets the granularity of the launch: a whole scene has
// thousands per SM, a rank's band of a sharded scene (2000 rows on 8 GPUs) would have less than one with full-length strips
// and leave most warps idle. Shorter strips re-read the Lanczos window overlap (73 of 780 columns at 8 n-tiles), so the longest
// strip that still gives every warp >= 2.3 walks wins. SARPRO_STRIP_NT overrides (measurement).
uint32_t choose_strip_nt(const sarpro_ctx* ctx, uint64_t rows, uint64_t out_cols, bool clahe) {
    if (const char* v = getenv("SARPRO_STRIP_NT")) return (uint32_t)std::max(1, std::min(32, atoi(v)));
    const uint64_t n_nt = (out_cols + 7) / 8, groups = (rows + 15) / 16;
    const uint64_t slots = (uint64_t)ctx->sm_count * hmma_warps(clahe);
    for (uint32_t nt = 32; nt > 4; nt /= 2)
        if (groups * ((n_nt + nt - 1) / nt) * 10 >= slots * 23) return nt;
    return 4;
}

// ---- horizontal pass dispatch -----------------------------------------------------------------------
// Table range for kernels_hmma.cu: the smallest hot such that every PRESENT DN >= hot-1 has the table word `top`
// (the tables are monotone and saturate above the window; the planner leaves absent DNs at 0, so presence comes
// from the histogram; hist == nullptr: every DN <= max_present_dn counts as present). 0 when that needs more
// than 2000 table entries.
uint32_t hmma_hot(const uint16_t* lut, const uint32_t* hist, uint32_t max_present_dn, uint32_t* top_out) {
    const uint32_t top = lut[max_present_dn] & 255u;
    uint32_t h = max_present_dn;
    for (uint32_t d = max_present_dn;; --d) {
        if (!hist || hist[d]) {
            if ((lut[d] & 255u) == top) h = d; else break;
        }
        if (d == 0) break;
    }
    *top_out = top;
    const uint32_t need = std::max(64u, (h + 1 + 7) & ~7u);
    return need <= 2000 ? need : 0;
}

// The same from a plan (the planner walked the present DNs already: no second scan of the histogram on the critical path).
uint32_t hmma_hot_from_plan(const BandPlan& plan, uint32_t* top_out) {
    *top_out = plan.lut[plan.max_present_dn] & 255u;
    const uint32_t need = std::max(64u, (plan.sat_from_dn + 1 + 7) & ~7u);
    return need <= 2000 ? need : 0;
}

// Piece lists of kernels_hmma.cu for band slot `slot`: equal-weight runs of (strip, rows) per persistent CTA; pieces never
// straddle a vertical CLAHE cell boundary (rows where floor(r/tile_h - 0.5) changes, autoscale.rs:308-310).
int prepare_pieces(sarpro_ctx* ctx, int slot, uint64_t rows, uint64_t row_off, bool clahe, AxisPlan* ah) {
    BandWs& w = ctx->band[slot];
    const uint64_t th = clahe ? ctx->clahe_tile_h : 0;
    if (w.pc_rows == rows && w.pc_row_off == row_off && w.pc_clahe == (int)clahe && w.pc_tile_h == th && w.pc_axis_id == ah->id &&
        w.pc_n_ctas && w.pc_spare == ctx->pair_spare)
        return 0;
    std::vector<uint64_t> cuts;
    cuts.push_back(0);
    if (clahe && th)
        for (uint64_t k = 0; k < kClaheTiles; ++k) {
            const uint64_t g = (th * (2 * k + 1) + 1) / 2; // first global row with 2r >= th*(2k+1)
            if (g > row_off && g < row_off + rows) cuts.push_back(g - row_off);
        }
    cuts.push_back(rows);
    // Whole rounds of 16-row groups, one group per warp (12 warps for CLAHE, 16 otherwise), so that only the last piece of a
    // vertical cell ends with a partly filled round. Short rasters (a rank's band of a sharded scene) do not have a round per
    // CTA: the unit shrinks to the groups a CTA gets, so that every SM still takes a share.
    const std::vector<HStrip>& weights = ah->m_weights_h;
    const uint32_t warps = hmma_warps(clahe);
    // band 0 of a pipelined pair leaves ctx->pair_spare SMs to band 1's planner and CLAHE statistics, which run beside it
    const uint32_t n_ctas = (uint32_t)std::max(1, ctx->sm_count - (slot == 0 ? ctx->pair_spare : 0));
    const uint64_t groups = (rows + 15) / 16 * std::max<size_t>(1, weights.size());
    const uint32_t per_cta = (uint32_t)std::max<uint64_t>(1, groups / n_ctas);
    const uint32_t unit = 16u * std::min(warps, per_cta);
    std::vector<uint32_t> pieces, first;
    uint32_t max_rows = 0;
    hmma_build_pieces(weights, cuts, n_ctas, unit, &pieces, &first, &max_rows);
    RC(reserve(ctx, w.pieces, std::max<size_t>(pieces.size() * 4, 16)));
    RC(reserve(ctx, w.cta_first, std::max<size_t>(first.size() * 4, 16)));
    CU(cudaMemcpyAsync(w.pieces.p, pieces.data(), pieces.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(w.cta_first.p, first.data(), first.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    w.pc_n_ctas = (uint32_t)first.size() - 1;
    w.pc_unit = unit;
    w.pc_spare = ctx->pair_spare;
    w.pc_rows = rows;
    w.pc_row_off = row_off;
    w.pc_clahe = clahe;
    w.pc_tile_h = th;
    w.pc_axis_id = ah->id;
    return 0;
}

// Horizontal pass with the generic exact kernel (u8 / u16 images, u16 samples, odd widths, small scale factors, the
// device-gated re-runs: a.skip / a.run_if).
int run_hpass_generic(sarpro_ctx* ctx, const HResizeArgs& a, int src_kind, int pix16, AxisPlan* ah) {
    KS(a.skip || a.run_if ? SARPRO_STAGE_OTHER : SARPRO_STAGE_APPLY,
       launch_hresize_planned(a, src_kind, pix16, (const HStrip*)ah->strips.p, ah->n_strips, ah->oxb, ah->rbw, ah->smem, ctx->sm_count,
                              ctx->stream));
    return 0;
}

// Horizontal pass of a DN band: the tensor-core kernel when the shape allows it (u8 samples, source width a multiple of 8,
// 16-byte aligned raster, at most three n-tiles per block), else the generic exact kernel. *gate is set when the tensor-core
// kernel was launched for a band whose table range is only known on the device (planned there): the kernel returns at once
// if the table does not fit (plan->use_generic), and the caller queues the generic kernel gated on that flag BEHIND the
// band's vertical pass (a gated launch still needs its shared memory: queued right behind the tensor-core kernel it would
// wait for the other band's persistent CTAs to leave and hold the vertical pass back).
int run_hpass(sarpro_ctx* ctx, int slot, const HResizeArgs& a, int src_kind, int pix16, AxisPlan* ah, uint64_t row_off, bool* gate) {
    BandWs& w = ctx->band[slot];
    if (gate) *gate = false;
    if (getenv("SARPRO_TRACE"))
        fprintf(stderr, "run_hpass: kind %d pix16 %d mma %d plan %p remap %p skip %p src%%16 %d smem %zu rows %u cols %u\n", src_kind, pix16,
                (int)ah->mma, (const void*)a.plan, (const void*)a.remap, (const void*)a.skip, (int)(reinterpret_cast<uintptr_t>(a.src) % 16),
                ah->mma ? hmma_smem_bytes(src_kind, ah->m_b_bytes) : 0, a.n_rows, a.src_cols);
    const bool common_ok = !pix16 && ctx->use_hmma && !ctx->force_exact && src_kind != HSRC_IMAGE && a.plan && !a.remap &&
                           (reinterpret_cast<uintptr_t>(a.src) % 16) == 0 && (uint64_t)a.src_rows * a.src_cols < (1ull << 32) &&
                           (w.dev_planned || w.hot);
    const bool shape_ok = common_ok && ah->mma && hmma_smem_bytes(src_kind, ah->m_b_bytes) <= 227 * 1024;
    if (shape_ok) {
        RC(prepare_pieces(ctx, slot, a.n_rows, row_off, src_kind == HSRC_DN_CLAHE, ah));
        KS(SARPRO_STAGE_APPLY, launch_hmma(a, src_kind, (const uint4*)ah->m_btab.p, (const int4*)ah->m_ntile.p, (const uint4*)ah->m_strips.p,
                                           (const uint32_t*)w.pieces.p, (const uint32_t*)w.cta_first.p, w.pc_n_ctas,
                                           ah->m_b_bytes, ctx->stream));
        if (gate) *gate = w.dev_planned;
        return 0;
    }
    return run_hpass_generic(ctx, a, src_kind, pix16, ah);
}

// ---- one band through the device passes -----------------------------------------------------------
OutGeom out_geometry(size_t cols, size_t rows, bool has_target, size_t target, bool pad) {
    OutGeom g;
    resize_output_dims(cols, rows, has_target, target, pad, &g.rc, &g.rr, &g.oc, &g.orr);
    g.resize = has_target && std::max(cols, rows) != target; // resize.rs:115-116
    g.pad = pad;
    g.pad_left = pad ? (g.oc - g.rc) / 2 : 0;
    g.pad_top = pad ? (g.orr - g.rr) / 2 : 0;
    g.meta.cols = g.oc;
    g.meta.rows = g.orr;
    g.meta.scale_x = g.resize ? (double)g.rc / (double)cols : 1.0; // resize.rs:169-170
    g.meta.scale_y = g.resize ? (double)g.rr / (double)rows : 1.0;
    g.meta.pad_left = g.pad_left;
    g.meta.pad_top = g.pad_top;
    return g;
}


// Pass A launches for band slot b (no synchronisation): per-tile DN histogram + totals.
int dn_pass_a_launch(sarpro_ctx* ctx, int b, const uint16_t* dn, uint64_t rows, uint64_t cols, bool clahe_units, int phase) {
    ShardGeom sg;
    sg.scene_rows = rows;
    sg.row_off = 0;
    sg.own0 = 0;
    sg.own1 = rows;
    return dn_pass_a_launch_sharded(ctx, b, dn, rows, cols, clahe_units, sg, phase);
}

int dn_pass_a_launch_sharded(sarpro_ctx* ctx, int b, const uint16_t* dn, uint64_t rows, uint64_t cols, bool clahe_units,
                             const ShardGeom& sg, int phase) {
    const uint64_t pitch = ctx->band[b].pitch ? ctx->band[b].pitch : cols; // row pitch of `dn` (a re-pitched raster: > cols)
    if (ctx->units_rows != rows || ctx->units_cols != cols || ctx->units_clahe != (int)clahe_units ||
        ctx->units_scene_rows != sg.scene_rows || ctx->units_row_off != sg.row_off || ctx->units_own0 != sg.own0 ||
        ctx->units_own1 != sg.own1 || ctx->units_pitch != pitch) {
        RC(prepare_units(ctx, rows, cols, clahe_units, sg.scene_rows, sg.row_off, sg.own0, sg.own1, pitch));
        ctx->units_pitch = pitch;
        ctx->units_rows = rows;
        ctx->units_cols = cols;
        ctx->units_clahe = clahe_units;
        ctx->units_scene_rows = sg.scene_rows;
        ctx->units_row_off = sg.row_off;
        ctx->units_own0 = sg.own0;
        ctx->units_own1 = sg.own1;
    }
    BandWs& w = ctx->band[b];
    if (phase != 2) {
        RC(reserve(ctx, w.tile_hist, (size_t)ctx->n_tiles * kDnBins * 4));
        RC(reserve(ctx, w.total, kDnBins * 4));
        RC(reserve(ctx, w.lut, kDnBins * 2));
        RC(reserve(ctx, w.scalars, 64));
        RC(reserve(ctx, w.present, kPresentWords * 4));
        CU(cudaMemsetAsync(w.tile_hist.p, 0, (size_t)ctx->n_tiles * kDnBins * 4, ctx->stream));
        const uint32_t init[4] = {0xffffffffu, 0u, 0u, 0u};
        std::memcpy(ctx->h_scalars + 8 * b, init, sizeof(init));
        CU(cudaMemcpyAsync(w.scalars.p, ctx->h_scalars + 8 * b, sizeof(init), cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaMemsetAsync((uint32_t*)w.scalars.p + 4, 0, 16, ctx->stream)); // [4]: work-unit counter of pass A, [5]: present-list allocator
    }
    if (phase == 1) return 0;
    // (persistent CTAs pulling work units: any grid works; ctx->pair_spare SMs are left to the other band's small kernels)
    const int grid_sms = std::max(1, ctx->sm_count - (b == 1 ? ctx->pair_spare : 0));
    const int variant = ctx->hist_variant >= 0 ? ctx->hist_variant : w.hist_auto;
    const sarpro_ctx::StreamedBand& sb = ctx->streamed[b < 2 ? b : 0];
    if (b < 2 && sb.n_chunks > 0 && ctx->units_r1.size() == ctx->n_units) {
        // the raster is still arriving (stage_band): one launch per uploaded chunk over the units that end inside it
        uint32_t done = 0;
        for (int c = 0; c < sb.n_chunks; ++c) {
            const uint32_t upto = (uint32_t)(std::upper_bound(ctx->units_r1.begin(), ctx->units_r1.end(), sb.row_end[c]) - ctx->units_r1.begin());
            CU(cudaStreamWaitEvent(ctx->stream, ctx->ev_chunk[sb.ev0 + c], 0));
            if (upto == done) continue;
            if (done) CU(cudaMemsetAsync((uint32_t*)w.scalars.p + 4, 0, 4, ctx->stream)); // the kernels' work-unit counter
            KS(SARPRO_STAGE_HIST, launch_dn_hist(dn, pitch, (const HistUnit*)ctx->units_by_row.p + done, upto - done, (uint32_t*)w.tile_hist.p,
                                                 grid_sms, variant, ctx->stream, (uint32_t*)w.scalars.p + 4));
            done = upto;
        }
    } else {
        KS(SARPRO_STAGE_HIST, launch_dn_hist(dn, pitch, (const HistUnit*)ctx->units.p, ctx->n_units, (uint32_t*)w.tile_hist.p, grid_sms, variant,
                                             ctx->stream, (uint32_t*)w.scalars.p + 4));
    }
    KS(SARPRO_STAGE_PLAN, launch_hist_total((const uint32_t*)w.tile_hist.p, ctx->n_tiles, (uint32_t*)w.total.p,
                                            (uint32_t*)w.scalars.p + 2, (uint32_t*)w.scalars.p + 5, (uint2*)w.present.p, kPresentCap,
                                            ctx->stream));
    return 0;
}

// Pass-A table shape for the next raster in this slot: the replicated shared histogram covers DN < 1024 (32 replicas,
// conflict-free), < 2048 (16) or < 4096 (8); brighter pixels take a per-pixel global atomic, which is only cheap
// when they are rare (< 4e-4 of the pixels).
int choose_hist_variant(const BandPlan& plan) {
    const uint64_t total = plan.px_total, ge1k = plan.px_ge1024, ge2k = plan.px_ge2048; // counted by the planner
    const uint64_t lim = total / 2500; // measured on the C3 co-pol band (2.1e-4 of the pixels at DN >= 1024): 0.174 ms with the 1024-DN table, 0.187 ms with the 2048-DN one
    return ge1k <= lim ? 22 : (ge2k <= lim ? 21 : 20);
}

// The plan of band slot b as the host planner left it (w.plan), shipped to the device for the kernels downstream.
int upload_plan_dev(sarpro_ctx* ctx, int b) {
    BandWs& w = ctx->band[b];
    RC(reserve(ctx, w.plan_dev, sizeof(PlanDev)));
    PlanDev& p = ctx->h_plan_up[b];
    std::memset(&p, 0, sizeof(p));
    p.any_valid = w.plan.any_valid;
    p.have_invalid = w.plan.have_invalid;
    p.max_present_dn = w.plan.max_present_dn;
    p.sat_from_dn = w.plan.sat_from_dn;
    p.hot = w.hot;
    p.hot_top = w.hot_top;
    p.use_generic = w.hot == 0;
    p.clahe = w.plan.clahe;
    p.pre_min = w.plan.pre_min;
    p.pre_max = w.plan.pre_max;
    p.px_total = w.plan.px_total;
    p.px_ge1024 = w.plan.px_ge1024;
    p.px_ge2048 = w.plan.px_ge2048;
    p.stats = w.plan.stats;
    CU(cudaMemcpyAsync(w.plan_dev.p, &p, sizeof(PlanDev), cudaMemcpyHostToDevice, ctx->stream));
    return 0;
}

// Host planning of band b from its histogram in pinned memory; ships the DN -> sample / bin table and the plan.
int plan_band_and_upload(sarpro_ctx* ctx, int b, const BandJob& job) {
    BandWs& w = ctx->band[b];
    const bool trace = getenv("SARPRO_TRACE") != nullptr;
    const double t_a = trace ? host_ms() - ctx->host_t0 : 0;
    g_plan_trace_on = trace;
    // the planner reads the device-compacted list of present DNs; the dense totals only when that list overflowed
    const uint32_t* hp = ctx->h_present + (size_t)b * kPresentWords;
    bool dense = false;
    if (getenv("SARPRO_DENSE_PLAN") ||
        !plan_from_present_list(hp, hp + 512, kPresentCap, job.bit_depth, job.strategy, job.kind, &w.plan)) {
        CU(cudaMemcpyAsync(ctx->h_hist + (size_t)b * kDnBins, w.total.p, kDnBins * 4, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        ctx->timing.host_syncs++;
        dense = true;
        plan_from_dn_histogram32(ctx->h_hist + (size_t)b * kDnBins, job.bit_depth, job.strategy, job.kind, &w.plan,
                                 (int)std::min<uint32_t>(ctx->h_scalars[8 * b + 6] + 1u, kDnBins));
    }
    const double t_b = trace ? host_ms() - ctx->host_t0 : 0;
    // only DNs up to the brightest present one are ever looked up (stale entries beyond it are never read)
    const size_t n_lut = getenv("SARPRO_FULL_LUT") ? (size_t)kDnBins : std::min<size_t>(kDnBins, ((size_t)w.plan.max_present_dn + 1 + 63) & ~size_t(63));
    std::memcpy(ctx->h_lut + (size_t)b * kDnBins, w.plan.lut.data(), n_lut * 2);
    w.hot = w.plan.any_valid ? hmma_hot_from_plan(w.plan, &w.hot_top) : 0;
    CU(cudaMemcpyAsync(w.lut.p, ctx->h_lut + (size_t)b * kDnBins, n_lut * 2, cudaMemcpyHostToDevice, ctx->stream));
    RC(upload_plan_dev(ctx, b));
    w.dev_planned = false;
    w.hist_auto_pending = true; // pass-A table shape for the next call: chosen after this band's pass B is queued
    if (trace) fprintf(stderr, "trace host: band %d histogram on the host at %.3f ms, planned at %.3f, table queued at %.3f ms; planner (%s) us: table cleared %.1f, scan %.1f, moments %.1f, percentiles %.1f, table %.1f\n", b, t_a, t_b, host_ms() - ctx->host_t0,
                       dense ? "dense totals" : "present list",
                       g_plan_trace_us[0], g_plan_trace_us[1], g_plan_trace_us[2], g_plan_trace_us[3], g_plan_trace_us[4]);
    return 0;
}

// Strategies planned on the device (kernels_plan.cu): no host round trip between pass A and pass B. SARPRO_HOST_PLAN=1 keeps
// every band on the host planner (validation: the tests compare the two).
bool plans_on_device(const sarpro_ctx* ctx, const BandJob& job) {
    return !ctx->host_plan && plan_on_device_supported(job.strategy, (int)job.kind);
}

// Queues the device planner for band b behind its k_hist_total on ctx->stream, and the copy of the plan to pinned host
// memory (read in end_call: statistics for the caller, the pass-A table shape of the next call).
int plan_bands_on_device(sarpro_ctx* ctx, const int* slots, const BandJob* jobs, int nb) {
    PlanJobs pj{};
    for (int k = 0; k < nb; ++k) {
        BandWs& w = ctx->band[slots[k]];
        RC(reserve(ctx, w.plan_dev, sizeof(PlanDev)));
        RC(reserve(ctx, w.plan_scratch, plan_scratch_bytes()));
        PlanParams pr;
        pr.strategy = jobs[k].strategy;
        pr.kind = (int)jobs[k].kind;
        pr.bit_depth = jobs[k].bit_depth;
        pr.clahe = uses_clahe(jobs[k]);
        pj.j[k] = PlanJob{(const uint32_t*)w.total.p, (uint16_t*)w.lut.p, (PlanDev*)w.plan_dev.p, w.plan_scratch.p, pr};
    }
    KS(SARPRO_STAGE_PLAN, launch_plan_bands(pj, nb, (const double*)ctx->db_table.p, ctx->stream));
    for (int k = 0; k < nb; ++k) {
        BandWs& w = ctx->band[slots[k]];
        CU(cudaMemcpyAsync(&ctx->h_plan[slots[k]], w.plan_dev.p, sizeof(PlanDev), cudaMemcpyDeviceToHost, ctx->stream));
        w.dev_planned = true;
        w.plan_copy_pending = true;
        w.plan.clahe = uses_clahe(jobs[k]);
        w.hist_auto_pending = true;
    }
    return 0;
}
int plan_band_on_device(sarpro_ctx* ctx, int b, const BandJob& job) { return plan_bands_on_device(ctx, &b, &job, 1); }

// CLAHE tile statistics (256-bin tile histograms through the table, then clip / redistribute / CDF) for the given band slots in
// one launch each. all_reduce: optional hook between the two (sharded scene: the tile histograms are merged over the ranks).
int run_clahe_stats_bands(sarpro_ctx* ctx, const int* slots, int nb, int (*all_reduce)(sarpro_ctx*, void*), void* arg) {
    ClaheStatJobs cj{};
    for (int k = 0; k < nb; ++k) {
        BandWs& w = ctx->band[slots[k]];
        RC(reserve(ctx, w.tile256, (size_t)ctx->n_tiles * 256 * 4));
        RC(reserve(ctx, w.cdf, (size_t)ctx->n_tiles * 256 * 8));
        RC(reserve(ctx, w.cdf32, (size_t)ctx->n_tiles * 256 * 4));
        CU(cudaMemsetAsync(w.tile256.p, 0, (size_t)ctx->n_tiles * 256 * 4, ctx->stream));
        cj.j[k] = ClaheStatJob{(const uint32_t*)w.tile_hist.p, (const uint16_t*)w.lut.p, (const PlanDev*)w.plan_dev.p, (uint32_t*)w.tile256.p,
                               (double*)w.cdf.p, (float*)w.cdf32.p};
    }
    KS(SARPRO_STAGE_PLAN, launch_clahe_tile256_2(cj, nb, ctx->n_tiles, ctx->stream));
    if (all_reduce) RC(all_reduce(ctx, arg));
    KS(SARPRO_STAGE_PLAN, launch_clahe_cdf_2(cj, nb, (const uint64_t*)ctx->tile_px.p, ctx->n_tiles, ctx->stream));
    return 0;
}

// Pass A launches (+ the histogram read-back of the bands the HOST will plan) for `nb` bands of identical geometry;
// ctx->ev[2 + b] marks band b's histogram as complete (device-planned bands) / on the host (host-planned bands).
int run_pass_a(sarpro_ctx* ctx, const BandJob* jobs, int nb) {
    bool any_clahe = false;
    for (int b = 0; b < nb; ++b) any_clahe |= uses_clahe(jobs[b]);
    // every band's cleared workspaces first, so that nothing but k_hist_total sits between two bands' histogram kernels
    for (int b = 0; b < nb; ++b) RC(dn_pass_a_launch(ctx, b, jobs[b].dn, jobs[b].rows, jobs[b].cols, any_clahe, 1));
    for (int b = 0; b < nb; ++b) {
        RC(dn_pass_a_launch(ctx, b, jobs[b].dn, jobs[b].rows, jobs[b].cols, any_clahe, 2));
        if (plans_on_device(ctx, jobs[b])) {
            CU(cudaEventRecord(ctx->ev[2 + b], ctx->stream));
            continue;
        }
        // The read-back of all but the last band goes through the side stream: on the main stream the two copies would
        // hold the next band's histogram kernel back.
        cudaStream_t rs = ctx->stream;
        if (b + 1 < nb && ctx->two_stream && ctx->stream2) {
            CU(cudaEventRecord(ctx->ev[4], ctx->stream));
            CU(cudaStreamWaitEvent(ctx->stream2, ctx->ev[4], 0));
            rs = ctx->stream2;
        }
        // the present DNs (k_hist_total) and the brightest one (bounds the planner's walk if it has to read the dense totals)
        CU(cudaMemcpyAsync(ctx->h_present + (size_t)b * kPresentWords, ctx->band[b].present.p, kPresentWords * 4, cudaMemcpyDeviceToHost, rs));
        CU(cudaMemcpyAsync(ctx->h_scalars + 8 * b + 6, (uint32_t*)ctx->band[b].scalars.p + 2, 4, cudaMemcpyDeviceToHost, rs));
        CU(cudaEventRecord(ctx->ev[2 + b], rs));
    }
    return 0;
}

// Plans band b once its histogram is complete. Device-planned strategies: the planner kernel is queued behind the histogram
// (a stream dependency, the host does not wait). Host-planned strategies (Standard, Adaptive): waits for the read-back and
// plans on the host while the device keeps working on whatever is queued behind it (the other band's pass A, or the
// previous band's pass B), then queues the table upload. When `side` is given, everything from here on is issued on that
// stream (ctx->stream is switched; the caller switches it back).
int plan_band_on(sarpro_ctx* ctx, int b, const BandJob& job, cudaStream_t side) {
    if (plans_on_device(ctx, job)) {
        if (side) {
            CU(cudaStreamWaitEvent(side, ctx->ev[2 + b], 0));
            ctx->stream = side;
        }
        return plan_band_on_device(ctx, b, job);
    }
    CU(cudaEventSynchronize(ctx->ev[2 + b]));
    ctx->timing.host_syncs++;
    if (side) ctx->stream = side;
    return plan_band_and_upload(ctx, b, job);
}

// Pass A for `nb` bands, then the planner for all of them (callers that need every plan before pass B).
int run_pass_a_and_plan(sarpro_ctx* ctx, const BandJob* jobs, int nb) {
    RC(run_pass_a(ctx, jobs, nb));
    for (int b = 0; b < nb; ++b) RC(plan_band_on(ctx, b, jobs[b], nullptr));
    return 0;
}

// CLAHE tile CDFs for band b (after the LUT is on the device)
int run_clahe_stats(sarpro_ctx* ctx, int b) { return run_clahe_stats_bands(ctx, &b, 1, nullptr, nullptr); }

// Reads back the CLAHE sample min/max of band b and returns whether scale_u16_to_u8 is the identity.
int clahe_minmax(sarpro_ctx* ctx, int b, uint32_t* mn, uint32_t* mx) {
    BandWs& w = ctx->band[b];
    CU(cudaMemcpyAsync(ctx->h_scalars + 8 * b, w.scalars.p, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->timing.host_syncs++;
    *mn = ctx->h_scalars[8 * b];
    *mx = ctx->h_scalars[8 * b + 1];
    if (*mn == 0xffffffffu) { *mn = 0; *mx = 0; }
    return 0;
}
int upload_remap(sarpro_ctx* ctx, int b, uint32_t mn, uint32_t mx) {
    BandWs& w = ctx->band[b];
    RC(reserve(ctx, w.remap, 256));
    make_u16_to_u8_remap((uint16_t)mn, (uint16_t)mx, 256, ctx->h_remap + 256 * b);
    CU(cudaMemcpyAsync(w.remap.p, ctx->h_remap + 256 * b, 256, cudaMemcpyHostToDevice, ctx->stream));
    return 0;
}
int reset_minmax(sarpro_ctx* ctx, int b) {
    BandWs& w = ctx->band[b];
    const uint32_t init[2] = {0xffffffffu, 0u};
    std::memcpy(ctx->h_scalars + 8 * b + 4, init, sizeof(init));
    CU(cudaMemcpyAsync(w.scalars.p, ctx->h_scalars + 8 * b + 4, sizeof(init), cudaMemcpyHostToDevice, ctx->stream));
    return 0;
}

// A band the HOST planned and found without a valid pixel (a device-planned band is not known here; its all-zero table
// gives the same all-zero output).
static inline bool known_all_invalid(const BandWs& w) { return !w.dev_planned && !w.plan.any_valid; }

// Pass B, full resolution: dst is a device buffer of rows*cols samples of the job's bit depth.
int run_pass_b_full(sarpro_ctx* ctx, int b, const BandJob& j, void* dst) {
    BandWs& w = ctx->band[b];
    const uint64_t n = j.rows * j.cols;
    const bool out8 = j.kind != PlanKind::Autoscale || j.bit_depth == SARPRO_U8;
    if (known_all_invalid(w)) { // all-zero output (autoscale.rs:376-378, 466-468, 716-718); a device-planned band gets there
                                // through its all-zero table
        CU(cudaMemsetAsync(dst, 0, n * (out8 ? 1 : 2), ctx->stream));
        return 0;
    }
    if (!uses_clahe(j)) {
        KS(SARPRO_STAGE_APPLY, launch_apply_lut(j.dn, n, (const uint16_t*)w.lut.p, out8 ? (uint8_t*)dst : nullptr,
                            out8 ? nullptr : (uint16_t*)dst, ctx->sm_count, ctx->stream));
        return 0;
    }
    RC(run_clahe_stats(ctx, b));
    KS(SARPRO_STAGE_APPLY, launch_apply_clahe(j.dn, (uint32_t)j.rows, (uint32_t)j.cols, (const uint16_t*)w.lut.p, clahe_dev(ctx, b),
                          out8 ? 255 : 65535, out8 ? (uint8_t*)dst : nullptr, out8 ? nullptr : (uint16_t*)dst,
                          (uint32_t*)w.scalars.p, ctx->sm_count, ctx->stream));
    if (out8) { // scale_u16_to_u8 over the blended samples (autoscale.rs:691-693): decided on the device, the remap pass
                // is always queued and returns at once when the re-stretch is the identity (no host round trip)
        RC(reserve(ctx, w.remap, 256 + 16));
        uint32_t* flag = reinterpret_cast<uint32_t*>((unsigned char*)w.remap.p + 256);
        KS(SARPRO_STAGE_PLAN, launch_clahe_remap_decide((const uint32_t*)w.scalars.p, (uint8_t*)w.remap.p, flag, ctx->stream));
        KS(SARPRO_STAGE_OTHER, launch_remap_u8((uint8_t*)dst, n, (const uint8_t*)w.remap.p, ctx->sm_count, ctx->stream, flag));
    }
    return 0;
}

// Pass B fused with the Lanczos resize: writes the resized band into `canvas` (pitch g.oc) at the pad offset.
int run_pass_b_resized(sarpro_ctx* ctx, int b, const BandJob& j, const OutGeom& g, void* canvas) {
    BandWs& w = ctx->band[b];
    const bool out8 = j.kind != PlanKind::Autoscale || j.bit_depth == SARPRO_U8;
    const int pix16 = out8 ? 0 : 1;
    const size_t esz = out8 ? 1 : 2;
    if (g.pad || known_all_invalid(w)) CU(cudaMemsetAsync(canvas, 0, g.oc * g.orr * esz, ctx->stream));
    if (g.rc == 0 || g.rr == 0) return 0;
    if (known_all_invalid(w)) return 0; // resize of an all-zero raster is all zero
    const bool clahe = uses_clahe(j);
    const int src_kind = clahe ? HSRC_DN_CLAHE : HSRC_DN_LUT;
    AxisPlan *ah, *av;
    // (a re-pitched raster, produce_bands: j.dn has rows w.pitch samples apart; taps and geometry stay those of j.cols)
    RC(get_axis(ctx, (uint32_t)j.cols, (uint32_t)g.rc, pix16, true, src_kind, &ah, choose_strip_nt(ctx, j.rows, g.rc, clahe), (uint32_t)w.pitch));
    RC(get_axis(ctx, (uint32_t)j.rows, (uint32_t)g.rr, pix16, false, 0, &av));
    RC(reserve(ctx, w.temp, (size_t)j.rows * g.rc * esz));
    if (clahe) RC(run_clahe_stats(ctx, b));
    HResizeArgs a{};
    a.src = j.dn;
    a.src_rows = (uint32_t)j.rows;
    a.src_cols = (uint32_t)(w.pitch ? w.pitch : j.cols);
    a.src_width = (uint32_t)j.cols;
    a.lut = (const uint16_t*)w.lut.p;
    a.plan = (const PlanDev*)w.plan_dev.p;
    a.remap = nullptr;
    if (clahe) a.clahe = clahe_dev(ctx, b);
    a.minmax = clahe ? (uint32_t*)w.scalars.p : nullptr;
    a.row0 = 0;
    a.n_rows = (uint32_t)j.rows;
    a.temp = w.temp.p;
    a.ax = ah->dev();
    unsigned char* dst = (unsigned char*)canvas + (g.pad_top * g.oc + g.pad_left) * esz;
    // 1. horizontal pass (tensor-core kernel or generic), 2. vertical pass
    bool gate = false;
    RC(run_hpass(ctx, b, a, src_kind, pix16, ah, 0, &gate));
    KS(SARPRO_STAGE_VRESIZE, launch_vresize(w.temp.p, 0, (uint32_t)g.rc, av->dev(), 0, (uint32_t)g.rr, dst, (uint32_t)g.oc, 0, pix16, ctx->stream));
    // 3. device-gated re-runs, all queued without a host round trip; each returns at once in the common case:
    //    - the generic horizontal kernel when the band's table did not fit the tensor-core kernel (plan->use_generic);
    //    - CLAHE u8: scale_u16_to_u8 (autoscale.rs:691-693) is the identity when the blended samples span exactly [0,255];
    //      the first run assumed so and tracked the true min / max. A one-block kernel builds the remap table and the
    //      flags; the horizontal pass is repeated with the remap when it is not the identity;
    //    - the vertical pass again when either of the two ran.
    const uint32_t* use_generic = gate ? &a.plan->use_generic : nullptr;
    if (gate) {
        HResizeArgs ag = a;
        ag.run_if = use_generic;
        RC(run_hpass_generic(ctx, ag, src_kind, pix16, ah));
    }
    if (clahe && out8) {
        RC(reserve(ctx, w.remap, 256 + 16));
        uint32_t* flag = reinterpret_cast<uint32_t*>((unsigned char*)w.remap.p + 256);
        KS(SARPRO_STAGE_PLAN, launch_clahe_remap_decide((const uint32_t*)w.scalars.p, (uint8_t*)w.remap.p, flag, ctx->stream,
                                                        gate ? a.plan : nullptr));
        HResizeArgs ar = a;
        ar.remap = (const uint8_t*)w.remap.p;
        ar.minmax = nullptr;
        ar.skip = flag;
        RC(run_hpass_generic(ctx, ar, src_kind, pix16, ah));
        KS(SARPRO_STAGE_OTHER, launch_vresize(w.temp.p, 0, (uint32_t)g.rc, av->dev(), 0, (uint32_t)g.rr, dst, (uint32_t)g.oc, 0, pix16,
                                              ctx->stream, flag + 1));
    } else if (gate) {
        KS(SARPRO_STAGE_OTHER, launch_vresize(w.temp.p, 0, (uint32_t)g.rc, av->dev(), 0, (uint32_t)g.rr, dst, (uint32_t)g.oc, 0, pix16,
                                              ctx->stream, nullptr, use_generic));
    }
    return 0;
}

// Pass B for a planned band: full resolution, pad only, or fused with the Lanczos resize.
int dn_run_pass_b(sarpro_ctx* ctx, int b, const BandJob& j, const OutGeom& g, void* canvas) {
    BandWs& w = ctx->band[b];
    const bool out8 = j.kind != PlanKind::Autoscale || j.bit_depth == SARPRO_U8;
    const size_t esz = out8 ? 1 : 2;
    if (!g.resize && !g.pad) return run_pass_b_full(ctx, b, j, canvas);
    if (!g.resize) { // pad only: full-res apply, then copy into the zeroed canvas
        RC(reserve(ctx, w.full, j.rows * j.cols * esz));
        RC(run_pass_b_full(ctx, b, j, w.full.p));
        CU(cudaMemsetAsync(canvas, 0, g.oc * g.orr * esz, ctx->stream));
        CU(cudaMemcpy2DAsync((unsigned char*)canvas + (g.pad_top * g.oc + g.pad_left) * esz, g.oc * esz, w.full.p,
                             j.cols * esz, j.cols * esz, j.rows, cudaMemcpyDeviceToDevice, ctx->stream));
        return 0;
    }
    return run_pass_b_resized(ctx, b, j, g, canvas);
}

// synthetic_rgb.rs:182-197 on two device-resident u8 bands -> ctx->rgb (interleaved RGB)
int synrgb_compose(sarpro_ctx* ctx, int strategy, const uint8_t* c1, const uint8_t* c2, size_t n) {
    RC(reserve(ctx, ctx->rgb, std::max<size_t>(n * 3, 16)));
    if (!n) return 0;
    const bool suppressed = strategy == SARPRO_STRATEGY_TAMED || strategy == SARPRO_STRATEGY_CLAHE; // synthetic_rgb.rs:189-195
    if (suppressed) {
        RC(reserve(ctx, ctx->hist256, 256 * 4));
        RC(reserve(ctx, ctx->rgbsel, 16));
        CU(cudaMemsetAsync(ctx->hist256.p, 0, 256 * 4, ctx->stream));
        KS(SARPRO_STAGE_RGB, launch_hist256_pair(c1, c2, n, (uint32_t*)ctx->hist256.p, ctx->stream));
        KS(SARPRO_STAGE_RGB, launch_synrgb_floor((const uint32_t*)ctx->hist256.p, n, (uint32_t*)ctx->rgbsel.p, ctx->stream));
        KS(SARPRO_STAGE_RGB, launch_synrgb(c1, c2, n, (const uint8_t*)ctx->rgb_luts.p, (const uint32_t*)ctx->rgbsel.p, 0, 1,
                                           (uint8_t*)ctx->rgb.p, ctx->stream));
    } else {
        KS(SARPRO_STAGE_RGB, launch_synrgb(c1, c2, n, (const uint8_t*)ctx->rgb_luts.p, nullptr, kSynRgbDefaultSet, 0,
                                           (uint8_t*)ctx->rgb.p, ctx->stream));
    }
    return 0;
}

// A band whose DN -> sample table is already known (general f32 rasters bridged through a key plane).
int dn_band_with_preset_lut(sarpro_ctx* ctx, int b, const BandJob& j, const uint16_t* lut_host, uint32_t max_key,
                            const OutGeom& g, void* canvas) {
    BandWs& w = ctx->band[b];
    RC(dn_pass_a_launch(ctx, b, j.dn, j.rows, j.cols, uses_clahe(j)));
    std::memcpy(ctx->h_lut + (size_t)b * kDnBins, lut_host, kDnBins * 2);
    CU(cudaMemcpyAsync(w.lut.p, ctx->h_lut + (size_t)b * kDnBins, kDnBins * 2, cudaMemcpyHostToDevice, ctx->stream));
    w.plan.any_valid = true;
    w.plan.have_invalid = true;
    w.plan.clahe = uses_clahe(j);
    w.plan.max_present_dn = max_key;
    w.plan.sat_from_dn = 0;
    w.hot = hmma_hot(lut_host, nullptr, max_key, &w.hot_top);
    w.dev_planned = false;
    RC(upload_plan_dev(ctx, b));
    return dn_run_pass_b(ctx, b, j, g, canvas);
}

// ---- inputs ----------------------------------------------------------------------------------------
// Narrowing upload of a large f32 host raster (see stage_band). Returns 0 when every chunk is queued on the copy stream,
// 1 when the raster has to go up as f32 (it is not u16-valued, or there is no pinned memory for the staging slots; the copy
// stream is drained, nothing of this band is left in flight), < 0 on errors.
static int narrow_and_stream(sarpro_ctx* ctx, int b, const sarpro_band* in, const uint16_t** dn_out) {
    BandWs& w = ctx->band[b];
    const uint64_t n = in->rows * in->cols;
    RC(ensure_upload_stream(ctx));
    const uint64_t per = ((in->rows + sarpro_ctx::kUploadChunks - 1) / sarpro_ctx::kUploadChunks + 127) & ~uint64_t(127);
    const size_t slot_bytes = per * in->cols * 2;
    if (ctx->narrow_ring_bytes < slot_bytes) {
        CU(cudaStreamSynchronize(ctx->stream_up));
        for (int s = 0; s < 2; ++s) {
            if (ctx->narrow_ring[s]) { cudaFreeHost(ctx->narrow_ring[s]); ctx->narrow_ring[s] = nullptr; }
            ctx->ring_busy[s] = false;
        }
        ctx->narrow_ring_bytes = 0;
        for (int s = 0; s < 2; ++s) {
            if (cudaHostAlloc(&ctx->narrow_ring[s], slot_bytes, cudaHostAllocDefault) != cudaSuccess) {
                // no pinned memory to spare (locked-memory limit): not an error of the call, the raster goes up as f32
                cudaGetLastError();
                ctx->narrow_ring[s] = nullptr;
                if (s == 1) { cudaFreeHost(ctx->narrow_ring[0]); ctx->narrow_ring[0] = nullptr; }
                ctx->narrow_upload = 0;
                return 1;
            }
            if (!ctx->ev_ring[s]) CU(cudaEventCreateWithFlags(&ctx->ev_ring[s], cudaEventDisableTiming));
        }
        ctx->narrow_ring_bytes = slot_bytes;
    }
    RC(reserve(ctx, w.dn, n * 2));
    sarpro_ctx::StreamedBand& sb = ctx->streamed[b];
    sb.n_chunks = 0;
    sb.ev0 = b * sarpro_ctx::kUploadChunks;
    ctx->upload_in_flight = true;
    const float* src = (const float*)in->data;
    int slot = 0;
    double conv_s = 0.0;
    for (uint64_t r = 0; r < in->rows; r += per, slot ^= 1) {
        const uint64_t r1 = std::min<uint64_t>(in->rows, r + per), cnt = (r1 - r) * in->cols;
        if (ctx->ring_busy[slot]) { CU(cudaEventSynchronize(ctx->ev_ring[slot])); ctx->ring_busy[slot] = false; }
        const auto t0 = std::chrono::steady_clock::now();
        const bool dn_valued = narrow_f32_to_dn(src + r * in->cols, (uint16_t*)ctx->narrow_ring[slot], cnt, ctx->valid_thresh);
        conv_s += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        if (!dn_valued) {
            CU(cudaStreamSynchronize(ctx->stream_up)); // the chunks already queued must not land behind the f32 path's writes
            ctx->ring_busy[0] = ctx->ring_busy[1] = false;
            return 1;
        }
        CU(cudaMemcpyAsync((char*)w.dn.p + r * in->cols * 2, ctx->narrow_ring[slot], cnt * 2, cudaMemcpyHostToDevice, ctx->stream_up));
        CU(cudaEventRecord(ctx->ev_chunk[sb.ev0 + sb.n_chunks], ctx->stream_up));
        CU(cudaEventRecord(ctx->ev_ring[slot], ctx->stream_up));
        ctx->ring_busy[slot] = true;
        sb.row_end[sb.n_chunks++] = (uint32_t)r1;
    }
    ctx->timing.h2d_bytes += n * 2;
    ctx->narrowed_bands++;
    // Narrowing only pays while the host threads read the f32 raster faster than PCIe would have carried it (~55 GB/s of pinned
    // host-to-device bandwidth on this platform, profiles/r02w_h2d_probe.log): a host that cannot keep that up goes back to
    // the f32 upload for the following calls of this context.
    ctx->narrow_gbs = conv_s > 0.0 ? (double)n * 4.0 / conv_s * 1e-9 : 0.0;
    if (ctx->narrow_upload == 1 && ctx->narrow_gbs < 60.0) ctx->narrow_upload = 0;
    *dn_out = (const uint16_t*)w.dn.p;
    return 0;
}

// Brings band `in` (optionally op(in, in2)) to a u16 DN raster on the device; *integral = false when the samples are not
// u16-valued (general f32 path). may_stream: large host rasters may arrive in row chunks on the copy stream (the caller's next
// consumer of the raster is pass A, which waits per chunk).
int stage_band(sarpro_ctx* ctx, int b, const sarpro_band* in, const sarpro_band* in2, int op, const uint16_t** dn_out,
               bool* integral, bool may_stream) {
    BandWs& w = ctx->band[b];
    const uint64_t n = in->rows * in->cols;
    *integral = true;
    if (in->dtype == SARPRO_DT_U16 && op < 0) {
        if (in->location == SARPRO_LOC_DEVICE) { *dn_out = (const uint16_t*)in->data; return 0; }
        RC(reserve(ctx, w.dn, n * 2));
        ctx->timing.h2d_bytes += n * 2;
        *dn_out = (const uint16_t*)w.dn.p;
        if (may_stream && ctx->stream_upload && b < 2 && n * 2 >= (64u << 20) && in->rows >= 1024) {
            // Streamed upload: row chunks on the copy stream, an event per chunk. Pass A consumes the chunks as they land and
            // the other band's plan / pass B run beside this band's upload (PCIe is the bottleneck of a call with host
            // rasters: 29 of 31 ms at C3). The previous call on this context ended with its streams drained, so nothing still
            // reads the staging buffer.
            RC(ensure_upload_stream(ctx));
            sarpro_ctx::StreamedBand& sb = ctx->streamed[b];
            const uint64_t per = ((in->rows + sarpro_ctx::kUploadChunks - 1) / sarpro_ctx::kUploadChunks + 127) & ~uint64_t(127);
            sb.n_chunks = 0;
            sb.ev0 = b * sarpro_ctx::kUploadChunks;
            ctx->upload_in_flight = true;
            for (uint64_t r = 0; r < in->rows; r += per) {
                const uint64_t r1 = std::min<uint64_t>(in->rows, r + per);
                CU(cudaMemcpyAsync((char*)w.dn.p + r * in->cols * 2, (const char*)in->data + r * in->cols * 2, (r1 - r) * in->cols * 2,
                                   cudaMemcpyHostToDevice, ctx->stream_up));
                CU(cudaEventRecord(ctx->ev_chunk[sb.ev0 + sb.n_chunks], ctx->stream_up));
                sb.row_end[sb.n_chunks++] = (uint32_t)r1;
            }
            return 0;
        }
        CU(cudaMemcpyAsync(w.dn.p, in->data, n * 2, cudaMemcpyHostToDevice, ctx->stream));
        return 0;
    }
    if (op >= 0 && in->dtype == SARPRO_DT_U16 && in2->dtype == SARPRO_DT_U16) {
        // raw DN pair through a polarization op: the loaders of the general path read the u16 rasters directly (2 B per sample
        // instead of the 4 B of the f32 the reference reads them as, gdal.rs:123; same values)
        const sarpro_band* bs[2] = {in, in2};
        DevBuf* stg[2] = {&w.f32a, &w.f32b};
        for (int k = 0; k < 2; ++k)
            if (bs[k]->location == SARPRO_LOC_HOST) {
                RC(reserve(ctx, *stg[k], n * 2));
                CU(cudaMemcpyAsync(stg[k]->p, bs[k]->data, n * 2, cudaMemcpyHostToDevice, ctx->stream));
                ctx->timing.h2d_bytes += n * 2;
            }
        *integral = false;
        *dn_out = nullptr;
        return 0;
    }
    if (in->dtype != SARPRO_DT_F32 || (op >= 0 && in2->dtype != SARPRO_DT_F32))
        return fail(ctx, SARPRO_ERR_INVALID_ARGUMENT, "polarization ops take two f32 bands or two u16 DN bands (ops.rs:4-44)");
    if (may_stream && op < 0 && in->location == SARPRO_LOC_HOST && ctx->narrow_upload && ctx->stream_upload && b < 2 && n * 2 >= (64u << 20) &&
        in->rows >= 1024) {
        // The reference's boundary: the f32 raster GDAL made of a u16 band (gdal.rs:123). PCIe is the whole end-to-end time of
        // such a call, so the host threads narrow each row chunk to DNs (narrow.cpp, the rule of k_f32_to_dn) into one of two
        // pinned slots while the previous chunk is on the wire, and the chunks then take the streamed path of a u16 raster: half
        // the bytes, pass A per chunk. A chunk with a sample that is not a DN ends the attempt: the raster goes up as f32 below.
        int rc = narrow_and_stream(ctx, b, in, dn_out);
        if (rc <= 0) return rc;          // 0: queued; < 0: error
        ctx->streamed[b].n_chunks = 0;   // 1: goes up as f32
    }
    const float* fa = (const float*)in->data;
    const float* fb = op >= 0 ? (const float*)in2->data : nullptr;
    if (in->location == SARPRO_LOC_HOST) {
        RC(reserve(ctx, w.f32a, n * 4));
        CU(cudaMemcpyAsync(w.f32a.p, in->data, n * 4, cudaMemcpyHostToDevice, ctx->stream));
        ctx->timing.h2d_bytes += n * 4;
        fa = (const float*)w.f32a.p;
    }
    if (op >= 0 && in2->location == SARPRO_LOC_HOST) {
        RC(reserve(ctx, w.f32b, n * 4));
        CU(cudaMemcpyAsync(w.f32b.p, in2->data, n * 4, cudaMemcpyHostToDevice, ctx->stream));
        ctx->timing.h2d_bytes += n * 4;
        fb = (const float*)w.f32b.p;
    }
    RC(reserve(ctx, w.dn, n * 2));
    RC(reserve(ctx, w.scalars, 64));
    CU(cudaMemsetAsync((uint32_t*)w.scalars.p + 3, 0, 4, ctx->stream));
    KS(SARPRO_STAGE_CONVERT, launch_f32_to_dn(fa, fb, op, n, ctx->valid_thresh, (uint16_t*)w.dn.p, (uint32_t*)w.scalars.p + 3, ctx->sm_count,
                        ctx->stream));
    CU(cudaMemcpyAsync(ctx->h_scalars + 8 * b + 3, (uint32_t*)w.scalars.p + 3, 4, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->timing.host_syncs++;
    *integral = ctx->h_scalars[8 * b + 3] == 0;
    *dn_out = (const uint16_t*)w.dn.p;
    return 0;
}

// Enum arguments of the C ABI (types.rs:8-14, 115-123, 170-173): anything outside the declared discriminants is an error,
// never a silent default. Pass -2 for an argument the entry point does not take.
int check_enums(sarpro_ctx* ctx, int op, int strategy, int bit_depth) {
    if (op != -2 && (op < -1 || op > SARPRO_OP_LOGRATIO))
        return fail(ctx, SARPRO_ERR_INVALID_ARGUMENT, "unknown polarization operation %d", op);
    if (strategy != -2 && (strategy < SARPRO_STRATEGY_STANDARD || strategy > SARPRO_STRATEGY_DEFAULT))
        return fail(ctx, SARPRO_ERR_INVALID_ARGUMENT, "unknown autoscale strategy %d", strategy);
    if (bit_depth != -2 && bit_depth != SARPRO_U8 && bit_depth != SARPRO_U16)
        return fail(ctx, SARPRO_ERR_INVALID_ARGUMENT, "unknown bit depth %d", bit_depth);
    return 0;
}

int check_band(sarpro_ctx* ctx, const sarpro_band* b) {
    if (!b) return fail(ctx, SARPRO_ERR_INVALID_ARGUMENT, "band descriptor is NULL");
    if (b->rows * b->cols > 0 && !b->data) return fail(ctx, SARPRO_ERR_INVALID_ARGUMENT, "band data is NULL");
    if (b->rows >= (1ull << 31) || b->cols >= (1ull << 31) || b->rows * b->cols >= (1ull << 32))
        return fail(ctx, SARPRO_ERR_TOO_LARGE, "raster %llux%llu exceeds 2^32 samples", (unsigned long long)b->rows,
                    (unsigned long long)b->cols);
    return 0;
}

double host_ms() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
int begin_call(sarpro_ctx* ctx) {
    if (!ctx) return SARPRO_ERR_INVALID_ARGUMENT;
    ctx->err.clear();
    CU(cudaSetDevice(ctx->device));
    if (ctx->axes.size() > 64) drop_axis_plans(ctx); // bounded cache; no plan pointer is held across calls
    std::memset(&ctx->timing, 0, sizeof(ctx->timing));
    ctx->pending_stats[0] = ctx->pending_stats[1] = nullptr;
    ctx->streamed[0].n_chunks = ctx->streamed[1].n_chunks = 0;
    ctx->band[0].pitch = ctx->band[1].pitch = 0;
    if (ctx->upload_in_flight && ctx->stream_up) { cudaStreamSynchronize(ctx->stream_up); ctx->upload_in_flight = false; }
    if (!ctx->keep_last)
        for (auto& l : ctx->last) l = sarpro_ctx::LastResult{}; // the buffers are about to be reused
    ctx->band[0].plan_copy_pending = ctx->band[1].plan_copy_pending = false;
    ctx->n_sev = 0;
    ctx->host_t0 = host_ms();
    CU(cudaEventRecord(ctx->ev[0], ctx->stream));
    return 0;
}
int end_call(sarpro_ctx* ctx) {
    CU(cudaEventRecord(ctx->ev[1], ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->upload_in_flight = false; // every chunk event was waited for by the work that just completed
    for (int b = 0; b < 2; ++b) {
        BandWs& w = ctx->band[b];
        if (w.plan_copy_pending) { // device-planned: the plan's host mirror is valid now
            const PlanDev& p = ctx->h_plan[b];
            w.plan.any_valid = p.any_valid != 0;
            w.plan.have_invalid = p.have_invalid != 0;
            w.plan.max_present_dn = p.max_present_dn;
            w.plan.sat_from_dn = p.sat_from_dn;
            w.plan.pre_min = (uint16_t)p.pre_min;
            w.plan.pre_max = (uint16_t)p.pre_max;
            w.plan.px_total = p.px_total;
            w.plan.px_ge1024 = p.px_ge1024;
            w.plan.px_ge2048 = p.px_ge2048;
            w.plan.stats = p.stats;
            w.hot = p.hot;
            w.hot_top = p.hot_top;
            w.plan_copy_pending = false;
        }
        if (ctx->pending_stats[b]) {
            *ctx->pending_stats[b] = w.plan.stats;
            ctx->pending_stats[b] = nullptr;
        }
        if (w.hist_auto_pending) { // pass-A table shape of the next call on this slot
            w.hist_auto = choose_hist_variant(w.plan);
            w.hist_auto_pending = false;
        }
    }
    float ms = 0;
    CU(cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]));
    ctx->timing.total_ms = ms;
    for (int i = 0; i < ctx->n_sev; ++i) {
        float t = 0;
        CU(cudaEventElapsedTime(&t, ctx->sev[2 * i], ctx->sev[2 * i + 1]));
        ctx->timing.stage_ms[ctx->sev_stage[i]] += t;
        ctx->timing.stage_launches[ctx->sev_stage[i]]++;
    }
    if (getenv("SARPRO_TRACE")) { // exploration: device timeline of the stage-timed launches of this call
        for (int i = 0; i < ctx->n_sev; ++i) {
            float t0 = 0, t1 = 0;
            cudaEventElapsedTime(&t0, ctx->ev[0], ctx->sev[2 * i]);
            cudaEventElapsedTime(&t1, ctx->ev[0], ctx->sev[2 * i + 1]);
            fprintf(stderr, "trace %2d stage %d  %8.3f -> %8.3f ms (%.3f)  issued by the host at %.3f\n", i, ctx->sev_stage[i], t0, t1, t1 - t0,
                    ctx->sev_host[i]);
        }
        fprintf(stderr, "trace total %.3f ms\n", ms);
    }
    return 0;
}

int deliver(sarpro_ctx* ctx, const void* dev_src, size_t bytes, sarpro_image* out) {
    if (out->location == SARPRO_LOC_NONE) return 0; // the result stays in the context (sarpro_encode_last_jpeg)
    if (!out->data && bytes) return fail(ctx, SARPRO_ERR_INVALID_ARGUMENT, "output buffer is NULL");
    if (out->capacity_bytes < bytes)
        return fail(ctx, SARPRO_ERR_INVALID_ARGUMENT, "output buffer holds %llu bytes, %llu needed",
                    (unsigned long long)out->capacity_bytes, (unsigned long long)bytes);
    if (!bytes) return 0;
    if (out->location == SARPRO_LOC_DEVICE) {
        if (out->data != dev_src) CU(cudaMemcpyAsync(out->data, dev_src, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
    } else {
        CU(cudaMemcpyAsync(out->data, dev_src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
        ctx->timing.d2h_bytes += bytes;
    }
    return 0;
}

void fill_image(sarpro_image* out, const OutGeom& g, int channels, int bit_depth) {
    out->cols = g.oc;
    out->rows = g.orr;
    out->channels = channels;
    out->bit_depth = bit_depth;
    out->meta = g.meta;
}

} // namespace sarpro

namespace {

// Produces one processed band (autoscale [+resize+pad]) into a device canvas owned by the ctx.
// Returns the canvas pointer through *canvas. `slot` selects the workspace.
int produce_bands(sarpro_ctx* ctx, const sarpro_band* const* ins, const sarpro_band* const* ins2, const int* ops, int nb,
                  const int* strategies, const int* bit_depths, const PlanKind* kinds, bool has_target, size_t target,
                  bool pad, void** canvases, OutGeom* geom, sarpro_stats* stats) {
    const uint64_t rows = ins[0]->rows, cols = ins[0]->cols;
    *geom = out_geometry(cols, rows, has_target, target, pad);
    BandJob jobs[2];
    bool integral[2] = {true, true};
    for (int b = 0; b < nb; ++b) {
        RC(check_band(ctx, ins[b]));
        if (ops[b] >= 0) RC(check_band(ctx, ins2[b]));
        if (ins[b]->rows != rows || ins[b]->cols != cols || (ops[b] >= 0 && (ins2[b]->rows != rows || ins2[b]->cols != cols)))
            return fail(ctx, SARPRO_ERR_INVALID_ARGUMENT, "bands differ in shape");
        jobs[b].rows = rows;
        jobs[b].cols = cols;
        jobs[b].strategy = strategies[b];
        jobs[b].bit_depth = bit_depths[b];
        jobs[b].kind = kinds[b];
    }
    const size_t n_out = geom->oc * geom->orr;
    if (rows * cols == 0) {
        for (int b = 0; b < nb; ++b) {
            const bool out8 = kinds[b] != PlanKind::Autoscale || bit_depths[b] == SARPRO_U8;
            RC(reserve(ctx, ctx->band[b].small, std::max<size_t>(n_out * (out8 ? 1 : 2), 16)));
            canvases[b] = ctx->band[b].small.p;
            if (stats) std::memset(&stats[b], 0, sizeof(sarpro_stats));
        }
        return 0;
    }
    // A resized output of a DN raster whose width is not a multiple of 8 would miss the tensor-core pass B, the third-generation
    // histogram kernel and the 128-bit loads of the generic kernel (all need 16-byte aligned rows): such rasters are re-pitched
    // to the next multiple of 8 columns, the edge sample replicated into the 1..7 padding columns (which carry no tap and are
    // kept out of every statistic). A host raster is re-pitched by its upload (a 2-D copy, no extra pass), a device raster by
    // one kernel (4 B per sample).
    bool want_pad[2] = {false, false};
    for (int b = 0; b < nb; ++b)
        want_pad[b] = geom->resize && (cols % 8) != 0 && cols >= 512 && ops[b] < 0 && ctx->repitch && !ctx->force_exact;
    const uint64_t pitch = (cols + 7) & ~uint64_t(7);
    for (int b = 0; b < nb; ++b) {
        BandWs& w = ctx->band[b];
        if (want_pad[b] && ins[b]->dtype == SARPRO_DT_U16 && ins[b]->location == SARPRO_LOC_HOST) {
            RC(reserve(ctx, w.dn_pad, rows * pitch * 2));
            CU(cudaMemcpy2DAsync(w.dn_pad.p, pitch * 2, ins[b]->data, cols * 2, cols * 2, rows, cudaMemcpyHostToDevice, ctx->stream));
            ctx->timing.h2d_bytes += rows * cols * 2;
            KL(launch_pad_cols((uint16_t*)w.dn_pad.p, (uint32_t)rows, (uint32_t)cols, (uint32_t)pitch, ctx->stream));
            jobs[b].dn = (const uint16_t*)w.dn_pad.p;
            w.pitch = pitch;
            continue;
        }
        // (a raster that still has to be re-pitched on the device is read by that kernel at once: no streamed chunks)
        RC(stage_band(ctx, b, ins[b], ins2[b], ops[b], &jobs[b].dn, &integral[b], !want_pad[b]));
        if (want_pad[b] && integral[b]) {
            RC(reserve(ctx, w.dn_pad, rows * pitch * 2));
            KL(launch_repitch(jobs[b].dn, (uint16_t*)w.dn_pad.p, (uint32_t)rows, (uint32_t)cols, (uint32_t)pitch, ctx->sm_count, ctx->stream));
            jobs[b].dn = (const uint16_t*)w.dn_pad.p;
            w.pitch = pitch;
        }
    }
    // u16-valued rasters share one pass A + planner round trip
    BandJob dnjobs[2];
    int dnidx[2], ndn = 0;
    for (int b = 0; b < nb; ++b)
        if (integral[b]) { dnjobs[ndn] = jobs[b]; dnidx[ndn] = b; ndn++; }
    const bool pipelined = ndn == nb && ndn > 0; // plan band b on the host while the device works on the other band
    ctx->pair_spare = (pipelined && nb == 2 && ctx->two_stream && ctx->stream2) ? ctx->spare_sms : 0;
    if (pipelined) {
        RC(run_pass_a(ctx, jobs, nb));
    } else {
        for (int k = 0; k < ndn; ++k) {
            // mixed case: plan band by band in its own slot
            if (dnidx[k] != 0) { std::swap(ctx->band[0], ctx->band[dnidx[k]]); std::swap(ctx->streamed[0], ctx->streamed[dnidx[k]]); }
            int rc = run_pass_a_and_plan(ctx, &dnjobs[k], 1);
            if (dnidx[k] != 0) { std::swap(ctx->band[0], ctx->band[dnidx[k]]); std::swap(ctx->streamed[0], ctx->streamed[dnidx[k]]); }
            RC(rc);
        }
    }
    bool join_pending = false;
    int break_rc = 0;
    for (int b = 0; b < nb; ++b) {
        BandWs& w = ctx->band[b];
        const bool out8 = kinds[b] != PlanKind::Autoscale || bit_depths[b] == SARPRO_U8;
        const size_t esz = out8 ? 1 : 2;
        if ((break_rc = reserve(ctx, w.small, std::max<size_t>(n_out * esz, 16)))) break;
        canvases[b] = w.small.p;
        if (!integral[b]) {
            const void* fa = ins[b]->location == SARPRO_LOC_HOST ? w.f32a.p : ins[b]->data;
            const void* fb = ops[b] >= 0 ? (ins2[b]->location == SARPRO_LOC_HOST ? w.f32b.p : ins2[b]->data) : nullptr;
            const int a16 = ins[b]->dtype == SARPRO_DT_U16, b16 = ops[b] >= 0 && ins2[b]->dtype == SARPRO_DT_U16;
            RC(f32_general_single(ctx, b, fa, fb, a16, b16, ops[b], rows, cols, bit_depths[b], strategies[b], kinds[b], *geom, w.small.p,
                                  stats ? &stats[b] : nullptr));
            continue;
        }
        // First band of a pipelined pair: its planner upload, CLAHE statistics, pass B and vertical pass go to the side
        // stream. Its pass A is complete (host-synced in wait_and_plan), so the table upload and the CLAHE statistics
        // run while the second band's pass A still streams on the main stream (they used to queue behind it: 60 us of
        // idle device between the two passes), and pass B's persistent CTAs move in as that pass A drains. The second
        // band's chain follows its pass A on the main stream; its persistent CTAs fill the SMs the first band's pass B
        // leaves early (the piece runs do not end together). The main stream joins after both chains are queued.
        const bool side = pipelined && nb == 2 && b == 0 && ctx->two_stream && ctx->stream2;
        cudaStream_t main_stream = ctx->stream;
        int rc = 0;
        if (pipelined) rc = plan_band_on(ctx, b, jobs[b], side ? ctx->stream2 : nullptr);
        if (!rc) {
            if (stats) {
                if (w.dev_planned) ctx->pending_stats[b] = &stats[b]; // filled in end_call from the plan's host mirror
                else stats[b] = w.plan.stats;
            }
            rc = dn_run_pass_b(ctx, b, jobs[b], *geom, w.small.p);
        }
        if (side) {
            cudaError_t e1 = cudaEventRecord(ctx->ev_join, ctx->stream2);
            ctx->stream = main_stream;
            join_pending = true;
            if (!rc) CU(e1);
        }
        if (rc) break_rc = rc;
        if (rc) break;
    }
    if (join_pending) { // also on the error path: the side stream must not run past this call unobserved
        cudaError_t e2 = cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0);
        if (!break_rc) CU(e2);
    }
    return break_rc;
}

} // namespace

// =================================================================================================
// C ABI
// =================================================================================================
extern "C" {

int sarpro_abi_version(void) { return SARPRO_GPU_ABI_VERSION; }

const char* sarpro_last_error(const sarpro_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int sarpro_ctx_create(sarpro_ctx** out, int device_id) {
    if (!out) return SARPRO_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return fail(nullptr, SARPRO_ERR_NO_DEVICE, "no CUDA device available (%s); libsarpro_gpu has no CPU fallback",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    if (device_id < 0 || device_id >= n) return fail(nullptr, SARPRO_ERR_INVALID_ARGUMENT, "device %d out of range [0,%d)", device_id, n);
    sarpro_ctx* ctx = new sarpro_ctx();
    ctx->device = device_id;
    auto bail = [&](const char* what, cudaError_t err) {
        fail(nullptr, SARPRO_ERR_CUDA, "%s failed: %s", what, cudaGetErrorString(err));
        delete ctx;
        return SARPRO_ERR_CUDA;
    };
    if ((e = cudaSetDevice(device_id)) != cudaSuccess) return bail("cudaSetDevice", e);
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device_id)) != cudaSuccess) return bail("cudaGetDeviceProperties", e);
    if (prop.major < 10) {
        fail(nullptr, SARPRO_ERR_NO_DEVICE, "device %d is sm_%d%d; this library is built for sm_100a only", device_id, prop.major, prop.minor);
        delete ctx;
        return SARPRO_ERR_NO_DEVICE;
    }
    ctx->sm_count = prop.multiProcessorCount;
    if ((e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess) return bail("cudaStreamCreate", e);
    if ((e = cudaStreamCreateWithFlags(&ctx->stream2, cudaStreamNonBlocking)) != cudaSuccess) return bail("cudaStreamCreate", e);
    if ((e = cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming)) != cudaSuccess) return bail("cudaEventCreate", e);
    for (auto& ev : ctx->ev)
        if ((e = cudaEventCreate(&ev)) != cudaSuccess) return bail("cudaEventCreate", e);
    for (auto& ev : ctx->sev)
        if ((e = cudaEventCreate(&ev)) != cudaSuccess) return bail("cudaEventCreate", e);
    if ((e = cudaMallocHost((void**)&ctx->h_hist, 2 * kDnBins * 4)) != cudaSuccess) return bail("cudaMallocHost", e);
    if ((e = cudaMallocHost((void**)&ctx->h_present, 2 * kPresentWords * 4)) != cudaSuccess) return bail("cudaMallocHost", e);
    if ((e = cudaMallocHost((void**)&ctx->h_lut, 2 * kDnBins * 2)) != cudaSuccess) return bail("cudaMallocHost", e);
    if ((e = cudaMallocHost((void**)&ctx->h_scalars, 2 * 8 * 4)) != cudaSuccess) return bail("cudaMallocHost", e);
    if ((e = cudaMallocHost((void**)&ctx->h_remap, 2 * 256)) != cudaSuccess) return bail("cudaMallocHost", e);
    if ((e = cudaMallocHost((void**)&ctx->h_plan, 2 * sizeof(PlanDev))) != cudaSuccess) return bail("cudaMallocHost", e);
    if ((e = cudaMallocHost((void**)&ctx->h_plan_up, 2 * sizeof(PlanDev))) != cudaSuccess) return bail("cudaMallocHost", e);
    ctx->valid_thresh = compute_valid_thresh();
    // the DN -> dB table of the device planner (pipeline.rs:19-20, evaluated with the host libm like the reference's)
    if ((e = cudaMalloc(&ctx->db_table.p, kDnBins * sizeof(double))) != cudaSuccess) return bail("cudaMalloc", e);
    ctx->db_table.cap = kDnBins * sizeof(double);
    if ((e = cudaMemcpy(ctx->db_table.p, dn_db_table(), kDnBins * sizeof(double), cudaMemcpyHostToDevice)) != cudaSuccess) return bail("cudaMemcpy", e);
    if (const char* v = getenv("SARPRO_HOST_PLAN")) ctx->host_plan = atoi(v);
    if (const char* v = getenv("SARPRO_F32_NO_GUARD")) ctx->f32_no_guard = atoi(v);
    if (const char* v = getenv("SARPRO_STREAM_UPLOAD")) ctx->stream_upload = atoi(v);
    if (const char* v = getenv("SARPRO_NARROW_UPLOAD")) ctx->narrow_upload = atoi(v);
    if (const char* v = getenv("SARPRO_REPITCH")) ctx->repitch = atoi(v);
    if (const char* v = getenv("SARPRO_HIST_VARIANT")) ctx->hist_variant = atoi(v);
    if (const char* v = getenv("SARPRO_FORCE_EXACT")) ctx->force_exact = atoi(v);
    if (const char* v = getenv("SARPRO_HMMA")) ctx->use_hmma = atoi(v);
    if (const char* v = getenv("SARPRO_TWO_STREAM")) ctx->two_stream = atoi(v);
    if (const char* v = getenv("SARPRO_SPARE_SMS")) ctx->spare_sms = std::max(0, std::min(16, atoi(v)));
    if (getenv("SARPRO_TRACE") || (getenv("SARPRO_STAGE_TIMING") && std::string(getenv("SARPRO_STAGE_TIMING")) == "all")) ctx->stage_mask = 0xffu;
    int rc = upload_rgb_luts(ctx);
    if (rc) {
        g_create_error = ctx->err;
        sarpro_ctx_destroy(ctx);
        return rc;
    }
    *out = ctx;
    return SARPRO_OK;
}

void sarpro_ctx_destroy(sarpro_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    if (ctx->stream2) cudaStreamSynchronize(ctx->stream2);
    sarpro_comm_destroy(ctx); // NCCL communicator + CommState
    jpeg_state_destroy(ctx);
    for (auto& w : ctx->band)
        for (DevBuf* b : {&w.dn, &w.dn_pad, &w.f32a, &w.f32b, &w.tile_hist, &w.total, &w.lut, &w.tile256, &w.cdf, &w.cdf32, &w.remap,
                          &w.temp, &w.small, &w.full, &w.scalars, &w.present, &w.edges, &w.hist4096, &w.f32scan, &w.pieces,
                          &w.cta_first, &w.plan_dev, &w.plan_scratch})
            release(*b);
    for (DevBuf* b : {&ctx->units, &ctx->tile_px, &ctx->col_dx, &ctx->col_omdx, &ctx->col_t, &ctx->row_dy, &ctx->row_omdy,
                      &ctx->row_t, &ctx->rgb, &ctx->hist256, &ctx->rgbsel, &ctx->rgb_luts, &ctx->col_m, &ctx->row_sat, &ctx->db_table})
        release(*b);
    for (auto& slot : ctx->batch_stage)
        for (DevBuf& b : slot) release(b);
    release(ctx->gather);
    release(ctx->units_by_row);
    if (ctx->stream_up) { cudaStreamSynchronize(ctx->stream_up); cudaStreamDestroy(ctx->stream_up); }
    for (int s = 0; s < 2; ++s) {
        if (ctx->narrow_ring[s]) cudaFreeHost(ctx->narrow_ring[s]);
        if (ctx->ev_ring[s]) cudaEventDestroy(ctx->ev_ring[s]);
    }
    for (auto& ev : ctx->ev_up)
        if (ev) cudaEventDestroy(ev);
    for (auto& ev : ctx->ev_chunk)
        if (ev) cudaEventDestroy(ev);
    drop_axis_plans(ctx);
    if (ctx->h_hist) cudaFreeHost(ctx->h_hist);
    if (ctx->h_present) cudaFreeHost(ctx->h_present);
    if (ctx->h_lut) cudaFreeHost(ctx->h_lut);
    if (ctx->h_scalars) cudaFreeHost(ctx->h_scalars);
    if (ctx->h_remap) cudaFreeHost(ctx->h_remap);
    if (ctx->h_plan) cudaFreeHost(ctx->h_plan);
    if (ctx->h_plan_up) cudaFreeHost(ctx->h_plan_up);
    for (auto& ev : ctx->ev)
        if (ev) cudaEventDestroy(ev);
    for (auto& ev : ctx->sev)
        if (ev) cudaEventDestroy(ev);
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    if (ctx->stream2) cudaStreamDestroy(ctx->stream2);
    if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
    delete ctx;
}

int sarpro_ctx_set_stream(sarpro_ctx* ctx, void* cuda_stream) {
    if (!ctx) return SARPRO_ERR_INVALID_ARGUMENT;
    if (ctx->own_stream && ctx->stream) { cudaStreamSynchronize(ctx->stream); cudaStreamDestroy(ctx->stream); }
    ctx->stream = (cudaStream_t)cuda_stream;
    ctx->own_stream = false;
    return SARPRO_OK;
}
int sarpro_ctx_synchronize(sarpro_ctx* ctx) {
    if (!ctx) return SARPRO_ERR_INVALID_ARGUMENT;
    CU(cudaStreamSynchronize(ctx->stream));
    return SARPRO_OK;
}
int sarpro_last_timing(const sarpro_ctx* ctx, sarpro_timing* out) {
    if (!ctx || !out) return SARPRO_ERR_INVALID_ARGUMENT;
    *out = ctx->timing;
    return SARPRO_OK;
}
int sarpro_host_alloc(void** out, size_t bytes) {
    if (!out) return SARPRO_ERR_INVALID_ARGUMENT;
    return cudaMallocHost(out, bytes) == cudaSuccess ? SARPRO_OK : SARPRO_ERR_OUT_OF_MEMORY;
}
void sarpro_host_free(void* p) { if (p) cudaFreeHost(p); }
int sarpro_host_register(void* p, size_t bytes) {
    return cudaHostRegister(p, bytes, cudaHostRegisterDefault) == cudaSuccess ? SARPRO_OK : SARPRO_ERR_CUDA;
}
void sarpro_host_unregister(void* p) { if (p) cudaHostUnregister(p); }

int sarpro_resize_output_dims(size_t cols, size_t rows, int has_target, size_t target, int pad, size_t* out_cols,
                              size_t* out_rows) {
    if (!out_cols || !out_rows) return SARPRO_ERR_INVALID_ARGUMENT;
    size_t rc, rr;
    resize_output_dims(cols, rows, has_target != 0, target, pad != 0, &rc, &rr, out_cols, out_rows);
    return SARPRO_OK;
}

int sarpro_lanczos_row_plan_check(const uint8_t* samples, size_t in_size, size_t out_size, size_t max_span, size_t strip_ntiles,
                                  uint8_t* out_direct, uint8_t* out_replay) {
    if (!samples || !out_direct || !out_replay || in_size == 0 || out_size == 0) return SARPRO_ERR_INVALID_ARGUMENT;
    ResampleAxis ax;
    build_lanczos3_axis((uint32_t)in_size, (uint32_t)out_size, false, &ax);
    for (uint32_t ox = 0; ox < out_size; ++ox) { // fast_image_resize u8 horizontal pass: i32 accumulate from 1 << (p - 1), >> p, clamp
        int acc = ax.precision > 0 ? (1 << (ax.precision - 1)) : 0;
        for (uint32_t k = 0; k < ax.size[ox]; ++k) acc += (int)samples[ax.start[ox] + k] * ax.coef[(size_t)ox * ax.window + k];
        int v = acc >> ax.precision;
        out_direct[ox] = (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
    }
    HMmaPlanHost plan;
    if (!hmma_build_plan(ax.start.data(), ax.size.data(), ax.coef.data(), ax.window, (uint32_t)out_size, (uint32_t)in_size, (uint32_t)max_span, &plan,
                         strip_ntiles ? (uint32_t)strip_ntiles : 32u))
        return 0;
    hmma_replay_row(plan, samples, (uint32_t)in_size, (uint32_t)out_size, ax.precision, out_replay);
    return 1;
}

int sarpro_lanczos_row_check_u16(const uint16_t* samples, size_t in_size, size_t out_size, uint16_t* out) {
    if (!samples || !out || in_size == 0 || out_size == 0) return SARPRO_ERR_INVALID_ARGUMENT;
    ResampleAxis ax;
    build_lanczos3_axis((uint32_t)in_size, (uint32_t)out_size, true, &ax); // the table k_hresize<., PIX16> / k_vresize read
    for (uint32_t ox = 0; ox < out_size; ++ox) { // fast_image_resize u16 pass: i32 taps, i64 accumulate from 1 << (p - 1), >> p, clamp
        long long acc = ax.precision > 0 ? (1ll << (ax.precision - 1)) : 0ll;
        for (uint32_t k = 0; k < ax.size[ox]; ++k) acc += (long long)ax.coef[(size_t)ox * ax.window + k] * (long long)samples[ax.start[ox] + k];
        const long long v = acc >> ax.precision;
        out[ox] = (uint16_t)(v < 0 ? 0 : (v > 65535 ? 65535 : v));
    }
    return SARPRO_OK;
}

int sarpro_f32_guard_params(double low_db, double range_db, uint32_t n, float min_v, float max_v, int* e0, float* f0, float* scale,
                            float* guard) {
    if (!e0 || !f0 || !scale || !guard) return SARPRO_ERR_INVALID_ARGUMENT;
    f32_guard(true, low_db, range_db, n, min_v, max_v, e0, f0, scale, guard);
    return SARPRO_OK;
}

int sarpro_synrgb_lut_check(int set, const uint32_t* hist256, uint64_t n_per_band, int* floor_with_cushion, uint8_t* lut_r, uint8_t* lut_g,
                            uint8_t* lut_b) {
    if (set < -1 || set > (int)kSynRgbDefaultSet || !lut_r || !lut_g || !lut_b || (set < 0 && !hist256)) return SARPRO_ERR_INVALID_ARGUMENT;
    if (set < 0) set = synrgb_floor_from_histogram(hist256, n_per_band); // the host mirror of k_synrgb_floor
    if (floor_with_cushion) *floor_with_cushion = set == (int)kSynRgbDefaultSet ? -1 : set;
    SynRgbLut l; // built exactly like the sets load_rgb_luts uploads
    if (set == (int)kSynRgbDefaultSet) build_synrgb_default_lut(&l);
    else build_synrgb_suppressed_lut(set, &l);
    std::memcpy(lut_r, l.r, 256);
    std::memcpy(lut_g, l.g, 256);
    std::memcpy(lut_b, l.b.data(), 65536);
    return SARPRO_OK;
}

int sarpro_plan_from_stat_histogram(const uint64_t* hist4096, uint64_t valid_count, float min_v, float max_v, double mean_db, double std_db,
                                    int strategy, int tamed_synrgb_kind, sarpro_stats* stats) {
    if (!hist4096 || !stats || strategy < SARPRO_STRATEGY_STANDARD || strategy > SARPRO_STRATEGY_DEFAULT || tamed_synrgb_kind < 0 ||
        tamed_synrgb_kind > 2)
        return SARPRO_ERR_INVALID_ARGUMENT;
    std::memset(stats, 0, sizeof(*stats));
    if (valid_count == 0) return SARPRO_OK;
    // the sequence of api_f32.cu after its second pass
    stats_from_stat_histogram(hist4096, valid_count, db_of_sample(min_v), db_of_sample(max_v), mean_db, std_db, stats);
    choose_window(strategy, tamed_synrgb_kind == 0 ? PlanKind::Autoscale : (tamed_synrgb_kind == 1 ? PlanKind::TamedSynRgbCopol : PlanKind::TamedSynRgbCross),
                  stats);
    return SARPRO_OK;
}

int sarpro_narrow_f32_check(const float* src, size_t n, uint16_t* dst, int* u16_valued) {
    if ((n && (!src || !dst)) || !u16_valued) return SARPRO_ERR_INVALID_ARGUMENT;
    *u16_valued = narrow_f32_to_dn(src, dst, n, valid_threshold()) ? 1 : 0;
    return SARPRO_OK;
}

int sarpro_f32_edges_check(int kind, double low_db, double high_db, double gamma, uint32_t n_levels, float min_v, float max_v,
                           uint32_t* n_analytic, uint32_t* n_mismatch, float* edges_out) {
    if (!n_analytic || !n_mismatch || !(min_v > 0.f) || !(max_v >= min_v) || kind < -1 || kind > 2) return SARPRO_ERR_INVALID_ARGUMENT;
    if (kind >= 0 && (n_levels == 0 || n_levels > 65535u)) return SARPRO_ERR_INVALID_ARGUMENT;
    std::vector<float> fast, slow;
    auto build = [&](std::vector<float>* e) {
        if (kind < 0) build_stat_edges(min_v, max_v, e);
        else build_level_edges((LevelKind)kind, low_db, high_db, gamma, n_levels, min_v, max_v, e, nullptr, nullptr);
    };
    const uint64_t h0 = f32_edges_analytic_hits();
    build(&fast);
    *n_analytic = (uint32_t)(f32_edges_analytic_hits() - h0);
    f32_edges_set_analytic(false);
    build(&slow);
    f32_edges_set_analytic(true);
    uint32_t bad = fast.size() != slow.size();
    for (size_t i = 0; i < fast.size() && i < slow.size(); ++i) bad += std::memcmp(&fast[i], &slow[i], 4) != 0;
    *n_mismatch = bad;
    if (edges_out) std::memcpy(edges_out, fast.data(), fast.size() * 4);
    return (int)fast.size();
}

int sarpro_plan_from_present_list(const uint32_t* blocks, const uint32_t* pairs, uint32_t cap, int bit_depth, int strategy,
                                  sarpro_stats* stats, uint16_t* lut16) {
    if (!blocks || !pairs) return SARPRO_ERR_INVALID_ARGUMENT;
    BandPlan p;
    if (!plan_from_present_list(blocks, pairs, cap, bit_depth, strategy, PlanKind::Autoscale, &p)) return 0;
    if (stats) *stats = p.stats;
    if (lut16) std::memcpy(lut16, p.lut.data(), kDnBins * 2);
    return 1;
}

int sarpro_plan_from_dn_histogram(const uint64_t* hist65536, int bit_depth, int strategy, sarpro_stats* stats,
                                  uint16_t* lut16) {
    if (!hist65536) return SARPRO_ERR_INVALID_ARGUMENT;
    BandPlan p;
    plan_from_dn_histogram(hist65536, bit_depth, strategy, PlanKind::Autoscale, &p);
    if (stats) *stats = p.stats;
    if (lut16) std::memcpy(lut16, p.lut.data(), kDnBins * 2);
    return SARPRO_OK;
}

int sarpro_plan_on_device(sarpro_ctx* ctx, const uint32_t* hist65536, int bit_depth, int strategy, int plan_kind,
                          sarpro_stats* stats, uint16_t* lut16, uint32_t* hot2) {
    RC(begin_call(ctx));
    if (!hist65536) return fail(ctx, SARPRO_ERR_INVALID_ARGUMENT, "NULL argument");
    RC(check_enums(ctx, -2, strategy, bit_depth));
    if (plan_kind < 0 || plan_kind > 2 || !plan_on_device_supported(strategy, plan_kind))
        return fail(ctx, SARPRO_ERR_INVALID_ARGUMENT, "strategy %d (kind %d) is planned on the host", strategy, plan_kind);
    BandWs& w = ctx->band[0];
    RC(reserve(ctx, w.total, kDnBins * 4));
    RC(reserve(ctx, w.lut, kDnBins * 2));
    CU(cudaMemcpyAsync(w.total.p, hist65536, kDnBins * 4, cudaMemcpyHostToDevice, ctx->stream));
    BandJob job;
    job.strategy = strategy;
    job.bit_depth = bit_depth;
    job.kind = (PlanKind)plan_kind;
    RC(plan_band_on_device(ctx, 0, job));
    if (lut16) CU(cudaMemcpyAsync(lut16, w.lut.p, kDnBins * 2, cudaMemcpyDeviceToHost, ctx->stream));
    RC(end_call(ctx));
    if (stats) *stats = w.plan.stats;
    if (hot2) { hot2[0] = w.hot; hot2[1] = w.hot_top; }
    return SARPRO_OK;
}

int sarpro_plan_kind_from_dn_histogram(const uint64_t* hist65536, int bit_depth, int strategy, int plan_kind,
                                       sarpro_stats* stats, uint16_t* lut16, uint32_t* hot2) {
    if (!hist65536 || plan_kind < 0 || plan_kind > 2) return SARPRO_ERR_INVALID_ARGUMENT;
    BandPlan p;
    plan_from_dn_histogram(hist65536, bit_depth, strategy, (PlanKind)plan_kind, &p);
    if (stats) *stats = p.stats;
    if (lut16) std::memcpy(lut16, p.lut.data(), kDnBins * 2);
    if (hot2) {
        hot2[1] = 0;
        hot2[0] = p.any_valid ? hmma_hot_from_plan(p, &hot2[1]) : 0;
    }
    return SARPRO_OK;
}

// ---- fused pipelines ---------------------------------------------------------------------------------
int sarpro_pipeline_single(sarpro_ctx* ctx, const sarpro_band* a, const sarpro_band* b, int op, int format, int bit_depth,
                           int strategy, int has_target, size_t target, int pad, sarpro_image* out, sarpro_stats* stats) {
    RC(begin_call(ctx));
    if (!a || !out) return fail(ctx, SARPRO_ERR_INVALID_ARGUMENT, "NULL argument");
    RC(check_enums(ctx, op, strategy, bit_depth));
    if (op >= 0 && !b) return fail(ctx, SARPRO_ERR_INVALID_ARGUMENT, "a polarization operation needs two bands");
    if (format != SARPRO_FORMAT_TIFF && format != SARPRO_FORMAT_JPEG) return fail(ctx, SARPRO_ERR_INVALID_ARGUMENT, "unknown output format %d", format);
    if (format == SARPRO_FORMAT_JPEG) bit_depth = SARPRO_U8; // save.rs:121
    const sarpro_band* ins[1] = {a};
    const sarpro_band* ins2[1] = {b};
    const int ops[1] = {op};
    const int strategies[1] = {strategy}, depths[1] = {bit_depth};
    const PlanKind kinds[1] = {PlanKind::Autoscale};
    void* canvas[1];
    OutGeom g;
    RC(produce_bands(ctx, ins, ins2, ops, 1, strategies, depths, kinds, has_target != 0, target, pad != 0, canvas, &g, stats));
    fill_image(out, g, 1, bit_depth);
    RC(deliver(ctx, canvas[0], g.oc * g.orr * (bit_depth == SARPRO_U8 ? 1 : 2), out));
    if (bit_depth == SARPRO_U8) ctx->last[1] = sarpro_ctx::LastResult{canvas[0], g.oc, g.orr};
    return end_call(ctx);
}

int sarpro_pipeline_multiband_tiff(sarpro_ctx* ctx, const sarpro_band* b1, const sarpro_band* b2, int bit_depth,
                                   int strategy, int has_target, size_t target, int pad, sarpro_image* out1,
                                   sarpro_image* out2, sarpro_stats* stats2) {
    RC(begin_call(ctx));
    if (!b1 || !b2 || !out1 || !out2) return fail(ctx, SARPRO_ERR_INVALID_ARGUMENT, "NULL argument");
    RC(check_enums(ctx, -2, strategy, bit_depth));
    const sarpro_band* ins[2] = {b1, b2};
    const sarpro_band* ins2[2] = {nullptr, nullptr};
    const int ops[2] = {-1, -1};
    const int strategies[2] = {strategy, strategy}, depths[2] = {bit_depth, bit_depth};
    const PlanKind kinds[2] = {PlanKind::Autoscale, PlanKind::Autoscale};
    void* canvas[2];
    OutGeom g;
    RC(produce_bands(ctx, ins, ins2, ops, 2, strategies, depths, kinds, has_target != 0, target, pad != 0, canvas, &g, stats2));
    const size_t bytes = g.oc * g.orr * (bit_depth == SARPRO_U8 ? 1 : 2);
    fill_image(out1, g, 1, bit_depth);
    fill_image(out2, g, 1, bit_depth);
    RC(deliver(ctx, canvas[0], bytes, out1));
    RC(deliver(ctx, canvas[1], bytes, out2));
    if (bit_depth == SARPRO_U8) {
        ctx->last[1] = sarpro_ctx::LastResult{canvas[0], g.oc, g.orr};
        ctx->last[2] = sarpro_ctx::LastResult{canvas[1], g.oc, g.orr};
    }
    return end_call(ctx);
}

int sarpro_pipeline_synrgb(sarpro_ctx* ctx, const sarpro_band* b1, const sarpro_band* b2, int strategy, int mode,
                           int has_target, size_t target, int pad, int tamed_band_step, sarpro_image* out,
                           sarpro_stats* stats2) {
    (void)mode; // all four SyntheticRgbMode values alias Default (synthetic_rgb.rs:72-79)
    RC(begin_call(ctx));
    if (!b1 || !b2 || !out) return fail(ctx, SARPRO_ERR_INVALID_ARGUMENT, "NULL argument");
    RC(check_enums(ctx, -2, strategy, -2));
    const sarpro_band* ins[2] = {b1, b2};
    const sarpro_band* ins2[2] = {nullptr, nullptr};
    const int ops[2] = {-1, -1};
    const int strategies[2] = {strategy, strategy}, depths[2] = {SARPRO_U8, SARPRO_U8}; // save.rs:321
    PlanKind kinds[2] = {PlanKind::Autoscale, PlanKind::Autoscale};
    if (tamed_band_step && strategy == SARPRO_STRATEGY_TAMED) { // save.rs:324-328, 347-351
        kinds[0] = PlanKind::TamedSynRgbCopol;
        kinds[1] = PlanKind::TamedSynRgbCross;
    }
    void* canvas[2];
    OutGeom g;
    RC(produce_bands(ctx, ins, ins2, ops, 2, strategies, depths, kinds, has_target != 0, target, pad != 0, canvas, &g, stats2));
    const size_t n = g.oc * g.orr;
    RC(synrgb_compose(ctx, strategy, (const uint8_t*)canvas[0], (const uint8_t*)canvas[1], n));
    fill_image(out, g, 3, SARPRO_U8);
    RC(deliver(ctx, ctx->rgb.p, n * 3, out));
    ctx->last[0] = sarpro_ctx::LastResult{ctx->rgb.p, g.oc, g.orr};
    ctx->last[1] = sarpro_ctx::LastResult{canvas[0], g.oc, g.orr};
    ctx->last[2] = sarpro_ctx::LastResult{canvas[1], g.oc, g.orr};
    return end_call(ctx);
}

// ---- stage level ---------------------------------------------------------------------------------------
static int stage_pipeline(sarpro_ctx* ctx, const void* data, int dtype, size_t rows, size_t cols, int bit_depth,
                          int strategy, PlanKind kind, uint8_t* out_u8, uint16_t* out_u16, sarpro_stats* stats) {
    RC(begin_call(ctx));
    RC(check_enums(ctx, -2, strategy, bit_depth));
    sarpro_band band{data, dtype, SARPRO_LOC_HOST, rows, cols};
    const sarpro_band* ins[1] = {&band};
    const sarpro_band* ins2[1] = {nullptr};
    const int ops[1] = {-1};
    const int strategies[1] = {strategy}, depths[1] = {bit_depth};
    const PlanKind kinds[1] = {kind};
    void* canvas[1];
    OutGeom g;
    RC(produce_bands(ctx, ins, ins2, ops, 1, strategies, depths, kinds, false, 0, false, canvas, &g, stats));
    const bool out8 = kind != PlanKind::Autoscale || bit_depth == SARPRO_U8;
    void* dst = out8 ? (void*)out_u8 : (void*)out_u16;
    const size_t bytes = rows * cols * (out8 ? 1 : 2);
    if (bytes) {
        if (!dst) return fail(ctx, SARPRO_ERR_INVALID_ARGUMENT, "output pointer for the requested bit depth is NULL");
        CU(cudaMemcpyAsync(dst, canvas[0], bytes, cudaMemcpyDeviceToHost, ctx->stream));
        ctx->timing.d2h_bytes += bytes;
    }
    return end_call(ctx);
}

int sarpro_process_scalar_data_pipeline(sarpro_ctx* ctx, const float* v, size_t rows, size_t cols, int bit_depth,
                                        int strategy, uint8_t* out_u8, uint16_t* out_u16, sarpro_stats* stats) {
    return stage_pipeline(ctx, v, SARPRO_DT_F32, rows, cols, bit_depth, strategy, PlanKind::Autoscale, out_u8, out_u16, stats);
}
int sarpro_process_dn_pipeline(sarpro_ctx* ctx, const uint16_t* dn, size_t rows, size_t cols, int bit_depth, int strategy,
                               uint8_t* out_u8, uint16_t* out_u16, sarpro_stats* stats) {
    return stage_pipeline(ctx, dn, SARPRO_DT_U16, rows, cols, bit_depth, strategy, PlanKind::Autoscale, out_u8, out_u16, stats);
}
int sarpro_autoscale_tamed_synrgb_u8(sarpro_ctx* ctx, const float* v, size_t rows, size_t cols, int is_copol, uint8_t* out) {
    return stage_pipeline(ctx, v, SARPRO_DT_F32, rows, cols, SARPRO_U8, SARPRO_STRATEGY_TAMED,
                          is_copol ? PlanKind::TamedSynRgbCopol : PlanKind::TamedSynRgbCross, out, nullptr, nullptr);
}

int sarpro_scale_u16_to_u8(sarpro_ctx* ctx, const uint16_t* data, size_t n, uint8_t* out) {
    RC(begin_call(ctx));
    if (n == 0) return end_call(ctx);
    if (!data || !out) return fail(ctx, SARPRO_ERR_INVALID_ARGUMENT, "NULL argument");
    BandWs& w = ctx->band[0];
    RC(reserve(ctx, w.dn, n * 2));
    RC(reserve(ctx, w.small, n));
    RC(reserve(ctx, w.scalars, 64));
    CU(cudaMemcpyAsync(w.dn.p, data, n * 2, cudaMemcpyHostToDevice, ctx->stream));
    RC(reset_minmax(ctx, 0));
    KL(launch_minmax_u16((const uint16_t*)w.dn.p, n, (uint32_t*)w.scalars.p, ctx->sm_count, ctx->stream));
    KL(launch_scale_u16_to_u8((const uint16_t*)w.dn.p, n, (const uint32_t*)w.scalars.p, (uint8_t*)w.small.p, ctx->sm_count, ctx->stream));
    CU(cudaMemcpyAsync(out, w.small.p, n, cudaMemcpyDeviceToHost, ctx->stream));
    return end_call(ctx);
}

int sarpro_pol_op(sarpro_ctx* ctx, int op, const float* a, const float* b, size_t rows, size_t cols, float* out) {
    RC(begin_call(ctx));
    const size_t n = rows * cols;
    if (n == 0) return end_call(ctx);
    if (!a || !b || !out) return fail(ctx, SARPRO_ERR_INVALID_ARGUMENT, "NULL argument");
    if (op < 0) return fail(ctx, SARPRO_ERR_INVALID_ARGUMENT, "unknown polarization operation %d", op);
    RC(check_enums(ctx, op, -2, -2));
    BandWs& w = ctx->band[0];
    RC(reserve(ctx, w.f32a, n * 4));
    RC(reserve(ctx, w.f32b, n * 4));
    RC(reserve(ctx, w.full, n * 4));
    CU(cudaMemcpyAsync(w.f32a.p, a, n * 4, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(w.f32b.p, b, n * 4, cudaMemcpyHostToDevice, ctx->stream));
    KS(SARPRO_STAGE_CONVERT, launch_pol_op((const float*)w.f32a.p, (const float*)w.f32b.p, op, n, (float*)w.full.p, ctx->sm_count, ctx->stream));
    CU(cudaMemcpyAsync(out, w.full.p, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    return end_call(ctx);
}

int sarpro_add_padding_to_square(sarpro_ctx* ctx, const uint8_t* u8_data, const uint16_t* u16_data, size_t cols,
                                 size_t rows, int bit_depth, uint8_t* out_u8, uint16_t* out_u16) {
    RC(begin_call(ctx));
    RC(check_enums(ctx, -2, -2, bit_depth));
    const void* src = bit_depth == SARPRO_U8 ? (const void*)u8_data : (const void*)u16_data;
    void* dst = bit_depth == SARPRO_U8 ? (void*)out_u8 : (void*)out_u16;
    if (bit_depth == SARPRO_U16 && !u16_data) return fail(ctx, SARPRO_ERR_U16_REQUIRED, "U16 data required for U16 bit depth");
    const size_t esz = bit_depth == SARPRO_U8 ? 1 : 2;
    const size_t m = std::max(cols, rows), n = m * m;
    if (n == 0) return end_call(ctx);
    if (!src || !dst) return fail(ctx, SARPRO_ERR_INVALID_ARGUMENT, "NULL argument");
    BandWs& w = ctx->band[0];
    RC(reserve(ctx, w.full, std::max<size_t>(rows * cols * esz, 16)));
    RC(reserve(ctx, w.small, n * esz));
    CU(cudaMemcpyAsync(w.full.p, src, rows * cols * esz, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemsetAsync(w.small.p, 0, n * esz, ctx->stream));
    const size_t pl = (m - cols) / 2, pt = (m - rows) / 2; // padding.rs:13-14
    if (rows && cols)
        CU(cudaMemcpy2DAsync((unsigned char*)w.small.p + (pt * m + pl) * esz, m * esz, w.full.p, cols * esz, cols * esz, rows,
                             cudaMemcpyDeviceToDevice, ctx->stream));
    CU(cudaMemcpyAsync(dst, w.small.p, n * esz, cudaMemcpyDeviceToHost, ctx->stream));
    return end_call(ctx);
}

int sarpro_resize_image_data_with_meta(sarpro_ctx* ctx, const uint8_t* u8_data, const uint16_t* u16_data, size_t cols,
                                       size_t rows, int has_target, size_t target, int bit_depth, int pad,
                                       uint8_t* out_u8, uint16_t* out_u16, sarpro_resize_meta* meta) {
    RC(begin_call(ctx));
    RC(check_enums(ctx, -2, -2, bit_depth));
    const int pix16 = bit_depth == SARPRO_U16;
    const size_t esz = pix16 ? 2 : 1;
    const OutGeom g = out_geometry(cols, rows, has_target != 0, target, pad != 0);
    if (meta) *meta = g.meta;
    const void* src = pix16 ? (const void*)u16_data : (const void*)u8_data;
    void* dst = pix16 ? (void*)out_u16 : (void*)out_u8;
    if (pix16 && !u16_data) return fail(ctx, SARPRO_ERR_U16_REQUIRED, "U16 data required for U16 bit depth");
    const size_t n_out = g.oc * g.orr;
    if (n_out == 0) return end_call(ctx);
    if (!dst || (rows != 0 && cols != 0 && !src)) return fail(ctx, SARPRO_ERR_INVALID_ARGUMENT, "NULL argument");
    if (rows >= (1ull << 31) || cols >= (1ull << 31)) return fail(ctx, SARPRO_ERR_TOO_LARGE, "raster too large");
    BandWs& w = ctx->band[0];
    RC(reserve(ctx, w.full, std::max<size_t>(rows * cols * esz, 16)));
    RC(reserve(ctx, w.small, n_out * esz));
    CU(cudaMemcpyAsync(w.full.p, src, rows * cols * esz, cudaMemcpyHostToDevice, ctx->stream));
    if (g.pad) CU(cudaMemsetAsync(w.small.p, 0, n_out * esz, ctx->stream));
    unsigned char* region = (unsigned char*)w.small.p + (g.pad_top * g.oc + g.pad_left) * esz;
    if (!g.resize) {
        if (rows && cols)
            CU(cudaMemcpy2DAsync(region, g.oc * esz, w.full.p, cols * esz, cols * esz, rows, cudaMemcpyDeviceToDevice, ctx->stream));
    } else if (g.rc && g.rr) {
        AxisPlan *ah, *av;
        RC(get_axis(ctx, (uint32_t)cols, (uint32_t)g.rc, pix16, true, HSRC_IMAGE, &ah));
        RC(get_axis(ctx, (uint32_t)rows, (uint32_t)g.rr, pix16, false, 0, &av));
        RC(reserve(ctx, w.temp, rows * g.rc * esz));
        HResizeArgs a{};
        a.src = w.full.p;
        a.src_rows = (uint32_t)rows;
        a.src_cols = (uint32_t)cols;
        a.row0 = 0;
        a.n_rows = (uint32_t)rows;
        a.temp = w.temp.p;
        a.ax = ah->dev();
        RC(run_hpass_generic(ctx, a, HSRC_IMAGE, pix16, ah));
        KS(SARPRO_STAGE_VRESIZE, launch_vresize(w.temp.p, 0, (uint32_t)g.rc, av->dev(), 0, (uint32_t)g.rr, region, (uint32_t)g.oc, 0, pix16, ctx->stream));
    }
    CU(cudaMemcpyAsync(dst, w.small.p, n_out * esz, cudaMemcpyDeviceToHost, ctx->stream));
    return end_call(ctx);
}

int sarpro_create_synthetic_rgb_by_mode_and_strategy(sarpro_ctx* ctx, int mode, int strategy, const uint8_t* band1,
                                                     const uint8_t* band2, size_t n, uint8_t* rgb) {
    (void)mode;
    RC(begin_call(ctx));
    RC(check_enums(ctx, -2, strategy, -2));
    if (n == 0) return end_call(ctx);
    if (!band1 || !band2 || !rgb) return fail(ctx, SARPRO_ERR_INVALID_ARGUMENT, "NULL argument");
    RC(reserve(ctx, ctx->band[0].small, n));
    RC(reserve(ctx, ctx->band[1].small, n));
    RC(reserve(ctx, ctx->rgb, n * 3));
    CU(cudaMemcpyAsync(ctx->band[0].small.p, band1, n, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(ctx->band[1].small.p, band2, n, cudaMemcpyHostToDevice, ctx->stream));
    const uint8_t* d1 = (const uint8_t*)ctx->band[0].small.p;
    const uint8_t* d2 = (const uint8_t*)ctx->band[1].small.p;
    if (strategy == SARPRO_STRATEGY_TAMED || strategy == SARPRO_STRATEGY_CLAHE) {
        RC(reserve(ctx, ctx->hist256, 256 * 4));
        RC(reserve(ctx, ctx->rgbsel, 16));
        CU(cudaMemsetAsync(ctx->hist256.p, 0, 256 * 4, ctx->stream));
        KS(SARPRO_STAGE_RGB, launch_hist256_pair(d1, d2, n, (uint32_t*)ctx->hist256.p, ctx->stream));
        KS(SARPRO_STAGE_RGB, launch_synrgb_floor((const uint32_t*)ctx->hist256.p, n, (uint32_t*)ctx->rgbsel.p, ctx->stream));
        KS(SARPRO_STAGE_RGB, launch_synrgb(d1, d2, n, (const uint8_t*)ctx->rgb_luts.p, (const uint32_t*)ctx->rgbsel.p, 0, 1, (uint8_t*)ctx->rgb.p, ctx->stream));
    } else {
        KS(SARPRO_STAGE_RGB, launch_synrgb(d1, d2, n, (const uint8_t*)ctx->rgb_luts.p, nullptr, kSynRgbDefaultSet, 0, (uint8_t*)ctx->rgb.p, ctx->stream));
    }
    CU(cudaMemcpyAsync(rgb, ctx->rgb.p, n * 3, cudaMemcpyDeviceToHost, ctx->stream));
    return end_call(ctx);
}

} // extern "C"
