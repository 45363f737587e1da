// ctx.h — internal definition of sarpro_ctx and helpers shared by api.cu / api_f32.cu / comm.cu.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <map>
#include <string>
#include <tuple>
#include <vector>

#include "../../include/sarpro_gpu.h"
#include "kernels.h"
#include "plan.h"

namespace sarpro {

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
};

void release(DevBuf& b);

// One Lanczos axis (plan.cpp build_lanczos3_axis) with its device tables. Owned by the context's cache; plans are only
// dropped between calls (begin_call), never while a call may still hold a pointer to one.
struct AxisPlan {
    ResampleAxis h; // host copy
    DevBuf start, size, coef, packed, strips;
    uint32_t pairs = 0, oxb = 0, rbw = 0, smem = 0, n_strips = 0;
    bool has_strips = false;
    // tensor-core kernel (kernels_hmma.cu)
    bool mma = false;
    DevBuf m_btab, m_ntile, m_strips;
    std::vector<HStrip> m_weights_h;
    uint32_t m_b_bytes = 0;
    uint64_t id = 0; // unique per plan: cache key of the piece lists (a recycled address must not match)
    AxisPlan() = default;
    AxisPlan(const AxisPlan&) = delete;
    AxisPlan& operator=(const AxisPlan&) = delete;
    ~AxisPlan() {
        for (DevBuf* b : {&start, &size, &coef, &packed, &strips, &m_btab, &m_ntile, &m_strips}) release(*b);
    }
    AxisDev dev() const {
        AxisDev d;
        d.start = (const uint32_t*)start.p;
        d.size = (const uint32_t*)size.p;
        d.coef = (const int32_t*)coef.p;
        d.packed = (const uint32_t*)packed.p;
        d.window = h.window;
        d.pairs = pairs;
        d.out_size = h.out_size;
        d.in_size = h.in_size;
        d.precision = h.precision;
        return d;
    }
};

struct AxisKey {
    uint32_t in, out;
    int wide, horiz, src_kind;
    uint32_t strip_nt; // n-tiles per strip of the tensor-core plan (horizontal axes only)
    uint32_t hmma_in;  // width the tensor-core plan walks (the pitch of a re-pitched raster; = in otherwise)
    bool operator<(const AxisKey& o) const {
        return std::tie(in, out, wide, horiz, src_kind, strip_nt, hmma_in) <
               std::tie(o.in, o.out, o.wide, o.horiz, o.src_kind, o.strip_nt, o.hmma_in);
    }
};

struct BandWs {
    DevBuf dn;         // uploaded / converted raster
    DevBuf dn_pad;     // the raster re-pitched to a multiple of 8 columns (resized u8 outputs of rasters of other widths)
    uint64_t pitch = 0; // row pitch of dn_pad in samples while the call uses it, else 0
    DevBuf f32a, f32b; // f32 staging
    DevBuf tile_hist, total, lut, tile256, cdf, cdf32, remap;
    DevBuf temp, small, full;
    DevBuf scalars; // [0..1] minmax, [2] max_dn, [3] flag, [4] pass-A work-unit counter, [5] present-list allocator
    DevBuf present; // k_hist_total: 256 {offset, count} + kPresentCap {dn, count} (see plan_from_present_list)
    DevBuf edges, hist4096, f32scan; // general f32 path
    BandPlan plan;
    int hist_auto = 21;            // pass-A table shape for the next call (see choose_hist_variant)
    bool hist_auto_pending = false; // h_hist of this slot still has to go through choose_hist_variant
    uint32_t hot = 0, hot_top = 0; // table range / saturated table word for kernels_hmma.cu (0 = not eligible)
    // Piece lists of the persistent pass-B kernel for this slot (cached per geometry). Per slot, because the two bands of a
    // pair run their pass B on different streams and may be cut differently: a shared list could be overwritten while the
    // other band's kernel still reads it.
    DevBuf plan_scratch;            // device planner: list of present DNs when it does not fit shared memory
    DevBuf plan_dev;                // PlanDev: the band's plan for the kernels downstream (device or host planner)
    bool dev_planned = false;       // planned by kernels_plan.cu in this call: w.plan / hot are only valid after end_call
    bool plan_copy_pending = false; // ctx->h_plan[slot] is being written by the device
    DevBuf pieces, cta_first;
    uint64_t pc_rows = 0, pc_row_off = 0, pc_tile_h = 0, pc_axis_id = 0;
    int pc_clahe = -1;
    uint32_t pc_n_ctas = 0, pc_unit = 0;
    int pc_spare = -1;
};

constexpr uint32_t kSynRgbSets = 42; // 0..40 suppressed by floor_with_cushion, 41 default
constexpr uint32_t kSynRgbDefaultSet = 41;

struct OutGeom {
    bool resize = false, pad = false;
    size_t rc = 0, rr = 0;  // resized dims
    size_t oc = 0, orr = 0; // output dims (after pad)
    size_t pad_left = 0, pad_top = 0;
    sarpro_resize_meta meta{};
};
OutGeom out_geometry(size_t cols, size_t rows, bool has_target, size_t target, bool pad);

struct CommState; // comm.cu
struct JpegState; // api_jpeg.cu
void jpeg_state_destroy(sarpro_ctx* ctx);

// One band through the DN passes
struct BandJob {
    const uint16_t* dn = nullptr; // device
    uint64_t rows = 0, cols = 0;
    int strategy = 0, bit_depth = 0;
    PlanKind kind = PlanKind::Autoscale;
};
// Row-band sharding of a scene: the local raster holds scene rows [row_off, row_off + rows); the rank owns
// (counts / produces statistics for) local rows [own0, own1); the rest is Lanczos halo.
struct ShardGeom {
    uint64_t scene_rows = 0, row_off = 0, own0 = 0, own1 = 0;
};
inline bool uses_clahe(const BandJob& j) { return j.kind == PlanKind::Autoscale && j.strategy == SARPRO_STRATEGY_CLAHE; }

} // namespace sarpro

struct sarpro_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t stream2 = nullptr; // side stream: the second band's pass B (see produce_bands)
    cudaEvent_t ev_join = nullptr;
    int two_stream = 1;             // SARPRO_TWO_STREAM=0: everything on `stream`
    int spare_sms = 2;              // SARPRO_SPARE_SMS: SMs the persistent kernels of a band pair leave to the other band's small kernels
    int pair_spare = 0;             // spare_sms while a pipelined pair is in flight, else 0
    bool own_stream = true;
    int sm_count = 148;
    std::string err;
    sarpro::BandWs band[2];
    sarpro::DevBuf units, tile_px, col_dx, col_omdx, col_t, row_dy, row_omdy, row_t, rgb, hist256, rgbsel, rgb_luts;
    sarpro::DevBuf col_m, row_sat;
    uint64_t clahe_tile_w = 0, clahe_tile_h = 0, clahe_rows = 0;
    uint64_t next_axis_id = 1;
    int use_hmma = 1;    // SARPRO_HMMA=0: the generic exact kernel (kernels_resize.cu) instead of kernels_hmma.cu (validation)
    int force_exact = 0; // SARPRO_FORCE_EXACT=1: generic kernels + exact f64 CLAHE everywhere (validation)
    // geometry caches
    uint64_t units_pitch = 0; // columns the CLAHE per-column tables were built for (the pitch of a re-pitched raster)
    uint64_t units_rows = 0, units_cols = 0, units_scene_rows = 0, units_row_off = 0, units_own0 = 0, units_own1 = 0;
    int units_clahe = -1;
    uint32_t n_units = 0, n_tiles = 0;
    std::map<sarpro::AxisKey, sarpro::AxisPlan*> axes;
    // pinned staging
    uint32_t* h_hist = nullptr;    // [2][65536]
    uint32_t* h_present = nullptr; // [2][2 * (256 + kPresentCap)]
    uint16_t* h_lut = nullptr;     // [2][65536]
    uint32_t* h_scalars = nullptr; // [2][8]
    uint8_t* h_remap = nullptr;    // [2][256]
    sarpro::PlanDev* h_plan = nullptr;    // [2] device -> host mirror of the plans (read in end_call)
    sarpro::PlanDev* h_plan_up = nullptr; // [2] staging of host-planned bands' plans on their way to the device
    sarpro_stats* pending_stats[2] = {nullptr, nullptr}; // caller's stats structs to fill in end_call (device-planned bands)
    sarpro::DevBuf db_table;       // [65536] f64: dB of every DN (device planner)
    bool shard_reduce = false;     // inside a sharded general-path call: merge scan / stat histogram over the ranks
    uint64_t shard_scene_px = 0;   // pixels of the whole scene in that call
    int host_plan = 0;             // SARPRO_HOST_PLAN=1: plan every band on the host (validation)
    int f32_no_guard = 0;          // SARPRO_F32_NO_GUARD=1: general f32 path compares thresholds for every sample (validation)
    // timing
    cudaEvent_t ev[6] = {};
    sarpro_timing timing{};
    // per-stage event pairs recorded during a call, resolved in end_call
    static constexpr int kMaxStageEvents = 64;
    cudaEvent_t sev[2 * kMaxStageEvents] = {};
    int sev_stage[kMaxStageEvents] = {};
    double sev_host[kMaxStageEvents] = {}; // host time of the launch (ms since begin_call; SARPRO_TRACE)
    double host_t0 = 0;
    int n_sev = 0;
    // stages that get CUDA event pairs (each pair costs ~3 us of host time per call): pass A, pass B and the collectives by
    // default; SARPRO_STAGE_TIMING=all (or SARPRO_TRACE) times every launch
    uint32_t stage_mask = (1u << SARPRO_STAGE_HIST) | (1u << SARPRO_STAGE_APPLY) | (1u << SARPRO_STAGE_COMM);
    int hist_variant = -1; // SARPRO_HIST_VARIANT; -1 = per band, from the tail of the previous histogram of that slot
    float valid_thresh = 0.f;
    sarpro::CommState* comm = nullptr;
    // batch / streamed uploads (api_batch.cu): a copy stream and two staging slots of a band pair, created on first use
    cudaStream_t stream_up = nullptr;
    cudaEvent_t ev_up[2] = {nullptr, nullptr};  // upload of staging slot s complete
    cudaEvent_t ev_chunk[16] = {};              // streamed upload: chunk c of band b has landed (index b * kUploadChunks + c)
    sarpro::DevBuf batch_stage[2][2];           // [slot][band]
    // streamed upload of a host u16 band: row chunks go out on the copy stream, pass A runs on each chunk as it
    // lands (units ordered by their last row), and the first band's plan / pass B overlap the second band's upload
    static constexpr int kUploadChunks = 8;
    struct StreamedBand { int n_chunks = 0, ev0 = 0; uint32_t row_end[kUploadChunks] = {}; }; // ev0: first of its events in ev_chunk
    StreamedBand streamed[2];
    bool upload_in_flight = false;              // chunks were queued in a call that has not completed (error path): drain first
    int repitch = 1;                            // SARPRO_REPITCH=0: rasters of odd widths stay on the generic kernels (measurement)
    int stream_upload = 1;                      // SARPRO_STREAM_UPLOAD=0: one copy on the main stream (measurement)
    // Host-side narrowing of large f32 host rasters (narrow.cpp): two pinned staging slots of one row chunk of DNs each; the
    // host threads fill one while the other is on the wire. ev_ring[s]: the copy out of slot s has completed.
    int narrow_upload = 1;                      // SARPRO_NARROW_UPLOAD=0: f32 rasters are uploaded as f32 and narrowed by k_f32_to_dn;
                                                // =2: always narrow; 1 (default): until the host proves slower than the f32 upload
    double narrow_gbs = 0.0;                    // f32 source bytes the host threads narrowed per second, last band (GB/s)
    void* narrow_ring[2] = {nullptr, nullptr};
    size_t narrow_ring_bytes = 0;
    cudaEvent_t ev_ring[2] = {nullptr, nullptr};
    bool ring_busy[2] = {false, false};
    uint64_t narrowed_bands = 0;                // bands that took the narrowing path (tests)
    sarpro::DevBuf units_by_row;                // the work units of pass A ordered by last row
    std::vector<uint32_t> units_r1;             // their last rows (ascending), host copy
    // u8 results of the last pipeline call that are still in the context's device buffers: [0] interleaved RGB, [1] / [2] the
    // gray bands (sarpro_encode_last_jpeg encodes them where they lie). Cleared when a call that does not produce them begins.
    struct LastResult { const void* dev = nullptr; size_t cols = 0, rows = 0; };
    LastResult last[3];
    bool keep_last = false; // set around calls that read `last` (begin_call clears it otherwise)
    sarpro::JpegState* jpeg = nullptr;
    sarpro::DevBuf gather;                      // sharded scene: all-gathered output rows + extrema of every rank (comm.cu)
};

namespace sarpro {

int fail(sarpro_ctx* c, int code, const char* fmt, ...);
double host_ms();
int reserve(sarpro_ctx* ctx, DevBuf& b, size_t bytes);
uint32_t hmma_hot(const uint16_t* lut_host, const uint32_t* hist_host, uint32_t max_present_dn, uint32_t* top_out);
uint32_t hmma_hot_from_plan(const BandPlan& plan, uint32_t* top_out);
// strip_nt: n-tiles (8 output columns) per strip of the tensor-core pass B; shorter strips = finer work units for short rasters
// hmma_in: width the tensor-core plan walks when the raster was re-pitched to a multiple of 8 columns (0 = in)
int get_axis(sarpro_ctx* ctx, uint32_t in, uint32_t out, bool wide, bool horiz, int src_kind, AxisPlan** res, uint32_t strip_nt = 32,
             uint32_t hmma_in = 0);
uint32_t choose_strip_nt(const sarpro_ctx* ctx, uint64_t rows, uint64_t out_cols, bool clahe);
int begin_call(sarpro_ctx* ctx);
int ensure_upload_stream(sarpro_ctx* ctx); // api_batch.cu
// slot: the band slot whose piece lists the tensor-core kernel uses; *gate: see api.cu
int run_hpass(sarpro_ctx* ctx, int slot, const HResizeArgs& a, int src_kind, int pix16, AxisPlan* ah, uint64_t row_off, bool* gate);
int run_hpass_generic(sarpro_ctx* ctx, const HResizeArgs& a, int src_kind, int pix16, AxisPlan* ah);
// phase: 0 = everything, 1 = only the preamble (workspaces, cleared histograms and counters), 2 = only the kernels
int dn_pass_a_launch(sarpro_ctx* ctx, int b, const uint16_t* dn, uint64_t rows, uint64_t cols, bool clahe_units, int phase = 0);
int dn_pass_a_launch_sharded(sarpro_ctx* ctx, int b, const uint16_t* dn, uint64_t rows, uint64_t cols, bool clahe_units,
                             const ShardGeom& sg, int phase = 0);
int dn_run_pass_b(sarpro_ctx* ctx, int b, const BandJob& j, const OutGeom& g, void* canvas);
int run_clahe_stats(sarpro_ctx* ctx, int b);
int clahe_minmax(sarpro_ctx* ctx, int b, uint32_t* mn, uint32_t* mx);
int upload_remap(sarpro_ctx* ctx, int b, uint32_t mn, uint32_t mx);
ClaheDev clahe_dev(sarpro_ctx* ctx, int b);
int deliver(sarpro_ctx* ctx, const void* dev_src, size_t bytes, sarpro_image* out);
void fill_image(sarpro_image* out, const OutGeom& g, int channels, int bit_depth);
int check_band(sarpro_ctx* ctx, const sarpro_band* b);
int check_enums(sarpro_ctx* ctx, int op, int strategy, int bit_depth); // -2 = not applicable
int synrgb_compose(sarpro_ctx* ctx, int strategy, const uint8_t* c1, const uint8_t* c2, size_t n);
int dn_band_with_preset_lut(sarpro_ctx* ctx, int b, const BandJob& j, const uint16_t* lut_host, uint32_t max_key,
                            const OutGeom& g, void* canvas);
int comm_reduce_f32_scan(sarpro_ctx* ctx, F32Scan* scan_dev, int nops);                                    // comm.cu
int comm_reduce_f32_hist(sarpro_ctx* ctx, unsigned long long* hist4096_dev, double* sums_dev, int nops);   // comm.cu
int f32_general(sarpro_ctx* ctx, int nops, const int* slots, const void* a_dev, const void* b_dev, int a_u16, int b_u16, const int* ops,
                uint64_t rows, uint64_t cols, int bit_depth, int strategy, PlanKind kind, const OutGeom& g, void* const* canvases,
                sarpro_stats* stats);
int f32_general_single(sarpro_ctx* ctx, int slot, const void* a_dev, const void* b_dev, int a_u16, int b_u16, int op, uint64_t rows,
                       uint64_t cols, int bit_depth, int strategy, PlanKind kind, const OutGeom& g, void* canvas_dev,
                       sarpro_stats* stats);
int end_call(sarpro_ctx* ctx);
struct BandJob;
bool plans_on_device(const sarpro_ctx* ctx, const BandJob& job);
int plan_band_on_device(sarpro_ctx* ctx, int b, const BandJob& job);
int plan_bands_on_device(sarpro_ctx* ctx, const int* slots, const BandJob* jobs, int nb);
int run_clahe_stats_bands(sarpro_ctx* ctx, const int* slots, int nb, int (*all_reduce)(sarpro_ctx*, void*), void* arg);
int upload_plan_dev(sarpro_ctx* ctx, int b);

#define CU(call)                                                                                             \
    do {                                                                                                     \
        cudaError_t e__ = (call);                                                                            \
        if (e__ != cudaSuccess)                                                                              \
            return sarpro::fail(ctx, e__ == cudaErrorMemoryAllocation ? SARPRO_ERR_OUT_OF_MEMORY : SARPRO_ERR_CUDA, \
                                "CUDA error %s at %s:%d (%s)", cudaGetErrorName(e__), __FILE__, __LINE__,    \
                                cudaGetErrorString(e__));                                                    \
    } while (0)
#define KL(call)                       \
    do {                               \
        CU(call);                      \
        ctx->timing.kernel_launches++; \
    } while (0)
// stage-timed kernel launch: CUDA events around the launch, attributed to sarpro_stage S
#define KS(S, call)                                                                  \
    do {                                                                             \
        const int si__ = (((ctx->stage_mask >> (S)) & 1u) && ctx->n_sev < sarpro_ctx::kMaxStageEvents) ? ctx->n_sev : -1; \
        if (si__ >= 0) CU(cudaEventRecord(ctx->sev[2 * si__], ctx->stream));         \
        KL(call);                                                                    \
        if (si__ >= 0) {                                                             \
            CU(cudaEventRecord(ctx->sev[2 * si__ + 1], ctx->stream));                \
            ctx->sev_stage[si__] = (S);                                              \
            ctx->sev_host[si__] = sarpro::host_ms() - ctx->host_t0;                  \
            ctx->n_sev++;                                                            \
        }                                                                            \
    } while (0)
#define RC(call)               \
    do {                       \
        int rc__ = (call);     \
        if (rc__) return rc__; \
    } while (0)

} // namespace sarpro
