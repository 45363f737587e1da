// comm.cu — multi-GPU row-band sharding of ONE scene (one process per GPU), SURVEY.md §8e.
//
// Each rank holds scene rows [h0,h1): its own band [r0,r1) plus the vertical Lanczos halo. All exchanged
// quantities are integers, so the sharded result is bit-identical to the single-GPU result:
//   1. DN histogram            all-reduce(sum)  2 x 65,536 u32 (one group) -> every rank plans redundantly (on the device)
//   2. CLAHE tile histograms   all-reduce(sum)  2 x 64 x 256 u32 (one group) -> every rank builds all 64 CDFs
//   3. resized rows + CLAHE sample min/max   ONE all-gather of a slot per rank: its own output rows of both bands and its
//      extrema (16 bytes); the scale_u16_to_u8 decision is speculated (identity) and repaired in the rare other case
// NCCL is dlopen()ed (libnccl.so.2, the copy torch ships) so the library has no link-time dependency; the
// host only moves the 128-byte ncclUniqueId between ranks.
#include <dlfcn.h>

#include <algorithm>
#include <cstring>
#include <thread>
#include <vector>

#include "ctx.h"

using namespace sarpro;

namespace sarpro {

typedef struct { char internal[128]; } nccl_unique_id;
typedef void* nccl_comm_t;
enum { kNcclUint8 = 1, kNcclUint32 = 3, kNcclUint64 = 5, kNcclFloat64 = 8 };
enum { kNcclSum = 0, kNcclMax = 2, kNcclMin = 3 };

struct NcclApi {
    void* handle = nullptr;
    int (*GetUniqueId)(nccl_unique_id*) = nullptr;
    int (*CommInitRank)(nccl_comm_t*, int, nccl_unique_id, int) = nullptr;
    int (*CommInitRankConfig)(nccl_comm_t*, int, nccl_unique_id, int, void*) = nullptr; // optional (NCCL >= 2.14)
    int (*CommDestroy)(nccl_comm_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
    int (*Broadcast)(const void*, void*, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, nccl_comm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    std::string error;
    bool ok = false;
};

static NcclApi& nccl() {
    static NcclApi api;
    static bool tried = false;
    if (tried) return api;
    tried = true;
    std::vector<std::string> names;
    if (const char* e = getenv("SARPRO_NCCL_LIB")) names.push_back(e);
    names.push_back("libnccl.so.2");
    names.push_back("libnccl.so");
    for (const auto& n : names) {
        api.handle = dlopen(n.c_str(), RTLD_NOW | RTLD_GLOBAL);
        if (api.handle) break;
    }
    if (!api.handle) {
        api.error = std::string("cannot dlopen libnccl.so.2 (set SARPRO_NCCL_LIB): ") + (dlerror() ? dlerror() : "");
        return api;
    }
#define SARPRO_SYM(field, name)                                                     \
    *(void**)(&api.field) = dlsym(api.handle, name);                                \
    if (!api.field) { api.error = std::string("missing NCCL symbol ") + name; return api; }
    SARPRO_SYM(GetUniqueId, "ncclGetUniqueId")
    SARPRO_SYM(CommInitRank, "ncclCommInitRank")
    *(void**)(&api.CommInitRankConfig) = dlsym(api.handle, "ncclCommInitRankConfig");
    SARPRO_SYM(CommDestroy, "ncclCommDestroy")
    SARPRO_SYM(AllReduce, "ncclAllReduce")
    SARPRO_SYM(Broadcast, "ncclBroadcast")
    SARPRO_SYM(AllGather, "ncclAllGather")
    SARPRO_SYM(GroupStart, "ncclGroupStart")
    SARPRO_SYM(GroupEnd, "ncclGroupEnd")
    SARPRO_SYM(GetErrorString, "ncclGetErrorString")
#undef SARPRO_SYM
    api.ok = true;
    return api;
}

constexpr int kCommSpareSms = 16; // SMs the big kernels of a sharded call leave to the collectives and the small kernels

struct CommState {
    nccl_comm_t comm = nullptr;
    int rank = 0, world = 1;
};

// CUDA events around a collective, attributed to SARPRO_STAGE_COMM (like KS in ctx.h)
#define COMM_BEGIN()                                                                   \
    int comm_si__ = (((ctx->stage_mask >> SARPRO_STAGE_COMM) & 1u) && ctx->n_sev < sarpro_ctx::kMaxStageEvents) ? ctx->n_sev : -1; \
    if (comm_si__ >= 0) CU(cudaEventRecord(ctx->sev[2 * comm_si__], ctx->stream))
#define COMM_END()                                                                     \
    if (comm_si__ >= 0) {                                                              \
        CU(cudaEventRecord(ctx->sev[2 * comm_si__ + 1], ctx->stream));                 \
        ctx->sev_stage[comm_si__] = SARPRO_STAGE_COMM;                                 \
        ctx->n_sev++;                                                                  \
    }
#define NC(call)                                                                                        \
    do {                                                                                                \
        int r__ = (call);                                                                               \
        if (r__ != 0) return fail(ctx, SARPRO_ERR_COMM, "NCCL error %d (%s) at %s:%d", r__,             \
                                  nccl().GetErrorString ? nccl().GetErrorString(r__) : "?", __FILE__, __LINE__); \
    } while (0)

// output rows [oy0, oy1) produced by `rank`: those whose window centre falls in the rank's row band
static void owned_output_rows(size_t scene_rows, size_t out_rows, size_t r0, size_t r1, size_t* oy0, size_t* oy1) {
    auto first_at_or_after = [&](size_t r) -> size_t { // smallest oy with floor((oy + 0.5) * scene/out) >= r
        if (r == 0) return 0;
        if (r >= scene_rows) return out_rows;
        const double scale = (double)scene_rows / (double)out_rows;
        size_t oy = (size_t)std::max(0.0, std::floor((double)r / scale - 0.5));
        while (oy > 0 && (size_t)std::floor(((double)(oy - 1) + 0.5) * scale) >= r) --oy;
        while (oy < out_rows && (size_t)std::floor(((double)oy + 0.5) * scale) < r) ++oy;
        return oy;
    };
    *oy0 = first_at_or_after(r0);
    *oy1 = first_at_or_after(r1);
}

// av: the vertical Lanczos axis of the scene when the caller has it cached (else it is built here).
static int shard_geometry(size_t scene_rows, size_t cols, bool has_target, size_t target, bool pad, int world, int rank,
                          bool clahe, size_t* r0, size_t* r1, size_t* h0, size_t* h1, size_t* oy0, size_t* oy1,
                          OutGeom* g_out, const ResampleAxis* av_cached = nullptr) {
    int rc = sarpro_shard_rows(scene_rows, world, rank, clahe, r0, r1);
    if (rc) return rc;
    const OutGeom g = out_geometry(cols, scene_rows, has_target, target, pad);
    if (g_out) *g_out = g;
    *h0 = *r0;
    *h1 = *r1;
    *oy0 = *oy1 = 0;
    if (g.resize && g.rr > 0 && g.rc > 0) {
        owned_output_rows(scene_rows, g.rr, *r0, *r1, oy0, oy1);
        if (*oy1 > *oy0) {
            ResampleAxis local;
            if (!av_cached) build_lanczos3_axis((uint32_t)scene_rows, (uint32_t)g.rr, false, &local);
            const ResampleAxis& av = av_cached ? *av_cached : local;
            *h0 = std::min<size_t>(*h0, av.start[*oy0]);
            *h1 = std::max<size_t>(*h1, (size_t)av.start[*oy1 - 1] + av.size[*oy1 - 1]);
        }
    }
    return 0;
}

// General (f32 / polarization-op) path of a sharded scene: merged scan {min key, max key, valid count} of nops operations ...
int comm_reduce_f32_scan(sarpro_ctx* ctx, F32Scan* scan_dev, int nops) {
    if (!ctx->comm) return fail(ctx, SARPRO_ERR_COMM, "sarpro_comm_init has not been called on this context");
    NcclApi& api = nccl();
    CommState* cs = ctx->comm;
    COMM_BEGIN();
    NC(api.GroupStart());
    for (int o = 0; o < nops; ++o) {
        NC(api.AllReduce(&scan_dev[o].min_key, &scan_dev[o].min_key, 1, kNcclUint32, kNcclMin, cs->comm, ctx->stream));
        NC(api.AllReduce(&scan_dev[o].max_key, &scan_dev[o].max_key, 1, kNcclUint32, kNcclMax, cs->comm, ctx->stream));
        NC(api.AllReduce(&scan_dev[o].valid_count, &scan_dev[o].valid_count, 1, kNcclUint64, kNcclSum, cs->comm, ctx->stream));
    }
    NC(api.GroupEnd());
    COMM_END();
    return 0;
}
// ... and the merged 4096-bin stat histograms (exact integers; nops contiguous tables) with the log sums (mean / std: log lines only)
int comm_reduce_f32_hist(sarpro_ctx* ctx, unsigned long long* hist4096_dev, double* sums_dev, int nops) {
    if (!ctx->comm) return fail(ctx, SARPRO_ERR_COMM, "sarpro_comm_init has not been called on this context");
    NcclApi& api = nccl();
    CommState* cs = ctx->comm;
    COMM_BEGIN();
    NC(api.GroupStart());
    NC(api.AllReduce(hist4096_dev, hist4096_dev, (size_t)nops * kStatBins, kNcclUint64, kNcclSum, cs->comm, ctx->stream));
    NC(api.AllReduce(sums_dev, sums_dev, (size_t)nops * 2, kNcclFloat64, kNcclSum, cs->comm, ctx->stream));
    NC(api.GroupEnd());
    COMM_END();
    return 0;
}

} // namespace sarpro

extern "C" {

int sarpro_pipeline_polops(sarpro_ctx* ctx, const sarpro_band* a, const sarpro_band* b, size_t scene_rows, int n_ops, const int* ops,
                           int bit_depth, int strategy, sarpro_image* outs, sarpro_stats* stats) {
    RC(begin_call(ctx));
    if (!a || !b || !outs || !ops) return fail(ctx, SARPRO_ERR_INVALID_ARGUMENT, "NULL argument");
    if (n_ops < 1 || n_ops > 2) return fail(ctx, SARPRO_ERR_INVALID_ARGUMENT, "one or two polarization operations per call");
    for (int o = 0; o < n_ops; ++o) {
        RC(check_enums(ctx, ops[o], strategy, bit_depth));
        if (ops[o] < 0) return fail(ctx, SARPRO_ERR_INVALID_ARGUMENT, "sarpro_pipeline_polops takes polarization operations (ops.rs:4-44)");
    }
    RC(check_band(ctx, a));
    RC(check_band(ctx, b));
    if (a->rows != b->rows || a->cols != b->cols || a->dtype != b->dtype)
        return fail(ctx, SARPRO_ERR_INVALID_ARGUMENT, "bands differ in shape or type");
    const uint64_t rows = a->rows, cols = a->cols, n = rows * cols;
    const bool sharded = scene_rows != 0 && scene_rows != rows;
    if (sharded && !ctx->comm) return fail(ctx, SARPRO_ERR_COMM, "sarpro_comm_init has not been called on this context");
    if (sharded && scene_rows < rows) return fail(ctx, SARPRO_ERR_INVALID_ARGUMENT, "scene_rows is smaller than the rank's row band");
    if (strategy == SARPRO_STRATEGY_CLAHE && (sharded || n_ops == 2))
        return fail(ctx, SARPRO_ERR_INVALID_ARGUMENT, "CLAHE needs the tile statistics of one whole band: use sarpro_pipeline_single per operation");
    const size_t esz = bit_depth == SARPRO_U8 ? 1 : 2;
    const int is16 = a->dtype == SARPRO_DT_U16;
    const size_t isz = is16 ? 2 : 4;
    BandWs& w = ctx->band[0];
    const void* pa = a->data;
    const void* pb = b->data;
    if (a->location == SARPRO_LOC_HOST) {
        RC(reserve(ctx, w.f32a, std::max<size_t>(n * isz, 16)));
        CU(cudaMemcpyAsync(w.f32a.p, a->data, n * isz, cudaMemcpyHostToDevice, ctx->stream));
        ctx->timing.h2d_bytes += n * isz;
        pa = w.f32a.p;
    }
    if (b->location == SARPRO_LOC_HOST) {
        RC(reserve(ctx, w.f32b, std::max<size_t>(n * isz, 16)));
        CU(cudaMemcpyAsync(w.f32b.p, b->data, n * isz, cudaMemcpyHostToDevice, ctx->stream));
        ctx->timing.h2d_bytes += n * isz;
        pb = w.f32b.p;
    }
    const OutGeom g = out_geometry(cols, rows, false, 0, false);
    const int slots[2] = {0, 1};
    void* canvases[2] = {nullptr, nullptr};
    for (int o = 0; o < n_ops; ++o) {
        // device-resident outputs are written in place (no staging copy of a full-resolution band)
        if (outs[o].data && outs[o].location == SARPRO_LOC_DEVICE) {
            if (outs[o].capacity_bytes < n * esz) return fail(ctx, SARPRO_ERR_INVALID_ARGUMENT, "output buffer too small");
            canvases[o] = outs[o].data;
        } else {
            RC(reserve(ctx, ctx->band[o].small, std::max<size_t>(n * esz, 16)));
            canvases[o] = ctx->band[o].small.p;
        }
    }
    ctx->shard_reduce = sharded;
    ctx->shard_scene_px = (uint64_t)(sharded ? scene_rows : rows) * cols;
    const int rc = f32_general(ctx, n_ops, slots, pa, pb, is16, is16, ops, rows, cols, bit_depth, strategy, PlanKind::Autoscale, g, canvases, stats);
    ctx->shard_reduce = false;
    RC(rc);
    for (int o = 0; o < n_ops; ++o) {
        void* dst = outs[o].data;
        fill_image(&outs[o], g, 1, bit_depth);
        if (dst && canvases[o] != dst) RC(deliver(ctx, canvases[o], n * esz, &outs[o]));
    }
    return end_call(ctx);
}

int sarpro_pipeline_single_sharded(sarpro_ctx* ctx, const sarpro_band* a, const sarpro_band* b, size_t scene_rows, int op, int bit_depth,
                                   int strategy, sarpro_image* out, sarpro_stats* stats) {
    return sarpro_pipeline_polops(ctx, a, b, scene_rows, 1, &op, bit_depth, strategy, out, stats);
}

int sarpro_shard_rows(size_t rows, int world, int rank, int clahe, size_t* r0, size_t* r1) {
    if (!r0 || !r1 || world < 1 || rank < 0 || rank >= world) return SARPRO_ERR_INVALID_ARGUMENT;
    if (clahe) {
        // align band edges to the CLAHE tile height ceil(rows/8) (autoscale.rs:235)
        const size_t tile_h = (rows + kClaheTiles - 1) / kClaheTiles;
        const size_t n_tile_rows = tile_h ? (rows + tile_h - 1) / tile_h : 0;
        const size_t a = n_tile_rows * (size_t)rank / (size_t)world, b = n_tile_rows * (size_t)(rank + 1) / (size_t)world;
        *r0 = std::min(rows, a * tile_h);
        *r1 = std::min(rows, b * tile_h);
    } else {
        *r0 = rows * (size_t)rank / (size_t)world;
        *r1 = rows * (size_t)(rank + 1) / (size_t)world;
    }
    return SARPRO_OK;
}

int sarpro_shard_halo_rows(size_t rows, size_t cols, int has_target, size_t target, int world, int rank, int clahe,
                           size_t* h0, size_t* h1) {
    if (!h0 || !h1) return SARPRO_ERR_INVALID_ARGUMENT;
    size_t r0, r1, oy0, oy1;
    return shard_geometry(rows, cols, has_target != 0, target, false, world, rank, clahe != 0, &r0, &r1, h0, h1, &oy0, &oy1,
                          nullptr);
}

int sarpro_comm_unique_id(void* out128) {
    if (!out128) return SARPRO_ERR_INVALID_ARGUMENT;
    NcclApi& api = nccl();
    if (!api.ok) return SARPRO_ERR_COMM;
    nccl_unique_id id;
    if (api.GetUniqueId(&id) != 0) return SARPRO_ERR_COMM;
    std::memcpy(out128, &id, 128);
    return SARPRO_OK;
}

int sarpro_comm_init(sarpro_ctx* ctx, const void* unique_id128, int rank, int world) {
    if (!ctx || !unique_id128 || world < 1 || rank < 0 || rank >= world) return SARPRO_ERR_INVALID_ARGUMENT;
    NcclApi& api = nccl();
    if (!api.ok) return fail(ctx, SARPRO_ERR_COMM, "%s", api.error.c_str());
    CU(cudaSetDevice(ctx->device));
    sarpro_comm_destroy(ctx);
    nccl_unique_id id;
    std::memcpy(&id, unique_id128, 128);
    CommState* cs = new CommState();
    cs->rank = rank;
    cs->world = world;
    // The collectives of a sharded scene are small and run BESIDE the persistent histogram / pass-B kernels, on the SMs those
    // leave free (kCommSpareSms = 16: with 4 a 256 KB all-reduce waited for the big kernel to drain, 0.09 ms instead of 0.03 ms).
    // SARPRO_NCCL_MAX_CTAS caps the communicator's CTAs instead (measurement: a cap of 8 made the early all-reduces fast on 8
    // spare SMs but tripled the final all-gather, 0.114 ms instead of 0.042 ms on 4 GPUs, so it is off by default).
    struct NcclConfigV22700 { // ncclConfig_t as of NCCL 2.27 (newer libraries accept older layouts by size / version)
        size_t size; unsigned int magic, version;
        int blocking, cgaClusterSize, minCTAs, maxCTAs; const char* netName; int splitShare, trafficClass; const char* commName;
        int collnetEnable, CTAPolicy, shrinkShare, nvlsCTAs;
    };
    constexpr int kUndef = (int)0x80000000; // NCCL_CONFIG_UNDEF_INT
    const int max_ctas = getenv("SARPRO_NCCL_MAX_CTAS") ? atoi(getenv("SARPRO_NCCL_MAX_CTAS")) : 0;
    NcclConfigV22700 cfg = {sizeof(NcclConfigV22700), 0xcafebeefu, 22700u, kUndef, kUndef, 1, max_ctas, nullptr, kUndef, kUndef, nullptr,
                            kUndef, kUndef, kUndef, kUndef};
    int r = -1;
    if (api.CommInitRankConfig && max_ctas > 0) r = api.CommInitRankConfig(&cs->comm, world, id, rank, &cfg);
    if (r != 0) r = api.CommInitRank(&cs->comm, world, id, rank);
    if (r != 0) {
        delete cs;
        return fail(ctx, SARPRO_ERR_COMM, "ncclCommInitRank failed: %s", api.GetErrorString(r));
    }
    ctx->comm = cs;
    return SARPRO_OK;
}

int sarpro_comm_destroy(sarpro_ctx* ctx) {
    if (!ctx) return SARPRO_ERR_INVALID_ARGUMENT;
    if (ctx->comm) {
        if (ctx->comm->comm && nccl().ok) nccl().CommDestroy(ctx->comm->comm);
        delete ctx->comm;
        ctx->comm = nullptr;
    }
    return SARPRO_OK;
}

int sarpro_pipeline_synrgb_sharded(sarpro_ctx* ctx, const sarpro_band* b1, const sarpro_band* b2, size_t scene_rows,
                                   int strategy, int mode, int has_target, size_t target, int pad, int tamed_band_step,
                                   sarpro_image* out) {
    (void)mode;
    RC(begin_call(ctx));
    if (!b1 || !b2) return fail(ctx, SARPRO_ERR_INVALID_ARGUMENT, "NULL argument");
    RC(check_enums(ctx, -2, strategy, -2));
    if (!ctx->comm) return fail(ctx, SARPRO_ERR_COMM, "sarpro_comm_init has not been called on this context");
    NcclApi& api = nccl();
    CommState* cs = ctx->comm;
    const sarpro_band* ins[2] = {b1, b2};
    for (int b = 0; b < 2; ++b) {
        RC(check_band(ctx, ins[b]));
        if (ins[b]->dtype != SARPRO_DT_U16) return fail(ctx, SARPRO_ERR_INVALID_ARGUMENT, "the sharded pipeline takes u16 DN bands");
        if (ins[b]->rows != b1->rows || ins[b]->cols != b1->cols) return fail(ctx, SARPRO_ERR_INVALID_ARGUMENT, "bands differ in shape");
    }
    const size_t cols = b1->cols;
    if (!has_target || std::max(cols, scene_rows) == target)
        return fail(ctx, SARPRO_ERR_INVALID_ARGUMENT, "the sharded pipeline needs a resize target below the long side");
    const bool clahe = strategy == SARPRO_STRATEGY_CLAHE;
    size_t r0, r1, h0, h1, oy0, oy1;
    OutGeom g = out_geometry(cols, scene_rows, true, target, pad != 0);
    AxisPlan *ah = nullptr, *av = nullptr;
    const int src_kind = clahe ? HSRC_DN_CLAHE : HSRC_DN_LUT;
    if (g.rc == 0 || g.rr == 0) return fail(ctx, SARPRO_ERR_INVALID_ARGUMENT, "resize target yields an empty image");
    // (the strip length follows the rows this rank holds, b1->rows: see choose_strip_nt)
    const uint32_t hmma_in = ((cols % 8) != 0 && cols >= 512 && ctx->repitch && !ctx->force_exact) ? (uint32_t)((cols + 7) & ~(size_t)7) : 0u;
    RC(get_axis(ctx, (uint32_t)cols, (uint32_t)g.rc, false, true, src_kind, &ah, choose_strip_nt(ctx, b1->rows, g.rc, clahe), hmma_in));
    RC(get_axis(ctx, (uint32_t)scene_rows, (uint32_t)g.rr, false, false, 0, &av));
    RC(shard_geometry(scene_rows, cols, true, target, pad != 0, cs->world, cs->rank, clahe, &r0, &r1, &h0, &h1, &oy0, &oy1, &g,
                      &av->h));
    if (b1->rows != h1 - h0)
        return fail(ctx, SARPRO_ERR_INVALID_ARGUMENT, "rank %d must hold scene rows [%zu,%zu) (%zu rows), got %llu", cs->rank, h0, h1,
                    h1 - h0, (unsigned long long)b1->rows);
    const uint64_t rows = h1 - h0;
    PlanKind kinds[2] = {PlanKind::Autoscale, PlanKind::Autoscale};
    if (tamed_band_step && strategy == SARPRO_STRATEGY_TAMED) {
        kinds[0] = PlanKind::TamedSynRgbCopol;
        kinds[1] = PlanKind::TamedSynRgbCross;
    }

    // ---- stage + pass A on the owned rows ------------------------------------------------------------
    BandJob jobs[2];
    ShardGeom sg;
    sg.scene_rows = scene_rows;
    sg.row_off = h0;
    sg.own0 = r0 - h0;
    sg.own1 = r1 - h0;
    // (a width that is not a multiple of 8: the rank's rows are re-pitched like a whole scene's, see produce_bands in api.cu)
    const bool repitch = (cols % 8) != 0 && cols >= 512 && ctx->repitch && !ctx->force_exact;
    const uint64_t pitch = repitch ? ((cols + 7) & ~uint64_t(7)) : 0;
    for (int b = 0; b < 2; ++b) {
        BandWs& w = ctx->band[b];
        const uint16_t* dn = (const uint16_t*)ins[b]->data;
        if (repitch) {
            RC(reserve(ctx, w.dn_pad, std::max<size_t>(rows * pitch * 2, 16)));
            if (ins[b]->location == SARPRO_LOC_HOST) {
                CU(cudaMemcpy2DAsync(w.dn_pad.p, pitch * 2, ins[b]->data, cols * 2, cols * 2, rows, cudaMemcpyHostToDevice, ctx->stream));
                ctx->timing.h2d_bytes += rows * cols * 2;
                KL(launch_pad_cols((uint16_t*)w.dn_pad.p, (uint32_t)rows, (uint32_t)cols, (uint32_t)pitch, ctx->stream));
            } else {
                KL(launch_repitch(dn, (uint16_t*)w.dn_pad.p, (uint32_t)rows, (uint32_t)cols, (uint32_t)pitch, ctx->sm_count, ctx->stream));
            }
            dn = (const uint16_t*)w.dn_pad.p;
            w.pitch = pitch;
        } else if (ins[b]->location == SARPRO_LOC_HOST) {
            RC(reserve(ctx, w.dn, rows * cols * 2));
            CU(cudaMemcpyAsync(w.dn.p, ins[b]->data, rows * cols * 2, cudaMemcpyHostToDevice, ctx->stream));
            ctx->timing.h2d_bytes += rows * cols * 2;
            dn = (const uint16_t*)w.dn.p;
        }
        jobs[b].dn = dn;
        jobs[b].rows = rows;
        jobs[b].cols = cols;
        jobs[b].strategy = strategy;
        jobs[b].bit_depth = SARPRO_U8;
        jobs[b].kind = kinds[b];
    }
    // ---- per band: pass A -> DN-histogram all-reduce -> plan -> CLAHE tile statistics (+ all-reduce) -> pass B. Band 0's chain
    // (everything after its pass A) runs on the side stream, so its two collectives, its planner and its CLAHE statistics
    // overlap band 1's pass A, and its pass B starts while band 1's chain is still at its collectives; band 1's persistent
    // pass-B CTAs then fill the SMs band 0's leave. The collectives are issued in the same order on every rank (program
    // order: both bands' first one, then both bands' second one), which is what NCCL needs of one communicator used from two streams.
    const size_t esz = 1;
    const size_t n_out = g.oc * g.orr;
    HResizeArgs args[2];
    bool gates[2] = {false, false};
    // SMs that band 1's pass A and band 0's pass B leave free: band 0's collectives (NCCL kernels need an SM slot each), planner
    // and CLAHE statistics run beside band 1's pass A, band 1's beside band 0's pass B (measured without: band 0's all-reduce
    // waited 0.06 ms for pass A's persistent CTAs to drain)
    ctx->pair_spare = std::max(ctx->spare_sms, kCommSpareSms);
    cudaStream_t main_stream = ctx->stream;
    const bool two = ctx->two_stream && ctx->stream2;
    for (int b = 0; b < 2; ++b) RC(dn_pass_a_launch_sharded(ctx, b, jobs[b].dn, rows, cols, clahe, sg, 1));
    RC(dn_pass_a_launch_sharded(ctx, 0, jobs[0].dn, rows, cols, clahe, sg, 2));
    if (two) CU(cudaEventRecord(ctx->ev[4], main_stream));
    RC(dn_pass_a_launch_sharded(ctx, 1, jobs[1].dn, rows, cols, clahe, sg, 2));
    if (two) CU(cudaStreamWaitEvent(ctx->stream2, ctx->ev[4], 0));
    struct Hook { // all-reduce of one band's CLAHE tile histograms: every rank then builds all 64 CDFs of that band
        static int reduce(sarpro_ctx* ctx, void* arg) {
            const int b = *(const int*)arg;
            NcclApi& api = nccl();
            CommState* cs = ctx->comm;
            COMM_BEGIN();
            NC(api.AllReduce(ctx->band[b].tile256.p, ctx->band[b].tile256.p, (size_t)ctx->n_tiles * 256, kNcclUint32, kNcclSum, cs->comm, ctx->stream));
            COMM_END();
            return 0;
        }
    };
    // Two phases, each issued for band 0 (side stream) and then band 1 (main stream): NCCL runs a communicator's collectives
    // in issue order, so band 1's first all-reduce is queued before band 0's second one and does not wait behind it.
    int rc_b = 0;
    for (int phase = 0; phase < 2 && !rc_b; ++phase)
    for (int b = 0; b < 2 && !rc_b; ++b) {
        BandWs& w = ctx->band[b];
        ctx->stream = (two && b == 0) ? ctx->stream2 : main_stream;
        auto body = [&]() -> int {
            if (phase == 0) {
            // 1. the band's DN histogram over all ranks
            {
                COMM_BEGIN();
                NC(api.AllReduce(w.total.p, w.total.p, kDnBins, kNcclUint32, kNcclSum, cs->comm, ctx->stream));
                COMM_END();
            }
            // plan: on the device for the gamma == 1 strategies (no host round trip; every rank plans redundantly from the
            // merged histogram, bit-identically), else on the host from the merged dense totals
            if (plans_on_device(ctx, jobs[b])) {
                RC(plan_band_on_device(ctx, b, jobs[b]));
            } else {
                CU(cudaMemcpyAsync(ctx->h_hist + (size_t)b * kDnBins, w.total.p, kDnBins * 4, cudaMemcpyDeviceToHost, ctx->stream));
                CU(cudaStreamSynchronize(ctx->stream));
                ctx->timing.host_syncs++;
                plan_from_dn_histogram32(ctx->h_hist + (size_t)b * kDnBins, SARPRO_U8, strategy, kinds[b], &w.plan);
                std::memcpy(ctx->h_lut + (size_t)b * kDnBins, w.plan.lut.data(), kDnBins * 2);
                w.hot = w.plan.any_valid ? hmma_hot_from_plan(w.plan, &w.hot_top) : 0;
                CU(cudaMemcpyAsync(w.lut.p, ctx->h_lut + (size_t)b * kDnBins, kDnBins * 2, cudaMemcpyHostToDevice, ctx->stream));
                RC(upload_plan_dev(ctx, b));
                w.dev_planned = false;
                w.hist_auto_pending = true;
            }
            return 0;
            } // phase 0
            // 2. CLAHE tile histograms through the table, merged over the ranks, then the 64 CDFs
            if (clahe) {
                int slot = b;
                RC(run_clahe_stats_bands(ctx, &slot, 1, &Hook::reduce, &slot));
            }
            // pass B: horizontal pass over the held rows
            RC(reserve(ctx, w.temp, std::max<size_t>((size_t)rows * g.rc * esz, 16)));
            RC(reserve(ctx, w.small, std::max<size_t>(n_out * esz, 16)));
            CU(cudaMemsetAsync(w.small.p, 0, std::max<size_t>(n_out * esz, 1), ctx->stream));
            HResizeArgs a{};
            a.src = jobs[b].dn;
            a.src_rows = (uint32_t)rows;
            a.src_cols = (uint32_t)(w.pitch ? w.pitch : cols);
            a.src_width = (uint32_t)cols;
            a.lut = (const uint16_t*)w.lut.p;
            a.plan = (const PlanDev*)w.plan_dev.p;
            a.remap = nullptr;
            if (clahe) a.clahe = clahe_dev(ctx, b);
            a.minmax = clahe ? (uint32_t*)w.scalars.p : nullptr;
            a.row0 = 0;
            a.n_rows = (uint32_t)rows;
            a.temp = w.temp.p;
            a.ax = ah->dev();
            args[b] = a;
            RC(run_hpass(ctx, b, a, src_kind, 0, ah, h0, &gates[b]));
            return 0;
        };
        rc_b = body();
    }
    ctx->stream = main_stream;
    if (two) { // also on the error path: the side stream must not run past this call unobserved
        cudaError_t e1 = cudaEventRecord(ctx->ev_join, ctx->stream2);
        cudaError_t e2 = cudaStreamWaitEvent(main_stream, ctx->ev_join, 0);
        if (!rc_b) { CU(e1); CU(e2); }
    }
    RC(rc_b);
    // ---- owned output rows of every rank: the row bands partition the resized rows. Each rank writes its rows of both
    // bands into its slot of one buffer, with its CLAHE sample extrema in the slot's tail, and ONE all-gather moves rows and
    // extrema together (the all-reduce of the extrema and the 2 x world broadcasts this replaces cost two more collective
    // latencies per scene).
    if (cs->world > 16) return fail(ctx, SARPRO_ERR_COMM, "the sharded pipeline takes at most 16 ranks");
    GatherGeom gg{};
    gg.world = (uint32_t)cs->world;
    gg.out_pitch = (uint32_t)g.oc;
    gg.pad_top = (uint32_t)g.pad_top;
    gg.out_rows = (uint32_t)g.rr;
    gg.clahe = clahe ? 1u : 0u;
    for (int r = 0; r < cs->world; ++r) {
        size_t rr0, rr1, hh0, hh1, o0, o1;
        RC(shard_geometry(scene_rows, cols, true, target, pad != 0, cs->world, r, clahe, &rr0, &rr1, &hh0, &hh1, &o0, &o1, nullptr, &av->h));
        gg.oy0[r] = (uint32_t)o0;
        gg.oy1[r] = (uint32_t)o1;
        gg.max_rows = std::max<uint32_t>(gg.max_rows, (uint32_t)(o1 - o0));
    }
    gg.slot_bytes = (uint32_t)(((2 * (size_t)gg.max_rows * g.oc + 15) & ~(size_t)15) + 16);
    RC(reserve(ctx, ctx->gather, (size_t)cs->world * gg.slot_bytes));
    unsigned char* const my_slot = (unsigned char*)ctx->gather.p + (size_t)cs->rank * gg.slot_bytes;
    CU(cudaMemsetAsync(my_slot, 0, gg.slot_bytes, ctx->stream)); // pad columns of the rows
    // vertical pass for the owned output rows, straight into the slot (first run: assumes the tensor-core kernel took the band
    // and the CLAHE re-stretch is the identity; the checks after the exchange cover the rest)
    auto vpass = [&](int b, int stage, const uint32_t* skip, const uint32_t* run_if) -> int {
        BandWs& w = ctx->band[b];
        if (oy1 <= oy0) return 0;
        // output row oy lands at local row oy - oy0 of the band's part of the slot (the kernel indexes by absolute output row)
        unsigned char* dst = my_slot + (size_t)b * gg.max_rows * g.oc + g.pad_left;
        dst -= (size_t)oy0 * g.oc;
        KS(stage, launch_vresize(w.temp.p, (uint32_t)h0, (uint32_t)g.rc, av->dev(), (uint32_t)oy0, (uint32_t)oy1, dst, (uint32_t)g.oc, 0, 0,
                                 ctx->stream, skip, run_if));
        return 0;
    };
    for (int b = 0; b < 2; ++b) RC(vpass(b, SARPRO_STAGE_VRESIZE, nullptr, nullptr));
    RC(reserve(ctx, ctx->rgbsel, 64));
    uint32_t* const nonident = (uint32_t*)ctx->rgbsel.p + 4; // set by the unpack kernel: some band's re-stretch is not the identity
    auto exchange = [&]() -> int {
        if (clahe)
            KS(SARPRO_STAGE_COMM, launch_gather_tail((const uint32_t*)ctx->band[0].scalars.p, (const uint32_t*)ctx->band[1].scalars.p,
                                                     (uint32_t*)(my_slot + gg.slot_bytes - 16), ctx->stream));
        CU(cudaMemsetAsync(nonident, 0, 4, ctx->stream));
        {
            COMM_BEGIN();
            NC(api.AllGather(my_slot, ctx->gather.p, gg.slot_bytes, kNcclUint8, cs->comm, ctx->stream));
            COMM_END();
        }
        KS(SARPRO_STAGE_COMM, launch_gather_unpack((const unsigned char*)ctx->gather.p, gg, (unsigned char*)ctx->band[0].small.p,
                                                   (unsigned char*)ctx->band[1].small.p, (uint32_t*)ctx->band[0].scalars.p,
                                                   (uint32_t*)ctx->band[1].scalars.p, nonident, ctx->stream));
        RC(synrgb_compose(ctx, strategy, (const uint8_t*)ctx->band[0].small.p, (const uint8_t*)ctx->band[1].small.p, n_out));
        ctx->last[0] = sarpro_ctx::LastResult{ctx->rgb.p, g.oc, g.orr};
        if (out) {
            fill_image(out, g, 3, SARPRO_U8);
            if (out->data) RC(deliver(ctx, ctx->rgb.p, n_out * 3, out));
        }
        return 0;
    };
    if (n_out) RC(exchange());
    // ---- what the first pass speculated, checked once the result is there (one small read-back in front of the sync every
    // call ends with; every rank sees the same values, so all of them take the same path):
    //  * a device-planned band whose DN table does not fit the tensor-core kernel (plan->use_generic: more than 2000 hot DNs)
    //    was skipped by it: the generic exact kernel takes the band and the exchange runs again;
    //  * scale_u16_to_u8 after CLAHE (autoscale.rs:348-364) needs the sample extrema of the WHOLE scene. The first pass assumed
    //    the identity (extrema 0 / 255: any scene with an invalid pixel and a saturated one); the merged extrema arrive with the
    //    rows, and otherwise every rank re-runs its rows through the remap table and the exchange runs once more.
    if (n_out) {
        auto read_flags = [&]() -> int {
            ctx->h_scalars[7] = 0;
            ctx->h_scalars[5] = ctx->h_scalars[8 + 5] = 0;
            if (clahe) CU(cudaMemcpyAsync(ctx->h_scalars + 7, nonident, 4, cudaMemcpyDeviceToHost, ctx->stream));
            for (int b = 0; b < 2; ++b)
                if (gates[b]) CU(cudaMemcpyAsync(ctx->h_scalars + 8 * b + 5, &args[b].plan->use_generic, 4, cudaMemcpyDeviceToHost, ctx->stream));
            CU(cudaStreamSynchronize(ctx->stream)); // (the sync every call ends with; end_call's then returns at once)
            return 0;
        };
        RC(read_flags());
        bool redo = false;
        for (int b = 0; b < 2; ++b)
            if (gates[b] && ctx->h_scalars[8 * b + 5]) {
                RC(run_hpass_generic(ctx, args[b], src_kind, 0, ah));
                RC(vpass(b, SARPRO_STAGE_OTHER, nullptr, nullptr));
                gates[b] = false;
                redo = true;
            }
        if (redo) {
            ctx->timing.host_syncs++;
            RC(exchange());
            RC(read_flags());
        }
        if (clahe && ctx->h_scalars[7]) {
            ctx->timing.host_syncs++;
            for (int b = 0; b < 2; ++b) {
                BandWs& w = ctx->band[b];
                RC(reserve(ctx, w.remap, 256 + 16));
                uint32_t* flag = reinterpret_cast<uint32_t*>((unsigned char*)w.remap.p + 256);
                KS(SARPRO_STAGE_PLAN, launch_clahe_remap_decide((const uint32_t*)w.scalars.p, (uint8_t*)w.remap.p, flag, ctx->stream, nullptr));
                HResizeArgs ar = args[b];
                ar.remap = (const uint8_t*)w.remap.p;
                ar.minmax = nullptr;
                ar.skip = flag; // a band whose own re-stretch is the identity keeps its rows
                RC(run_hpass_generic(ctx, ar, src_kind, 0, ah));
                RC(vpass(b, SARPRO_STAGE_OTHER, flag, nullptr));
            }
            RC(exchange());
        }
    }
    return end_call(ctx);
}

} // extern "C"
