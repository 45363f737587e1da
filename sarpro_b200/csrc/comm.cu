// comm.cu — multi-GPU row-band sharding (one process per GPU).
#include <algorithm>

#include "ctx.h"

using namespace sarpro;

extern "C" {

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
    size_t r0, r1;
    int rc = sarpro_shard_rows(rows, world, rank, clahe, &r0, &r1);
    if (rc) return rc;
    if (!h0 || !h1) return SARPRO_ERR_INVALID_ARGUMENT;
    *h0 = r0;
    *h1 = r1;
    (void)cols; (void)has_target; (void)target;
    return SARPRO_OK;
}

int sarpro_comm_unique_id(void* out128) { (void)out128; return SARPRO_ERR_COMM; }
int sarpro_comm_init(sarpro_ctx* ctx, const void* unique_id128, int rank, int world) {
    (void)unique_id128; (void)rank; (void)world;
    return fail(ctx, SARPRO_ERR_COMM, "multi-GPU support not built in this revision");
}
int sarpro_comm_destroy(sarpro_ctx* ctx) { (void)ctx; return SARPRO_OK; }
int sarpro_pipeline_synrgb_sharded(sarpro_ctx* ctx, const sarpro_band* b1, const sarpro_band* b2, size_t scene_rows,
                                   int strategy, int mode, int has_target, size_t target, int pad, int tamed_band_step,
                                   sarpro_image* out) {
    (void)b1; (void)b2; (void)scene_rows; (void)strategy; (void)mode; (void)has_target; (void)target; (void)pad;
    (void)tamed_band_step; (void)out;
    return fail(ctx, SARPRO_ERR_COMM, "multi-GPU support not built in this revision");
}

} // extern "C"
