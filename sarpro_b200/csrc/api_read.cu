// api_read.cu — downsample-on-read entry points: the reader's `--size` flow (sentinel1.rs:1074-1109 ->
// gdal.rs:145-177) on the GPU, so that the host uploads the raw u16 raster once and the resampler runs at HBM speed instead of
// inside GDAL's RasterIO. The result is the f32 raster the rest of the reference pipeline starts from.
#include <algorithm>
#include <vector>

#include "ctx.h"

using namespace sarpro;

extern "C" {

int sarpro_read_dims_for_target(size_t cols, size_t rows, size_t target, size_t* out_cols, size_t* out_rows, int* alg) {
    if (!out_cols || !out_rows || cols == 0 || rows == 0 || target == 0) return SARPRO_ERR_INVALID_ARGUMENT;
    uint64_t oc = 0, orr = 0;
    int a = 0;
    read_dims_for_target(cols, rows, target, &oc, &orr, &a);
    *out_cols = (size_t)oc;
    *out_rows = (size_t)orr;
    if (alg) *alg = a;
    return SARPRO_OK;
}

int sarpro_read_row_plan_check(const uint16_t* samples, size_t in_size, size_t out_size, int alg, float* out) {
    if (!samples || !out || in_size == 0 || out_size == 0 || out_size > in_size) return SARPRO_ERR_INVALID_ARGUMENT;
    if (alg == SARPRO_RESAMPLE_AVERAGE) {
        ReadAverageAxisHost ax;
        build_read_average_axis(in_size, out_size, &ax);
        for (size_t d = 0; d < out_size; ++d) { // k_read_average with one source row (wy == 1)
            double total = 0.0, wsum = 0.0;
            for (int x = ax.start[d]; x < ax.end[d]; ++x) {
                const double w = x == ax.start[d] ? ax.w_first[d] : (x + 1 == ax.end[d] ? ax.w_last[d] : 1.0);
                total += (double)samples[x] * w;
                wsum += w;
            }
            out[d] = (float)(total / wsum);
        }
        return SARPRO_OK;
    }
    if (alg == SARPRO_RESAMPLE_LANCZOS) {
        ReadLanczosAxisHost ax;
        build_read_lanczos_axis(in_size, out_size, &ax);
        for (size_t d = 0; d < out_size; ++d) { // k_read_conv_h, then the identity vertical pass of a one-row raster
            double v = 0.0;
            for (int k = 0; k < ax.count[d]; ++k) v += (double)samples[ax.start[d] + k] * ax.w[d * (size_t)ax.window + k];
            out[d] = (float)v;
        }
        return SARPRO_OK;
    }
    return SARPRO_ERR_INVALID_ARGUMENT;
}

int sarpro_read_band_resampled(sarpro_ctx* ctx, const sarpro_band* in, size_t out_cols, size_t out_rows, int alg, float* out,
                               int out_location) {
    RC(begin_call(ctx));
    RC(check_band(ctx, in));
    if (!out) return fail(ctx, SARPRO_ERR_INVALID_ARGUMENT, "NULL argument");
    if (alg != SARPRO_RESAMPLE_AVERAGE && alg != SARPRO_RESAMPLE_LANCZOS) return fail(ctx, SARPRO_ERR_INVALID_ARGUMENT, "unknown resampler %d", alg);
    if (out_location != SARPRO_LOC_HOST && out_location != SARPRO_LOC_DEVICE) return fail(ctx, SARPRO_ERR_INVALID_ARGUMENT, "unknown location %d", out_location);
    const uint64_t rows = in->rows, cols = in->cols;
    if (rows == 0 || cols == 0 || out_cols == 0 || out_rows == 0) return fail(ctx, SARPRO_ERR_INVALID_ARGUMENT, "empty raster");
    if (out_cols > cols || out_rows > rows)
        return fail(ctx, SARPRO_ERR_INVALID_ARGUMENT, "downsample-on-read never enlarges (sentinel1.rs:1086: scale = min(target / long side, 1))");
    const int is16 = in->dtype == SARPRO_DT_U16;
    const size_t isz = is16 ? 2 : 4, n = rows * cols, n_out = out_cols * out_rows;
    BandWs& w = ctx->band[0];
    const void* src = in->data;
    if (in->location == SARPRO_LOC_HOST) {
        RC(reserve(ctx, w.f32a, std::max<size_t>(n * isz, 16)));
        CU(cudaMemcpyAsync(w.f32a.p, in->data, n * isz, cudaMemcpyHostToDevice, ctx->stream));
        ctx->timing.h2d_bytes += n * isz;
        src = w.f32a.p;
    }
    float* dst = out;
    if (out_location == SARPRO_LOC_HOST) {
        RC(reserve(ctx, w.small, std::max<size_t>(n_out * 4, 16)));
        dst = (float*)w.small.p;
    }
    // axis tables: small (one entry per output column / row), rebuilt per call; the buffers persist in the context
    auto up = [&](DevBuf& b, size_t off, const void* p, size_t bytes) -> int {
        CU(cudaMemcpyAsync((char*)b.p + off, p, bytes, cudaMemcpyHostToDevice, ctx->stream));
        return 0;
    };
    if (alg == SARPRO_RESAMPLE_AVERAGE) {
        ReadAverageAxisHost hx, hy;
        build_read_average_axis(cols, out_cols, &hx);
        build_read_average_axis(rows, out_rows, &hy);
        const size_t ex = out_cols, ey = out_rows;
        const size_t o_xs = 0, o_xe = o_xs + ex * 4, o_ys = o_xe + ex * 4, o_ye = o_ys + ey * 4;
        const size_t o_d = (o_ye + ey * 4 + 7) & ~(size_t)7; // doubles: xwf, xwl, ywf, ywl
        RC(reserve(ctx, w.edges, o_d + (2 * ex + 2 * ey) * 8));
        RC(up(w.edges, o_xs, hx.start.data(), ex * 4));
        RC(up(w.edges, o_xe, hx.end.data(), ex * 4));
        RC(up(w.edges, o_ys, hy.start.data(), ey * 4));
        RC(up(w.edges, o_ye, hy.end.data(), ey * 4));
        RC(up(w.edges, o_d, hx.w_first.data(), ex * 8));
        RC(up(w.edges, o_d + ex * 8, hx.w_last.data(), ex * 8));
        RC(up(w.edges, o_d + 2 * ex * 8, hy.w_first.data(), ey * 8));
        RC(up(w.edges, o_d + (2 * ex + ey) * 8, hy.w_last.data(), ey * 8));
        char* b = (char*)w.edges.p;
        const ReadAvgAxis ax{(const int*)(b + o_xs), (const int*)(b + o_xe), (const double*)(b + o_d), (const double*)(b + o_d + ex * 8)};
        const ReadAvgAxis ay{(const int*)(b + o_ys), (const int*)(b + o_ye), (const double*)(b + o_d + 2 * ex * 8),
                             (const double*)(b + o_d + (2 * ex + ey) * 8)};
        KS(SARPRO_STAGE_CONVERT, launch_read_average(src, is16, (uint32_t)rows, (uint32_t)cols, ax, ay, dst, (uint32_t)out_rows,
                                                     (uint32_t)out_cols, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream)); // the host tables are temporaries
    } else {
        ReadLanczosAxisHost hx, hy;
        build_read_lanczos_axis(cols, out_cols, &hx);
        build_read_lanczos_axis(rows, out_rows, &hy);
        const size_t ex = out_cols, ey = out_rows;
        const size_t o_xs = 0, o_xc = o_xs + ex * 4, o_ys = o_xc + ex * 4, o_yc = o_ys + ey * 4;
        const size_t o_wx = (o_yc + ey * 4 + 7) & ~(size_t)7, o_wy = o_wx + hx.w.size() * 8;
        RC(reserve(ctx, w.edges, o_wy + hy.w.size() * 8));
        RC(reserve(ctx, w.full, std::max<size_t>(rows * out_cols * 8, 16))); // f64 intermediate of the horizontal pass
        RC(up(w.edges, o_xs, hx.start.data(), ex * 4));
        RC(up(w.edges, o_xc, hx.count.data(), ex * 4));
        RC(up(w.edges, o_ys, hy.start.data(), ey * 4));
        RC(up(w.edges, o_yc, hy.count.data(), ey * 4));
        RC(up(w.edges, o_wx, hx.w.data(), hx.w.size() * 8));
        RC(up(w.edges, o_wy, hy.w.data(), hy.w.size() * 8));
        char* b = (char*)w.edges.p;
        const ReadConvAxis ax{(const int*)(b + o_xs), (const int*)(b + o_xc), (const double*)(b + o_wx), hx.window};
        const ReadConvAxis ay{(const int*)(b + o_ys), (const int*)(b + o_yc), (const double*)(b + o_wy), hy.window};
        KS(SARPRO_STAGE_CONVERT, launch_read_lanczos(src, is16, (uint32_t)rows, (uint32_t)cols, ax, ay, (double*)w.full.p, dst,
                                                     (uint32_t)out_rows, (uint32_t)out_cols, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
    }
    if (out_location == SARPRO_LOC_HOST) {
        CU(cudaMemcpyAsync(out, dst, n_out * 4, cudaMemcpyDeviceToHost, ctx->stream));
        ctx->timing.d2h_bytes += n_out * 4;
    }
    return end_call(ctx);
}

} // extern "C"
