// api_jpeg.cu — encoder hand-off: write_gray_jpeg / write_rgb_jpeg (src/io/writers/jpeg.rs:6-30: the
// jpeg_encoder crate at quality 100, 4:4:4, baseline, standard Huffman tables) on the GPU with nvJPEG, so that after a
// 1.5 ms raster stage the 12.6 MB RGB image is encoded where it lies and only the JPEG stream crosses PCIe; the host's
// writer then just writes bytes. nvJPEG is a library encoder (like calling cuBLAS): it is dlopen()ed from the CUDA toolkit,
// not linked, and the call fails loudly when it is missing. The stream is not byte-identical to jpeg_encoder's (both are
// baseline JPEG with all-ones quantisation tables at quality 100; entropy coding and the colour conversion's fixed-point
// details differ): the parity bar for this row is the decoded image, within the +-2 levels two conforming q=100 codecs differ by.
#include <dlfcn.h>
#include <nvjpeg.h>

#include <algorithm>
#include <string>

#include "ctx.h"

namespace sarpro {

struct JpegApi {
    void* handle = nullptr;
    nvjpegStatus_t (*CreateSimple)(nvjpegHandle_t*) = nullptr;
    nvjpegStatus_t (*Destroy)(nvjpegHandle_t) = nullptr;
    nvjpegStatus_t (*EncoderStateCreate)(nvjpegHandle_t, nvjpegEncoderState_t*, cudaStream_t) = nullptr;
    nvjpegStatus_t (*EncoderStateDestroy)(nvjpegEncoderState_t) = nullptr;
    nvjpegStatus_t (*EncoderParamsCreate)(nvjpegHandle_t, nvjpegEncoderParams_t*, cudaStream_t) = nullptr;
    nvjpegStatus_t (*EncoderParamsDestroy)(nvjpegEncoderParams_t) = nullptr;
    nvjpegStatus_t (*SetQuality)(nvjpegEncoderParams_t, const int, cudaStream_t) = nullptr;
    nvjpegStatus_t (*SetEncoding)(nvjpegEncoderParams_t, nvjpegJpegEncoding_t, cudaStream_t) = nullptr;
    nvjpegStatus_t (*SetOptimizedHuffman)(nvjpegEncoderParams_t, const int, cudaStream_t) = nullptr;
    nvjpegStatus_t (*SetSamplingFactors)(nvjpegEncoderParams_t, const nvjpegChromaSubsampling_t, cudaStream_t) = nullptr;
    nvjpegStatus_t (*EncodeImage)(nvjpegHandle_t, nvjpegEncoderState_t, const nvjpegEncoderParams_t, const nvjpegImage_t*,
                                  nvjpegInputFormat_t, int, int, cudaStream_t) = nullptr;
    nvjpegStatus_t (*EncodeYUV)(nvjpegHandle_t, nvjpegEncoderState_t, const nvjpegEncoderParams_t, const nvjpegImage_t*,
                                nvjpegChromaSubsampling_t, int, int, cudaStream_t) = nullptr;
    nvjpegStatus_t (*RetrieveBitstream)(nvjpegHandle_t, nvjpegEncoderState_t, unsigned char*, size_t*, cudaStream_t) = nullptr;
    std::string error;
    bool ok = false;
};

static JpegApi& jpeg_api() {
    static JpegApi api;
    static bool tried = false;
    if (tried) return api;
    tried = true;
    const char* names[] = {getenv("SARPRO_NVJPEG_LIB"), "libnvjpeg.so.12", "/usr/local/cuda/lib64/libnvjpeg.so.12", "libnvjpeg.so"};
    for (const char* n : names) {
        if (!n) continue;
        api.handle = dlopen(n, RTLD_NOW | RTLD_LOCAL);
        if (api.handle) break;
    }
    if (!api.handle) {
        api.error = std::string("cannot dlopen libnvjpeg.so.12 (set SARPRO_NVJPEG_LIB): ") + (dlerror() ? dlerror() : "");
        return api;
    }
#define SARPRO_JSYM(field, name)                                                       \
    *(void**)(&api.field) = dlsym(api.handle, name);                                   \
    if (!api.field) { api.error = std::string("missing nvJPEG symbol ") + name; return api; }
    SARPRO_JSYM(CreateSimple, "nvjpegCreateSimple")
    SARPRO_JSYM(Destroy, "nvjpegDestroy")
    SARPRO_JSYM(EncoderStateCreate, "nvjpegEncoderStateCreate")
    SARPRO_JSYM(EncoderStateDestroy, "nvjpegEncoderStateDestroy")
    SARPRO_JSYM(EncoderParamsCreate, "nvjpegEncoderParamsCreate")
    SARPRO_JSYM(EncoderParamsDestroy, "nvjpegEncoderParamsDestroy")
    SARPRO_JSYM(SetQuality, "nvjpegEncoderParamsSetQuality")
    SARPRO_JSYM(SetEncoding, "nvjpegEncoderParamsSetEncoding")
    SARPRO_JSYM(SetOptimizedHuffman, "nvjpegEncoderParamsSetOptimizedHuffman")
    SARPRO_JSYM(SetSamplingFactors, "nvjpegEncoderParamsSetSamplingFactors")
    SARPRO_JSYM(EncodeImage, "nvjpegEncodeImage")
    SARPRO_JSYM(EncodeYUV, "nvjpegEncodeYUV")
    SARPRO_JSYM(RetrieveBitstream, "nvjpegEncodeRetrieveBitstream")
#undef SARPRO_JSYM
    api.ok = true;
    return api;
}

struct JpegState {
    nvjpegHandle_t handle = nullptr;
    nvjpegEncoderState_t state = nullptr;
    nvjpegEncoderParams_t params = nullptr;
};

void jpeg_state_destroy(sarpro_ctx* ctx) {
    JpegState* js = ctx->jpeg;
    if (!js) return;
    JpegApi& api = jpeg_api();
    if (api.ok) {
        if (js->params) api.EncoderParamsDestroy(js->params);
        if (js->state) api.EncoderStateDestroy(js->state);
        if (js->handle) api.Destroy(js->handle);
    }
    delete js;
    ctx->jpeg = nullptr;
}

#define NJ(call)                                                                                                  \
    do {                                                                                                          \
        nvjpegStatus_t s__ = (call);                                                                              \
        if (s__ != NVJPEG_STATUS_SUCCESS) return fail(ctx, SARPRO_ERR_INTERNAL, "nvJPEG error %d at %s:%d", (int)s__, __FILE__, __LINE__); \
    } while (0)

static int jpeg_state(sarpro_ctx* ctx, JpegState** out) {
    JpegApi& api = jpeg_api();
    if (!api.ok) return fail(ctx, SARPRO_ERR_INTERNAL, "%s", api.error.c_str());
    if (!ctx->jpeg) {
        JpegState* js = new JpegState();
        ctx->jpeg = js;
        nvjpegStatus_t st = api.CreateSimple(&js->handle);
        if (st == NVJPEG_STATUS_SUCCESS) st = api.EncoderStateCreate(js->handle, &js->state, ctx->stream);
        if (st == NVJPEG_STATUS_SUCCESS) st = api.EncoderParamsCreate(js->handle, &js->params, ctx->stream);
        if (st != NVJPEG_STATUS_SUCCESS) {
            jpeg_state_destroy(ctx); // never leave a half-built encoder behind
            return fail(ctx, SARPRO_ERR_INTERNAL, "nvJPEG encoder setup failed with status %d", (int)st);
        }
    }
    *out = ctx->jpeg;
    return 0;
}

// dev: u8 image on the device, row-major, channels 1 (gray) or 3 (interleaved RGB)
static int encode_device_image(sarpro_ctx* ctx, const unsigned char* dev, size_t cols, size_t rows, int channels, int quality, void* out,
                               size_t capacity, size_t* out_bytes) {
    if (!out_bytes) return fail(ctx, SARPRO_ERR_INVALID_ARGUMENT, "NULL argument");
    if (quality < 1 || quality > 100) return fail(ctx, SARPRO_ERR_INVALID_ARGUMENT, "JPEG quality %d outside 1..100", quality);
    if (cols == 0 || rows == 0 || cols > 65535 || rows > 65535) // jpeg.rs:15,28: `cols as u16`, `rows as u16`
        return fail(ctx, SARPRO_ERR_INVALID_ARGUMENT, "JPEG dimensions %zux%zu outside 1..65535", cols, rows);
    JpegApi& api = jpeg_api();
    JpegState* js = nullptr;
    RC(jpeg_state(ctx, &js));
    NJ(api.SetQuality(js->params, quality, ctx->stream));
    NJ(api.SetEncoding(js->params, NVJPEG_ENCODING_BASELINE_DCT, ctx->stream));
    NJ(api.SetOptimizedHuffman(js->params, 0, ctx->stream));
    nvjpegImage_t img{};
    if (channels == 3) {
        // jpeg_encoder picks 4:4:4 from quality 90 up and 4:2:0 below (Encoder::new)
        NJ(api.SetSamplingFactors(js->params, quality >= 90 ? NVJPEG_CSS_444 : NVJPEG_CSS_420, ctx->stream));
        img.channel[0] = const_cast<unsigned char*>(dev);
        img.pitch[0] = cols * 3;
        NJ(api.EncodeImage(js->handle, js->state, js->params, &img, NVJPEG_INPUT_RGBI, (int)cols, (int)rows, ctx->stream));
    } else {
        NJ(api.SetSamplingFactors(js->params, NVJPEG_CSS_GRAY, ctx->stream));
        img.channel[0] = const_cast<unsigned char*>(dev);
        img.pitch[0] = cols;
        NJ(api.EncodeYUV(js->handle, js->state, js->params, &img, NVJPEG_CSS_GRAY, (int)cols, (int)rows, ctx->stream));
    }
    size_t len = 0;
    NJ(api.RetrieveBitstream(js->handle, js->state, nullptr, &len, ctx->stream));
    *out_bytes = len;
    if (!out) return 0; // size query
    if (len > capacity) return fail(ctx, SARPRO_ERR_INVALID_ARGUMENT, "JPEG stream is %zu bytes, the buffer holds %zu", len, capacity);
    NJ(api.RetrieveBitstream(js->handle, js->state, (unsigned char*)out, &len, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->timing.d2h_bytes += len;
    *out_bytes = len;
    return 0;
}

} // namespace sarpro

using namespace sarpro;

extern "C" {

int sarpro_encode_jpeg(sarpro_ctx* ctx, const sarpro_image* img, int quality, void* out, size_t capacity, size_t* out_bytes) {
    RC(begin_call(ctx));
    if (!img || !img->data) return fail(ctx, SARPRO_ERR_INVALID_ARGUMENT, "NULL argument");
    if (img->bit_depth != SARPRO_U8 || (img->channels != 1 && img->channels != 3))
        return fail(ctx, SARPRO_ERR_INVALID_ARGUMENT, "JPEG takes u8 gray or interleaved RGB (save.rs:121, 321 force U8 for JPEG)");
    const size_t bytes = (size_t)img->cols * img->rows * img->channels;
    const unsigned char* dev = (const unsigned char*)img->data;
    if (img->location == SARPRO_LOC_HOST) {
        RC(reserve(ctx, ctx->band[0].full, std::max<size_t>(bytes, 16)));
        CU(cudaMemcpyAsync(ctx->band[0].full.p, img->data, bytes, cudaMemcpyHostToDevice, ctx->stream));
        ctx->timing.h2d_bytes += bytes;
        dev = (const unsigned char*)ctx->band[0].full.p;
    }
    RC(encode_device_image(ctx, dev, img->cols, img->rows, img->channels, quality, out, capacity, out_bytes));
    return end_call(ctx);
}

int sarpro_encode_last_jpeg(sarpro_ctx* ctx, int which, int quality, void* out, size_t capacity, size_t* out_bytes) {
    if (!ctx) return SARPRO_ERR_INVALID_ARGUMENT;
    ctx->keep_last = true;
    const int rc0 = begin_call(ctx);
    ctx->keep_last = false;
    RC(rc0);
    if (which < 0 || which > 2) return fail(ctx, SARPRO_ERR_INVALID_ARGUMENT, "which = 0 (RGB), 1 or 2 (gray band)");
    const sarpro_ctx::LastResult& lr = ctx->last[which];
    if (!lr.dev || lr.cols == 0 || lr.rows == 0)
        return fail(ctx, SARPRO_ERR_INVALID_ARGUMENT, "the context holds no u8 %s result of a pipeline call", which == 0 ? "RGB" : "gray");
    RC(encode_device_image(ctx, (const unsigned char*)lr.dev, lr.cols, lr.rows, which == 0 ? 3 : 1, quality, out, capacity, out_bytes));
    return end_call(ctx);
}

} // extern "C"
