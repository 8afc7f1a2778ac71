// (f4) JPEG ingest on the device — SURVEY §8 f4.
//
// The reference decodes every input with PIL on the host (read_image_as_pil -> Image.open(path).convert("RGB"), [EXT sahi.utils.cv],
// reached from docs sahi/predict.py:229 and again at docs sahi/prediction.py:173), and its pipelines hand intermediate images from
// stage to stage as temporary JPEG files (pipeline_v4_yolo/1_Inference.py:328-330).  fsd_jpeg_decode decodes a baseline JPEG with
// nvJPEG straight into a slot of the image pool (interleaved RGB, or BGR for cv2.imread parity) on the caller's stream, so the
// compressed bytes are the only H2D traffic.  libnvjpeg is dlopen'ed on first use: the library itself has no link-time dependency
// on it, and a box without it gets FSD_ERR_ARG with a message from this entry point only.
//
// NOT bit-exact with PIL: nvJPEG's IDCT / chroma up-sampling differ from libjpeg-turbo's by a few LSB, so this path is opt-in
// (`ImagePool.upload_jpeg`) and sits outside the bit-exact parity claims (tests/test_jpeg_ingest_gpu.py states the tolerance).
#include <dlfcn.h>
#include <nvjpeg.h>

#include "fsd_common.cuh"

namespace fsd {

struct NvJpegApi {
    void* lib = nullptr;
    nvjpegStatus_t (*create)(nvjpegHandle_t*) = nullptr;
    nvjpegStatus_t (*destroy)(nvjpegHandle_t) = nullptr;
    nvjpegStatus_t (*state_create)(nvjpegHandle_t, nvjpegJpegState_t*) = nullptr;
    nvjpegStatus_t (*state_destroy)(nvjpegJpegState_t) = nullptr;
    nvjpegStatus_t (*info)(nvjpegHandle_t, const unsigned char*, size_t, int*, nvjpegChromaSubsampling_t*, int*, int*) = nullptr;
    nvjpegStatus_t (*decode)(nvjpegHandle_t, nvjpegJpegState_t, const unsigned char*, size_t, nvjpegOutputFormat_t, nvjpegImage_t*,
                             cudaStream_t) = nullptr;
    nvjpegHandle_t handle = nullptr;
    nvjpegJpegState_t state = nullptr;
    bool ok = false;
};

static NvJpegApi g_nvjpeg;  // one per process (nvjpegCreateSimple binds to the current device's context lazily)
static std::mutex g_nvjpeg_mu;

static bool nvjpeg_ready() {
    NvJpegApi& a = g_nvjpeg;
    if (a.ok) return true;
    if (!a.lib) {
        for (const char* name : {"libnvjpeg.so.12", "libnvjpeg.so", "/usr/local/cuda/lib64/libnvjpeg.so.12"}) {
            a.lib = dlopen(name, RTLD_NOW | RTLD_LOCAL);
            if (a.lib) break;
        }
        if (!a.lib) { set_error("fsd_jpeg: libnvjpeg could not be loaded (%s)", dlerror()); return false; }
    }
#define FSD_SYM(field, sym)                                                                  \
    a.field = reinterpret_cast<decltype(a.field)>(dlsym(a.lib, sym));                       \
    if (!a.field) { set_error("fsd_jpeg: %s missing from libnvjpeg", sym); return false; }
    FSD_SYM(create, "nvjpegCreateSimple")
    FSD_SYM(destroy, "nvjpegDestroy")
    FSD_SYM(state_create, "nvjpegJpegStateCreate")
    FSD_SYM(state_destroy, "nvjpegJpegStateDestroy")
    FSD_SYM(info, "nvjpegGetImageInfo")
    FSD_SYM(decode, "nvjpegDecode")
#undef FSD_SYM
    if (a.create(&a.handle) != NVJPEG_STATUS_SUCCESS) { set_error("fsd_jpeg: nvjpegCreateSimple failed"); return false; }
    if (a.state_create(a.handle, &a.state) != NVJPEG_STATUS_SUCCESS) { set_error("fsd_jpeg: nvjpegJpegStateCreate failed"); return false; }
    a.ok = true;
    return true;
}

}  // namespace fsd

using namespace fsd;

extern "C" int fsd_jpeg_info(const uint8_t* jpeg, int64_t len, int* width, int* height, int* channels) {
    FSD_CHECK_ARG(jpeg && len > 0 && width && height && channels, "fsd_jpeg_info: null argument");
    std::lock_guard<std::mutex> lock(g_nvjpeg_mu);
    if (!nvjpeg_ready()) return FSD_ERR_ARG;
    int comps = 0, ws[NVJPEG_MAX_COMPONENT] = {0}, hs[NVJPEG_MAX_COMPONENT] = {0};
    nvjpegChromaSubsampling_t sub;
    if (g_nvjpeg.info(g_nvjpeg.handle, jpeg, (size_t)len, &comps, &sub, ws, hs) != NVJPEG_STATUS_SUCCESS) {
        set_error("fsd_jpeg_info: not a JPEG stream nvJPEG can parse");
        return FSD_ERR_ARG;
    }
    *width = ws[0]; *height = hs[0]; *channels = comps;
    return FSD_OK;
}

extern "C" int fsd_jpeg_decode(fsd_handle_t h, const uint8_t* jpeg, int64_t len, int bgr, uint8_t* dst, int64_t row_pitch, int H,
                               int W, void* stream_) {
    FSD_CHECK_ARG(h && jpeg && len > 0 && dst && H > 0 && W > 0 && row_pitch >= (int64_t)W * 3, "fsd_jpeg_decode: bad argument");
    std::lock_guard<std::mutex> lock(g_nvjpeg_mu);
    FSD_CUDA(cudaSetDevice(h->device));
    if (!nvjpeg_ready()) return FSD_ERR_ARG;
    int comps = 0, ws[NVJPEG_MAX_COMPONENT] = {0}, hs[NVJPEG_MAX_COMPONENT] = {0};
    nvjpegChromaSubsampling_t sub;
    if (g_nvjpeg.info(g_nvjpeg.handle, jpeg, (size_t)len, &comps, &sub, ws, hs) != NVJPEG_STATUS_SUCCESS) {
        set_error("fsd_jpeg_decode: not a JPEG stream nvJPEG can parse");
        return FSD_ERR_ARG;
    }
    if (ws[0] != W || hs[0] != H) {
        set_error("fsd_jpeg_decode: the stream is %dx%d, the destination slot %dx%d", ws[0], hs[0], W, H);
        return FSD_ERR_ARG;
    }
    nvjpegImage_t out;
    for (int c = 0; c < NVJPEG_MAX_COMPONENT; ++c) { out.channel[c] = nullptr; out.pitch[c] = 0; }
    out.channel[0] = dst;
    out.pitch[0] = (size_t)row_pitch;
    const nvjpegStatus_t st = g_nvjpeg.decode(g_nvjpeg.handle, g_nvjpeg.state, jpeg, (size_t)len,
                                              bgr ? NVJPEG_OUTPUT_BGRI : NVJPEG_OUTPUT_RGBI, &out, (cudaStream_t)stream_);
    if (st != NVJPEG_STATUS_SUCCESS) {
        set_error("fsd_jpeg_decode: nvjpegDecode failed with status %d", (int)st);
        return FSD_ERR_CUDA;
    }
    h->launches += 1;  // (nvJPEG's own kernels: counted as one library call)
    return FSD_OK;
}
