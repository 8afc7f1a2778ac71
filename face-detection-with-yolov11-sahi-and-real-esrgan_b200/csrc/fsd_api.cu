// Handle management, error reporting and the host-side integer planners of the C ABI.
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include "fsd_common.cuh"

namespace fsd {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
}  // namespace fsd

extern "C" {

const char* fsd_version(void) { return "fsd_b200 0.1 (sm_100a)"; }
const char* fsd_last_error(void) { return fsd::g_err; }

int fsd_create(int device, fsd_handle_t* out) {
    FSD_CHECK_ARG(out != nullptr, "fsd_create: out is null");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0 || device < 0 || device >= n) {
        fsd::set_error("fsd_create: no usable CUDA device %d (%s); this library has no CPU fallback", device,
                       e == cudaSuccess ? "device index out of range" : cudaGetErrorString(e));
        return FSD_ERR_NO_DEVICE;
    }
    cudaDeviceProp prop;
    FSD_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        fsd::set_error("fsd_create: device %d is sm_%d%d; kernels are built for sm_100a only", device,
                       prop.major, prop.minor);
        return FSD_ERR_NO_DEVICE;
    }
    FSD_CUDA(cudaSetDevice(device));
    fsd_context* c = new fsd_context();
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    cudaDriverEntryPointQueryResult qres;
    void* fn = nullptr;
    e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || fn == nullptr) {
        delete c;
        fsd::set_error("fsd_create: cuTensorMapEncodeTiled not available from the driver");
        return FSD_ERR_CUDA;
    }
    c->encode_tiled = fn;
    *out = c;
    return FSD_OK;
}

int fsd_destroy(fsd_handle_t h) {
    if (!h) return FSD_OK;
    cudaSetDevice(h->device);
    for (auto& kv : h->resize_tables)
        if (kv.second.dev) cudaFree(kv.second.dev);
    for (void* d : h->dev_allocs) cudaFree(d);
    for (auto& t : h->timing_samples) { cudaEventDestroy(t.e0); cudaEventDestroy(t.e1); }
    for (cudaEvent_t e : h->event_pool) cudaEventDestroy(e);
    delete h;
    return FSD_OK;
}

int64_t fsd_launch_count(fsd_handle_t h) { return h ? h->launches : 0; }

int fsd_kernel_timing_enable(fsd_handle_t h, unsigned kernel_mask) {
    FSD_CHECK_ARG(h != nullptr, "fsd_kernel_timing_enable: null handle");
    FSD_CUDA(cudaSetDevice(h->device));
    for (auto& t : h->timing_samples) { h->event_pool.push_back(t.e0); h->event_pool.push_back(t.e1); }
    h->timing_samples.clear();
    h->timing = kernel_mask;
    return FSD_OK;
}

int fsd_kernel_timing_read(fsd_handle_t h, double* samples, int cap, int* n) {
    FSD_CHECK_ARG(h != nullptr && n != nullptr, "fsd_kernel_timing_read: null argument");
    FSD_CUDA(cudaSetDevice(h->device));
    const int total = (int)h->timing_samples.size();
    *n = total;
    if (!samples) return FSD_OK;
    for (int i = 0; i < total && i < cap; ++i) {
        auto& t = h->timing_samples[i];
        FSD_CUDA(cudaEventSynchronize(t.e1));
        float ms = 0.f;
        FSD_CUDA(cudaEventElapsedTime(&ms, t.e0, t.e1));
        samples[4 * i + 0] = (double)t.kernel; samples[4 * i + 1] = (double)t.units;
        samples[4 * i + 2] = (double)t.tag; samples[4 * i + 3] = (double)ms;
    }
    if (total > cap) { fsd::set_error("fsd_kernel_timing_read: %d samples exceed capacity %d", total, cap); return FSD_ERR_CAPACITY; }
    return FSD_OK;
}

// (a2) sahi.slicing.get_slice_bboxes restated: overlap = int(ratio*slice) truncation,
// row-major grid, border slices shifted back inside the image.
int fsd_slice_plan(int image_h, int image_w, int slice_h, int slice_w, double overlap_h_ratio,
                   double overlap_w_ratio, int32_t* boxes, int cap, int* n) {
    FSD_CHECK_ARG(n != nullptr, "fsd_slice_plan: n is null");
    FSD_CHECK_ARG(image_h > 0 && image_w > 0 && slice_h > 0 && slice_w > 0, "fsd_slice_plan: sizes must be > 0");
    int y_overlap = (int)(overlap_h_ratio * (double)slice_h);
    int x_overlap = (int)(overlap_w_ratio * (double)slice_w);
    FSD_CHECK_ARG(y_overlap < slice_h && x_overlap < slice_w && y_overlap >= 0 && x_overlap >= 0,
                  "fsd_slice_plan: overlap ratio must be in [0,1)");
    int count = 0;
    int y_max = 0, y_min = 0;
    while (y_max < image_h) {
        int x_min = 0, x_max = 0;
        y_max = y_min + slice_h;
        while (x_max < image_w) {
            x_max = x_min + slice_w;
            int bx0, by0, bx1, by1;
            if (y_max > image_h || x_max > image_w) {
                bx1 = x_max < image_w ? x_max : image_w;
                by1 = y_max < image_h ? y_max : image_h;
                bx0 = bx1 - slice_w > 0 ? bx1 - slice_w : 0;
                by0 = by1 - slice_h > 0 ? by1 - slice_h : 0;
            } else {
                bx0 = x_min; by0 = y_min; bx1 = x_max; by1 = y_max;
            }
            if (boxes && count < cap) {
                boxes[4 * count + 0] = bx0; boxes[4 * count + 1] = by0;
                boxes[4 * count + 2] = bx1; boxes[4 * count + 3] = by1;
            }
            ++count;
            x_min = x_max - x_overlap;
        }
        y_min = y_max - y_overlap;
    }
    *n = count;
    if (boxes && count > cap) {
        fsd::set_error("fsd_slice_plan: %d slices exceed capacity %d", count, cap);
        return FSD_ERR_CAPACITY;
    }
    return FSD_OK;
}

// (a4) ultralytics LetterBox(new_shape=imgsz, auto=True, scaleup=True, center=True, stride) geometry
// (SURVEY App. A.3).  Python's round() on a double is round-half-even == nearbyint().
int fsd_letterbox_geometry(int src_h, int src_w, int imgsz, int stride, int32_t geom[8], double* gain) {
    FSD_CHECK_ARG(geom != nullptr, "fsd_letterbox_geometry: geom is null");
    FSD_CHECK_ARG(src_h > 0 && src_w > 0 && imgsz > 0 && stride > 0, "fsd_letterbox_geometry: sizes must be > 0");
    double r = fmin((double)imgsz / (double)src_h, (double)imgsz / (double)src_w);
    int new_w = (int)nearbyint((double)src_w * r);
    int new_h = (int)nearbyint((double)src_h * r);
    double dw = (double)((imgsz - new_w) % stride + ((imgsz - new_w) % stride < 0 ? stride : 0));
    double dh = (double)((imgsz - new_h) % stride + ((imgsz - new_h) % stride < 0 ? stride : 0));
    dw /= 2.0; dh /= 2.0;
    int top = (int)nearbyint(dh - 0.1), bottom = (int)nearbyint(dh + 0.1);
    int left = (int)nearbyint(dw - 0.1), right = (int)nearbyint(dw + 0.1);
    int mode = 1;
    if (new_w == src_w && new_h == src_h) mode = 0;
    else if (src_w == 2 * new_w && src_h == 2 * new_h) mode = 2;  // cv2 INTER_LINEAR -> 2x2 area fast path
    geom[0] = new_w; geom[1] = new_h; geom[2] = left; geom[3] = top;
    geom[4] = new_w + left + right; geom[5] = new_h + top + bottom; geom[6] = mode; geom[7] = 0;
    if (gain) {
        // ultralytics scale_boxes: gain = min(img1_h/img0_h, img1_w/img0_w) on the PADDED network shape
        double g0 = (double)geom[5] / (double)src_h, g1 = (double)geom[4] / (double)src_w;
        *gain = g0 < g1 ? g0 : g1;
    }
    return FSD_OK;
}

}  // extern "C"
