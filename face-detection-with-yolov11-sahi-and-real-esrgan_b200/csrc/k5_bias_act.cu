// Fused conv epilogue for the PyTorch backbones: y = act(x + bias[c]) in place, one pass over HBM.
// cuDNN (through torch) produces the raw convolution; eager PyTorch would then launch a broadcast bias-add and a
// SiLU / LeakyReLU kernel (two more read+write passes, the bias-add non-vectorised).  The ncu launch list of the
// sliced-detection step showed those two passes costing ~3x the convolutions themselves (profiles/r1_launches_*).
// Layout: channels-last ([N,H,W,C] dense, C % 8 == 0 for fp16 / C % 4 == 0 for fp32) so the bias index is idx % C.
#include "fsd_common.cuh"

namespace fsd {

constexpr int K5_THREADS = 256;

template <int ACT> __device__ __forceinline__ float activate(float v, float slope) {
    if (ACT == 1) return __fdividef(v, 1.0f + __expf(-v));  // SiLU
    if (ACT == 2) return v > 0.f ? v : v * slope;           // LeakyReLU
    return v;
}

template <int ACT>
__global__ void __launch_bounds__(K5_THREADS)
k5_bias_act_half_kernel(uint4* __restrict__ x, const uint4* __restrict__ bias, size_t n_vec, int c_vec, float slope) {
    for (size_t i = (size_t)blockIdx.x * K5_THREADS + threadIdx.x; i < n_vec; i += (size_t)gridDim.x * K5_THREADS) {
        uint4 v = x[i];
        const uint4 b = __ldg(bias + (i % c_vec));
        __half2* hv = reinterpret_cast<__half2*>(&v);
        const __half2* hb = reinterpret_cast<const __half2*>(&b);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float2 f = __half22float2(hv[k]);
            const float2 g = __half22float2(hb[k]);
            hv[k] = __floats2half2_rn(activate<ACT>(f.x + g.x, slope), activate<ACT>(f.y + g.y, slope));
        }
        x[i] = v;
    }
}

template <int ACT>
__global__ void __launch_bounds__(K5_THREADS)
k5_bias_act_float_kernel(float4* __restrict__ x, const float4* __restrict__ bias, size_t n_vec, int c_vec, float slope) {
    for (size_t i = (size_t)blockIdx.x * K5_THREADS + threadIdx.x; i < n_vec; i += (size_t)gridDim.x * K5_THREADS) {
        float4 v = x[i];
        const float4 b = __ldg(bias + (i % c_vec));
        v.x = activate<ACT>(v.x + b.x, slope); v.y = activate<ACT>(v.y + b.y, slope);
        v.z = activate<ACT>(v.z + b.z, slope); v.w = activate<ACT>(v.w + b.w, slope);
        x[i] = v;
    }
}

// out[n,y,x,:] = concat(a[n,y/2,x/2,:], b[n,y,x,:]) — the FPN "nearest 2x up-sample, then concat" of the YOLO neck in ONE
// pass (torch runs an up-sample kernel, then a concat kernel: the up-sampled tensor is written and read once more).
__global__ void __launch_bounds__(K5_THREADS)
k5_upsample2x_concat_kernel(const uint4* __restrict__ a, const uint4* __restrict__ b, uint4* __restrict__ out,
                            int N, int h, int w, int ca_vec, int cb_vec) {
    const int H = 2 * h, W = 2 * w, cv = ca_vec + cb_vec;
    const size_t total = (size_t)N * H * W * cv;
    for (size_t i = (size_t)blockIdx.x * K5_THREADS + threadIdx.x; i < total; i += (size_t)gridDim.x * K5_THREADS) {
        const int c = (int)(i % cv);
        const size_t pix = i / cv;
        if (c < ca_vec) {
            const int x = (int)(pix % W), y = (int)((pix / W) % H);
            const size_t n = pix / ((size_t)W * H);
            out[i] = __ldg(a + ((n * h + (y >> 1)) * w + (x >> 1)) * ca_vec + c);
        } else {
            out[i] = __ldg(b + pix * cb_vec + (c - ca_vec));
        }
    }
}

}  // namespace fsd

using namespace fsd;

extern "C" int fsd_upsample2x_concat(fsd_handle_t h, const void* a, const void* b, void* out, int N, int ah, int aw,
                                     int ca, int cb, int dtype, void* stream_) {
    FSD_CHECK_ARG(h && a && b && out, "fsd_upsample2x_concat: null argument");
    FSD_CHECK_ARG(dtype == FSD_F16 || dtype == FSD_F32, "fsd_upsample2x_concat: bad dtype");
    const int per_vec = dtype == FSD_F16 ? 8 : 4;
    FSD_CHECK_ARG(N >= 0 && ah > 0 && aw > 0 && ca > 0 && cb > 0 && ca % per_vec == 0 && cb % per_vec == 0,
                  "fsd_upsample2x_concat: channel counts must be positive multiples of %d", per_vec);
    if (((uintptr_t)a & 15) || ((uintptr_t)b & 15) || ((uintptr_t)out & 15)) { set_error("fsd_upsample2x_concat: pointers must be 16-byte aligned"); return FSD_ERR_ALIGN; }
    if (N == 0) return FSD_OK;
    const size_t total = (size_t)N * 4 * ah * aw * ((ca + cb) / per_vec);
    const size_t want = (total + K5_THREADS - 1) / K5_THREADS;
    const int grid = (int)(want < (size_t)h->sm_count * 32 ? want : (size_t)h->sm_count * 32);
    FSD_CUDA(cudaSetDevice(h->device));
    k5_upsample2x_concat_kernel<<<grid, K5_THREADS, 0, (cudaStream_t)stream_>>>((const uint4*)a, (const uint4*)b, (uint4*)out, N, ah, aw,
                                                                               ca / per_vec, cb / per_vec);
    FSD_CUDA(cudaGetLastError());
    h->launches += 1;
    return FSD_OK;
}


extern "C" int fsd_bias_act_inplace(fsd_handle_t h, void* x, const void* bias, int64_t n_pixels, int channels, int act,
                                    float slope, int dtype, void* stream_) {
    FSD_CHECK_ARG(h && x && bias, "fsd_bias_act_inplace: null argument");
    FSD_CHECK_ARG(n_pixels >= 0 && channels > 0 && act >= 0 && act <= 2, "fsd_bias_act_inplace: bad sizes / activation");
    FSD_CHECK_ARG(dtype == FSD_F16 || dtype == FSD_F32, "fsd_bias_act_inplace: bad dtype");
    const int per_vec = dtype == FSD_F16 ? 8 : 4;
    FSD_CHECK_ARG(channels % per_vec == 0, "fsd_bias_act_inplace: channels (%d) must be a multiple of %d", channels, per_vec);
    if (((uintptr_t)x & 15) || ((uintptr_t)bias & 15)) { set_error("fsd_bias_act_inplace: pointers must be 16-byte aligned"); return FSD_ERR_ALIGN; }
    if (n_pixels == 0) return FSD_OK;
    const size_t n_vec = (size_t)n_pixels * channels / per_vec;
    const int c_vec = channels / per_vec;
    const size_t want = (n_vec + K5_THREADS - 1) / K5_THREADS;
    const int grid = (int)(want < (size_t)h->sm_count * 16 ? want : (size_t)h->sm_count * 16);  // persistent-ish grid-stride
    cudaStream_t s = (cudaStream_t)stream_;
    FSD_CUDA(cudaSetDevice(h->device));
#define LAUNCH(K, T) \
    if (act == 0) K<0><<<grid, K5_THREADS, 0, s>>>((T*)x, (const T*)bias, n_vec, c_vec, slope); \
    else if (act == 1) K<1><<<grid, K5_THREADS, 0, s>>>((T*)x, (const T*)bias, n_vec, c_vec, slope); \
    else K<2><<<grid, K5_THREADS, 0, s>>>((T*)x, (const T*)bias, n_vec, c_vec, slope);
    if (dtype == FSD_F16) { LAUNCH(k5_bias_act_half_kernel, uint4) } else { LAUNCH(k5_bias_act_float_kernel, float4) }
#undef LAUNCH
    FSD_CUDA(cudaGetLastError());
    h->launches += 1;
    return FSD_OK;
}
