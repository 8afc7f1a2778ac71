// Fused conv epilogue for the PyTorch backbones: y = act(x + bias[c]) in place, one pass over HBM.
// cuDNN (through torch) produces the raw convolution; eager PyTorch would then launch a broadcast bias-add and a
// SiLU / LeakyReLU kernel (two more read+write passes, the bias-add non-vectorised).  The ncu launch list of the
// sliced-detection step showed those two passes costing ~3x the convolutions themselves (profiles/r1_launches_*).
// Layout: channels-last ([N,H,W,C] dense, C % 8 == 0 for fp16 / C % 4 == 0 for fp32) so the bias index is idx % C.
#include <algorithm>

#include "fsd_common.cuh"

namespace fsd {

constexpr int K5_THREADS = 256;

template <int ACT> __device__ __forceinline__ float activate(float v, float slope) {
    if (ACT == 3) return tanh_silu(v);  // SiLU, one-MUFU form (FSD_SILU=tanh)
    if (ACT == 1) return fast_silu(v);  // SiLU
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

// General epilogue: out[pix, c] = act(x[pix, c] + bias[c]) (+ res[pix, c]), where `out` and `res` may be channel slots of
// wider channels-last buffers (own pixel stride), and channels >= out2_c0 are additionally copied to a second
// destination.  This lets the conv epilogue write straight into the concat buffer of C3k2 / SPPF / C2PSA (torch.cat was
// 14.5 % of the step in profiles/r1_launches_bench_b32_final.txt) and folds the bottleneck's residual add.
struct K5GenArgs {
    const uint4* x; const uint4* bias; uint4* out; const uint4* res; uint4* out2;
    unsigned n_vec; int c_vec; int out_stride; int res_stride; int out2_stride; int out2_c0; float slope;
    uint4* up; int up_stride; int H, W;  // optional nearest-2x up-sampled destination ([N, 2H, 2W] pixels, own stride)
};

// one 8-channel vector: bias + activation (+ residual) -> out slot (+ second destination, + 2x up-sampled copies)
template <int ACT>
__device__ __forceinline__ void k5_apply(const K5GenArgs& a, unsigned pix, int c, uint4 v, uint4 r) {
    const uint4 b = __ldg(a.bias + c);
    __half2* hv = reinterpret_cast<__half2*>(&v);
    const __half2* hb = reinterpret_cast<const __half2*>(&b);
    const __half2* hr = reinterpret_cast<const __half2*>(&r);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float2 f = __half22float2(hv[k]);
        const float2 g = __half22float2(hb[k]);
        __half2 o = __floats2half2_rn(activate<ACT>(f.x + g.x, a.slope), activate<ACT>(f.y + g.y, a.slope));
        // the residual is added in fp16 after the activation is rounded, exactly as torch's `x + act(conv)` does
        if (a.res) o = __hadd2(o, hr[k]);
        hv[k] = o;
    }
    a.out[(size_t)pix * a.out_stride + c] = v;
    if (a.out2 && c >= a.out2_c0) a.out2[(size_t)pix * a.out2_stride + (c - a.out2_c0)] = v;
    if (a.up) {  // FPN: the next stage concatenates the nearest-2x up-sampled map — store the four copies here
        const unsigned hw = (unsigned)(a.H * a.W);
        const unsigned n = pix / hw, rem = pix - n * hw;
        const unsigned y = rem / (unsigned)a.W, x = rem - y * (unsigned)a.W;
        uint4* d = a.up + (((size_t)n * 2 * a.H + 2 * y) * 2 * a.W + 2 * x) * a.up_stride + c;
        const size_t row = (size_t)2 * a.W * a.up_stride;
        d[0] = v; d[a.up_stride] = v; d[row] = v; d[row + a.up_stride] = v;
    }
}

// Two independent vectors per thread and trip: with one 16-byte load in flight per thread a fully occupied SM holds 32 KB in flight,
// short of the ~46 KB that 23 B/clk/SM x ~2000 cycles of loaded-HBM latency ask for (top stall: long scoreboard, r1_k5_bias_act.summary)
template <int ACT>
__global__ void __launch_bounds__(K5_THREADS, 2048 / K5_THREADS * 3 / 4) k5_bias_act_general_half_kernel(const K5GenArgs a) {
    const unsigned stride = gridDim.x * K5_THREADS;
    unsigned i = blockIdx.x * K5_THREADS + threadIdx.x;
    const uint4 zero = make_uint4(0, 0, 0, 0);
    for (; i + stride < a.n_vec; i += 2 * stride) {  // (the host keeps n_vec below 2^32 - 2^28: no wrap)
        const unsigned j = i + stride;
        const unsigned pix0 = i / (unsigned)a.c_vec, pix1 = j / (unsigned)a.c_vec;
        const int c0 = (int)(i - pix0 * (unsigned)a.c_vec), c1 = (int)(j - pix1 * (unsigned)a.c_vec);
        const uint4 v0 = __ldcs(a.x + i), v1 = __ldcs(a.x + j);
        uint4 r0 = zero, r1 = zero;
        if (a.res) {
            r0 = __ldg(a.res + (size_t)pix0 * a.res_stride + c0);
            r1 = __ldg(a.res + (size_t)pix1 * a.res_stride + c1);
        }
        k5_apply<ACT>(a, pix0, c0, v0, r0);
        k5_apply<ACT>(a, pix1, c1, v1, r1);
    }
    for (; i < a.n_vec; i += stride) {
        const unsigned pix = i / (unsigned)a.c_vec;
        const int c = (int)(i - pix * (unsigned)a.c_vec);
        const uint4 v = __ldcs(a.x + i);
        const uint4 r = a.res ? __ldg(a.res + (size_t)pix * a.res_stride + c) : zero;
        k5_apply<ACT>(a, pix, c, v, r);
    }
}

// SPPF pooling: ultralytics SPPF = cat(y, m(y), m(m(y)), m(m(m(y)))) with m = MaxPool2d(5, 1, 2).  The input is channel
// slot 0 of the [N,H,W,4c] concat buffer; this kernel fills slots 1..3.  One CTA owns one image x one 8-channel vector:
// the H*W map lives in shared memory and each 5x5 max is done separably (row pass, column pass); borders clip exactly
// like -inf padding.  torch ran 3 max_pool launches + a concat (5.4 % + part of 14.5 % of the step).
__device__ __forceinline__ uint4 hmax8(uint4 a, uint4 b) {
    __half2* x = reinterpret_cast<__half2*>(&a);
    const __half2* y = reinterpret_cast<const __half2*>(&b);
#pragma unroll
    for (int k = 0; k < 4; ++k) x[k] = __hmax2_nan(x[k], y[k]);
    return a;
}

__global__ void __launch_bounds__(K5_THREADS) k5_sppf_pool_kernel(uint4* __restrict__ buf, int H, int W, int c_vec) {
    extern __shared__ uint4 k5_smem[];
    const int P = H * W;
    uint4* A = k5_smem;      // current map
    uint4* T = k5_smem + P;  // row-max of A
    const int cv = blockIdx.x % c_vec;
    const size_t n = blockIdx.x / c_vec;
    const int ct = 4 * c_vec;  // pixel stride of the concat buffer in vectors
    uint4* base = buf + n * (size_t)P * ct + cv;
    for (int p = threadIdx.x; p < P; p += K5_THREADS) A[p] = base[(size_t)p * ct];
    __syncthreads();
    for (int rep = 1; rep <= 3; ++rep) {
        for (int p = threadIdx.x; p < P; p += K5_THREADS) {
            const int x = p % W;
            uint4 m = A[p];
            if (x >= 1) m = hmax8(m, A[p - 1]);
            if (x >= 2) m = hmax8(m, A[p - 2]);
            if (x + 1 < W) m = hmax8(m, A[p + 1]);
            if (x + 2 < W) m = hmax8(m, A[p + 2]);
            T[p] = m;
        }
        __syncthreads();
        for (int p = threadIdx.x; p < P; p += K5_THREADS) {
            const int y = p / W;
            uint4 m = T[p];
            if (y >= 1) m = hmax8(m, T[p - W]);
            if (y >= 2) m = hmax8(m, T[p - 2 * W]);
            if (y + 1 < H) m = hmax8(m, T[p + W]);
            if (y + 2 < H) m = hmax8(m, T[p + 2 * W]);
            A[p] = m;  // safe: the column pass reads T only
            base[(size_t)p * ct + rep * c_vec] = m;
        }
        __syncthreads();
    }
}

// out[n,y,x,:] = concat(a[n,y/2,x/2,:], b[n,y,x,:]) — the FPN "nearest 2x up-sample, then concat" of the YOLO neck in ONE
// pass (torch runs an up-sample kernel, then a concat kernel: the up-sampled tensor is written and read once more).
__global__ void __launch_bounds__(K5_THREADS)
k5_upsample2x_concat_kernel(const uint4* __restrict__ a, const uint4* __restrict__ b, uint4* __restrict__ out,
                            int N, int h, int w, int ca_vec, int cb_vec) {
    // blockIdx.y = output row (n, y); threads run over the row's W * (ca+cb)/8 vectors: one 32-bit division per thread
    const int H = 2 * h, W = 2 * w, cv = ca_vec + cb_vec;
    for (unsigned row = blockIdx.y; row < (unsigned)N * H; row += gridDim.y) {  // row = n * H + y
        const unsigned n = row / (unsigned)H, y = row - n * (unsigned)H;
        const uint4* arow = a + ((size_t)n * h + (y >> 1)) * w * ca_vec;
        const uint4* brow = b + (size_t)row * W * cb_vec;
        uint4* orow = out + (size_t)row * W * cv;
        for (unsigned i = blockIdx.x * K5_THREADS + threadIdx.x; i < (unsigned)(W * cv); i += gridDim.x * K5_THREADS) {
            const unsigned x = i / (unsigned)cv, c = i - x * (unsigned)cv;
            orow[i] = c < (unsigned)ca_vec ? __ldg(arow + (x >> 1) * ca_vec + c) : __ldcs(brow + x * cb_vec + (c - ca_vec));
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
    FSD_CHECK_ARG((int64_t)N * 2 * ah < (1ll << 31), "fsd_upsample2x_concat: too many rows for one launch");
    const int row_vecs = 2 * aw * ((ca + cb) / per_vec);
    const int rows = N * 2 * ah;
    FSD_CUDA(cudaSetDevice(h->device));
    {
        dim3 grid(std::min((row_vecs + K5_THREADS - 1) / K5_THREADS, 8), std::min(rows, 65535));
        k5_upsample2x_concat_kernel<<<grid, K5_THREADS, 0, (cudaStream_t)stream_>>>((const uint4*)a, (const uint4*)b, (uint4*)out, N, ah, aw,
                                                                                   ca / per_vec, cb / per_vec);
    }
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
    else if (act == 1 && silu_tanh_mode()) K<3><<<grid, K5_THREADS, 0, s>>>((T*)x, (const T*)bias, n_vec, c_vec, slope); \
    else if (act == 1) K<1><<<grid, K5_THREADS, 0, s>>>((T*)x, (const T*)bias, n_vec, c_vec, slope); \
    else K<2><<<grid, K5_THREADS, 0, s>>>((T*)x, (const T*)bias, n_vec, c_vec, slope);
    TimedLaunch timed(h, FSD_KERNEL_BIAS_ACT, (int64_t)n_vec * 32, channels, s);  // read + write once
    if (dtype == FSD_F16) { LAUNCH(k5_bias_act_half_kernel, uint4) } else { LAUNCH(k5_bias_act_float_kernel, float4) }
#undef LAUNCH
    FSD_CUDA(cudaGetLastError());
    h->launches += 1;
    return FSD_OK;
}

extern "C" int fsd_bias_act(fsd_handle_t h, const void* x, const void* bias, void* out, int64_t out_pixel_stride,
                            const void* residual, int64_t residual_pixel_stride, void* out2, int64_t out2_pixel_stride,
                            int out2_first_channel, void* up2x, int64_t up2x_pixel_stride, int height, int width,
                            int64_t n_pixels, int channels, int act, float slope, int dtype, void* stream_) {
    FSD_CHECK_ARG(h && x && bias && out, "fsd_bias_act: null argument");
    FSD_CHECK_ARG(dtype == FSD_F16, "fsd_bias_act: only fp16 is implemented (use fsd_bias_act_inplace for fp32)");
    FSD_CHECK_ARG(n_pixels >= 0 && channels > 0 && channels % 8 == 0 && act >= 0 && act <= 2, "fsd_bias_act: bad sizes / activation");
    FSD_CHECK_ARG(out_pixel_stride >= channels && out_pixel_stride % 8 == 0, "fsd_bias_act: out pixel stride must be >= channels and a multiple of 8");
    FSD_CHECK_ARG(!residual || (residual_pixel_stride >= channels && residual_pixel_stride % 8 == 0), "fsd_bias_act: bad residual stride");
    FSD_CHECK_ARG(!out2 || (out2_first_channel >= 0 && out2_first_channel < channels && out2_first_channel % 8 == 0 &&
                            out2_pixel_stride >= channels - out2_first_channel && out2_pixel_stride % 8 == 0),
                  "fsd_bias_act: bad second destination");
    FSD_CHECK_ARG(!up2x || (height > 0 && width > 0 && n_pixels % ((int64_t)height * width) == 0 && up2x_pixel_stride >= channels &&
                            up2x_pixel_stride % 8 == 0),
                  "fsd_bias_act: the up-sampled destination needs height * width dividing n_pixels and a stride >= channels");
    if (((uintptr_t)x & 15) || ((uintptr_t)bias & 15) || ((uintptr_t)out & 15) || ((uintptr_t)residual & 15) || ((uintptr_t)out2 & 15) ||
        ((uintptr_t)up2x & 15)) {
        set_error("fsd_bias_act: pointers must be 16-byte aligned");
        return FSD_ERR_ALIGN;
    }
    if (n_pixels == 0) return FSD_OK;
    const size_t n_vec = (size_t)n_pixels * channels / 8;
    FSD_CHECK_ARG(n_vec < 0xf0000000ull, "fsd_bias_act: tensor too large for one launch (%zu vectors)", n_vec);
    K5GenArgs a;
    a.x = (const uint4*)x; a.bias = (const uint4*)bias; a.out = (uint4*)out; a.res = (const uint4*)residual; a.out2 = (uint4*)out2;
    a.n_vec = (unsigned)n_vec; a.c_vec = channels / 8; a.out_stride = (int)(out_pixel_stride / 8);
    a.res_stride = (int)(residual_pixel_stride / 8); a.out2_stride = (int)(out2_pixel_stride / 8);
    a.out2_c0 = out2_first_channel / 8; a.slope = slope;
    a.up = (uint4*)up2x; a.up_stride = (int)(up2x_pixel_stride / 8); a.H = height; a.W = width;
    const size_t want = (n_vec + K5_THREADS - 1) / K5_THREADS;
    const int grid = (int)(want < (size_t)h->sm_count * 16 ? want : (size_t)h->sm_count * 16);
    cudaStream_t s = (cudaStream_t)stream_;
    FSD_CUDA(cudaSetDevice(h->device));
    {
        // units = bytes this launch moves (read x [+ residual], write out [+ out2])
        const int64_t bytes = (int64_t)n_pixels * 2 * ((residual ? 3 : 2) * channels + (out2 ? channels - out2_first_channel : 0) + (up2x ? 4 * channels : 0));
        TimedLaunch timed(h, FSD_KERNEL_BIAS_ACT, bytes, channels, s);
        if (act == 0) k5_bias_act_general_half_kernel<0><<<grid, K5_THREADS, 0, s>>>(a);
        else if (act == 1 && silu_tanh_mode()) k5_bias_act_general_half_kernel<3><<<grid, K5_THREADS, 0, s>>>(a);
        else if (act == 1) k5_bias_act_general_half_kernel<1><<<grid, K5_THREADS, 0, s>>>(a);
        else k5_bias_act_general_half_kernel<2><<<grid, K5_THREADS, 0, s>>>(a);
    }
    FSD_CUDA(cudaGetLastError());
    h->launches += 1;
    return FSD_OK;
}

extern "C" int fsd_sppf_pool(fsd_handle_t h, void* buf, int N, int H, int W, int c, int dtype, void* stream_) {
    FSD_CHECK_ARG(h && buf, "fsd_sppf_pool: null argument");
    FSD_CHECK_ARG(dtype == FSD_F16, "fsd_sppf_pool: only fp16 is implemented");
    FSD_CHECK_ARG(N >= 0 && H > 0 && W > 0 && c > 0 && c % 8 == 0, "fsd_sppf_pool: bad sizes (c must be a multiple of 8)");
    if ((uintptr_t)buf & 15) { set_error("fsd_sppf_pool: pointer must be 16-byte aligned"); return FSD_ERR_ALIGN; }
    const size_t smem = (size_t)2 * H * W * sizeof(uint4);
    FSD_CHECK_ARG(smem <= 200 * 1024, "fsd_sppf_pool: map of %d x %d does not fit shared memory", H, W);
    if (N == 0) return FSD_OK;
    FSD_CUDA(cudaSetDevice(h->device));
    if (smem > 48 * 1024) FSD_CUDA(cudaFuncSetAttribute(k5_sppf_pool_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    TimedLaunch timed(h, FSD_KERNEL_SPPF, (int64_t)N * H * W * c * 4 * 2, c, (cudaStream_t)stream_);  // read slot 0, write slots 1..3
    k5_sppf_pool_kernel<<<N * (c / 8), K5_THREADS, smem, (cudaStream_t)stream_>>>((uint4*)buf, H, W, c / 8);
    FSD_CUDA(cudaGetLastError());
    h->launches += 1;
    return FSD_OK;
}
