// Kernel 11: depth-wise 3x3 convolution (stride 1, pad 1) + bias + activation -> channel slot, channels-last fp16.
//
// ultralytics DWConv(c, c, 3) in the YOLO11 head's cv3 branches and C2PSA's positional convolution (run from utils/yolo_wrapper.py:72).
// cuDNN's grouped convolution + the epilogue pass reach 0.13-0.33 of the HBM peak on these layers (two passes over the activations, a
// generic grouped kernel); a depth-wise convolution has 9 MACs per element, i.e. it is a stencil that should move at memory speed.
//
// One thread = one pixel x 8 channels: nine 16-byte neighbour loads (a warp covers 512 contiguous bytes of each of the nine pixel rows, so
// eight of the nine come out of L1/L2), weights and bias as fp32 in shared memory (staged once per CTA, indexed by the channel vector:
// conflict-free), fp32 accumulation in tap order ky, kx (the order of a direct convolution), bias + activation, one 16-byte store into the
// destination slot.  Out-of-image taps contribute zero (the padding).
#include "fsd_common.cuh"

namespace fsd {

constexpr int K11_THREADS = 256;

struct K11Params {
    const uint4* x;
    const __half* w;     // tap-major [9][C]
    const __half* bias;  // [C]
    uint4* out;
    int rows;            // n_images * H
    int H, W, C, cv;     // cv = C / 8
    int x_stride, out_stride;  // pixel strides in 16-byte vectors
    float slope;
};

template <int ACT>
__device__ __forceinline__ float k11_act(float v, float slope) {
    if (ACT == 1) return fast_silu(v);
    if (ACT == 2) return v > 0.f ? v : v * slope;
    return v;
}

// A CTA stages the weights once and walks image rows (row = image * H + y); its threads walk the row's (pixel, channel-vector) pairs.
// All index arithmetic is 32-bit (the first version's three 64-bit divisions per thread cost more than the convolution).
template <int ACT>
__global__ void __launch_bounds__(K11_THREADS) k11_dwconv3x3_kernel(const K11Params p) {
    extern __shared__ float k11_smem[];  // [9][C] weights, then [C] bias
    float* ws = k11_smem;
    float* bs = k11_smem + 9 * p.C;
    for (int i = threadIdx.x; i < 9 * p.C; i += K11_THREADS) ws[i] = __half2float(__ldg(p.w + i));
    for (int i = threadIdx.x; i < p.C; i += K11_THREADS) bs[i] = __half2float(__ldg(p.bias + i));
    __syncthreads();
    const int row_vecs = p.W * p.cv;
    for (int r = blockIdx.x; r < p.rows; r += gridDim.x) {
        const int y = r % p.H;
        const bool has_up = y > 0, has_down = y + 1 < p.H;
        const uint4* xrow = p.x + (size_t)r * p.W * p.x_stride;
        uint4* orow = p.out + (size_t)r * p.W * p.out_stride;
        for (int t = threadIdx.x; t < row_vecs; t += K11_THREADS) {
            const int x = t / p.cv, c = t - x * p.cv;
            float acc[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[e] = 0.f;
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
                if ((ky == 0 && !has_up) || (ky == 2 && !has_down)) continue;
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    const int xx = x + kx - 1;
                    if (xx < 0 || xx >= p.W) continue;
                    const uint4 v = __ldg(xrow + ((ky - 1) * p.W + xx) * p.x_stride + c);
                    const float4 w0 = *reinterpret_cast<const float4*>(ws + (ky * 3 + kx) * p.C + 8 * c);
                    const float4 w1 = *reinterpret_cast<const float4*>(ws + (ky * 3 + kx) * p.C + 8 * c + 4);
                    const __half2* hv = reinterpret_cast<const __half2*>(&v);
                    const float2 f0 = __half22float2(hv[0]), f1 = __half22float2(hv[1]), f2 = __half22float2(hv[2]), f3 = __half22float2(hv[3]);
                    acc[0] = fmaf(f0.x, w0.x, acc[0]); acc[1] = fmaf(f0.y, w0.y, acc[1]);
                    acc[2] = fmaf(f1.x, w0.z, acc[2]); acc[3] = fmaf(f1.y, w0.w, acc[3]);
                    acc[4] = fmaf(f2.x, w1.x, acc[4]); acc[5] = fmaf(f2.y, w1.y, acc[5]);
                    acc[6] = fmaf(f3.x, w1.z, acc[6]); acc[7] = fmaf(f3.y, w1.w, acc[7]);
                }
            }
            const float4 b0 = *reinterpret_cast<const float4*>(bs + 8 * c), b1 = *reinterpret_cast<const float4*>(bs + 8 * c + 4);
            const __half2 o0 = __floats2half2_rn(k11_act<ACT>(acc[0] + b0.x, p.slope), k11_act<ACT>(acc[1] + b0.y, p.slope));
            const __half2 o1 = __floats2half2_rn(k11_act<ACT>(acc[2] + b0.z, p.slope), k11_act<ACT>(acc[3] + b0.w, p.slope));
            const __half2 o2 = __floats2half2_rn(k11_act<ACT>(acc[4] + b1.x, p.slope), k11_act<ACT>(acc[5] + b1.y, p.slope));
            const __half2 o3 = __floats2half2_rn(k11_act<ACT>(acc[6] + b1.z, p.slope), k11_act<ACT>(acc[7] + b1.w, p.slope));
            uint4 o;
            o.x = *reinterpret_cast<const uint32_t*>(&o0); o.y = *reinterpret_cast<const uint32_t*>(&o1);
            o.z = *reinterpret_cast<const uint32_t*>(&o2); o.w = *reinterpret_cast<const uint32_t*>(&o3);
            orow[x * p.out_stride + c] = o;
        }
    }
}

}  // namespace fsd

using namespace fsd;

extern "C" int fsd_dwconv3x3(fsd_handle_t h, const void* x, int64_t x_pixel_stride, int n_images, int H, int W, const void* weight_taps,
                             const void* bias, void* out, int64_t out_pixel_stride, int channels, int act, float slope, int dtype,
                             void* stream_) {
    FSD_CHECK_ARG(h && x && weight_taps && bias && out, "fsd_dwconv3x3: null argument");
    FSD_CHECK_ARG(dtype == FSD_F16, "fsd_dwconv3x3: only fp16 is implemented");
    FSD_CHECK_ARG(n_images >= 0 && H > 0 && W > 0 && act >= 0 && act <= 2, "fsd_dwconv3x3: bad sizes / activation");
    FSD_CHECK_ARG(channels >= 8 && channels % 8 == 0 && channels <= 1024, "fsd_dwconv3x3: channels must be a multiple of 8 in [8, 1024]");
    FSD_CHECK_ARG(x_pixel_stride >= channels && x_pixel_stride % 8 == 0, "fsd_dwconv3x3: bad input stride");
    FSD_CHECK_ARG(out_pixel_stride >= channels && out_pixel_stride % 8 == 0, "fsd_dwconv3x3: bad output stride");
    if (((uintptr_t)x & 15) || ((uintptr_t)out & 15) || ((uintptr_t)weight_taps & 1) || ((uintptr_t)bias & 1)) {
        set_error("fsd_dwconv3x3: x / out must be 16-byte aligned");
        return FSD_ERR_ALIGN;
    }
    if (n_images == 0) return FSD_OK;
    K11Params p;
    p.x = (const uint4*)x; p.w = (const __half*)weight_taps; p.bias = (const __half*)bias; p.out = (uint4*)out;
    p.H = H; p.W = W; p.C = channels; p.cv = channels / 8;
    FSD_CHECK_ARG((int64_t)n_images * H < (1LL << 31) && (int64_t)W * (x_pixel_stride > out_pixel_stride ? x_pixel_stride : out_pixel_stride) < (1LL << 31),
                  "fsd_dwconv3x3: tensor too large for one launch");
    p.rows = n_images * H;
    p.x_stride = (int)(x_pixel_stride / 8); p.out_stride = (int)(out_pixel_stride / 8); p.slope = slope;
    const size_t smem = (size_t)10 * channels * sizeof(float);
    const int grid = p.rows < h->sm_count * 8 ? p.rows : h->sm_count * 8;  // 8 resident CTAs of 256 threads per SM
    cudaStream_t s = (cudaStream_t)stream_;
    FSD_CUDA(cudaSetDevice(h->device));
    {
        // algorithmic bytes: input + output once
        TimedLaunch timed(h, FSD_KERNEL_DWCONV, (int64_t)n_images * H * W * channels * 4, channels, s);
        if (act == 0) k11_dwconv3x3_kernel<0><<<grid, K11_THREADS, smem, s>>>(p);
        else if (act == 1) k11_dwconv3x3_kernel<1><<<grid, K11_THREADS, smem, s>>>(p);
        else k11_dwconv3x3_kernel<2><<<grid, K11_THREADS, smem, s>>>(p);
    }
    FSD_CUDA(cudaGetLastError());
    h->launches += 1;
    return FSD_OK;
}
