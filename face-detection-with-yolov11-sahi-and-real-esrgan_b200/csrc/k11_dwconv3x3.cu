// Kernel 11: depth-wise 3x3 convolution (stride 1, pad 1) + bias + activation -> channel slot, channels-last fp16.
//
// ultralytics DWConv(c, c, 3) in the YOLO11 head's cv3 branches and C2PSA's positional convolution (run from utils/yolo_wrapper.py:72).
// cuDNN's grouped convolution + the epilogue pass reach 0.13-0.33 of the HBM peak on these layers (two passes over the activations, a
// generic grouped kernel); a depth-wise convolution has 9 MACs per element, i.e. it is a stencil that should move at memory speed.
//
// One thread = two neighbouring pixels x 8 channels (see the kernel): fp32 accumulation in tap order ky, kx (the order of a direct
// convolution), bias + activation, 16-byte stores into the destination slot.  Out-of-image taps contribute zero (the padding).
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

__device__ __forceinline__ void k11_unpack(const uint4& v, float (&f)[8]) {
    const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float2 t = __half22float2(h[k]);
        f[2 * k] = t.x; f[2 * k + 1] = t.y;
    }
}

// A CTA stages the weights (fp16, tap-major) once and walks image rows (row = image * H + y); a thread owns one channel vector and TWO
// neighbouring pixels of the row: per filter row it reads three 16-byte weight vectors from shared memory (the eight lanes of a
// quarter-warp read 128 contiguous bytes: conflict-free; the first version's 18 fp32 reads per pixel were its bound) and four input
// vectors (columns x0-1 .. x0+2, zero outside the image), i.e. 6 instead of 9 global loads and 4.5 instead of 18 shared loads per
// output vector.  Control flow is uniform (out-of-image taps are zero vectors, not branches), index arithmetic 32-bit.
template <int ACT>
__global__ void __launch_bounds__(K11_THREADS) k11_dwconv3x3_kernel(const K11Params p) {
    extern __shared__ __align__(16) uint8_t k11_raw[];
    uint4* ws = reinterpret_cast<uint4*>(k11_raw);  // [9][cv] vectors of 8 halves
    float* bs = reinterpret_cast<float*>(k11_raw + (size_t)9 * p.C * 2);
    for (int i = threadIdx.x; i < 9 * p.cv; i += K11_THREADS) ws[i] = __ldg(reinterpret_cast<const uint4*>(p.w) + i);
    for (int i = threadIdx.x; i < p.C; i += K11_THREADS) bs[i] = __half2float(__ldg(p.bias + i));
    __syncthreads();
    const int pairs = (p.W + 1) >> 1;
    const int row_items = pairs * p.cv;
    const uint4 zero = make_uint4(0, 0, 0, 0);
    // a CTA owns a band of consecutive rows, so two of the three input rows of every output row are already in this SM's L1
    const int band = (p.rows + gridDim.x - 1) / gridDim.x;
    const int r_end = min(p.rows, ((int)blockIdx.x + 1) * band);
    for (int r = blockIdx.x * band; r < r_end; ++r) {
        const int y = r % p.H;
        const uint4* xrow = p.x + (size_t)r * p.W * p.x_stride;
        uint4* orow = p.out + (size_t)r * p.W * p.out_stride;
        for (int t = threadIdx.x; t < row_items; t += K11_THREADS) {
            const int pr = t / p.cv, c = t - pr * p.cv;
            const int x0 = 2 * pr;
            float acc0[8], acc1[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) { acc0[e] = 0.f; acc1[e] = 0.f; }
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
                const int yy = y + ky - 1;
                const bool row_ok = yy >= 0 && yy < p.H;
                const uint4* src = xrow + (ky - 1) * p.W * p.x_stride + c;
                uint4 v[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int xx = x0 - 1 + j;
                    const bool ok = row_ok && xx >= 0 && xx < p.W;
                    v[j] = ok ? __ldg(src + xx * p.x_stride) : zero;
                }
                float f[4][8];
#pragma unroll
                for (int j = 0; j < 4; ++j) k11_unpack(v[j], f[j]);
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    float w[8];
                    k11_unpack(ws[(ky * 3 + kx) * p.cv + c], w);
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        acc0[e] = fmaf(f[kx][e], w[e], acc0[e]);
                        acc1[e] = fmaf(f[kx + 1][e], w[e], acc1[e]);
                    }
                }
            }
            const float4 b0 = *reinterpret_cast<const float4*>(bs + 8 * c), b1 = *reinterpret_cast<const float4*>(bs + 8 * c + 4);
            const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
            uint32_t o0[4], o1[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const __half2 h0 = __floats2half2_rn(k11_act<ACT>(acc0[2 * k] + bb[2 * k], p.slope), k11_act<ACT>(acc0[2 * k + 1] + bb[2 * k + 1], p.slope));
                const __half2 h1 = __floats2half2_rn(k11_act<ACT>(acc1[2 * k] + bb[2 * k], p.slope), k11_act<ACT>(acc1[2 * k + 1] + bb[2 * k + 1], p.slope));
                o0[k] = *reinterpret_cast<const uint32_t*>(&h0);
                o1[k] = *reinterpret_cast<const uint32_t*>(&h1);
            }
            orow[x0 * p.out_stride + c] = make_uint4(o0[0], o0[1], o0[2], o0[3]);
            if (x0 + 1 < p.W) orow[(x0 + 1) * p.out_stride + c] = make_uint4(o1[0], o1[1], o1[2], o1[3]);
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
    if (((uintptr_t)x & 15) || ((uintptr_t)out & 15) || ((uintptr_t)weight_taps & 15) || ((uintptr_t)bias & 1)) {
        set_error("fsd_dwconv3x3: x / out / weight_taps must be 16-byte aligned");
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
    const size_t smem = (size_t)9 * channels * 2 + (size_t)channels * sizeof(float);
    int grid = h->sm_count * 3;  // 3 resident CTAs of 256 threads per SM (78 registers), each walking a band of rows
    if (p.rows < grid * 4) grid = (p.rows + 3) / 4;
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
