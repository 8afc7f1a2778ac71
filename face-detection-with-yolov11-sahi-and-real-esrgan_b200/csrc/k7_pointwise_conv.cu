// Kernel 7 — 1x1 convolution + bias + activation (+ residual) straight into a concat slot, for the LOW-INTENSITY layers.
//
// A 1x1 convolution over a channels-last tensor is the GEMM out[P, N] = x[P, K] * W[N, K]^T.  For the high-resolution
// layers of YOLO11n (K, N <= 192/128: C3k2.cv1/cv2 of the 256^2 and 128^2 stages, the P3 neck and head) it performs
// 16-77 flop per byte: it is HBM-bound, and the cuDNN convolution + fsd_bias_act pair moves the output three times
// (conv writes it raw, the epilogue reads and re-writes it).  This kernel moves it once:
//   * the whole weight matrix (<= 64 KB) sits in shared memory for the lifetime of a persistent CTA;
//   * each warp owns 32-pixel tiles: cp.async brings the 32 x K activations into the warp's private staging rows,
//     ldmatrix + mma.sync.m16n8k16 (f16 x f16 -> f32) produce the 32 x N accumulators — the tensor core only has to keep up
//     with HBM here, which is why the legacy warp-level MMA is enough and tcgen05/TMEM would buy nothing;
//   * the epilogue (+ bias, SiLU, fp16) goes back through the staging rows so that the final stores — and the residual
//     loads — are 16-byte, row-contiguous accesses into the destination slot (own pixel stride) and the optional second
//     destination.
// Layers with K > 128 or N > 128 stay on cuDNN's tcgen05 kernels, which are compute-efficient there.
#include "fsd_common.cuh"

namespace fsd {

constexpr int K7_THREADS = 256;
constexpr int K7_WARPS = K7_THREADS / 32;

struct K7Params {
    const __half* x; const __half* w; const __half* bias;
    __half* out; const __half* res; __half* out2;
    long long P;
    int K, x_stride, out_stride, res_stride, out2_stride, out2_c0, act;
    float slope;
};

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool pred) {
    const uint32_t s = smem_u32(smem);
    const int bytes = pred ? 16 : 0;  // src-size 0 -> zero fill
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "r"(bytes));
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_but_one() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }

__device__ __forceinline__ void ldsm4(uint32_t (&r)[4], const void* p) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_u32(p)));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <int ACT> __device__ __forceinline__ float k7_act(float v, float slope) {
    if (ACT == 1) return fast_silu(v);
    if (ACT == 2) return v > 0.f ? v : v * slope;
    return v;
}

// MT = 16-pixel m-tiles per warp tile.  MT = 2 halves the weight-fragment traffic per pixel; MT = 1 halves the accumulator
// registers (N = 64: 128 -> ~80 per thread), which doubles the resident warps — tried because the N = 64 variant runs at 23 %
// occupancy (profiles/r1_k7_pointwise_48x64.summary.txt), but it was slower: the kernel is MUFU/issue-bound, not latency-bound.
template <int N, int ACT, int MT>
__global__ void __launch_bounds__(K7_THREADS, N <= 32 ? 3 : (N <= 64 ? 2 : 1)) k7_pointwise_conv_kernel(const K7Params p) {
    constexpr int K7_TILE = 16 * MT;  // pixels per warp tile
    extern __shared__ __align__(16) uint8_t k7_smem[];
    const int K = p.K;
    const int wpitch = K + 8;                       // halfs; +16 B keeps ldmatrix rows on distinct banks
    const int spitch = (K > N ? K : N) + 8;         // staging rows hold the activations, then the outputs
    __half* Ws = reinterpret_cast<__half*>(k7_smem);
    // two staging buffers per warp: the next tile's activations stream in (cp.async) while this tile is computed and stored
    __half* stage0 = Ws + (size_t)N * wpitch + (size_t)(threadIdx.x >> 5) * 2 * K7_TILE * spitch;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    // ---- weights: once per CTA --------------------------------------------------------------------------------------
    const int kchunks = K >> 3;  // 16-byte chunks per row
    for (int i = threadIdx.x; i < N * kchunks; i += K7_THREADS) {
        const int n = i / kchunks, c = i - n * kchunks;
        cp_async16(Ws + (size_t)n * wpitch + c * 8, p.w + (size_t)n * K + c * 8, true);
    }
    cp_async_wait_all();
    __syncthreads();
    const int g = lane >> 2, t = lane & 3;
    float bias[N / 8][2];
#pragma unroll
    for (int j = 0; j < N / 8; ++j) {
        bias[j][0] = __half2float(__ldg(p.bias + 8 * j + 2 * t));
        bias[j][1] = __half2float(__ldg(p.bias + 8 * j + 2 * t + 1));
    }

    const long long tiles = (p.P + K7_TILE - 1) / K7_TILE;
    const long long tile_step = (long long)gridDim.x * K7_WARPS;
    auto prefetch = [&](long long tile, __half* dst) {
        if (tile < tiles) {
            const long long pix0 = tile * K7_TILE;
            for (int i = lane; i < K7_TILE * kchunks; i += 32) {
                const int r = i / kchunks, c = i - r * kchunks;
                const bool ok = pix0 + r < p.P;  // zero-filled past P
                cp_async16(dst + (size_t)r * spitch + c * 8, p.x + (size_t)(ok ? pix0 + r : 0) * p.x_stride + c * 8, ok);
            }
        }
        cp_async_commit();  // (an empty group keeps the wait_group arithmetic uniform)
    };
    long long tile = (long long)blockIdx.x * K7_WARPS + warp;
    int cur = 0;
    prefetch(tile, stage0);
    for (; tile < tiles; tile += tile_step, cur ^= 1) {
        const long long pix0 = tile * K7_TILE;
        __half* stage = stage0 + (size_t)cur * K7_TILE * spitch;
        prefetch(tile + tile_step, stage0 + (size_t)(cur ^ 1) * K7_TILE * spitch);
        cp_async_wait_but_one();
        __syncwarp();

        float acc[MT][N / 8][4];
#pragma unroll
        for (int m = 0; m < MT; ++m)
#pragma unroll
            for (int j = 0; j < N / 8; ++j)
#pragma unroll
                for (int q = 0; q < 4; ++q) acc[m][j][q] = 0.f;
        for (int ks = 0; ks < K; ks += 16) {
            uint32_t a[MT][4];
#pragma unroll
            for (int m = 0; m < MT; ++m)  // lanes 0-15 -> rows, lanes 16-31 -> the k+8 half
                ldsm4(a[m], stage + (size_t)(16 * m + (lane & 15)) * spitch + ks + 8 * (lane >> 4));
#pragma unroll
            for (int jp = 0; jp < N / 16; ++jp) {
                // W rows n0..n0+15: matrices (n 0-7,k 0-7) (n 0-7,k 8-15) (n 8-15,k 0-7) (n 8-15,k 8-15)
                uint32_t b[4];
                ldsm4(b, Ws + (size_t)(16 * jp + (lane & 7) + 8 * (lane >> 4)) * wpitch + ks + 8 * ((lane >> 3) & 1));
#pragma unroll
                for (int m = 0; m < MT; ++m) {
                    mma16816(acc[m][2 * jp], a[m], b[0], b[1]);
                    mma16816(acc[m][2 * jp + 1], a[m], b[2], b[3]);
                }
            }
        }
        __syncwarp();  // every lane is done reading the activations: the staging rows now take the outputs
#pragma unroll
        for (int m = 0; m < MT; ++m)
#pragma unroll
            for (int j = 0; j < N / 8; ++j)
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                    const __half2 o = __floats2half2_rn(k7_act<ACT>(acc[m][j][2 * hh] + bias[j][0], p.slope),
                                                        k7_act<ACT>(acc[m][j][2 * hh + 1] + bias[j][1], p.slope));
                    *reinterpret_cast<__half2*>(stage + (size_t)(16 * m + g + 8 * hh) * spitch + 8 * j + 2 * t) = o;
                }
        __syncwarp();
        // ---- row-contiguous 16-byte stores (+ residual) ----------------------------------------------------------------
        constexpr int NCH = N / 8;
        for (int i = lane; i < K7_TILE * NCH; i += 32) {
            const int r = i / NCH, c = i - r * NCH;
            const long long pix = pix0 + r;
            if (pix >= p.P) continue;
            uint4 v = *reinterpret_cast<const uint4*>(stage + (size_t)r * spitch + c * 8);
            if (p.res) {
                const uint4 rr = __ldg(reinterpret_cast<const uint4*>(p.res + (size_t)pix * p.res_stride + c * 8));
                __half2* hv = reinterpret_cast<__half2*>(&v);
                const __half2* hr = reinterpret_cast<const __half2*>(&rr);
#pragma unroll
                for (int q = 0; q < 4; ++q) hv[q] = __hadd2(hv[q], hr[q]);
            }
            *reinterpret_cast<uint4*>(p.out + (size_t)pix * p.out_stride + c * 8) = v;
            if (p.out2 && c * 8 >= p.out2_c0) *reinterpret_cast<uint4*>(p.out2 + (size_t)pix * p.out2_stride + (c * 8 - p.out2_c0)) = v;
        }
        __syncwarp();  // the tile after next is prefetched into these rows
    }
}

template <int N>
static int k7_launch(fsd_context* h, const K7Params& p, cudaStream_t s) {
    constexpr int MT = 2;  // MT = 1 (half the accumulators, twice the resident warps) measured slower for N = 64: 0.50 vs 0.54 of peak
    constexpr int K7_TILE = 16 * MT;
    const int K = p.K;
    const size_t smem = ((size_t)N * (K + 8) + (size_t)2 * K7_WARPS * K7_TILE * ((K > N ? K : N) + 8)) * sizeof(__half);
    const long long tiles = (p.P + K7_TILE - 1) / K7_TILE;
    int per_sm = (int)((200 * 1024) / (smem + 1024));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 4) per_sm = 4;
    long long want = (tiles + K7_WARPS - 1) / K7_WARPS;
    const int grid = (int)(want < (long long)h->sm_count * per_sm ? want : (long long)h->sm_count * per_sm);
#define K7_GO(ACT)                                                                                                   \
    {                                                                                                                \
        auto kern = k7_pointwise_conv_kernel<N, ACT, MT>;                                                                \
        FSD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));               \
        kern<<<grid, K7_THREADS, smem, s>>>(p);                                                                      \
    }
    // algorithmic bytes: input + output (+ residual, + second destination) once
    TimedLaunch timed(h, FSD_KERNEL_POINTWISE, (int64_t)p.P * (K + N + (p.res ? N : 0) + (p.out2 ? N - p.out2_c0 : 0)) * 2, N, s);
    if (p.act == 0) K7_GO(0) else if (p.act == 1) K7_GO(1) else K7_GO(2)
#undef K7_GO
    FSD_CUDA(cudaGetLastError());
    h->launches += 1;
    return FSD_OK;
}

}  // namespace fsd

namespace fsd {
bool k10_supported(int K, int N);
int launch_pointwise_tc(fsd_context* h, const void* x, int64_t x_stride, const void* w, const void* bias, void* out, int64_t out_stride,
                        const void* res, int64_t res_stride, void* out2, int64_t out2_stride, int out2_c0, int64_t P, int K, int N,
                        int act, float slope, cudaStream_t stream, bool* taken);
}  // namespace fsd

using namespace fsd;

// 0: not supported; 1: the mma.sync kernel only (K <= 128, N in 16/32/64/128); 2: the tcgen05 kernel takes it (it is preferred when both do)
extern "C" int fsd_pointwise_conv_supported(int in_channels, int out_channels) {
    if (k10_supported(in_channels, out_channels)) return 2;
    const bool k7_shape = in_channels >= 16 && in_channels % 16 == 0 && in_channels <= 128 &&
                          (out_channels == 16 || out_channels == 32 || out_channels == 64 || out_channels == 128);
    return k7_shape ? 1 : 0;
}

extern "C" int fsd_pointwise_conv(fsd_handle_t h, const void* x, int64_t x_pixel_stride, const void* weight, const void* bias,
                                  void* out, int64_t out_pixel_stride, const void* residual, int64_t residual_pixel_stride,
                                  void* out2, int64_t out2_pixel_stride, int out2_first_channel, int64_t n_pixels,
                                  int in_channels, int out_channels, int act, float slope, int dtype, void* stream_) {
    FSD_CHECK_ARG(h && x && weight && bias && out, "fsd_pointwise_conv: null argument");
    FSD_CHECK_ARG(dtype == FSD_F16, "fsd_pointwise_conv: only fp16 is implemented");
    FSD_CHECK_ARG(n_pixels >= 0 && act >= 0 && act <= 2, "fsd_pointwise_conv: bad sizes / activation");
    const bool k7_shape = in_channels >= 16 && in_channels % 16 == 0 && in_channels <= 128 &&
                          (out_channels == 16 || out_channels == 32 || out_channels == 64 || out_channels == 128);
    FSD_CHECK_ARG(k7_shape || k10_supported(in_channels, out_channels),
                  "fsd_pointwise_conv: unsupported shape %d -> %d (channels must be multiples of 16; see fsd_pointwise_conv_supported)",
                  in_channels, out_channels);
    FSD_CHECK_ARG(x_pixel_stride >= in_channels && x_pixel_stride % 8 == 0, "fsd_pointwise_conv: bad input stride");
    FSD_CHECK_ARG(out_pixel_stride >= out_channels && out_pixel_stride % 8 == 0, "fsd_pointwise_conv: bad output stride");
    FSD_CHECK_ARG(!residual || (residual_pixel_stride >= out_channels && residual_pixel_stride % 8 == 0), "fsd_pointwise_conv: bad residual stride");
    FSD_CHECK_ARG(!out2 || (out2_first_channel >= 0 && out2_first_channel < out_channels && out2_first_channel % 8 == 0 &&
                            out2_pixel_stride >= out_channels - out2_first_channel && out2_pixel_stride % 8 == 0),
                  "fsd_pointwise_conv: bad second destination");
    if (((uintptr_t)x & 15) || ((uintptr_t)weight & 15) || ((uintptr_t)out & 15) || ((uintptr_t)residual & 15) || ((uintptr_t)out2 & 15)) {
        set_error("fsd_pointwise_conv: pointers must be 16-byte aligned");
        return FSD_ERR_ALIGN;
    }
    if (n_pixels == 0) return FSD_OK;
    cudaStream_t s = (cudaStream_t)stream_;
    FSD_CUDA(cudaSetDevice(h->device));
    // tensor-core path (k10_pointwise_tc.cu: TMA + tcgen05.mma + tensor memory) for every shape it takes; FSD_K7_NO_TC=1 keeps the
    // mma.sync kernel below (the parity tests run both)
    if (k10_supported(in_channels, out_channels) && !getenv("FSD_K7_NO_TC")) {
        bool taken = false;
        const int rc = launch_pointwise_tc(h, x, x_pixel_stride, weight, bias, out, out_pixel_stride, residual, residual_pixel_stride, out2,
                                           out2_pixel_stride, out2_first_channel, n_pixels, in_channels, out_channels, act, slope, s, &taken);
        if (rc != FSD_OK || taken) return rc;
    }
    if (!k7_shape) {
        set_error("fsd_pointwise_conv: %d -> %d needs the tensor-core path, which is unavailable here", in_channels, out_channels);
        return FSD_ERR_ARG;
    }
    K7Params p;
    p.x = (const __half*)x; p.w = (const __half*)weight; p.bias = (const __half*)bias; p.out = (__half*)out;
    p.res = (const __half*)residual; p.out2 = (__half*)out2; p.P = n_pixels; p.K = in_channels;
    p.x_stride = (int)x_pixel_stride; p.out_stride = (int)out_pixel_stride; p.res_stride = (int)residual_pixel_stride;
    p.out2_stride = (int)out2_pixel_stride; p.out2_c0 = out2_first_channel; p.act = act; p.slope = slope;
    switch (out_channels) {
        case 16: return k7_launch<16>(h, p, s);
        case 32: return k7_launch<32>(h, p, s);
        case 64: return k7_launch<64>(h, p, s);
        default: return k7_launch<128>(h, p, s);
    }
}
