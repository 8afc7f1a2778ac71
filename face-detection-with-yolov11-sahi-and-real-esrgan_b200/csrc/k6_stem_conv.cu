// Kernel 6 — YOLO stem: 3x3 stride-2 convolution of the 3-channel network input + bias + SiLU in ONE pass.
//
// ultralytics' first layer (Conv(3, c, 3, 2): yolo11-pose.yaml layer 0, run by model.predict at utils/yolo_wrapper.py:72)
// is the one convolution cuDNN serves badly: a 3-channel channels-last input matches none of its tensor-core NHWC kernels
// (profiles/r1_launches_bench_b32_slotcat.txt: 0.84 ms per 96 inputs, 8.8 % of the step, plus 0.28 ms for the epilogue
// pass) although the layer moves only 1.4 GB (0.22 ms at HBM speed).  Here:
//   * a CTA owns an 8 x 128 tile of output pixels; the 17 x 258 input pixels under it are staged in shared memory as
//     [row][col][4 halfs] (channel 3 = 0), so the 3 x 4(cols) x 4(ch) window of an output pixel's kernel row is 16
//     contiguous halfs and consecutive output pixels are 16 bytes apart: one `ldmatrix.x4` yields the 16 x 16 A tile
//     (16 output pixels x one kernel row) without any bank conflict;
//   * tensor cores via mma.sync.m16n8k16 (f16 x f16 -> f32): K = 3 kernel rows x 16, N = 16 output channels = 2 n-tiles
//     -> 6 MMAs per 16 output pixels, weights live in 12 registers per thread (the 4th column / 4th channel are zero);
//     this layer is 0.2 % of a tcgen05-sized GEMM — it is HBM-bound, the tensor core only keeps the FMA work off the
//     issue slots;
//   * epilogue in registers: + bias, SiLU (same fp32 formula as fsd_bias_act), fp16, channels-last stores.
#include "fsd_common.cuh"

namespace fsd {

constexpr int K6_THREADS = 256;
constexpr int K6_TH = 8;     // output rows per CTA (one per warp)
constexpr int K6_TW = 128;   // output columns per CTA
constexpr int K6_ROWS = 2 * K6_TH + 1;          // input rows staged
constexpr int K6_PAIRS = K6_TW + 2;             // input column pairs staged: columns 2*x0-2 .. 2*x0+257
constexpr int K6_PITCH = 2 * K6_PAIRS + 2;      // smem columns per row (index 0 unused; 8 bytes each; 2096 B = 131 * 16)

struct K6Params {
    const __half* x;      // [E, H, W, 3]
    const __half* w;      // [16, 3, 3, 3] (o, c, ky, kx)
    const __half* bias;   // [16]
    __half* out;          // [E, OH, OW, 16], or the space-to-depth form [E, OH/2 + 1, OW/2 + 1, 64] (s2d)
    int H, W, OH, OW, s2d;
};

__device__ __forceinline__ float k6_silu(float v) { return fast_silu(v); }

__global__ void __launch_bounds__(K6_THREADS) k6_stem_conv_kernel(const K6Params p) {
    __shared__ __align__(16) uint2 tile[K6_ROWS][K6_PITCH];
    const int e = blockIdx.z;
    const int x0 = blockIdx.x * K6_TW, y0 = blockIdx.y * K6_TH;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    // ---- weights -> B fragments.  k = kx * 4 + c; zero for kx == 3 or c == 3.  Built once per CTA in shared memory
    // ([fragment][lane], conflict-free), then 12 LDS per thread (each thread used to assemble its 24 halfs itself: a
    // quarter of the kernel's instructions, profiles/r1_k6_stem.summary.txt)
    __shared__ uint32_t s_bfrag[12][32];
    const int g = lane >> 2, t = lane & 3;
    for (int i = tid; i < 12 * 32; i += K6_THREADS) {
        const int f = i >> 5, ln = i & 31;          // f = (ky * 2 + j) * 2 + r
        const int ky = f >> 2, j = (f >> 1) & 1, r = f & 1, n = 8 * j + (ln >> 2);
        uint32_t word = 0;
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int k = 2 * (ln & 3) + 8 * r + q, kx = k >> 2, c = k & 3;
            if (kx < 3 && c < 3) word |= (uint32_t)__half_as_ushort(__ldg(p.w + ((n * 3 + c) * 3 + ky) * 3 + kx)) << (16 * q);
        }
        s_bfrag[f][ln] = word;
    }
    float bias[2][2];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        bias[j][0] = __half2float(__ldg(p.bias + 8 * j + 2 * t));
        bias[j][1] = __half2float(__ldg(p.bias + 8 * j + 2 * t + 1));
    }

    // ---- stage the input tile: pixel pairs (even column first) = 3 words -> two [c0,c1,c2,0] cells -----------------------
    const __half* img = p.x + (size_t)e * p.H * p.W * 3;
    const int r_in0 = 2 * y0 - 1, c_in0 = 2 * x0 - 2;
    for (int it = tid; it < K6_ROWS * K6_PAIRS; it += K6_THREADS) {
        const int r = it / K6_PAIRS, pr = it - r * K6_PAIRS;
        const int gy = r_in0 + r, gx = c_in0 + 2 * pr;
        uint32_t w0 = 0, w1 = 0, w2 = 0;
        if (gy >= 0 && gy < p.H && gx >= 0 && gx < p.W) {  // W is even, so a pair never straddles the right edge
            const uint32_t* src = reinterpret_cast<const uint32_t*>(img + ((size_t)gy * p.W + gx) * 3);
            w0 = __ldg(src); w1 = __ldg(src + 1); w2 = __ldg(src + 2);
        }
        // halfs: w0 = (a0,a1) w1 = (a2,b0) w2 = (b1,b2)
        tile[r][1 + 2 * pr] = make_uint2(w0, w1 & 0xffffu);
        tile[r][2 + 2 * pr] = make_uint2((w1 >> 16) | (w2 << 16), w2 >> 16);
    }
    if (tid < K6_ROWS) tile[tid][0] = make_uint2(0u, 0u);
    __syncthreads();
    uint32_t bfrag[3][2][2];
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int r = 0; r < 2; ++r) bfrag[ky][j][r] = s_bfrag[(ky * 2 + j) * 2 + r][lane];

    // ---- one warp per output row; K6_TW / 16 groups of 16 output pixels ---------------------------------------------------------
    const int y = y0 + warp;
    if (y >= p.OH) return;
    // plain: pixel (y, x) -> out[e][y][x][16].  s2d: -> out[e][y/2 + 1][x/2 + 1][((y&1)*2 + (x&1))*16 ..]: the layout in which
    // the following 3x3 stride-2 convolution is a 2x2 stride-1 convolution over 64 channels (row 0 / column 0 = its zero padding)
    const int OW2 = p.OW / 2 + 1;
    __half* orow = p.s2d ? p.out + (((size_t)e * (p.OH / 2 + 1) + (y >> 1) + 1) * OW2) * 64 + (y & 1) * 32
                         : p.out + ((size_t)e * p.OH + y) * p.OW * 16;
#pragma unroll
    for (int grp = 0; grp < K6_TW / 16; ++grp) {
        const int xg = grp * 16;
        if (x0 + xg >= p.OW) break;
        float acc[2][4];
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int q = 0; q < 4; ++q) acc[j][q] = 0.f;
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
            // A row m = output pixel xg+m: input columns 2(x0+xg+m)-1 .. +2 (+1 zero-weighted) = smem columns 2(xg+m)+2 ..
            const uint2* a_ptr = &tile[2 * warp + ky][2 * (xg + (lane & 15)) + 2 + 2 * (lane >> 4)];
            uint32_t a0, a1, a2, a3;
            asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                         : "=r"(a0), "=r"(a1), "=r"(a2), "=r"(a3) : "r"(smem_u32(a_ptr)));
#pragma unroll
            for (int j = 0; j < 2; ++j)
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(acc[j][0]), "+f"(acc[j][1]), "+f"(acc[j][2]), "+f"(acc[j][3])
                             : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(bfrag[ky][j][0]), "r"(bfrag[ky][j][1]));
        }
        // epilogue: acc[j][0..1] = pixel g, channels 8j+2t, +1; acc[j][2..3] = pixel g+8
#pragma unroll
        for (int half_ = 0; half_ < 2; ++half_) {
            const int x = x0 + xg + g + 8 * half_;
            if (x < p.OW) {
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const __half2 o = __floats2half2_rn(k6_silu(acc[j][2 * half_] + bias[j][0]), k6_silu(acc[j][2 * half_ + 1] + bias[j][1]));
                    const size_t px = p.s2d ? (size_t)((x >> 1) + 1) * 64 + (x & 1) * 16 : (size_t)x * 16;
                    *reinterpret_cast<__half2*>(orow + px + 8 * j + 2 * t) = o;
                }
            }
        }
    }
}

}  // namespace fsd

using namespace fsd;

extern "C" int fsd_stem_conv(fsd_handle_t h, const void* x, int E, int H, int W, const void* weight, const void* bias,
                             int out_channels, int dtype, int space_to_depth, void* out, void* stream_) {
    FSD_CHECK_ARG(h && x && weight && bias && out, "fsd_stem_conv: null argument");
    FSD_CHECK_ARG(dtype == FSD_F16, "fsd_stem_conv: only fp16 is implemented");
    FSD_CHECK_ARG(out_channels == 16, "fsd_stem_conv: the kernel is specialised for 16 output channels (YOLO11n), got %d", out_channels);
    FSD_CHECK_ARG(E >= 0 && H > 0 && W > 0 && W % 2 == 0, "fsd_stem_conv: bad sizes (W must be even)");
    if (((uintptr_t)x & 3) || ((uintptr_t)out & 3)) { set_error("fsd_stem_conv: pointers must be 4-byte aligned"); return FSD_ERR_ALIGN; }
    if (E == 0) return FSD_OK;
    FSD_CHECK_ARG(E <= 65535, "fsd_stem_conv: at most 65535 inputs per launch");
    K6Params p;
    p.x = (const __half*)x; p.w = (const __half*)weight; p.bias = (const __half*)bias; p.out = (__half*)out;
    p.H = H; p.W = W; p.OH = (H - 1) / 2 + 1; p.OW = (W - 1) / 2 + 1; p.s2d = space_to_depth ? 1 : 0;
    FSD_CHECK_ARG(!space_to_depth || (p.OH % 2 == 0 && p.OW % 2 == 0), "fsd_stem_conv: the space-to-depth output needs even output sizes");
    dim3 grid((p.OW + K6_TW - 1) / K6_TW, (p.OH + K6_TH - 1) / K6_TH, E);
    FSD_CUDA(cudaSetDevice(h->device));
    // algorithmic bytes: the 3-channel input once + the 16-channel output once
    TimedLaunch timed(h, FSD_KERNEL_STEM, (int64_t)E * ((int64_t)H * W * 3 + (int64_t)p.OH * p.OW * 16) * 2, 16, (cudaStream_t)stream_);
    k6_stem_conv_kernel<<<grid, K6_THREADS, 0, (cudaStream_t)stream_>>>(p);
    FSD_CUDA(cudaGetLastError());
    h->launches += 1;
    return FSD_OK;
}
