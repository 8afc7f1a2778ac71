// Kernel 1, exact-2x fast path: slice gather + cv2 INTER_LINEAR 2x up-scale + uint8 -> fp16 normalise.
//
// SAHI slices whose side is half the network size (BASELINE config 2: 512x512 slices at imgsz 1024) resize by exactly
// 2.  cv2's fixed-point weights are then the constants 0.75/0.25 (1536/512 of 2048) on both axes, and the two passes
// collapse to small-integer arithmetic (all exact, derived from the general formula):
//     s      = 3*p[j] + p[j+1]   (odd output column)   |   p[j] + 3*p[j+1]   (even output column)      <= 1020
//     h      = (1536*p0 + 512*p1) >> 4 = 32*s
//     (1536*h) >> 16 = (3*s) >> 2 =: A        (512*h) >> 16 = s >> 2 =: B
//     out(2k+1) = (A[k] + B[k+1] + 2) >> 2      out(2k+2) = (B[k] + A[k+1] + 2) >> 2
// with source indices clamped at the borders (which reproduces cv2's border rules: x weights collapse to (2048, 0),
// y indices clip).  Everything fits 16-bit lanes, so one 32-bit register carries two output columns through every step:
// about 6 instructions per output element instead of 17 in the general kernel, which makes this path HBM-bound.
// One thread produces 8 output columns x 3 channels and marches down the source rows, keeping A/B of the previous row
// in registers; source bytes come straight from L1/L2 (they are 6 % of the traffic), stores are 128-bit.
#include <vector>

#include "fsd_common.cuh"

namespace fsd {

constexpr int UP2_THREADS = 128;
constexpr int UP2_ROWS = 16;  // source rows marched by one thread (32 output rows)

struct Up2Params {
    const uint8_t* images;
    int64_t row_pitch, image_pitch;
    const int32_t* entries;  // [B,3] image_index, x0, y0
    __half* out;
    int src_w, src_h, reverse;
};

__device__ __forceinline__ uint32_t norm255_pair(uint32_t w) {  // (vB << 16 | vA), 8-bit values -> half2(vA/255, vB/255)
    const __half2 magic = __halves2half2(__ushort_as_half(0x6400), __ushort_as_half(0x6400));
    uint32_t m = w | 0x64006400u;
    const __half2 v = __hsub2(*reinterpret_cast<__half2*>(&m), magic);
    const __half2 c_hi = __halves2half2(__ushort_as_half(0x1C04), __ushort_as_half(0x1C04));  // fp16(1/255)
    const __half2 c_lo = __halves2half2(__ushort_as_half(0x0001), __ushort_as_half(0x0001));  // 2^-24
    const __half2 r = __hfma2(v, c_hi, __hmul2(v, c_lo));  // == half(float(v)/255) for all 256 inputs (exhaustive search)
    return *reinterpret_cast<const uint32_t*>(&r);
}

// t = two 16-bit lanes holding A + B + 2 (<= 1022 each; stray bits above bit 9 of the low lane allowed): the cv2 result
// v = t >> 2 (8 bits) normalised to half(v / 255), without ever forming v as an integer: (t & 0x03fc) | 0x6400 is the half
// 1024 + 4v (exact, ulp 1), one FMA turns it into v (x 0.25 - 256, exact), then the two-term product of norm255_pair.
// One LOP3 + three half2 operations per pair instead of SHF + LOP + LOP + three: the integer ops run on the half-rate ALU
// pipe that bounds these kernels (profiles/r1_k1_sixteenths_c1.summary.txt: ALU 54 %, issue 71 %).
__device__ __forceinline__ uint32_t quant_norm_pair(uint32_t t) {
    uint32_t m = (t & 0x03fc03fcu) | 0x64006400u;
    const __half2 q = __halves2half2(__ushort_as_half(0x3400), __ushort_as_half(0x3400));      // 0.25
    const __half2 z = __halves2half2(__ushort_as_half(0xDC00), __ushort_as_half(0xDC00));      // -256
    const __half2 v = __hfma2(*reinterpret_cast<__half2*>(&m), q, z);
    const __half2 c_hi = __halves2half2(__ushort_as_half(0x1C04), __ushort_as_half(0x1C04));  // fp16(1/255)
    const __half2 c_lo = __halves2half2(__ushort_as_half(0x0001), __ushort_as_half(0x0001));  // 2^-24
    const __half2 r = __hfma2(v, c_hi, __hmul2(v, c_lo));
    return *reinterpret_cast<const uint32_t*>(&r);
}

__device__ __forceinline__ uint32_t mad_hi(uint32_t a, uint32_t b, uint32_t c) {  // hi32(a * b) + c, one IMAD.HI (FMA pipe)
    uint32_t d;
    asm("mad.hi.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ uint32_t pack16(uint32_t lo, uint32_t hi) {  // lo | hi << 16 as one IMAD (FMA pipe, not the ALU pipe)
    uint32_t d;
    asm("mad.lo.u32 %0, %1, 65536, %2;" : "=r"(d) : "r"(hi), "r"(lo));
    return d;
}

// A/B words of one source row for one channel: 4 words = output columns (2j..2j+7), two columns per word
struct RowAB { uint32_t a[4], b[4]; };

__device__ __forceinline__ void row_ab(const uint32_t (&p)[6], RowAB& r) {
    // p[i] = source pixel j-1+i (one channel).  word i = (se_i, so_{i+1}) = (p[i] + 3 p[i+1], 3 p[i+1] + p[i+2])
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const uint32_t s = p[i] + p[i + 1] * 0x00030003u + (p[i + 2] << 16);
        r.a[i] = (((s * 3u) >> 2) & 0x0fff0fffu) + 0x00020002u;  // + the rounding constant of the vertical pass (A <= 765)
        r.b[i] = (s >> 2) & 0x03ff03ffu;
    }
}

__device__ __forceinline__ void out_row(const uint32_t (&x)[4], const uint32_t (&y)[4], uint32_t (&o)[4]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) o[i] = quant_norm_pair(x[i] + y[i]);  // exactly one of x, y is an A word: it carries the + 2
}

// MB = minimum resident CTAs per SM the register allocation must allow: 5 -> 88-91 registers (30 % occupancy), 6 -> 80, no
// spills, 7 -> 72 with one 4-byte spill.  The kernel waits on scattered byte loads (top stall: long scoreboard), so more
// resident warps hide more latency; FSD_K1_MINBLOCKS selects the variant for the measurement (benchmarks/kernels.py k1).
template <bool NHWC, int MB>
__global__ void __launch_bounds__(UP2_THREADS, MB)
k1_upscale2x_kernel(const Up2Params p) {
    const int b = blockIdx.z;
    const int v = blockIdx.x * UP2_THREADS + threadIdx.x;  // 8-column output vector index
    const int out_w = 2 * p.src_w, out_h = 2 * p.src_h;
    if ((v & ~31) * 8 >= out_w) return;  // whole warp beyond the row (warp-uniform)
    const bool active = v * 8 < out_w;   // a partially filled warp keeps its idle lanes for the warp-level store transpose
    const int img = __ldg(p.entries + 3 * b + 0), x0 = __ldg(p.entries + 3 * b + 1), y0 = __ldg(p.entries + 3 * b + 2);
    const uint8_t* base = p.images + (size_t)img * p.image_pitch + (size_t)y0 * p.row_pitch + (size_t)x0 * 3;
    const int j = v * 4;  // first source column of this vector (output columns 2j .. 2j+7 use source j-1 .. j+4)
    int off[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) off[i] = min(max(j - 1 + i, 0), p.src_w - 1) * 3;
    const int c0 = p.reverse ? 2 : 0, c2 = p.reverse ? 0 : 2;  // plane <- source channel
    const size_t plane = (size_t)out_h * out_w;
    __half* out = p.out + (size_t)b * 3 * plane + (NHWC ? (size_t)v * 24 : (size_t)v * 8);
    // channels-last: first byte (within an output row) and base pointer of this WARP's 32 vectors
    const int warp_byte0 = (v & ~31) * 48;
    __half* out_warp = p.out + (size_t)b * 3 * plane + (size_t)(v & ~31) * 24;
    __shared__ uint4 s_stage[NHWC ? UP2_THREADS * 3 : 1];

    auto load_row = [&](int k, RowAB (&r)[3]) {
        const uint8_t* row = base + (size_t)min(max(k, 0), p.src_h - 1) * p.row_pitch;
        uint32_t px[3][6];
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            px[0][i] = __ldg(row + off[i] + c0);
            px[1][i] = __ldg(row + off[i] + 1);
            px[2][i] = __ldg(row + off[i] + c2);
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) row_ab(px[c], r[c]);
    };
    auto store_row = [&](int Y, const uint32_t (&o)[3][4]) {
        if (NHWC) {
            uint32_t w[12];
#pragma unroll
            for (int k = 0; k < 4; ++k) {  // word k of a plane holds pixels (2k, 2k+1)
                w[3 * k + 0] = __byte_perm(o[0][k], o[1][k], 0x5410);
                w[3 * k + 1] = __byte_perm(o[2][k], o[0][k], 0x7610);
                w[3 * k + 2] = __byte_perm(o[1][k], o[2][k], 0x7632);
            }
            // each thread owns 48 contiguous bytes; a direct store would make every STG.128 of the warp hit 32 chunks
            // 48 B apart (half-filled sectors).  Transpose through the warp's smem slab so that every store instruction
            // writes one contiguous 512-byte run (conflict-free: 48-byte stride -> distinct banks per quarter-warp).
            uint4* slab = s_stage + (threadIdx.x & ~31) * 3;
            const int lane = threadIdx.x & 31;
            slab[lane * 3 + 0] = make_uint4(w[0], w[1], w[2], w[3]);
            slab[lane * 3 + 1] = make_uint4(w[4], w[5], w[6], w[7]);
            slab[lane * 3 + 2] = make_uint4(w[8], w[9], w[10], w[11]);
            __syncwarp();
            uint4* d = reinterpret_cast<uint4*>(out_warp + (size_t)Y * out_w * 3);
            const int nchunk = min(96, (out_w * 3 * 2 - warp_byte0) / 16);  // the last warp of a row may be partial
#pragma unroll
            for (int k = 0; k < 3; ++k)
                if (k * 32 + lane < nchunk) d[k * 32 + lane] = slab[k * 32 + lane];
            __syncwarp();
        } else if (active) {
#pragma unroll
            for (int c = 0; c < 3; ++c)
                *reinterpret_cast<uint4*>(out + c * plane + (size_t)Y * out_w) = make_uint4(o[c][0], o[c][1], o[c][2], o[c][3]);
        }
    };

    // march k = k_lo .. k_hi-1: source rows (k, k+1) -> output rows 2k+1 and 2k+2 (k = -1 and k = src_h-1 are half steps)
    const int k_lo = -1 + (int)blockIdx.y * UP2_ROWS;
    const int k_hi = min(k_lo + UP2_ROWS, p.src_h);
    RowAB prev[3], cur[3];
    load_row(k_lo, prev);
    for (int k = k_lo; k < k_hi; ++k) {
        load_row(k + 1, cur);
        uint32_t o[3][4];
        if (2 * k + 1 >= 0) {
#pragma unroll
            for (int c = 0; c < 3; ++c) out_row(prev[c].a, cur[c].b, o[c]);
            store_row(2 * k + 1, o);
        }
        if (2 * k + 2 < out_h) {
#pragma unroll
            for (int c = 0; c < 3; ++c) out_row(prev[c].b, cur[c].a, o[c]);
            store_row(2 * k + 2, o);
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) prev[c] = cur[c];
    }
}

// called by fsd_gather_letterbox when the geometry is an exact 2x up-scale without letterbox border and fp16 output
int launch_upscale2x(fsd_context* h, const uint8_t* images, int64_t row_pitch, int64_t image_pitch, const int32_t* entries,
                     int B, int src_w, int src_h, int reverse, int nhwc, void* out, cudaStream_t stream) {
    Up2Params p;
    p.images = images; p.row_pitch = row_pitch; p.image_pitch = image_pitch; p.entries = entries;
    p.out = reinterpret_cast<__half*>(out); p.src_w = src_w; p.src_h = src_h; p.reverse = reverse;
    const int vecs = (2 * src_w + 7) / 8;
    dim3 grid((vecs + UP2_THREADS - 1) / UP2_THREADS, (src_h + 1 + UP2_ROWS - 1) / UP2_ROWS, B);
    {
        TimedLaunch timed(h, FSD_KERNEL_GATHER, B, src_w, stream);
        const int mb = getenv("FSD_K1_MINBLOCKS") ? atoi(getenv("FSD_K1_MINBLOCKS")) : 6;
#define UP2_GO(MBV)                                                                  \
        if (nhwc) k1_upscale2x_kernel<true, MBV><<<grid, UP2_THREADS, 0, stream>>>(p); \
        else k1_upscale2x_kernel<false, MBV><<<grid, UP2_THREADS, 0, stream>>>(p);
        if (mb <= 5) { UP2_GO(5) } else if (mb == 6) { UP2_GO(6) } else { UP2_GO(7) }
#undef UP2_GO
    }
    FSD_CUDA(cudaGetLastError());
    h->launches += 1;
    return FSD_OK;
}


// ---- "sixteenths" path: any border-less resize whose cv2 coefficients are all multiples of 128 (= k/16 weights) ----------
// Ratios 8/5 (SAHI's default 640^2 slices at imgsz 1024), 4, 2, 1 (the full-image pass when the image is already a
// multiple of the stride) have fractional source positions in sixteenths, so cv2's 11-bit coefficients are 128*k and the
// two fixed-point passes reduce to
//     s   = S[sx]*k0 + S[sx+1]*(16-k0)                          (<= 4080, horizontal pass; h = 8*s exactly)
//     out = (((s[sy]*m0) >> 6) + ((s[sy1]*m1) >> 6) + 2) >> 2    (m0 + m1 = 16; every product < 65536)
// i.e. the vertical pass and everything after it run on 16-bit lanes, two output columns per register, exactly as in
// the 2x kernel above; the horizontal pass is evaluated once per source row and reused by the 1-4 output rows that
// need it.  Indices and weights come from the cv2-exact host tables (border rules included), re-packed per column/row.
struct K16Params {
    const uint8_t* images;
    int64_t row_pitch, image_pitch;
    const int32_t* entries;
    const int32_t* xk;  // [out_w]  byte offset of S0 | (S1 is 3 bytes further ? 1 : 0) << 20 | k0 << 24
    const int32_t* yk;  // [out_h]  sy | (sy1 - sy) << 16 | m0 << 24
    __half* out;
    int out_w, out_h, reverse, rows_per_cta, src_w;
};

// Horizontal pattern of an 8-column output group when out_w * Q == 8 * src_w (ratios 8/Q: 8, 4, 8/3, 2, 8/5, 4/3, 8/7, 1 up or
// copy; 4/5, 2/3, 8/15 down):
// every group maps to Q source pixels with the SAME relative indices and weights, so they are compile-time constants and
// the group's Q + 2 source pixels are fetched as a few aligned 32-bit words (a byte load per tap costs ~4 L1 wavefronts
// because neighbouring lanes are 3*Q bytes apart; the byte-load version of this kernel was bound by exactly that).
template <int Q> struct K16Pat {
    static constexpr int n(int i) { return (2 * i + 1) * Q - 8; }                       // source position in sixteenths
    static constexpr int fl(int i) { return n(i) >= 0 ? n(i) / 16 : -((-n(i) + 15) / 16); }
    static constexpr int rel(int i) { return fl(i) + 1; }                               // tap 0 = pixel rel(i) of the span (span starts at pixel v*Q - 1)
    static constexpr int k1(int i) { return n(i) - 16 * fl(i); }                        // weight of tap 1 in sixteenths
    static constexpr int SPAN = (Q + 2) * 3;                                            // bytes
    static constexpr int NS = (SPAN + 3) / 4;                                           // words after alignment
};

template <bool NHWC, int Q>
__global__ void __launch_bounds__(UP2_THREADS)
k1_sixteenths_kernel(const K16Params p) {
    const int b = blockIdx.z;
    const int v = blockIdx.x * UP2_THREADS + threadIdx.x;  // 8-column output vector index
    const int out_w = p.out_w, out_h = p.out_h;
    if ((v & ~31) * 8 >= out_w) return;
    const bool active = v * 8 < out_w;
    const int img = __ldg(p.entries + 3 * b + 0), x0 = __ldg(p.entries + 3 * b + 1), y0 = __ldg(p.entries + 3 * b + 2);
    const uint8_t* base = p.images + (size_t)img * p.image_pitch + (size_t)y0 * p.row_pitch + (size_t)x0 * 3;
    int off0[Q == 0 ? 8 : 1], d1[Q == 0 ? 8 : 1], k0[Q == 0 ? 8 : 1];
    if (Q == 0) {
#pragma unroll
        for (int i = 0; i < (Q == 0 ? 8 : 1); ++i) {
            const int t = active ? __ldg(p.xk + v * 8 + i) : 0;
            off0[i] = t & 0xfffff; d1[i] = ((t >> 20) & 1) * 3; k0[i] = (t >> 24) & 31;
        }
    }
    // Q > 0: this thread's span = source pixels v*Q - 1 .. v*Q + Q (clamped to the slice: that IS cv2's border rule).  The
    // word path needs the whole span inside the slice and its aligned window inside the row pitch.
    using Pat = K16Pat<(Q > 0 ? Q : 1)>;
    const int px_first = v * Q - 1;
    const long long span_byte0 = (long long)x0 * 3 + (long long)px_first * 3;  // offset within the image row
    const bool word_path = Q > 0 && active && px_first >= 0 && px_first + Q + 1 <= p.src_w - 1 &&
                           (span_byte0 & ~3ll) + 4 * (Pat::NS + 1) <= p.row_pitch;
    const int c0 = p.reverse ? 2 : 0, c2 = p.reverse ? 0 : 2;
    const size_t plane = (size_t)out_h * out_w;
    __half* out = p.out + (size_t)b * 3 * plane + (NHWC ? (size_t)v * 24 : (size_t)v * 8);
    const int warp_byte0 = (v & ~31) * 48;
    __half* out_warp = p.out + (size_t)b * 3 * plane + (size_t)(v & ~31) * 24;
    __shared__ uint4 s_stage[NHWC ? UP2_THREADS * 3 : 1];

    auto hrow = [&](int r, uint32_t (&dst)[3][4]) {  // packed s values of source row r for this thread's 8 columns
        const uint8_t* row = base + (size_t)r * p.row_pitch;
        if (Q == 0) {
#pragma unroll
            for (int c = 0; c < 3; ++c)
#pragma unroll
                for (int j = 0; j < 4; ++j) dst[c][j] = 0;
#pragma unroll
            for (int i = 0; i < (Q == 0 ? 8 : 1); ++i) {
                const uint8_t* px = row + off0[i];
                const int k1 = 16 - k0[i];
                const uint32_t sA = __ldg(px + c0) * k0[i] + __ldg(px + d1[i] + c0) * k1;
                const uint32_t sB = __ldg(px + 1) * k0[i] + __ldg(px + d1[i] + 1) * k1;
                const uint32_t sC = __ldg(px + c2) * k0[i] + __ldg(px + d1[i] + c2) * k1;
                dst[0][i >> 1] |= sA << (16 * (i & 1));
                dst[1][i >> 1] |= sB << (16 * (i & 1));
                dst[2][i >> 1] |= sC << (16 * (i & 1));
            }
            return;
        }
        // ---- Q > 0: the span's bytes as a word stream st[] (byte 3*t + c = pixel t of the span, source channel c) ----------
        uint32_t st[Pat::NS + 1];
        if (word_path) {
            const uint8_t* p0 = row + (long long)px_first * 3;
            const uintptr_t a = reinterpret_cast<uintptr_t>(p0);
            const uint32_t* wp = reinterpret_cast<const uint32_t*>(a & ~(uintptr_t)3);
            const int sh = (int)(a & 3) * 8;
            uint32_t w[Pat::NS + 1];
#pragma unroll
            for (int k = 0; k <= Pat::NS; ++k) w[k] = __ldg(wp + k);
#pragma unroll
            for (int k = 0; k < Pat::NS; ++k) st[k] = __funnelshift_r(w[k], w[k + 1], sh);
        } else {  // border vectors (and idle lanes): clamped per-pixel byte loads, same stream layout
#pragma unroll
            for (int k = 0; k < Pat::NS; ++k) st[k] = 0;
#pragma unroll
            for (int t = 0; t < Q + 2; ++t) {
                const int col = min(max(px_first + t, 0), p.src_w - 1);
#pragma unroll
                for (int c = 0; c < 3; ++c) st[(3 * t + c) >> 2] |= (uint32_t)__ldg(row + col * 3 + c) << (8 * ((3 * t + c) & 3));
            }
        }
        st[Pat::NS] = 0;
        // s = S0*k0 + S1*k1 with the taps 3 bytes apart: ONE dp4a over the 4-byte window starting at tap 0, weights (k0,0,0,k1)
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            constexpr int dummy = 0; (void)dummy;
            const int k1 = Pat::k1(i), k0w = 16 - k1;
            const uint32_t wts = (uint32_t)k0w | ((uint32_t)k1 << 24);
            uint32_t sv[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const int b0 = 3 * Pat::rel(i) + c;  // stream byte of tap 0
                const uint32_t win = (b0 & 3) == 0 ? st[b0 >> 2] : __funnelshift_r(st[b0 >> 2], st[(b0 >> 2) + 1], 8 * (b0 & 3));
                sv[c] = __dp4a(win, wts, 0u);
            }
            // plane 0 <- source channel c0, plane 1 <- 1, plane 2 <- c2
            const uint32_t pA = p.reverse ? sv[2] : sv[0], pC = p.reverse ? sv[0] : sv[2];
            if (i & 1) {
                dst[0][i >> 1] = pack16(dst[0][i >> 1], pA); dst[1][i >> 1] = pack16(dst[1][i >> 1], sv[1]); dst[2][i >> 1] = pack16(dst[2][i >> 1], pC);
            } else {
                dst[0][i >> 1] = pA; dst[1][i >> 1] = sv[1]; dst[2][i >> 1] = pC;
            }
        }
    };
    auto store_row = [&](int Y, const uint32_t (&o)[3][4]) {
        if (NHWC) {
            uint32_t w[12];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                w[3 * k + 0] = __byte_perm(o[0][k], o[1][k], 0x5410);
                w[3 * k + 1] = __byte_perm(o[2][k], o[0][k], 0x7610);
                w[3 * k + 2] = __byte_perm(o[1][k], o[2][k], 0x7632);
            }
            uint4* slab = s_stage + (threadIdx.x & ~31) * 3;
            const int lane = threadIdx.x & 31;
            slab[lane * 3 + 0] = make_uint4(w[0], w[1], w[2], w[3]);
            slab[lane * 3 + 1] = make_uint4(w[4], w[5], w[6], w[7]);
            slab[lane * 3 + 2] = make_uint4(w[8], w[9], w[10], w[11]);
            __syncwarp();
            uint4* d = reinterpret_cast<uint4*>(out_warp + (size_t)Y * out_w * 3);
            const int nchunk = min(96, (out_w * 3 * 2 - warp_byte0) / 16);
#pragma unroll
            for (int k = 0; k < 3; ++k)
                if (k * 32 + lane < nchunk) d[k * 32 + lane] = slab[k * 32 + lane];
            __syncwarp();
        } else if (active) {
#pragma unroll
            for (int c = 0; c < 3; ++c)
                *reinterpret_cast<uint4*>(out + c * plane + (size_t)Y * out_w) = make_uint4(o[c][0], o[c][1], o[c][2], o[c][3]);
        }
    };

    const int y_lo = blockIdx.y * p.rows_per_cta, y_hi = min(y_lo + p.rows_per_cta, out_h);
    uint32_t cur[3][4], nxt[3][4];
    int cur_i = -1, nxt_i = -1;
    for (int y = y_lo; y < y_hi; ++y) {  // every branch below is uniform over the CTA (it depends on y only)
        const int t = __ldg(p.yk + y);
        const int sy = t & 0xffff, sy1 = sy + ((t >> 16) & 1);
        const uint32_t m0 = (uint32_t)(t >> 24) & 31u, m1 = 16u - m0;
        const uint32_t f0 = m0 << 26, f1 = m1 << 26;  // (x * m) >> 6 == umulhi(x, m << 26): multiply and shift in one IMAD.HI
        if (sy != cur_i) {
            if (sy == nxt_i) {
#pragma unroll
                for (int c = 0; c < 3; ++c)
#pragma unroll
                    for (int j = 0; j < 4; ++j) cur[c][j] = nxt[c][j];
            } else {
                hrow(sy, cur);
            }
            cur_i = sy;
        }
        if (sy1 != nxt_i) {
            if (sy1 == cur_i) {
#pragma unroll
                for (int c = 0; c < 3; ++c)
#pragma unroll
                    for (int j = 0; j < 4; ++j) nxt[c][j] = cur[c][j];
            } else {
                hrow(sy1, nxt);
            }
            nxt_i = sy1;
        }
        uint32_t o[3][4];
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                // A is masked to its 10-bit lanes; B may keep the <= 6 stray bits the shift moved into bits 10..15 of the
                // low lane: A + B + 2 <= 1023 + (63 << 10) cannot carry into the high lane, and the final mask drops them
                // hi32(x * f) + addend is ONE IMAD.HI (mad.hi): the rounding constant rides on the first product (A <= 1020 per
                // lane, so + 2 cannot carry out of its 10 bits), the masked A + 2 on the second
                const uint32_t A2 = mad_hi(cur[c][j], f0, 0x00020002u) & 0x03ff03ffu;
                o[c][j] = quant_norm_pair(mad_hi(nxt[c][j], f1, A2));
            }
        store_row(y, o);
    }
}

// Packs the cv2-exact host tables for the sixteenths path; false when some coefficient is not a multiple of 128.
bool pack_sixteenths_tables(const std::vector<int32_t>& xt, const std::vector<int32_t>& yt, int src_w, int src_h, int out_w,
                            int out_h, std::vector<int32_t>& packed) {
    packed.resize((size_t)out_w + out_h);
    for (int x = 0; x < out_w; ++x) {
        const int s = xt[2 * x], w0 = xt[2 * x + 1] & 0xffff, w1 = (xt[2 * x + 1] >> 16) & 0xffff;
        if (w0 % 128 || w1 % 128 || w0 + w1 != 2048 || s < 0 || s >= src_w || s * 3 > 0xfffff) return false;
        const int s1 = s + 1 < src_w ? s + 1 : src_w - 1;
        packed[x] = (s * 3) | ((s1 != s ? 1 : 0) << 20) | ((w0 / 128) << 24);
    }
    for (int y = 0; y < out_h; ++y) {
        const int s = yt[4 * y], s1 = yt[4 * y + 1], w0 = yt[4 * y + 2], w1 = yt[4 * y + 3];
        if (w0 % 128 || w1 % 128 || w0 + w1 != 2048 || s < 0 || s >= src_h || s > 0xffff || (s1 != s && s1 != s + 1)) return false;
        packed[out_w + y] = s | ((s1 - s) << 16) | ((w0 / 128) << 24);
    }
    return true;
}

int launch_sixteenths(fsd_context* h, const uint8_t* images, int64_t row_pitch, int64_t image_pitch, const int32_t* entries,
                      int B, int src_w, const int32_t* packed_dev, int out_w, int out_h, int reverse, int nhwc, void* out,
                      cudaStream_t stream) {
    K16Params p;
    p.images = images; p.row_pitch = row_pitch; p.image_pitch = image_pitch; p.entries = entries;
    p.xk = packed_dev; p.yk = packed_dev + out_w; p.out = reinterpret_cast<__half*>(out);
    p.out_w = out_w; p.out_h = out_h; p.reverse = reverse; p.src_w = src_w;
    const int vecs = out_w / 8;
    // rows per CTA: a CTA re-uses each horizontally filtered source row for the 1-4 output rows that need it, so taller CTAs
    // save work when up-scaling; but a small launch (the full-image pass: one entry per image) must still fill several
    // waves of the 148 SMs x ~8 resident CTAs, or the tail wave costs 30 % (1536 CTAs = 1.3 waves at 32 rows)
    p.rows_per_cta = 32;
    if (getenv("FSD_K16_ROWS")) p.rows_per_cta = atoi(getenv("FSD_K16_ROWS"));
    else
        while (p.rows_per_cta > 8 &&
               (long long)((vecs + UP2_THREADS - 1) / UP2_THREADS) * ((out_h + p.rows_per_cta - 1) / p.rows_per_cta) * B < 6LL * 148 * 8)
            p.rows_per_cta >>= 1;
    if (p.rows_per_cta < 1) p.rows_per_cta = 1;
    dim3 grid((vecs + UP2_THREADS - 1) / UP2_THREADS, (out_h + p.rows_per_cta - 1) / p.rows_per_cta, B);
    // uniform 8-column pattern (out_w * Q == 8 * src_w) -> compile-time taps; anything else -> the table-driven variant
    int q = 0;
    if ((8LL * src_w) % out_w == 0 && !getenv("FSD_K1_TABLE16")) q = (int)(8LL * src_w / out_w);
    {
        TimedLaunch timed(h, FSD_KERNEL_GATHER, B, src_w, stream);
#define K16_GO(QQ)                                                                         \
        if (nhwc) k1_sixteenths_kernel<true, QQ><<<grid, UP2_THREADS, 0, stream>>>(p);    \
        else k1_sixteenths_kernel<false, QQ><<<grid, UP2_THREADS, 0, stream>>>(p);
        switch (q) {
            case 1: K16_GO(1) break;
            case 2: K16_GO(2) break;
            case 3: K16_GO(3) break;
            case 4: K16_GO(4) break;
            case 5: K16_GO(5) break;
            case 6: K16_GO(6) break;
            case 7: K16_GO(7) break;
            case 8: K16_GO(8) break;
            case 10: K16_GO(10) break;  // down-scales: 1.25x
            case 12: K16_GO(12) break;  //              1.5x
            case 15: K16_GO(15) break;  //              1.875x: 1920x1080 -> 1024x576, the full-image pass of config 1
            default: K16_GO(0) break;
        }
#undef K16_GO
    }
    FSD_CUDA(cudaGetLastError());
    h->launches += 1;
    return FSD_OK;
}

}  // namespace fsd
