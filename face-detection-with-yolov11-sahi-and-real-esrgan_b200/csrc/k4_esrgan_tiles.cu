// Kernel 4 — Real-ESRGAN tile crop / stitch.
//
// crop:   u8 HWC BGR image -> packed fp16|fp32 [3,ph,pw] RGB tiles (value/255), each tile = its interior plus a
//         <= tile_pad halo clamped at the borders; the right/bottom 'reflect' pre-pad and mod-pad of
//         RealESRGANer.pre_process are folded into the index map, so no padded copy of the image ever exists.
// stitch: packed [3,ph*s,pw*s] network outputs -> u8 HWC BGR [H*s, W*s]: interior only (halo discarded, no
//         blending), clamp(0,1) * 255, round-half-even, RGB->BGR — RealESRGANer.tile_process/post_process/enhance.
// Both are pure streaming kernels (HBM-bound): 8 (crop) / 16 (stitch) pixels per thread, 128-bit stores whenever
// the destination is 16-byte aligned, scalar tails otherwise.
#include "fsd_common.cuh"

namespace fsd {

constexpr int K4_THREADS = 256;
constexpr int TT = 12;  // ints per tile-table row

__device__ __forceinline__ int reflect_index(int i, int n_pre, int n) {
    if (i >= n_pre) i = 2 * (n_pre - 1) - i;  // mod-pad reflects the pre-padded image
    if (i >= n) i = 2 * (n - 1) - i;          // pre-pad reflects the original image
    return i;
}

__device__ __forceinline__ int64_t off64(const int32_t* t, int k) {
    return (int64_t)(uint32_t)t[k] | ((int64_t)t[k + 1] << 32);
}

template <typename OutT> __device__ __forceinline__ OutT to_out(float v);
template <> __device__ __forceinline__ float to_out<float>(float v) { return v; }
template <> __device__ __forceinline__ __half to_out<__half>(float v) { return __float2half_rn(v); }

// v/255 in the precision the reference produces: fp32 IEEE division (astype(float32)/255), then .half() for fp16
// tiles.  half(v * (1/255)) == half(v / 255) for all 256 inputs; fp32 needs the Newton-corrected quotient.
template <typename OutT> __device__ __forceinline__ float unit255(uint32_t v) {
    const float rcp = 1.0f / 255.0f;
    const float fv = (float)v;
    const float q = fv * rcp;
    if (sizeof(OutT) == 2) return q;
    const float r = fmaf(-q, 255.0f, fv);
    return fmaf(r, rcp, q);
}

template <typename OutT>
__device__ __forceinline__ void store_vec8(OutT* o, const float (&v)[8], int nx) {
    if (nx == 8 && (reinterpret_cast<uintptr_t>(o) & 15) == 0) {
        if (sizeof(OutT) == 2) {
            __half2 h0 = __floats2half2_rn(v[0], v[1]), h1 = __floats2half2_rn(v[2], v[3]);
            __half2 h2 = __floats2half2_rn(v[4], v[5]), h3 = __floats2half2_rn(v[6], v[7]);
            uint4 pk;
            pk.x = *reinterpret_cast<uint32_t*>(&h0); pk.y = *reinterpret_cast<uint32_t*>(&h1);
            pk.z = *reinterpret_cast<uint32_t*>(&h2); pk.w = *reinterpret_cast<uint32_t*>(&h3);
            *reinterpret_cast<uint4*>(o) = pk;
        } else {
            float* f = reinterpret_cast<float*>(o);
            *reinterpret_cast<float4*>(f) = make_float4(v[0], v[1], v[2], v[3]);
            *reinterpret_cast<float4*>(f + 4) = make_float4(v[4], v[5], v[6], v[7]);
        }
    } else {
        for (int e = 0; e < nx; ++e) o[e] = to_out<OutT>(v[e]);
    }
}

// (vB << 16 | vA), two 8-bit values -> half2(vA/255, vB/255): 0x6400|v is the half 1024+v; the two-term product equals
// half(float(v)/255) — what `img.astype(float32)/255 -> .half()` yields — for all 256 inputs (exhaustive search, K1).
__device__ __forceinline__ uint32_t k4_norm255_pair(uint32_t w) {
    const __half2 magic = __halves2half2(__ushort_as_half(0x6400), __ushort_as_half(0x6400));
    uint32_t m = w | 0x64006400u;
    const __half2 v = __hsub2(*reinterpret_cast<__half2*>(&m), magic);
    const __half2 c_hi = __halves2half2(__ushort_as_half(0x1C04), __ushort_as_half(0x1C04));  // fp16(1/255)
    const __half2 c_lo = __halves2half2(__ushort_as_half(0x0001), __ushort_as_half(0x0001));  // 2^-24
    const __half2 r = __hfma2(v, c_hi, __hmul2(v, c_lo));
    return *reinterpret_cast<const uint32_t*>(&r);
}

// 8 halfs to a 2-byte-aligned address with the widest stores its alignment allows (tile rows are pw halfs long, so a row
// start is 16-byte aligned only when pw % 8 == 0; the scalar fallback used to cost 8 stores per plane)
__device__ __forceinline__ void store8h(__half* o, const uint32_t (&w)[4]) {
    const uintptr_t a = reinterpret_cast<uintptr_t>(o);
    if ((a & 15) == 0) {
        *reinterpret_cast<uint4*>(o) = make_uint4(w[0], w[1], w[2], w[3]);
    } else if ((a & 7) == 0) {
        reinterpret_cast<uint2*>(o)[0] = make_uint2(w[0], w[1]);
        reinterpret_cast<uint2*>(o)[1] = make_uint2(w[2], w[3]);
    } else if ((a & 3) == 0) {
#pragma unroll
        for (int k = 0; k < 4; ++k) reinterpret_cast<uint32_t*>(o)[k] = w[k];
    } else {
        unsigned short* u = reinterpret_cast<unsigned short*>(o);
#pragma unroll
        for (int k = 0; k < 4; ++k) { u[2 * k] = (unsigned short)(w[k] & 0xffffu); u[2 * k + 1] = (unsigned short)(w[k] >> 16); }
    }
}

template <typename OutT>
__global__ void __launch_bounds__(K4_THREADS)
k4_crop_kernel(const uint8_t* __restrict__ img, int H, int W, int64_t pitch, int H_pre, int W_pre,
               const int32_t* __restrict__ table, OutT* __restrict__ tiles, int64_t image_pitch, int64_t tiles_image_stride) {
    const int32_t* t = table + (size_t)blockIdx.y * TT;
    img += (size_t)blockIdx.z * image_pitch;        // blockIdx.z = image of the batch
    tiles += (size_t)blockIdx.z * tiles_image_stride;
    const int px0 = t[0], py0 = t[1], pw = t[2], ph = t[3];
    OutT* dst = tiles + off64(t, 8);
    const int vecs = (pw + 7) >> 3;
    const int items = ph * vecs;
    for (int it = blockIdx.x * K4_THREADS + threadIdx.x; it < items; it += gridDim.x * K4_THREADS) {
        const int y = it / vecs, x = (it - y * vecs) << 3;
        const int sy = reflect_index(py0 + y, H_pre, H);
        const uint8_t* row = img + (size_t)sy * pitch;
        const int nx = min(8, pw - x);
        const int gx = px0 + x;
        const bool fast = nx == 8 && gx + 8 <= W && sy + 1 < H;
        uint32_t s[6];
        if (fast) {
            // the 24 source bytes are contiguous and an aligned 28-byte window stays inside the image buffer
            const uintptr_t addr = reinterpret_cast<uintptr_t>(row + (size_t)gx * 3);
            const uint32_t* wp = reinterpret_cast<const uint32_t*>(addr & ~(uintptr_t)3);
            const int sh = (int)(addr & 3) * 8;
            uint32_t w[7];
#pragma unroll
            for (int k = 0; k < 7; ++k) w[k] = __ldg(wp + k);
#pragma unroll
            for (int k = 0; k < 6; ++k) s[k] = __funnelshift_r(w[k], w[k + 1], sh);
        }
        if (fast && sizeof(OutT) == 2) {
            // packed path: pixel pair (2k, 2k+1) of BGR channel c = stream bytes 6k+c and 6k+3+c -> one half2
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                uint32_t o[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int ia = 6 * k + c, ib = ia + 3;
                    const uint32_t sel = (uint32_t)(ia & 3) | ((uint32_t)(4 * ((ib >> 2) - (ia >> 2)) + (ib & 3)) << 8);
                    // selector nibble 0 -> stream byte ia, nibble 2 -> stream byte ib (bytes 1 and 3 are masked off)
                    const uint32_t pr = __byte_perm(s[ia >> 2], s[ib >> 2], sel);
                    o[k] = k4_norm255_pair(pr & 0x00ff00ffu);
                }
                store8h(reinterpret_cast<__half*>(dst) + ((size_t)(2 - c) * ph + y) * pw + x, o);
            }
            continue;
        }
        float v[3][8];
        if (fast) {
#pragma unroll
            for (int e = 0; e < 8; ++e) {
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const int i = 3 * e + c;  // byte i of the stream = pixel e, BGR channel c
                    v[2 - c][e] = unit255<OutT>((s[i >> 2] >> (8 * (i & 3))) & 0xffu);
                }
            }
        } else {
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                if (e < nx) {
                    const int sx = reflect_index(gx + e, W_pre, W);
                    const uint8_t* px = row + (size_t)sx * 3;
                    v[0][e] = unit255<OutT>(__ldg(px + 2));
                    v[1][e] = unit255<OutT>(__ldg(px + 1));
                    v[2][e] = unit255<OutT>(__ldg(px + 0));
                } else {
                    v[0][e] = v[1][e] = v[2][e] = 0.f;
                }
            }
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) store_vec8<OutT>(dst + ((size_t)c * ph + y) * pw + x, v[c], nx);
    }
}

template <typename T> __device__ __forceinline__ float ld_as_float(const T* p);
template <> __device__ __forceinline__ float ld_as_float<float>(const float* p) { return __ldg(p); }
template <> __device__ __forceinline__ float ld_as_float<__half>(const __half* p) { return __half2float(__ldg(p)); }

// 16 consecutive plane values as floats: 128-/64-bit loads when the address allows, scalar otherwise
template <typename T> __device__ __forceinline__ void load16(const T* p, float (&f)[16]);
template <> __device__ __forceinline__ void load16<__half>(const __half* p, float (&f)[16]) {
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    if ((a & 15) == 0) {
        const uint4 v0 = __ldg(reinterpret_cast<const uint4*>(p)), v1 = __ldg(reinterpret_cast<const uint4*>(p) + 1);
        const uint32_t w[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const float2 t = __half22float2(*reinterpret_cast<const __half2*>(&w[k]));
            f[2 * k] = t.x; f[2 * k + 1] = t.y;
        }
    } else if ((a & 7) == 0) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const uint2 v = __ldg(reinterpret_cast<const uint2*>(p) + q);
            const float2 t0 = __half22float2(*reinterpret_cast<const __half2*>(&v.x));
            const float2 t1 = __half22float2(*reinterpret_cast<const __half2*>(&v.y));
            f[4 * q] = t0.x; f[4 * q + 1] = t0.y; f[4 * q + 2] = t1.x; f[4 * q + 3] = t1.y;
        }
    } else {
#pragma unroll
        for (int e = 0; e < 16; ++e) f[e] = __half2float(__ldg(p + e));
    }
}
template <> __device__ __forceinline__ void load16<float>(const float* p, float (&f)[16]) {
    if ((reinterpret_cast<uintptr_t>(p) & 15) == 0) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(p) + q);
            f[4 * q] = v.x; f[4 * q + 1] = v.y; f[4 * q + 2] = v.z; f[4 * q + 3] = v.w;
        }
    } else {
#pragma unroll
        for (int e = 0; e < 16; ++e) f[e] = __ldg(p + e);
    }
}

// clamp_(0,1) -> *255.0 -> numpy round (half-even) -> uint8
__device__ __forceinline__ uint32_t quant255(float v) {
    v = fminf(fmaxf(v, 0.f), 1.f);
    return (uint32_t)__float2int_rn(__fmul_rn(v, 255.0f));
}

template <typename InT>
__global__ void __launch_bounds__(K4_THREADS)
k4_stitch_kernel(const InT* __restrict__ tiles_out, const int32_t* __restrict__ table, int scale,
                 uint8_t* __restrict__ out, int out_h, int out_w, int64_t out_pitch, int64_t tiles_image_stride,
                 int64_t out_image_pitch) {
    const int32_t* t = table + (size_t)blockIdx.y * TT;
    tiles_out += (size_t)blockIdx.z * tiles_image_stride;  // blockIdx.z = image of the batch
    out += (size_t)blockIdx.z * out_image_pitch;
    const int px0 = t[0], py0 = t[1], pw = t[2], ph = t[3], ix0 = t[4], iy0 = t[5], iw = t[6], ih = t[7];
    const InT* src = tiles_out + off64(t, 10);
    const int tw = pw * scale, th = ph * scale;
    const int ox0 = ix0 * scale, oy0 = iy0 * scale;
    const int ow = min(iw * scale, out_w - ox0), oh = min(ih * scale, out_h - oy0);  // post_process strips the pads
    if (ow <= 0 || oh <= 0) return;
    const int tx0 = (ix0 - px0) * scale, ty0 = (iy0 - py0) * scale;
    const int vecs = (ow + 15) >> 4;
    const int items = oh * vecs;
    const size_t plane = (size_t)th * tw;
    for (int it = blockIdx.x * K4_THREADS + threadIdx.x; it < items; it += gridDim.x * K4_THREADS) {
        const int y = it / vecs, x = (it - y * vecs) << 4;
        const int nx = min(16, ow - x);
        const InT* r = src + (size_t)(ty0 + y) * tw + tx0 + x;  // R plane; G at +plane, B at +2*plane
        uint8_t* o = out + (size_t)(oy0 + y) * out_pitch + (size_t)(ox0 + x) * 3;
        if (sizeof(InT) == 2 && nx == 16 && (reinterpret_cast<uintptr_t>(o) & 15) == 0 && (reinterpret_cast<uintptr_t>(r) & 3) == 0 && (plane & 1) == 0) {
            // packed-half path: clamp in half2 (exact), then ONE fp16 FMA v*255 + 1024 — the product is exact inside the FMA
            // and the sum is rounded once at ulp 1 (round-half-even), i.e. exactly numpy's (clamp(v)*255.0).round(); the
            // integer is the low byte of the result (0x6400 + n).  Bytes are gathered into BGR order with PRMT.
            uint32_t hw[3][8];  // hw[ch][m] = quantised pixels (2m, 2m+1) of BGR channel ch
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) {
                const __half* pp = reinterpret_cast<const __half*>(r) + (size_t)(2 - ch) * plane;
                const uintptr_t a = reinterpret_cast<uintptr_t>(pp);
                uint32_t raw[8];
                if ((a & 15) == 0) {
                    const uint4 v0 = __ldg(reinterpret_cast<const uint4*>(pp)), v1 = __ldg(reinterpret_cast<const uint4*>(pp) + 1);
                    raw[0] = v0.x; raw[1] = v0.y; raw[2] = v0.z; raw[3] = v0.w; raw[4] = v1.x; raw[5] = v1.y; raw[6] = v1.z; raw[7] = v1.w;
                } else if ((a & 7) == 0) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) { const uint2 v = __ldg(reinterpret_cast<const uint2*>(pp) + q); raw[2 * q] = v.x; raw[2 * q + 1] = v.y; }
                } else {
#pragma unroll
                    for (int q = 0; q < 8; ++q) raw[q] = __ldg(reinterpret_cast<const uint32_t*>(pp) + q);
                }
                const __half2 zero = __float2half2_rn(0.f), one = __float2half2_rn(1.f);
                const __half2 k255 = __float2half2_rn(255.f), k1024 = __float2half2_rn(1024.f);
#pragma unroll
                for (int m = 0; m < 8; ++m) {
                    __half2 v = *reinterpret_cast<const __half2*>(&raw[m]);
                    v = __hmin2(__hmax2(v, zero), one);
                    v = __hfma2(v, k255, k1024);
                    hw[ch][m] = *reinterpret_cast<const uint32_t*>(&v);
                }
            }
            uint32_t w[12];
#pragma unroll
            for (int k = 0; k < 12; ++k) {
                // output byte i = 4k+j: pixel i/3, BGR channel i%3 -> byte (px&1)*2 of hw[ch][px>>1]
                const int i0 = 4 * k, i1 = i0 + 1, i2 = i0 + 2, i3 = i0 + 3;
                const uint32_t lo = __byte_perm(hw[i0 % 3][(i0 / 3) >> 1], hw[i1 % 3][(i1 / 3) >> 1],
                                                (uint32_t)(((i0 / 3) & 1) * 2) | ((uint32_t)(4 + ((i1 / 3) & 1) * 2) << 4));
                const uint32_t hi = __byte_perm(hw[i2 % 3][(i2 / 3) >> 1], hw[i3 % 3][(i3 / 3) >> 1],
                                                (uint32_t)(((i2 / 3) & 1) * 2) | ((uint32_t)(4 + ((i3 / 3) & 1) * 2) << 4));
                w[k] = __byte_perm(lo, hi, 0x5410);
            }
            uint4* o4 = reinterpret_cast<uint4*>(o);
            o4[0] = make_uint4(w[0], w[1], w[2], w[3]);
            o4[1] = make_uint4(w[4], w[5], w[6], w[7]);
            o4[2] = make_uint4(w[8], w[9], w[10], w[11]);
        } else if (nx == 16 && (reinterpret_cast<uintptr_t>(o) & 15) == 0) {
            float f[3][16];
            load16<InT>(r, f[2]);               // R plane -> byte 2 of each BGR pixel
            load16<InT>(r + plane, f[1]);       // G
            load16<InT>(r + 2 * plane, f[0]);   // B
            uint32_t w[12];
#pragma unroll
            for (int k = 0; k < 12; ++k) {
                uint32_t word = 0;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int i = 4 * k + j;    // output byte i = pixel i/3, BGR channel i%3
                    word |= quant255(f[i % 3][i / 3]) << (8 * j);
                }
                w[k] = word;
            }
            uint4* o4 = reinterpret_cast<uint4*>(o);
            o4[0] = make_uint4(w[0], w[1], w[2], w[3]);
            o4[1] = make_uint4(w[4], w[5], w[6], w[7]);
            o4[2] = make_uint4(w[8], w[9], w[10], w[11]);
        } else {
            for (int e = 0; e < nx; ++e) {
                o[3 * e + 0] = (uint8_t)quant255(ld_as_float<InT>(r + 2 * plane + e));
                o[3 * e + 1] = (uint8_t)quant255(ld_as_float<InT>(r + plane + e));
                o[3 * e + 2] = (uint8_t)quant255(ld_as_float<InT>(r + e));
            }
        }
    }
}

// ---- (f2) key-point attach (utils/yolo_wrapper.py:168-217): one CTA per image, one WARP per merged box ----
// lanes stride over the image's detections; the reference's sequential rule (exact key -> LAST detection with that box;
// else the FIRST detection reaching the maximum IoU, if > 0.5, then the LAST detection sharing that box) is recovered
// by warp reductions on (iou, -index) / max(index).
__global__ void __launch_bounds__(256)
attach_keypoints_kernel(const float* __restrict__ merged, int merged_stride,
                        const int32_t* __restrict__ m_off, const int32_t* __restrict__ m_cnt,
                        const float* __restrict__ dets, int det_stride,
                        const int32_t* __restrict__ d_off, const int32_t* __restrict__ d_cnt,
                        int32_t* __restrict__ src_index, int smem_boxes) {
    extern __shared__ float4 s_det[];  // the image's detection boxes, staged once (rows are 96 B apart in HBM: a warp-wide
                                       // scan straight from global memory touches one sector per lane and per box)
    const int s = blockIdx.x;
    const int mo = m_off[s], mn = m_cnt[s], dof = d_off[s], dn = d_cnt[s];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    if ((int)(blockIdx.y * nwarps) >= mn) return;  // nothing for this CTA (uniform)
    const bool staged = dn <= smem_boxes;  // uniform over the CTA
    if (staged) {
        for (int j = threadIdx.x; j < dn; j += blockDim.x)
            s_det[j] = *reinterpret_cast<const float4*>(dets + (size_t)(dof + j) * det_stride);
        __syncthreads();
    }
    auto det_box = [&](int j) -> float4 {
        return staged ? s_det[j] : *reinterpret_cast<const float4*>(dets + (size_t)(dof + j) * det_stride);
    };
    // blockIdx.y splits one image's merged boxes over several CTAs (an image with thousands of boxes must not be one CTA's job)
    for (int i = blockIdx.y * nwarps + warp; i < mn; i += nwarps * gridDim.y) {
        const float4 mb = *reinterpret_cast<const float4*>(merged + (size_t)(mo + i) * merged_stride);
        const double bx1 = mb.x, by1 = mb.y, bx2 = mb.z, by2 = mb.w;
        int exact = -1, best = -1;
        double best_iou = 0.0;
        for (int j = lane; j < dn; j += 32) {
            const float4 d = det_box(j);
            if (d.x == mb.x && d.y == mb.y && d.z == mb.z && d.w == mb.w) { exact = j; continue; }
            if (fminf(mb.z, d.z) < fmaxf(mb.x, d.x) || fminf(mb.w, d.w) < fmaxf(mb.y, d.y)) continue;  // IoU 0 never wins
            const double ix1 = fmax(bx1, (double)d.x), iy1 = fmax(by1, (double)d.y);
            const double ix2 = fmin(bx2, (double)d.z), iy2 = fmin(by2, (double)d.w);
            const double inter = (ix2 - ix1) * (iy2 - iy1);
            const double uni = (bx2 - bx1) * (by2 - by1) + ((double)d.z - (double)d.x) * ((double)d.w - (double)d.y) - inter;
            const double iou = uni > 0 ? inter / uni : 0.0;
            if (iou > best_iou) { best_iou = iou; best = j; }  // ascending j per lane: keeps the lane's first maximum
        }
        exact = __reduce_max_sync(0xffffffffu, exact);
        int res = -1;
        if (exact >= 0) {
            res = exact;
        } else {
#pragma unroll
            for (int o = 16; o >= 1; o >>= 1) {
                const double oi = __shfl_xor_sync(0xffffffffu, best_iou, o);
                const int ob = __shfl_xor_sync(0xffffffffu, best, o);
                if (oi > best_iou || (oi == best_iou && ob >= 0 && (best < 0 || ob < best))) { best_iou = oi; best = ob; }
            }
            if (best >= 0 && best_iou > 0.5) {
                // the cache is a dict keyed by the box: the value is the LAST detection inserted with that box
                const float4 d = det_box(best);
                int last = best;
                for (int j = best + 1 + lane; j < dn; j += 32) {
                    const float4 e = det_box(j);
                    if (e.x == d.x && e.y == d.y && e.z == d.z && e.w == d.w) last = j;
                }
                res = __reduce_max_sync(0xffffffffu, last);
            }
        }
        if (lane == 0) src_index[mo + i] = res < 0 ? -1 : dof + res;
    }
}

int launch_crop_tma(fsd_context* h, const uint8_t* bgr, int H, int W, int64_t row_pitch, int pre_h, int pre_w,
                    const int32_t* table_dev, const int32_t* table_host, int T, void* tiles, int n_images, int64_t image_pitch,
                    int64_t tiles_image_stride, int64_t units, cudaStream_t stream, bool* taken);  // k4_crop_tma.cu

}  // namespace fsd

using namespace fsd;

// RealESRGANer.pre_process + tile_process index arithmetic; all integers, host side.
extern "C" int fsd_esrgan_tile_table(int H, int W, int scale, int tile, int tile_pad, int pre_pad, int32_t* table,
                                     int cap, int* n_tiles, int32_t padded_hw[2]) {
    FSD_CHECK_ARG(n_tiles != nullptr, "fsd_esrgan_tile_table: n_tiles is null");
    FSD_CHECK_ARG(H > 0 && W > 0 && scale > 0 && tile > 0 && tile_pad >= 0 && pre_pad >= 0, "fsd_esrgan_tile_table: bad sizes");
    int Hp = H + pre_pad, Wp = W + pre_pad;
    const int mod = scale == 2 ? 2 : (scale == 1 ? 4 : 0);
    if (mod) {
        if (Hp % mod) Hp += mod - Hp % mod;
        if (Wp % mod) Wp += mod - Wp % mod;
    }
    FSD_CHECK_ARG(Hp - H < H && Wp - W < W, "fsd_esrgan_tile_table: reflect padding needs pad < image size");
    if (padded_hw) { padded_hw[0] = Hp; padded_hw[1] = Wp; }
    const int tiles_x = (Wp + tile - 1) / tile, tiles_y = (Hp + tile - 1) / tile;
    int count = 0;
    int64_t in_off = 0, out_off = 0;
    for (int y = 0; y < tiles_y; ++y)
        for (int x = 0; x < tiles_x; ++x) {
            const int ix0 = x * tile, iy0 = y * tile;
            const int ix1 = ix0 + tile < Wp ? ix0 + tile : Wp, iy1 = iy0 + tile < Hp ? iy0 + tile : Hp;
            const int px0 = ix0 - tile_pad > 0 ? ix0 - tile_pad : 0, py0 = iy0 - tile_pad > 0 ? iy0 - tile_pad : 0;
            const int px1 = ix1 + tile_pad < Wp ? ix1 + tile_pad : Wp, py1 = iy1 + tile_pad < Hp ? iy1 + tile_pad : Hp;
            if (table && count < cap) {
                int32_t* t = table + (size_t)count * TT;
                t[0] = px0; t[1] = py0; t[2] = px1 - px0; t[3] = py1 - py0;
                t[4] = ix0; t[5] = iy0; t[6] = ix1 - ix0; t[7] = iy1 - iy0;
                t[8] = (int32_t)(in_off & 0xffffffff); t[9] = (int32_t)(in_off >> 32);
                t[10] = (int32_t)(out_off & 0xffffffff); t[11] = (int32_t)(out_off >> 32);
            }
            const int64_t e = 3LL * (px1 - px0) * (py1 - py0);
            in_off += (e + 7) / 8 * 8;                        // keep every tile 16-byte aligned (fp16)
            out_off += (e * scale * scale + 7) / 8 * 8;
            ++count;
        }
    *n_tiles = count;
    if (table && count > cap) { set_error("fsd_esrgan_tile_table: %d tiles exceed capacity %d", count, cap); return FSD_ERR_CAPACITY; }
    return FSD_OK;
}

extern "C" int fsd_esrgan_crop(fsd_handle_t h, const uint8_t* bgr, int H, int W, int64_t row_pitch, int pre_h,
                               int pre_w, const int32_t* table_dev, const int32_t* table_host, int T, int dtype,
                               void* tiles, int n_images, int64_t image_pitch, int64_t tiles_image_stride,
                               void* stream_) {
    FSD_CHECK_ARG(h && bgr && table_dev && table_host && tiles, "fsd_esrgan_crop: null argument");
    FSD_CHECK_ARG(H > 0 && W > 0 && T >= 0 && row_pitch >= (int64_t)W * 3 && pre_h >= H && pre_w >= W, "fsd_esrgan_crop: bad sizes");
    FSD_CHECK_ARG(dtype == FSD_F16 || dtype == FSD_F32, "fsd_esrgan_crop: bad dtype");
    FSD_CHECK_ARG(n_images >= 0 && n_images <= 65535 && (n_images <= 1 || (image_pitch >= row_pitch * H && tiles_image_stride > 0 && tiles_image_stride % 8 == 0)),
                  "fsd_esrgan_crop: bad batch strides (tile stride must be a positive multiple of 8 elements)");
    if (T == 0 || n_images == 0) return FSD_OK;
    int max_items = 1;
    for (int i = 0; i < T; ++i) {
        const int items = table_host[i * TT + 3] * ((table_host[i * TT + 2] + 7) / 8);
        if (items > max_items) max_items = items;
    }
    dim3 grid((max_items + K4_THREADS - 1) / K4_THREADS, T, n_images);
    FSD_CUDA(cudaSetDevice(h->device));
    int64_t crop_bytes = (int64_t)H * W * 3;  // algorithmic: source once + every tile once
    for (int i = 0; i < T; ++i) crop_bytes += (int64_t)3 * table_host[i * TT + 2] * table_host[i * TT + 3] * (dtype == FSD_F16 ? 2 : 4);
    if (dtype == FSD_F16) {  // bulk-copy pipeline (TMA boxes in, bulk stores out) when every tile qualifies: k4_crop_tma.cu
        bool taken = false;
        const int rc = launch_crop_tma(h, bgr, H, W, row_pitch, pre_h, pre_w, table_dev, table_host, T, tiles, n_images, image_pitch,
                                       tiles_image_stride, crop_bytes * n_images, (cudaStream_t)stream_, &taken);
        if (rc != FSD_OK || taken) return rc;
    }
    TimedLaunch timed(h, FSD_KERNEL_ESRGAN_CROP, crop_bytes * n_images, n_images, (cudaStream_t)stream_);
    if (dtype == FSD_F16) k4_crop_kernel<__half><<<grid, K4_THREADS, 0, (cudaStream_t)stream_>>>(bgr, H, W, row_pitch, pre_h, pre_w, table_dev, (__half*)tiles, image_pitch, tiles_image_stride);
    else k4_crop_kernel<float><<<grid, K4_THREADS, 0, (cudaStream_t)stream_>>>(bgr, H, W, row_pitch, pre_h, pre_w, table_dev, (float*)tiles, image_pitch, tiles_image_stride);
    FSD_CUDA(cudaGetLastError());
    h->launches += 1;
    return FSD_OK;
}

extern "C" int fsd_esrgan_stitch(fsd_handle_t h, const void* tiles_out, const int32_t* table_dev,
                                 const int32_t* table_host, int T, int scale, int dtype, uint8_t* out_bgr, int out_h,
                                 int out_w, int64_t out_pitch, int n_images, int64_t tiles_image_stride,
                                 int64_t out_image_pitch, void* stream_) {
    FSD_CHECK_ARG(h && tiles_out && table_dev && table_host && out_bgr, "fsd_esrgan_stitch: null argument");
    FSD_CHECK_ARG(T >= 0 && scale > 0 && out_h > 0 && out_w > 0 && out_pitch >= (int64_t)out_w * 3, "fsd_esrgan_stitch: bad sizes");
    FSD_CHECK_ARG(dtype == FSD_F16 || dtype == FSD_F32, "fsd_esrgan_stitch: bad dtype");
    FSD_CHECK_ARG(n_images >= 0 && n_images <= 65535 && (n_images <= 1 || (out_image_pitch >= out_pitch * out_h && tiles_image_stride > 0 && tiles_image_stride % 8 == 0)),
                  "fsd_esrgan_stitch: bad batch strides (tile stride must be a positive multiple of 8 elements)");
    if (T == 0 || n_images == 0) return FSD_OK;
    int max_items = 1;
    for (int i = 0; i < T; ++i) {
        const int items = table_host[i * TT + 7] * scale * ((table_host[i * TT + 6] * scale + 15) / 16);
        if (items > max_items) max_items = items;
    }
    dim3 grid((max_items + K4_THREADS - 1) / K4_THREADS, T, n_images);
    FSD_CUDA(cudaSetDevice(h->device));
    int64_t stitch_bytes = (int64_t)out_h * out_w * 3;  // algorithmic: every tile interior once + the output once
    for (int i = 0; i < T; ++i) stitch_bytes += (int64_t)3 * table_host[i * TT + 6] * scale * table_host[i * TT + 7] * scale * (dtype == FSD_F16 ? 2 : 4);
    TimedLaunch timed(h, FSD_KERNEL_ESRGAN_STITCH, stitch_bytes * n_images, n_images, (cudaStream_t)stream_);
    if (dtype == FSD_F16) k4_stitch_kernel<__half><<<grid, K4_THREADS, 0, (cudaStream_t)stream_>>>((const __half*)tiles_out, table_dev, scale, out_bgr, out_h, out_w, out_pitch, tiles_image_stride, out_image_pitch);
    else k4_stitch_kernel<float><<<grid, K4_THREADS, 0, (cudaStream_t)stream_>>>((const float*)tiles_out, table_dev, scale, out_bgr, out_h, out_w, out_pitch, tiles_image_stride, out_image_pitch);
    FSD_CUDA(cudaGetLastError());
    h->launches += 1;
    return FSD_OK;
}

extern "C" int fsd_attach_keypoints(fsd_handle_t h, const float* merged, int merged_stride, const int32_t* m_off,
                                    const int32_t* m_cnt, const float* dets, int det_stride, const int32_t* d_off,
                                    const int32_t* d_cnt, int S, int32_t* src_index, void* stream_) {
    FSD_CHECK_ARG(h && merged && m_off && m_cnt && dets && d_off && d_cnt && src_index, "fsd_attach_keypoints: null argument");
    FSD_CHECK_ARG(S >= 0 && merged_stride >= 4 && det_stride >= 4 && merged_stride % 4 == 0 && det_stride % 4 == 0,
                  "fsd_attach_keypoints: row strides must be multiples of 4 floats (16-byte rows)");
    if (S == 0) return FSD_OK;
    FSD_CUDA(cudaSetDevice(h->device));
    const int smem_boxes = 4096;  // 64 KB: images with more per-slice detections scan global memory instead (config 3's 9900)
    FSD_CUDA(cudaFuncSetAttribute(attach_keypoints_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_boxes * 16));
    TimedLaunch timed(h, FSD_KERNEL_ATTACH, S, 0, (cudaStream_t)stream_);
    attach_keypoints_kernel<<<dim3(S, 8), 256, smem_boxes * 16, (cudaStream_t)stream_>>>(merged, merged_stride, m_off, m_cnt, dets, det_stride, d_off, d_cnt, src_index, smem_boxes);
    FSD_CUDA(cudaGetLastError());
    h->launches += 1;
    return FSD_OK;
}
