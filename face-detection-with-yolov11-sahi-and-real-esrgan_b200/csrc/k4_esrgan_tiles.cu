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

template <typename OutT>
__global__ void __launch_bounds__(K4_THREADS)
k4_crop_kernel(const uint8_t* __restrict__ img, int H, int W, int64_t pitch, int H_pre, int W_pre,
               const int32_t* __restrict__ table, OutT* __restrict__ tiles) {
    const int32_t* t = table + (size_t)blockIdx.y * TT;
    const int px0 = t[0], py0 = t[1], pw = t[2], ph = t[3];
    OutT* dst = tiles + off64(t, 8);
    const int vecs = (pw + 7) >> 3;
    const int items = ph * vecs;
    for (int it = blockIdx.x * K4_THREADS + threadIdx.x; it < items; it += gridDim.x * K4_THREADS) {
        const int y = it / vecs, x = (it - y * vecs) << 3;
        const int sy = reflect_index(py0 + y, H_pre, H);
        const uint8_t* row = img + (size_t)sy * pitch;
        const int nx = min(8, pw - x);
        float v[3][8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            if (e < nx) {
                const int sx = reflect_index(px0 + x + e, W_pre, W);
                const uint8_t* px = row + (size_t)sx * 3;
                // astype(float32) / 255 (IEEE divide), BGR -> RGB
                v[0][e] = __fdiv_rn((float)__ldg(px + 2), 255.0f);
                v[1][e] = __fdiv_rn((float)__ldg(px + 1), 255.0f);
                v[2][e] = __fdiv_rn((float)__ldg(px + 0), 255.0f);
            } else {
                v[0][e] = v[1][e] = v[2][e] = 0.f;
            }
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            OutT* o = dst + ((size_t)c * ph + y) * pw + x;
            if (nx == 8 && (reinterpret_cast<uintptr_t>(o) & 15) == 0 && sizeof(OutT) == 2) {
                __half2 h0 = __floats2half2_rn(v[c][0], v[c][1]), h1 = __floats2half2_rn(v[c][2], v[c][3]);
                __half2 h2 = __floats2half2_rn(v[c][4], v[c][5]), h3 = __floats2half2_rn(v[c][6], v[c][7]);
                uint4 pk;
                pk.x = *reinterpret_cast<uint32_t*>(&h0); pk.y = *reinterpret_cast<uint32_t*>(&h1);
                pk.z = *reinterpret_cast<uint32_t*>(&h2); pk.w = *reinterpret_cast<uint32_t*>(&h3);
                *reinterpret_cast<uint4*>(o) = pk;
            } else {
                for (int e = 0; e < nx; ++e) o[e] = to_out<OutT>(v[c][e]);
            }
        }
    }
}

template <typename T> __device__ __forceinline__ float ld_as_float(const T* p);
template <> __device__ __forceinline__ float ld_as_float<float>(const float* p) { return __ldg(p); }
template <> __device__ __forceinline__ float ld_as_float<__half>(const __half* p) { return __half2float(__ldg(p)); }

// clamp_(0,1) -> *255.0 -> numpy round (half-even) -> uint8
__device__ __forceinline__ uint32_t quant255(float v) {
    v = fminf(fmaxf(v, 0.f), 1.f);
    return (uint32_t)__float2int_rn(__fmul_rn(v, 255.0f));
}

template <typename InT>
__global__ void __launch_bounds__(K4_THREADS)
k4_stitch_kernel(const InT* __restrict__ tiles_out, const int32_t* __restrict__ table, int scale,
                 uint8_t* __restrict__ out, int out_h, int out_w, int64_t out_pitch) {
    const int32_t* t = table + (size_t)blockIdx.y * TT;
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
        if (nx == 16 && (reinterpret_cast<uintptr_t>(o) & 15) == 0) {
            uint32_t q[48];
#pragma unroll
            for (int e = 0; e < 16; ++e) {
                q[3 * e + 0] = quant255(ld_as_float<InT>(r + 2 * plane + e));  // B
                q[3 * e + 1] = quant255(ld_as_float<InT>(r + plane + e));      // G
                q[3 * e + 2] = quant255(ld_as_float<InT>(r + e));              // R
            }
            uint32_t w[12];
#pragma unroll
            for (int k = 0; k < 12; ++k) w[k] = q[4 * k] | (q[4 * k + 1] << 8) | (q[4 * k + 2] << 16) | (q[4 * k + 3] << 24);
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

// ---- (f1) WIDER-FACE bbox_overlaps with the "+1" pixel convention ------------------------------------
__global__ void bbox_overlaps_p1_kernel(const double* __restrict__ boxes, int N, const double* __restrict__ query,
                                        int K, double* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)N * K) return;
    const int n = (int)(i / K), k = (int)(i % K);
    const double* b = boxes + 4 * (size_t)n;
    const double* q = query + 4 * (size_t)k;
    const double qa = (q[2] - q[0] + 1) * (q[3] - q[1] + 1);
    const double iw = fmin(b[2], q[2]) - fmax(b[0], q[0]) + 1;
    double v = 0.0;
    if (iw > 0) {
        const double ih = fmin(b[3], q[3]) - fmax(b[1], q[1]) + 1;
        if (ih > 0) {
            const double ua = (b[2] - b[0] + 1) * (b[3] - b[1] + 1) + qa - iw * ih;
            v = iw * ih / ua;
        }
    }
    out[i] = v;
}

// ---- (f2) key-point attach (utils/yolo_wrapper.py:168-217), one CTA per image ---------------------------
__global__ void attach_keypoints_kernel(const float* __restrict__ merged, int merged_stride,
                                        const int32_t* __restrict__ m_off, const int32_t* __restrict__ m_cnt,
                                        const float* __restrict__ dets, int det_stride,
                                        const int32_t* __restrict__ d_off, const int32_t* __restrict__ d_cnt,
                                        int32_t* __restrict__ src_index) {
    const int s = blockIdx.x;
    const int mo = m_off[s], mn = m_cnt[s], dof = d_off[s], dn = d_cnt[s];
    for (int i = threadIdx.x; i < mn; i += blockDim.x) {
        const float* mb = merged + (size_t)(mo + i) * merged_stride;
        const double bx1 = mb[0], by1 = mb[1], bx2 = mb[2], by2 = mb[3];
        int exact = -1, best = -1;
        double best_iou = 0.0;
        for (int j = 0; j < dn; ++j) {
            const float* d = dets + (size_t)(dof + j) * det_stride;
            const double dx1 = d[0], dy1 = d[1], dx2 = d[2], dy2 = d[3];
            if (dx1 == bx1 && dy1 == by1 && dx2 == bx2 && dy2 == by2) { exact = j; continue; }
            const double ix1 = fmax(bx1, dx1), iy1 = fmax(by1, dy1), ix2 = fmin(bx2, dx2), iy2 = fmin(by2, dy2);
            double iou = 0.0;
            if (!(ix2 < ix1 || iy2 < iy1)) {
                const double inter = (ix2 - ix1) * (iy2 - iy1);
                const double uni = (bx2 - bx1) * (by2 - by1) + (dx2 - dx1) * (dy2 - dy1) - inter;
                iou = uni > 0 ? inter / uni : 0.0;
            }
            if (iou > best_iou) { best_iou = iou; best = j; }
        }
        int res = -1;
        if (exact >= 0) res = exact;
        else if (best >= 0 && best_iou > 0.5) {
            // the cache is a dict keyed by the box: the value is the LAST detection inserted with that box
            const float* d = dets + (size_t)(dof + best) * det_stride;
            res = best;
            for (int j = best + 1; j < dn; ++j) {
                const float* e = dets + (size_t)(dof + j) * det_stride;
                if (e[0] == d[0] && e[1] == d[1] && e[2] == d[2] && e[3] == d[3]) res = j;
            }
        }
        src_index[mo + i] = res < 0 ? -1 : dof + res;
    }
}

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
                               void* tiles, void* stream_) {
    FSD_CHECK_ARG(h && bgr && table_dev && table_host && tiles, "fsd_esrgan_crop: null argument");
    FSD_CHECK_ARG(H > 0 && W > 0 && T >= 0 && row_pitch >= (int64_t)W * 3 && pre_h >= H && pre_w >= W, "fsd_esrgan_crop: bad sizes");
    FSD_CHECK_ARG(dtype == FSD_F16 || dtype == FSD_F32, "fsd_esrgan_crop: bad dtype");
    if (T == 0) return FSD_OK;
    int max_items = 1;
    for (int i = 0; i < T; ++i) {
        const int items = table_host[i * TT + 3] * ((table_host[i * TT + 2] + 7) / 8);
        if (items > max_items) max_items = items;
    }
    dim3 grid((max_items + K4_THREADS - 1) / K4_THREADS, T);
    FSD_CUDA(cudaSetDevice(h->device));
    if (dtype == FSD_F16) k4_crop_kernel<__half><<<grid, K4_THREADS, 0, (cudaStream_t)stream_>>>(bgr, H, W, row_pitch, pre_h, pre_w, table_dev, (__half*)tiles);
    else k4_crop_kernel<float><<<grid, K4_THREADS, 0, (cudaStream_t)stream_>>>(bgr, H, W, row_pitch, pre_h, pre_w, table_dev, (float*)tiles);
    FSD_CUDA(cudaGetLastError());
    h->launches += 1;
    return FSD_OK;
}

extern "C" int fsd_esrgan_stitch(fsd_handle_t h, const void* tiles_out, const int32_t* table_dev,
                                 const int32_t* table_host, int T, int scale, int dtype, uint8_t* out_bgr, int out_h,
                                 int out_w, int64_t out_pitch, void* stream_) {
    FSD_CHECK_ARG(h && tiles_out && table_dev && table_host && out_bgr, "fsd_esrgan_stitch: null argument");
    FSD_CHECK_ARG(T >= 0 && scale > 0 && out_h > 0 && out_w > 0 && out_pitch >= (int64_t)out_w * 3, "fsd_esrgan_stitch: bad sizes");
    FSD_CHECK_ARG(dtype == FSD_F16 || dtype == FSD_F32, "fsd_esrgan_stitch: bad dtype");
    if (T == 0) return FSD_OK;
    int max_items = 1;
    for (int i = 0; i < T; ++i) {
        const int items = table_host[i * TT + 7] * scale * ((table_host[i * TT + 6] * scale + 15) / 16);
        if (items > max_items) max_items = items;
    }
    dim3 grid((max_items + K4_THREADS - 1) / K4_THREADS, T);
    FSD_CUDA(cudaSetDevice(h->device));
    if (dtype == FSD_F16) k4_stitch_kernel<__half><<<grid, K4_THREADS, 0, (cudaStream_t)stream_>>>((const __half*)tiles_out, table_dev, scale, out_bgr, out_h, out_w, out_pitch);
    else k4_stitch_kernel<float><<<grid, K4_THREADS, 0, (cudaStream_t)stream_>>>((const float*)tiles_out, table_dev, scale, out_bgr, out_h, out_w, out_pitch);
    FSD_CUDA(cudaGetLastError());
    h->launches += 1;
    return FSD_OK;
}

extern "C" int fsd_bbox_overlaps_p1(fsd_handle_t h, const double* boxes, int N, const double* query, int K,
                                    double* overlaps, void* stream_) {
    FSD_CHECK_ARG(h && N >= 0 && K >= 0, "fsd_bbox_overlaps_p1: bad arguments");
    if (N == 0 || K == 0) return FSD_OK;
    FSD_CHECK_ARG(boxes && query && overlaps, "fsd_bbox_overlaps_p1: null argument");
    const int64_t total = (int64_t)N * K;
    FSD_CUDA(cudaSetDevice(h->device));
    bbox_overlaps_p1_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream_>>>(boxes, N, query, K, overlaps);
    FSD_CUDA(cudaGetLastError());
    h->launches += 1;
    return FSD_OK;
}

extern "C" int fsd_attach_keypoints(fsd_handle_t h, const float* merged, int merged_stride, const int32_t* m_off,
                                    const int32_t* m_cnt, const float* dets, int det_stride, const int32_t* d_off,
                                    const int32_t* d_cnt, int S, int32_t* src_index, void* stream_) {
    FSD_CHECK_ARG(h && merged && m_off && m_cnt && dets && d_off && d_cnt && src_index, "fsd_attach_keypoints: null argument");
    FSD_CHECK_ARG(S >= 0 && merged_stride >= 4 && det_stride >= 4, "fsd_attach_keypoints: bad sizes");
    if (S == 0) return FSD_OK;
    FSD_CUDA(cudaSetDevice(h->device));
    attach_keypoints_kernel<<<S, 128, 0, (cudaStream_t)stream_>>>(merged, merged_stride, m_off, m_cnt, dets, det_stride, d_off, d_cnt, src_index);
    FSD_CUDA(cudaGetLastError());
    h->launches += 1;
    return FSD_OK;
}
