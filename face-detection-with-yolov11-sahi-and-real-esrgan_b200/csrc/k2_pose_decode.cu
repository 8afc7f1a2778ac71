// Kernel 2 — fused YOLO pose-head decode.
//
// 2a  fsd_pose_decode:   pass 1: per anchor `cls logit >= gate` (the logit form of sigmoid(cls) > conf, found on the device
//                        with the same sigmoid) + warp-aggregated compaction: survivors get their row slot, one atomic per
//                        warp; pass 2: ONE WARP PER SURVIVOR over the whole grid: 32 lanes fetch the 64 DFL logits + 15
//                        key-point values in one round trip, 16-lane shuffle softmax-expectation, anchor/stride decode,
//                        xywh -> xyxy exactly in ultralytics' operation order (list order is fixed later by the anchor index).
// 2b  fsd_finalize_dets: for the rows Kernel 3 stage 1 kept: scale_boxes / scale_coords / clip, int()
//                        truncation (utils/yolo_wrapper.py:138), sahi clamp, + slice shift, packed per image in
//                        (entry, rank) order — the order the reference appends ObjectPredictions in.
// All arithmetic is fp32 with explicitly rounded (non-contracted) operations so results track torch's fp32 ops.
#include <cstring>

#include "fsd_common.cuh"

namespace fsd {

constexpr int K2_THREADS = 256;
constexpr int ROW = 24;  // floats per candidate / detection row

struct K2Levels {
    const void* box[3];
    const void* cls[3];
    const void* kpt[3];
    int h[3], w[3];
    int a_begin[4];  // first anchor index of each level, a_begin[3] = A
};

template <typename T> __device__ __forceinline__ float ldf(const T* p, size_t i);
template <> __device__ __forceinline__ float ldf<float>(const float* p, size_t i) { return __ldg(p + i); }
template <> __device__ __forceinline__ float ldf<__half>(const __half* p, size_t i) { return __half2float(__ldg(p + i)); }

// exp evaluated in double and rounded once: (almost always) the correctly rounded fp32 value, which is what a
// 1-ulp libm such as the one behind torch's CPU kernels returns in the vast majority of cases too
__device__ __forceinline__ float exp_cr(float x) { return (float)exp((double)x); }
__device__ __forceinline__ float sigmoid_rn(float x) { return __fdiv_rn(1.0f, __fadd_rn(1.0f, exp_cr(-x))); }

// element (b, c, a) of a [B,C,hw] (planar) or [B,hw,C] (channels-last) level tensor
template <int LAYOUT> __device__ __forceinline__ size_t idx(int b, int c, int a, int C, int hw) {
    return LAYOUT == FSD_PLANAR ? ((size_t)b * C + c) * hw + a : ((size_t)b * hw + a) * C + c;
}

// ---- Kernel 2a, pass 1: confidence gate + compaction -------------------------------------------------------------------
// `sigmoid_rn(x) > conf` is monotone in x, so it equals `x >= x_gate` for the smallest float x_gate that passes; the host
// entry point finds x_gate ON THE DEVICE with the very same sigmoid_rn (k2_find_gate_kernel: bisection over the ordered
// float bit patterns), which makes the gate one compare per anchor instead of a double-precision exp — and provably the
// same survivor set.  Survivors get their final row slot here (one atomic per warp and entry); only the anchor index is
// written.  Pass 2 decodes them with every warp of the grid, so a slice with thousands of survivors no longer serialises
// them inside the few warps that scanned its anchors.
__device__ __forceinline__ uint32_t float_order_key(float x) {  // monotone float -> uint32
    const uint32_t u = __float_as_uint(x);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float float_from_order_key(uint32_t k) {
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

__global__ void k2_find_gate_kernel(float conf, float* __restrict__ x_gate) {
    // smallest float (excluding NaN) with sigmoid_rn(x) > conf; +inf when none does (conf >= 1)
    uint32_t lo = float_order_key(-INFINITY), hi = float_order_key(INFINITY);  // invariant: answer in (lo, hi] or none
    if (!(sigmoid_rn(INFINITY) > conf)) { *x_gate = INFINITY; return; }
    if (sigmoid_rn(-INFINITY) > conf) { *x_gate = -INFINITY; return; }
    while (hi - lo > 1) {
        const uint32_t mid = lo + (hi - lo) / 2;
        if (sigmoid_rn(float_from_order_key(mid)) > conf) hi = mid; else lo = mid;
    }
    *x_gate = float_from_order_key(hi);
}

template <typename T>
__global__ void __launch_bounds__(K2_THREADS)
k2_gate_kernel(const K2Levels L, int B, const float* __restrict__ x_gate_p, float* __restrict__ cand, int cap,
               int* __restrict__ count) {
    const int A = L.a_begin[3];
    const int b = blockIdx.y;
    const int lane = threadIdx.x & 31;
    const int a = blockIdx.x * K2_THREADS + threadIdx.x;  // global anchor index of this thread
    const float x_gate = __ldg(x_gate_p);
    bool pass = false;
    if (a < A) {
        int lvl = 0;
        if (a >= L.a_begin[1]) lvl = 1;
        if (a >= L.a_begin[2]) lvl = 2;
        const int hw = L.h[lvl] * L.w[lvl];
        // one class: the class tensor is [B, hw] in either layout
        pass = ldf<T>(reinterpret_cast<const T*>(L.cls[lvl]), (size_t)b * hw + (a - L.a_begin[lvl])) >= x_gate;
    }
    const unsigned ballot = __ballot_sync(0xffffffffu, pass);
    if (ballot == 0) return;
    int base = 0;
    if (lane == 0) base = atomicAdd(count + b, __popc(ballot));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (pass) {
        const int pos = base + __popc(ballot & ((1u << lane) - 1u));
        if (pos < cap) cand[((size_t)b * cap + pos) * ROW + 5] = __int_as_float(a);
    }
}

// ---- Kernel 2a, pass 2: one warp per survivor, taken from every entry's list by the whole grid -------------------------
template <typename T, int LAYOUT>
__global__ void __launch_bounds__(K2_THREADS)
k2_pose_decode_kernel(const K2Levels L, int B, float* __restrict__ cand, int cap, const int* __restrict__ count) {
    const int lane = threadIdx.x & 31;
    const int gw = (blockIdx.x * K2_THREADS + threadIdx.x) >> 5, nw = (gridDim.x * K2_THREADS) >> 5;
    // warps walk the concatenation of all lists: survivor t of the grid = entry b, slot t - start(b).  The entry is found with
    // lane-parallel loads of 32 counters + a warp scan per step (a serial walk over the counters is a chain of dependent L2
    // loads: ~50 us for 96 entries, more than the decode itself)
    for (int t = gw;; t += nw) {
        int b = -1, slot = 0, start = 0;
        for (int base = 0; base < B; base += 32) {
            const int c = base + lane < B ? min(__ldg(count + base + lane), cap) : 0;
            int incl = c;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += v;
            }
            const int tot = __shfl_sync(0xffffffffu, incl, 31);
            if (t < start + tot) {
                const unsigned m = __ballot_sync(0xffffffffu, t < start + incl);
                const int l = __ffs(m) - 1;
                b = base + l;
                slot = t - (start + __shfl_sync(0xffffffffu, incl - c, l));
                break;
            }
            start += tot;
        }
        if (b < 0) break;  // t is past the last survivor (uniform over the warp)
        float* row = cand + ((size_t)b * cap + slot) * ROW;
        const int sa = __float_as_int(row[5]);
        int sl = 0;
        if (sa >= L.a_begin[1]) sl = 1;
        if (sa >= L.a_begin[2]) sl = 2;
        const int al = sa - L.a_begin[sl];
        const int w = L.w[sl], hw = L.h[sl] * w;
        const float stride = (float)(8 << sl);
        const float ax = (float)(al % w) + 0.5f, ay = (float)(al / w) + 0.5f;
        const T* boxp = reinterpret_cast<const T*>(L.box[sl]);
        // lane l holds DFL logits of channel l (sides l,t: bins 0..15 each) and channel l+32 (sides r,b)
        const float v0 = ldf<T>(boxp, idx<LAYOUT>(b, lane, al, 64, hw));
        const float v1 = ldf<T>(boxp, idx<LAYOUT>(b, lane + 32, al, 64, hw));
        float kv = 0.f;
        if (lane < 15) kv = ldf<T>(reinterpret_cast<const T*>(L.kpt[sl]), idx<LAYOUT>(b, lane, al, 15, hw));
        const float sscore = sigmoid_rn(ldf<T>(reinterpret_cast<const T*>(L.cls[sl]), (size_t)b * hw + al));
        // softmax-expectation over 16-lane groups
        float m0 = v0, m1 = v1;
#pragma unroll
        for (int o = 8; o >= 1; o >>= 1) {
            m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, o));
            m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, o));
        }
        const float e0 = exp_cr(__fsub_rn(v0, m0)), e1 = exp_cr(__fsub_rn(v1, m1));
        // torch's CPU softmax / sum accumulate the 16 bins sequentially (bin 0 first): keep that order so the
        // only remaining difference to the oracle is the exp() implementation
        const int gbase = lane & 16;
        float s0 = 0.f, s1 = 0.f;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            s0 = __fadd_rn(s0, __shfl_sync(0xffffffffu, e0, gbase + i));
            s1 = __fadd_rn(s1, __shfl_sync(0xffffffffu, e1, gbase + i));
        }
        const float bin = (float)(lane & 15);
        const float t0 = __fmul_rn(__fdiv_rn(e0, s0), bin), t1 = __fmul_rn(__fdiv_rn(e1, s1), bin);
        float d0 = 0.f, d1 = 0.f;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            d0 = __fadd_rn(d0, __shfl_sync(0xffffffffu, t0, gbase + i));
            d1 = __fadd_rn(d1, __shfl_sync(0xffffffffu, t1, gbase + i));
        }
        const float dl = __shfl_sync(0xffffffffu, d0, 0), dt = __shfl_sync(0xffffffffu, d0, 16);
        const float dr = __shfl_sync(0xffffffffu, d1, 0), db = __shfl_sync(0xffffffffu, d1, 16);
        // dist2bbox(xywh) * stride, then NMS's xywh2xyxy — same operation order as ultralytics
        const float x1 = __fsub_rn(ax, dl), y1 = __fsub_rn(ay, dt), x2 = __fadd_rn(ax, dr), y2 = __fadd_rn(ay, db);
        const float cx = __fmul_rn(__fdiv_rn(__fadd_rn(x1, x2), 2.0f), stride);
        const float cy = __fmul_rn(__fdiv_rn(__fadd_rn(y1, y2), 2.0f), stride);
        const float bw = __fmul_rn(__fsub_rn(x2, x1), stride), bh = __fmul_rn(__fsub_rn(y2, y1), stride);
        const float hw2 = __fdiv_rn(bw, 2.0f), hh2 = __fdiv_rn(bh, 2.0f);
        // key-points: lane c<15 -> (x, y, conf) of point c/3
        float kout;
        {
            const int comp = lane % 3;
            const float anc = comp == 0 ? ax : ay;
            const float xy = __fmul_rn(__fadd_rn(__fmul_rn(kv, 2.0f), __fsub_rn(anc, 0.5f)), stride);
            kout = comp == 2 ? sigmoid_rn(kv) : xy;
        }
        float val = 0.f;
        if (lane == 0) val = __fsub_rn(cx, hw2);
        else if (lane == 1) val = __fsub_rn(cy, hh2);
        else if (lane == 2) val = __fadd_rn(cx, hw2);
        else if (lane == 3) val = __fadd_rn(cy, hh2);
        else if (lane == 4) val = sscore;
        else if (lane == 5) val = __int_as_float(sa);
        const float kshift = __shfl_sync(0xffffffffu, kout, (lane + 26) & 31);  // lanes 6..20 <- kout of lanes 0..14
        if (lane >= 6 && lane < 21) val = kshift;
        if (lane < ROW) row[lane] = val;
    }
}

struct K2bParams {
    const float* cand;
    const int32_t* keep;
    const int32_t* keep_count;
    const int32_t* entry_geom;   // [B,8] shift_x, shift_y, src_w, src_h, pad_x, pad_y, full_w, full_h
    const float* entry_fgeom;    // [B,4] gain, kpt_pad_x, kpt_pad_y, 0
    const int32_t* group_range;  // [G,2] first entry, one-past-last entry
    const int32_t* group_offsets;
    float* det;
    int32_t* out_count;
    int cap, det_cap, truncate;
};

__global__ void __launch_bounds__(K2_THREADS) k2_finalize_kernel(const K2bParams p) {
    const int g = blockIdx.x;
    const int e0 = p.group_range[2 * g], e1 = p.group_range[2 * g + 1];
    int base = p.out_count[g];
    for (int e = e0; e < e1; ++e) {
        int n = p.keep_count[e];
        const int* geo = p.entry_geom + 8 * e;
        const float gain = p.entry_fgeom[4 * e], kpx = p.entry_fgeom[4 * e + 1], kpy = p.entry_fgeom[4 * e + 2];
        const float sw = (float)geo[2], sh = (float)geo[3], padx = (float)geo[4], pady = (float)geo[5];
        const int shx = geo[0], shy = geo[1], fw = geo[6], fh = geo[7];
        for (int r = threadIdx.x; r < n; r += K2_THREADS) {
            const int dst_i = base + r;
            if (dst_i >= p.det_cap) continue;
            const int src_row = p.keep[(size_t)e * p.cap + r];
            const float* c = p.cand + (size_t)src_row * ROW;
            float* d = p.det + ((size_t)p.group_offsets[g] + dst_i) * ROW;
            // scale_boxes: -= pad, /= gain, clip to the source (slice) shape
            float x1 = fminf(fmaxf(__fdiv_rn(__fsub_rn(c[0], padx), gain), 0.f), sw);
            float y1 = fminf(fmaxf(__fdiv_rn(__fsub_rn(c[1], pady), gain), 0.f), sh);
            float x2 = fminf(fmaxf(__fdiv_rn(__fsub_rn(c[2], padx), gain), 0.f), sw);
            float y2 = fminf(fmaxf(__fdiv_rn(__fsub_rn(c[3], pady), gain), 0.f), sh);
            if (p.truncate) {
                // astype(int) truncation, sahi ObjectAnnotation clamp (vs the FULL shape), then + shift
                int ix1 = max((int)x1, 0), iy1 = max((int)y1, 0);
                int ix2 = (int)x2, iy2 = (int)y2;
                if (fw > 0) { ix2 = min(ix2, fw); iy2 = min(iy2, fh); }
                x1 = (float)(ix1 + shx); y1 = (float)(iy1 + shy); x2 = (float)(ix2 + shx); y2 = (float)(iy2 + shy);
            } else {
                x1 = __fadd_rn(x1, (float)shx); y1 = __fadd_rn(y1, (float)shy);
                x2 = __fadd_rn(x2, (float)shx); y2 = __fadd_rn(y2, (float)shy);
            }
            d[0] = x1; d[1] = y1; d[2] = x2; d[3] = y2;
            d[4] = c[4];
            d[5] = __int_as_float(e);
#pragma unroll
            for (int k = 0; k < 5; ++k) {  // scale_coords + clip, then + shift in fp32 (numpy float32 + int)
                float kx = fminf(fmaxf(__fdiv_rn(__fsub_rn(c[6 + 3 * k], kpx), gain), 0.f), sw);
                float ky = fminf(fmaxf(__fdiv_rn(__fsub_rn(c[7 + 3 * k], kpy), gain), 0.f), sh);
                d[6 + 3 * k] = __fadd_rn(kx, (float)shx);
                d[7 + 3 * k] = __fadd_rn(ky, (float)shy);
                d[8 + 3 * k] = c[8 + 3 * k];
            }
            d[21] = __int_as_float(src_row);
            d[22] = 0.f; d[23] = 0.f;
        }
        base += n;
    }
    __syncthreads();
    if (threadIdx.x == 0) p.out_count[g] = min(base, p.det_cap);
}

// ---- result packing: merged boxes + attached key-points of all images into one contiguous row list ---------
__global__ void __launch_bounds__(K2_THREADS)
k2_pack_kernel(const float* __restrict__ det, const int32_t* __restrict__ group_offsets,
               const int32_t* __restrict__ keep, const int32_t* __restrict__ keep_count,
               const float* __restrict__ mboxes, const float* __restrict__ mscores,
               const int32_t* __restrict__ src_index, int G, float* __restrict__ out, int32_t* __restrict__ out_offsets) {
    __shared__ int s_base;
    const int g = blockIdx.x;
    if (threadIdx.x == 0) {
        int base = 0;
        for (int i = 0; i < g; ++i) base += keep_count[i];
        s_base = base;
        out_offsets[g] = base;
        if (g == G - 1) out_offsets[G] = base + keep_count[g];
    }
    __syncthreads();
    const int base = s_base, n = keep_count[g], off = group_offsets[g];
    for (int i = threadIdx.x; i < n; i += K2_THREADS) {
        float* o = out + (size_t)(base + i) * ROW;
        const float* mb = mboxes + (size_t)(off + i) * 4;
        o[0] = mb[0]; o[1] = mb[1]; o[2] = mb[2]; o[3] = mb[3];
        o[4] = mscores[off + i];
        const int src = src_index ? src_index[off + i] : -1;
        o[5] = __int_as_float(src);
        if (src >= 0) {
            const float* d = det + (size_t)src * ROW;
#pragma unroll
            for (int k = 0; k < 15; ++k) o[6 + k] = d[6 + k];
        } else {
#pragma unroll
            for (int k = 0; k < 15; ++k) o[6 + k] = 0.f;
        }
        o[21] = __int_as_float(keep[off + i]);
        o[22] = __int_as_float(g);
        o[23] = 0.f;
    }
}

}  // namespace fsd

using namespace fsd;

extern "C" int fsd_pack_results(fsd_handle_t h, const float* det, const int32_t* group_offsets, const int32_t* keep,
                                const int32_t* keep_count, const float* merged_boxes, const float* merged_scores,
                                const int32_t* src_index, int G, float* out, int32_t* out_offsets, void* stream_) {
    FSD_CHECK_ARG(h && det && group_offsets && keep && keep_count && merged_boxes && merged_scores && out && out_offsets,
                  "fsd_pack_results: null argument");
    FSD_CHECK_ARG(G >= 0, "fsd_pack_results: bad G");
    if (G == 0) return FSD_OK;
    FSD_CUDA(cudaSetDevice(h->device));
    TimedLaunch timed(h, FSD_KERNEL_PACK, G, 0, (cudaStream_t)stream_);
    k2_pack_kernel<<<G, K2_THREADS, 0, (cudaStream_t)stream_>>>(det, group_offsets, keep, keep_count, merged_boxes,
                                                               merged_scores, src_index, G, out, out_offsets);
    FSD_CUDA(cudaGetLastError());
    h->launches += 1;
    return FSD_OK;
}

extern "C" int fsd_pose_decode(fsd_handle_t h, const void* const box[3], const void* const cls[3],
                               const void* const kpt[3], const int32_t level_hw[6], int B, int layout, int dtype,
                               float conf, float* cand, int cap_per_entry, int32_t* count, void* stream_) {
    FSD_CHECK_ARG(h && box && cls && kpt && level_hw && cand && count, "fsd_pose_decode: null argument");
    FSD_CHECK_ARG(B >= 0 && B <= 65535 && cap_per_entry > 0, "fsd_pose_decode: bad B or capacity");
    FSD_CHECK_ARG(layout == FSD_PLANAR || layout == FSD_CHANNELS_LAST, "fsd_pose_decode: bad layout");
    FSD_CHECK_ARG(dtype == FSD_F16 || dtype == FSD_F32, "fsd_pose_decode: bad dtype");
    if (B == 0) return FSD_OK;
    K2Levels L;
    int a = 0;
    for (int l = 0; l < 3; ++l) {
        FSD_CHECK_ARG(box[l] && cls[l] && kpt[l] && level_hw[2 * l] > 0 && level_hw[2 * l + 1] > 0,
                      "fsd_pose_decode: level %d is empty", l);
        L.box[l] = box[l]; L.cls[l] = cls[l]; L.kpt[l] = kpt[l];
        L.h[l] = level_hw[2 * l]; L.w[l] = level_hw[2 * l + 1];
        L.a_begin[l] = a;
        a += L.h[l] * L.w[l];
    }
    L.a_begin[3] = a;
    cudaStream_t stream = (cudaStream_t)stream_;
    FSD_CUDA(cudaSetDevice(h->device));
    FSD_CUDA(cudaMemsetAsync(count, 0, sizeof(int32_t) * B, stream));
    // the gate logit for this confidence, found once on the device with the kernel's own sigmoid (cached in the handle)
    float* x_gate = nullptr;
    {
        std::lock_guard<std::mutex> lock(h->mu);
        uint32_t bits;
        memcpy(&bits, &conf, 4);
        auto it = h->decode_gates.find(bits);
        if (it == h->decode_gates.end()) {
            cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
            FSD_CUDA(cudaStreamIsCapturing(stream, &cap));
            FSD_CHECK_ARG(cap == cudaStreamCaptureStatusNone,
                          "fsd_pose_decode: first call for conf=%g inside a stream capture (call it once eagerly first)", conf);
            FSD_CUDA(cudaMalloc(&x_gate, sizeof(float)));
            h->dev_allocs.push_back(x_gate);
            k2_find_gate_kernel<<<1, 1, 0, stream>>>(conf, x_gate);
            FSD_CUDA(cudaGetLastError());
            // once per confidence value and handle: finished before anyone (any stream, any later capture) can use the cached value
            FSD_CUDA(cudaStreamSynchronize(stream));
            h->decode_gates.emplace(bits, x_gate);
        } else {
            x_gate = it->second;
        }
    }
    TimedLaunch timed(h, FSD_KERNEL_DECODE, (int64_t)B * a * 80 * (dtype == FSD_F16 ? 2 : 4), a, stream);
    dim3 grid((a + K2_THREADS - 1) / K2_THREADS, B);
    if (dtype == FSD_F16) k2_gate_kernel<__half><<<grid, K2_THREADS, 0, stream>>>(L, B, x_gate, cand, cap_per_entry, count);
    else k2_gate_kernel<float><<<grid, K2_THREADS, 0, stream>>>(L, B, x_gate, cand, cap_per_entry, count);
    FSD_CUDA(cudaGetLastError());
    // pass 2: one warp per survivor over a grid of a few CTAs per SM (the survivor count lives on the device)
    const int grid2 = h->sm_count * 4;
    if (dtype == FSD_F16) {
        if (layout == FSD_PLANAR) k2_pose_decode_kernel<__half, FSD_PLANAR><<<grid2, K2_THREADS, 0, stream>>>(L, B, cand, cap_per_entry, count);
        else k2_pose_decode_kernel<__half, FSD_CHANNELS_LAST><<<grid2, K2_THREADS, 0, stream>>>(L, B, cand, cap_per_entry, count);
    } else {
        if (layout == FSD_PLANAR) k2_pose_decode_kernel<float, FSD_PLANAR><<<grid2, K2_THREADS, 0, stream>>>(L, B, cand, cap_per_entry, count);
        else k2_pose_decode_kernel<float, FSD_CHANNELS_LAST><<<grid2, K2_THREADS, 0, stream>>>(L, B, cand, cap_per_entry, count);
    }
    FSD_CUDA(cudaGetLastError());
    h->launches += 2;
    return FSD_OK;
}

extern "C" int fsd_finalize_dets(fsd_handle_t h, const float* cand, int cap_per_entry, const int32_t* keep,
                                 const int32_t* keep_count, int B, const int32_t* entry_geom,
                                 const float* entry_fgeom, const int32_t* group_range,
                                 const int32_t* group_offsets, int G, int truncate, float* det,
                                 int det_cap_per_group, int32_t* out_count, void* stream_) {
    FSD_CHECK_ARG(h && cand && keep && keep_count && entry_geom && entry_fgeom && group_range && group_offsets && det && out_count,
                  "fsd_finalize_dets: null argument");
    FSD_CHECK_ARG(B >= 0 && G >= 0 && cap_per_entry > 0 && det_cap_per_group > 0, "fsd_finalize_dets: bad sizes");
    if (G == 0 || B == 0) return FSD_OK;
    K2bParams p;
    p.cand = cand; p.keep = keep; p.keep_count = keep_count; p.entry_geom = entry_geom; p.entry_fgeom = entry_fgeom;
    p.group_range = group_range; p.group_offsets = group_offsets; p.det = det; p.out_count = out_count;
    p.cap = cap_per_entry; p.det_cap = det_cap_per_group; p.truncate = truncate;
    FSD_CUDA(cudaSetDevice(h->device));
    TimedLaunch timed(h, FSD_KERNEL_FINALIZE, B, G, (cudaStream_t)stream_);
    k2_finalize_kernel<<<G, K2_THREADS, 0, (cudaStream_t)stream_>>>(p);
    FSD_CUDA(cudaGetLastError());
    h->launches += 1;
    return FSD_OK;
}
