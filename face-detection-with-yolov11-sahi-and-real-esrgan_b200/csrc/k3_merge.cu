// Kernel 3 — batched segment NMS / GREEDYNMM / NMM with IOU | IOS match metrics.
//
// One CTA per segment (a slice for stage 1 = torchvision.ops.nms semantics inside ultralytics' NMS; an image for
// stage 2 = sahi.postprocess.combine semantics).  Phases, all on-chip for segments up to 4096 boxes:
//   1. rank:   64-bit keys (score descending, tie-break key ascending) sorted by an in-CTA bitonic network;
//   2. scan:   greedy suppression in chunks of 64 ranks: the 64x64 in-chunk match bits come from warp ballots,
//              one thread resolves the chunk with 64-bit mask operations, then every thread sweeps the not yet
//              removed lower ranks against the chunk's (<= 64) new keeps — no N x N mask is ever materialised;
//              NMM instead walks every rank and propagates claims transitively (A.2.4 `nmm`);
//   3. replay: a second bitonic sort groups merge candidates by keep in append order, one thread per keep folds
//              them with the STRICT has_match re-check against the growing union box (A.2.5), in float64.
// The match test is bit-faithful: fp64 divide/compare for sahi (exact on integral boxes), fp32 divide with the
// result promoted to double for the torchvision rule.
#include "fsd_common.cuh"

namespace fsd {

constexpr int K3_SMEM_MAX_P = 4096;   // segments up to this many boxes live entirely in shared memory
constexpr int K3_BYTES_PER_BOX = 48;  // keys 8 + vals 4 + box 16 + parent 4 + step 4 + cat 4 + keep 4 + run 4

struct K3Params {
    const float* boxes; int box_stride;
    const float* scores; int score_stride;
    const int32_t* cats; int cat_stride;
    const int32_t* tie; int tie_stride;
    const int32_t* seg_offsets;
    const int32_t* seg_counts;
    int seg_cap;
    int type, metric, cmp_strict, precision, class_agnostic, pre_cap, max_keep;
    double thr;
    int32_t* keep; int32_t* keep_count; int32_t* parent;
    float* merged_boxes; float* merged_scores; int32_t* merged_cats;
    uint8_t* workspace; size_t ws_per_segment;
    int P;  // power of two >= seg_cap
    int use_global;
};

__device__ __forceinline__ uint32_t score_key_desc(float s) {
    uint32_t u = __float_as_uint(s);
    u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);  // ascending-orderable
    return ~u;                                      // descending
}

struct MatchCfg {
    int metric, cmp_strict, precision, class_agnostic;
    double thr;
};

__device__ __forceinline__ bool match_pair(const float4 a, const float4 b, int ca, int cb, const MatchCfg& m) {
    if (!m.class_agnostic && ca != cb) return false;
    // disjoint boxes have intersection 0 -> metric 0 -> no match for any positive threshold: decided with four fp32
    // min/max and two compares (exact: comparisons only), so the fp64 path below runs for overlapping pairs only
    if (m.thr > 0.0 && (fminf(a.z, b.z) <= fmaxf(a.x, b.x) || fminf(a.w, b.w) <= fmaxf(a.y, b.y))) return false;
    if (m.precision == 1) {
        // torchvision nms_kernel: fp32 areas / intersection, quotient compared against the double threshold
        const float iw = fmaxf(__fsub_rn(fminf(a.z, b.z), fmaxf(a.x, b.x)), 0.f);
        const float ih = fmaxf(__fsub_rn(fminf(a.w, b.w), fmaxf(a.y, b.y)), 0.f);
        const float inter = __fmul_rn(iw, ih);
        if (inter <= 0.f && m.thr > 0.0) return false;
        const float aa = __fmul_rn(__fsub_rn(a.z, a.x), __fsub_rn(a.w, a.y));
        const float ab = __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y));
        const float den = m.metric == FSD_IOU ? __fsub_rn(__fadd_rn(aa, ab), inter) : fminf(aa, ab);
        const double v = (double)__fdiv_rn(inter, den);
        return m.cmp_strict ? v > m.thr : v >= m.thr;
    }
    // sahi 0.11.34: float64 on the float32-valued coordinates; metric := 0 when the denominator is not positive
    const double iw = fmax(fmin((double)a.z, (double)b.z) - fmax((double)a.x, (double)b.x), 0.0);
    const double ih = fmax(fmin((double)a.w, (double)b.w) - fmax((double)a.y, (double)b.y), 0.0);
    const double inter = iw * ih;
    if (inter <= 0.0 && m.thr > 0.0) return false;
    const double aa = ((double)a.z - (double)a.x) * ((double)a.w - (double)a.y);
    const double ab = ((double)b.z - (double)b.x) * ((double)b.w - (double)b.y);
    const double den = m.metric == FSD_IOU ? aa + ab - inter : fmin(aa, ab);
    const double v = den > 0.0 ? inter / den : 0.0;
    return m.cmp_strict ? v > m.thr : v >= m.thr;
}

// sahi.postprocess.utils.has_match: numpy float64, STRICT >, nan (0/0) compares false
__device__ __forceinline__ bool has_match_f64(const double (&k)[4], const float4 c, int metric, double thr) {
    const double iw = fmax(fmin(k[2], (double)c.z) - fmax(k[0], (double)c.x), 0.0);
    const double ih = fmax(fmin(k[3], (double)c.w) - fmax(k[1], (double)c.y), 0.0);
    const double inter = iw * ih;
    const double ak = (k[2] - k[0]) * (k[3] - k[1]);
    const double ac = ((double)c.z - (double)c.x) * ((double)c.w - (double)c.y);
    const double den = metric == FSD_IOU ? ak + ac - inter : fmin(ak, ac);
    if (den == 0.0) return false;
    return inter / den > thr;
}

__device__ void bitonic_sort(uint64_t* keys, uint32_t* vals, int P) {
    for (int k = 2; k <= P; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = threadIdx.x; t < (P >> 1); t += blockDim.x) {
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const int l = i | j;
                const bool up = (i & k) == 0;
                const uint64_t ki = keys[i], kl = keys[l];
                if ((ki > kl) == up) {
                    keys[i] = kl; keys[l] = ki;
                    const uint32_t vi = vals[i]; vals[i] = vals[l]; vals[l] = vi;
                }
            }
            __syncthreads();
        }
    }
}

__global__ void __launch_bounds__(512) k3_merge_kernel(const K3Params p) {
    extern __shared__ __align__(16) uint8_t k3_smem[];
    __shared__ uint32_t s_diag[128];   // 64 rows x 2 halves of in-chunk match bits
    __shared__ int s_klist[64];
    __shared__ int s_kcount, s_ktotal, s_stop;

    const int s = blockIdx.x;
    const int tid = threadIdx.x, T = blockDim.x;
    const int off = p.seg_offsets[s];
    int n = p.seg_counts ? p.seg_counts[s] : p.seg_cap;
    n = min(max(n, 0), p.seg_cap);
    if (n == 0) {
        if (tid == 0) p.keep_count[s] = 0;
        return;
    }
    int P = 64;
    while (P < n) P <<= 1;

    uint8_t* base = p.use_global ? p.workspace + (size_t)s * p.ws_per_segment : k3_smem;
    const size_t PP = p.use_global ? (size_t)p.P : (size_t)P;
    float4* sbox = reinterpret_cast<float4*>(base);
    uint64_t* keys = reinterpret_cast<uint64_t*>(base + 16 * PP);
    uint32_t* vals = reinterpret_cast<uint32_t*>(base + 24 * PP);
    int* parent = reinterpret_cast<int*>(base + 28 * PP);   // rank of the claiming keep, own rank for keeps, -1 none
    int* step = reinterpret_cast<int*>(base + 32 * PP);     // NMM: rank whose visit produced the claim
    int* scat = reinterpret_cast<int*>(base + 36 * PP);
    int* keepr = reinterpret_cast<int*>(base + 40 * PP);    // ranks of keeps in output order
    int* runs = reinterpret_cast<int*>(base + 44 * PP);     // first replay-list position of each keep rank

    MatchCfg mc;
    mc.metric = p.metric; mc.cmp_strict = p.cmp_strict; mc.precision = p.precision;
    mc.class_agnostic = p.class_agnostic || p.cats == nullptr; mc.thr = p.thr;

    // ---- 1. rank ------------------------------------------------------------------------------------
    for (int i = tid; i < P; i += T) {
        uint64_t key = ~0ull;
        uint32_t v = 0xffffffffu;
        if (i < n) {
            const float sc = p.scores[(size_t)(off + i) * p.score_stride];
            const uint32_t tb = p.tie ? (uint32_t)p.tie[(size_t)(off + i) * p.tie_stride] : (uint32_t)i;
            key = ((uint64_t)score_key_desc(sc) << 32) | tb;
            v = (uint32_t)i;
        }
        keys[i] = key; vals[i] = v;
    }
    __syncthreads();
    bitonic_sort(keys, vals, P);
    int m = n;
    if (p.pre_cap > 0) m = min(m, p.pre_cap);
    for (int r = tid; r < n; r += T) {
        const int g = off + (int)vals[r];
        if (r < m) {
            const float* bp = p.boxes + (size_t)g * p.box_stride;
            sbox[r] = make_float4(bp[0], bp[1], bp[2], bp[3]);
            scat[r] = p.cats ? p.cats[(size_t)g * p.cat_stride] : 0;
            parent[r] = -1;
            runs[r] = -1;
        } else {
            p.parent[g] = -1;  // cut by the pre-NMS cap
        }
    }
    if (tid == 0) { s_ktotal = 0; s_stop = 0; }
    __syncthreads();

    if (p.type != FSD_NMM) {
        // ---- 2a. greedy scan in chunks of 64 ranks --------------------------------------------------
        // `step` doubles as the removed-bit array (32 ranks per word)
        uint32_t* rem = reinterpret_cast<uint32_t*>(step);
        for (int i = tid; i < (m + 31) / 32; i += T) rem[i] = 0;
        __syncthreads();
        for (int c0 = 0; c0 < m; c0 += 64) {
            const int cn = min(64, m - c0);
            // in-chunk match bits: 64 threads per row (two warps = two 32-bit halves)
            for (int r = tid >> 6; r < 64; r += T >> 6) {
                const int q = tid & 63;
                bool bit = false;
                if (r < cn && q < cn && q > r) bit = match_pair(sbox[c0 + r], sbox[c0 + q], scat[c0 + r], scat[c0 + q], mc);
                const uint32_t bal = __ballot_sync(0xffffffffu, bit);
                if ((tid & 31) == 0) s_diag[r * 2 + (q >> 5)] = bal;
            }
            __syncthreads();
            if (tid == 0) {
                uint64_t remw = (uint64_t)rem[c0 >> 5] | ((uint64_t)(((c0 >> 5) + 1) < (m + 31) / 32 ? rem[(c0 >> 5) + 1] : 0u) << 32);
                int kc = 0, kt = s_ktotal;
                for (int b = 0; b < cn; ++b) {
                    if ((remw >> b) & 1ull) continue;
                    if (p.max_keep > 0 && kt + kc >= p.max_keep) { s_stop = 1; break; }
                    const int kr = c0 + b;
                    s_klist[kc++] = kr;
                    parent[kr] = kr;
                    const uint64_t row = (uint64_t)s_diag[2 * b] | ((uint64_t)s_diag[2 * b + 1] << 32);
                    uint64_t fresh = row & ~remw;
                    remw |= row;
                    while (fresh) {
                        const int q = __ffsll((long long)fresh) - 1;
                        fresh &= fresh - 1;
                        parent[c0 + q] = kr;
                    }
                }
                rem[c0 >> 5] = (uint32_t)remw;
                if (((c0 >> 5) + 1) < (m + 31) / 32) rem[(c0 >> 5) + 1] = (uint32_t)(remw >> 32);
                for (int i = 0; i < kc; ++i) keepr[kt + i] = s_klist[i];
                s_kcount = kc;
                s_ktotal = kt + kc;
            }
            __syncthreads();
            const int kc = s_kcount;
            if (s_stop) break;
            // sweep: every not-yet-removed lower rank against this chunk's new keeps (first match claims it)
            for (int j = c0 + cn + tid; j < m; j += T) {
                if ((rem[j >> 5] >> (j & 31)) & 1u) continue;
                const float4 bj = sbox[j];
                const int cj = scat[j];
                for (int k = 0; k < kc; ++k) {
                    const int kr = s_klist[k];
                    if (match_pair(sbox[kr], bj, scat[kr], cj, mc)) {
                        atomicOr(&rem[j >> 5], 1u << (j & 31));
                        parent[j] = kr;
                        break;
                    }
                }
            }
            __syncthreads();
        }
    } else {
        // ---- 2b. NMM: visit every rank; a claimed rank forwards its unclaimed matches to its keep -------
        for (int i = 0; i < m; ++i) {
            if (tid == 0 && parent[i] == -1) {
                parent[i] = i;
                keepr[s_ktotal] = i;
                s_ktotal = s_ktotal + 1;
            }
            __syncthreads();
            const int k = parent[i];
            const float4 bi = sbox[i];
            const int ci = scat[i];
            for (int j = tid; j < m; j += T) {
                if (j == i || parent[j] != -1) continue;
                if (match_pair(bi, sbox[j], ci, scat[j], mc)) { parent[j] = k; step[j] = i; }
            }
            __syncthreads();
        }
    }
    __syncthreads();
    const int K = s_ktotal;

    // ---- 3. outputs + merge replay ------------------------------------------------------------------
    if (p.parent) {
        for (int r = tid; r < m; r += T) {
            const int pr = parent[r];
            p.parent[off + (int)vals[r]] = pr < 0 ? -1 : off + (int)vals[pr];
        }
    }
    if (tid == 0) p.keep_count[s] = K;
    // NOTE: vals[] (rank -> local row) is still needed below, so the replay sort uses the `runs`-adjacent scratch:
    // keys[] is free after ranking and is reused for the replay keys; replay values go to `step`'s upper half is
    // not possible for NMM, hence replay values are packed into the low 32 bits of the key itself.
    if (p.type == FSD_NMS) {
        for (int i = tid; i < K; i += T) {
            const int kr = keepr[i];
            const int g = off + (int)vals[kr];
            p.keep[off + i] = g;
            const float4 b = sbox[kr];
            float* mb = p.merged_boxes + (size_t)(off + i) * 4;
            mb[0] = b.x; mb[1] = b.y; mb[2] = b.z; mb[3] = b.w;
            p.merged_scores[off + i] = p.scores[(size_t)g * p.score_stride];
            if (p.merged_cats) p.merged_cats[off + i] = scat[kr];
        }
        return;
    }
    __syncthreads();
    // replay key = (keep rank : 15 bits | append sequence : 30 bits | candidate rank : 15 bits), ascending
    for (int r = tid; r < P; r += T) {
        uint64_t key = ~0ull;
        if (r < m) {
            const int pr = parent[r];
            if (pr >= 0 && pr != r) {
                const uint64_t seq = p.type == FSD_NMM ? (uint64_t)step[r] * 32768ull + (uint64_t)(32767 - r) : (uint64_t)r;
                key = ((uint64_t)pr << 45) | (seq << 15) | (uint64_t)r;
            }
        }
        keys[r] = key;
    }
    __syncthreads();
    {   // value-less bitonic sort of the replay keys (the candidate rank is packed in the key)
        for (int k = 2; k <= P; k <<= 1)
            for (int j = k >> 1; j > 0; j >>= 1) {
                for (int t = tid; t < (P >> 1); t += T) {
                    const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                    const int l = i | j;
                    const bool up = (i & k) == 0;
                    const uint64_t ki = keys[i], kl = keys[l];
                    if ((ki > kl) == up) { keys[i] = kl; keys[l] = ki; }
                }
                __syncthreads();
            }
    }
    for (int q = tid; q < P; q += T) {
        const uint64_t key = keys[q];
        if (key == ~0ull) continue;
        const int pr = (int)(key >> 45);
        if (q == 0 || (int)(keys[q - 1] >> 45) != pr) runs[pr] = q;
    }
    __syncthreads();
    for (int i = tid; i < K; i += T) {
        const int kr = keepr[i];
        const int g = off + (int)vals[kr];
        const float4 b = sbox[kr];
        double kb[4] = {(double)b.x, (double)b.y, (double)b.z, (double)b.w};
        const float kscore = p.scores[(size_t)g * p.score_stride];
        int kcat = scat[kr];
        int q = runs[kr];
        if (q >= 0) {
            for (; q < P; ++q) {
                const uint64_t key = keys[q];
                if (key == ~0ull || (int)(key >> 45) != kr) break;
                const int cr = (int)(key & 32767ull);
                const float4 c = sbox[cr];
                if (has_match_f64(kb, c, p.metric, p.thr)) {
                    kb[0] = fmin(kb[0], (double)c.x); kb[1] = fmin(kb[1], (double)c.y);
                    kb[2] = fmax(kb[2], (double)c.z); kb[3] = fmax(kb[3], (double)c.w);
                    // merged category: the keep's unless the candidate's score is not lower (sahi get_merged_category)
                    const float cscore = p.scores[(size_t)(off + (int)vals[cr]) * p.score_stride];
                    if (!(kscore > cscore)) kcat = scat[cr];
                }
            }
        }
        p.keep[off + i] = g;
        float* mb = p.merged_boxes + (size_t)(off + i) * 4;
        mb[0] = (float)kb[0]; mb[1] = (float)kb[1]; mb[2] = (float)kb[2]; mb[3] = (float)kb[3];
        p.merged_scores[off + i] = kscore;
        if (p.merged_cats) p.merged_cats[off + i] = kcat;
    }
}

static int pow2_at_least(int n) {
    int P = 64;
    while (P < n) P <<= 1;
    return P;
}

}  // namespace fsd

using namespace fsd;

extern "C" int64_t fsd_merge_workspace_bytes(int64_t N, int S, int max_segment) {
    (void)N;
    if (max_segment <= K3_SMEM_MAX_P) return 256;  // unused, but keep the pointer non-null for callers
    return (int64_t)S * pow2_at_least(max_segment) * K3_BYTES_PER_BOX + 256;
}

extern "C" int fsd_merge(fsd_handle_t h, const float* boxes, int box_stride, const float* scores, int score_stride,
                         const int32_t* cats, int cat_stride, const int32_t* tie, int tie_stride,
                         const int32_t* seg_offsets, const int32_t* seg_counts, int S, int max_segment, int type,
                         int metric, double thr, int cmp_strict, int precision, int class_agnostic, int pre_cap,
                         int max_keep, int32_t* keep, int32_t* keep_count, int32_t* parent, float* merged_boxes,
                         float* merged_scores, int32_t* merged_cats, void* workspace, int64_t workspace_bytes,
                         void* stream_) {
    FSD_CHECK_ARG(h && boxes && scores && seg_offsets && keep && keep_count && merged_boxes && merged_scores,
                  "fsd_merge: null argument");
    FSD_CHECK_ARG(type == FSD_NMS || type == FSD_GREEDYNMM || type == FSD_NMM, "fsd_merge: unknown merge type %d", type);
    FSD_CHECK_ARG(metric == FSD_IOU || metric == FSD_IOS, "fsd_merge: unknown match metric %d", metric);
    FSD_CHECK_ARG(S >= 0 && max_segment >= 0 && box_stride >= 4 && score_stride >= 1, "fsd_merge: bad sizes");
    if (S == 0 || max_segment == 0) {
        if (S > 0) FSD_CUDA(cudaMemsetAsync(keep_count, 0, sizeof(int32_t) * S, (cudaStream_t)stream_));
        return FSD_OK;
    }
    if (max_segment > 32768) {
        set_error("fsd_merge: segments of more than 32768 boxes are not supported (got %d)", max_segment);
        return FSD_ERR_CAPACITY;
    }
    K3Params p;
    p.boxes = boxes; p.box_stride = box_stride; p.scores = scores; p.score_stride = score_stride;
    p.cats = cats; p.cat_stride = cat_stride > 0 ? cat_stride : 1; p.tie = tie; p.tie_stride = tie_stride > 0 ? tie_stride : 1;
    p.seg_offsets = seg_offsets; p.seg_counts = seg_counts; p.seg_cap = max_segment;
    p.type = type; p.metric = metric; p.cmp_strict = cmp_strict; p.precision = precision;
    p.class_agnostic = class_agnostic; p.pre_cap = pre_cap; p.max_keep = max_keep; p.thr = thr;
    p.keep = keep; p.keep_count = keep_count; p.parent = parent; p.merged_boxes = merged_boxes;
    p.merged_scores = merged_scores; p.merged_cats = merged_cats;
    p.P = pow2_at_least(max_segment);
    p.use_global = p.P > K3_SMEM_MAX_P;
    p.workspace = reinterpret_cast<uint8_t*>(workspace);
    p.ws_per_segment = (size_t)p.P * K3_BYTES_PER_BOX;
    if (p.use_global) {
        FSD_CHECK_ARG(workspace && workspace_bytes >= fsd_merge_workspace_bytes(0, S, max_segment),
                      "fsd_merge: workspace too small (%lld bytes needed)", (long long)fsd_merge_workspace_bytes(0, S, max_segment));
        if (((uintptr_t)workspace & 15) != 0) { set_error("fsd_merge: workspace must be 16-byte aligned"); return FSD_ERR_ALIGN; }
    }
    const size_t smem = p.use_global ? 0 : (size_t)p.P * K3_BYTES_PER_BOX;
    const int threads = p.P <= 256 ? 128 : (p.P <= 1024 ? 256 : 512);
    FSD_CUDA(cudaSetDevice(h->device));
    FSD_CUDA(cudaFuncSetAttribute(k3_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, K3_SMEM_MAX_P * K3_BYTES_PER_BOX));
    k3_merge_kernel<<<S, threads, smem, (cudaStream_t)stream_>>>(p);
    FSD_CUDA(cudaGetLastError());
    h->launches += 1;
    return FSD_OK;
}
