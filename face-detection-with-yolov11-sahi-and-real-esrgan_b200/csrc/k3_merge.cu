// Kernel 3 — batched segment NMS / GREEDYNMM / NMM with IOU | IOS match metrics.
//
// One CTA per segment (a slice for stage 1 = torchvision.ops.nms semantics inside ultralytics' NMS; an image for
// stage 2 = sahi.postprocess.combine semantics); segments above 4096 boxes get a cluster of 8 CTAs.  The launch is sized
// by the caller's CAPACITY, the work by each segment's actual count read on the device: a CTA trims itself to the warps its
// segment needs, and the two kernels (shared-memory / cluster) are launched back to back, each leaving the other's segments
// alone — no host round trip decides the path.  Phases of the shared-memory kernel (k3_merge_kernel, <= 4096 boxes):
//   1. rank:   64-bit keys (score descending, tie-break key ascending), bitonic network whose stages with partner distance
//              < 32 run in registers through warp shuffles (20 block barriers for 1024 keys instead of 55);
//   2. scan:   greedy suppression in chunks of 64 ranks, TWO block barriers per chunk: one warp resolves the chunk from its
//              64x64 match bits held in registers (64 shuffle steps, no memory on the dependent chain), then all threads
//              assign in-chunk parents, sweep the not yet removed lower ranks against the chunk's (<= 64) new keeps (first
//              match claims) AND compute the next chunk's match bits in the same phase — no N x N mask is ever materialised;
//              NMM instead walks every rank and propagates claims transitively (A.2.4 `nmm`);
//   3. replay: every keep folds its candidates in append order with the STRICT has_match re-check against the growing
//              union box (A.2.5), in float64: one thread per keep, all threads scanning the claim array in lockstep
//              (broadcast reads) — no second sort for GREEDYNMM; NMM's step-major order still uses one.
// tie_rule 1 = sahi 0.11.34's lexicographic box rule for equal scores (SURVEY A.2.4 variant N, oracle/postprocess.py
// `_candidates`): an equal-score candidate whose coordinate tuple is larger is not tested, so both boxes can be kept, and
// the later keep claims the earlier equal-score keep into its merge list (then folds that keep's MERGED box).
// The match test is bit-faithful: fp64 divide/compare for sahi (exact on integral boxes), fp32 divide with the
// result promoted to double for the torchvision rule.
#include <cooperative_groups.h>

#include "fsd_common.cuh"

namespace fsd {

constexpr int K3_SMEM_MAX_P = 4096;   // segments up to this many boxes live entirely in shared memory
constexpr int K3_BYTES_PER_BOX = 48;      // shared-memory kernel: box 16 + key 8 + val 4 + parent 4 + step 4 + cat 4 + keep 4 + run 4
constexpr int K3_WS_BYTES_PER_BOX = 52;   // workspace of the cluster / fallback kernels: the same + the cluster kernel's NMM step array
constexpr int K3_CLUSTER = 8;          // CTAs (SMs) that share one large segment
constexpr int K3_SCRATCH_BYTES = 2048;  // per-segment broadcast area of the cluster kernel (after the box arrays)  // keys 8 + vals 4 + box 16 + parent 4 + step 4 + cat 4 + keep 4 + run 4

struct K3Params {
    const float* boxes; int box_stride;
    const float* scores; int score_stride;
    const int32_t* cats; int cat_stride;
    const int32_t* tie; int tie_stride;
    const int32_t* seg_offsets;
    const int32_t* seg_counts;
    int seg_cap;
    int type, metric, cmp_strict, precision, class_agnostic, pre_cap, max_keep, tie_rule;
    double thr;
    int32_t* keep; int32_t* keep_count; int32_t* parent;
    float* merged_boxes; float* merged_scores; int32_t* merged_cats;
    uint8_t* workspace; size_t ws_per_segment;
    int P;  // power of two >= seg_cap
    int use_global;
    int n_lo, n_hi;  // the shared-memory kernel takes the segments with n_lo < n <= n_hi (tiered launches, see fsd_merge)
    int cluster_min; // segments with more boxes than this go to the cluster kernel (<= K3_SMEM_MAX_P)
};

__device__ __forceinline__ uint32_t score_key_desc(float s) {
    uint32_t u = __float_as_uint(s);
    u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);  // ascending-orderable
    return ~u;                                      // descending
}

struct MatchCfg {
    int metric, cmp_strict, precision, class_agnostic, tie_rule;
    double thr;
};

__device__ __forceinline__ MatchCfg match_cfg(const K3Params& p) {
    MatchCfg mc;
    mc.metric = p.metric; mc.cmp_strict = p.cmp_strict; mc.precision = p.precision;
    mc.class_agnostic = p.class_agnostic || p.cats == nullptr; mc.thr = p.thr;
    mc.tie_rule = (p.tie_rule != 0 && p.type != FSD_NMM) ? 1 : 0;  // variant N's nmm has no tie rule
    return mc;
}

// tuple(a) > tuple(b), python tuple comparison of (x1, y1, x2, y2)
__device__ __forceinline__ bool lex_greater(const float4 a, const float4 b) {
    if (a.x != b.x) return a.x > b.x;
    if (a.y != b.y) return a.y > b.y;
    if (a.z != b.z) return a.z > b.z;
    return a.w > b.w;
}

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

// What the greedy loop applies to (current box, candidate): the match test, minus variant N's exclusion of an equal-score
// candidate whose coordinate tuple is lexicographically larger.  kcur / kcand = the score halves of the rank keys.
__device__ __forceinline__ bool suppresses(const float4 cur, const float4 cand, int ccur, int ccand, uint32_t kcur,
                                           uint32_t kcand, const MatchCfg& m) {
    if (m.tie_rule && kcur == kcand && lex_greater(cand, cur)) return false;
    return match_pair(cur, cand, ccur, ccand, m);
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


constexpr int K3_HASB = 1 << 30;  // step[] flag: this keep has an earlier equal-score keep in its merge list (tie_rule 1)

// barrier over the first T threads of the CTA (the warps a small segment does not need have already returned)
__device__ __forceinline__ void k3_bar(int T) { asm volatile("bar.sync 1, %0;" ::"r"(T) : "memory"); }

// Bitonic sort of P keys (power of two >= 64) by T threads.  Stages with partner distance j < 32 stay inside one 32-key
// group: a warp holds the group in registers and exchanges through shuffles, all stages j = 16..1 of a phase back to back
// with no barrier; only the stages with j >= 32 go through shared memory with one barrier each.
__device__ void bitonic_sort_hybrid(uint64_t* keys, uint32_t* vals, int P, int tid, int T) {
    const int lane = tid & 31, warp = tid >> 5, warps = T >> 5, groups = P >> 5;
    auto reg_pass = [&](int k_first, int k_last) {
        for (int g = warp; g < groups; g += warps) {
            const int i = (g << 5) | lane;
            uint64_t key = keys[i];
            uint32_t val = vals ? vals[i] : 0u;
            for (int k = k_first; k <= k_last; k <<= 1) {
                const bool up = (i & k) == 0;
                for (int j = min(k >> 1, 16); j > 0; j >>= 1) {
                    const uint64_t ok = __shfl_xor_sync(0xffffffffu, key, j);
                    const uint32_t ov = __shfl_xor_sync(0xffffffffu, val, j);
                    const bool take_min = ((lane & j) == 0) == up;
                    if (take_min ? (ok < key) : (ok > key)) { key = ok; val = ov; }
                }
            }
            keys[i] = key;
            if (vals) vals[i] = val;
        }
        k3_bar(T);
    };
    reg_pass(2, 32);
    for (int k = 64; k <= P; k <<= 1) {
        for (int j = k >> 1; j >= 32; j >>= 1) {
            for (int t = tid; t < (P >> 1); t += T) {
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const int l = i | j;
                const bool up = (i & k) == 0;
                const uint64_t ki = keys[i], kl = keys[l];
                if ((ki > kl) == up) {
                    keys[i] = kl; keys[l] = ki;
                    if (vals) { const uint32_t vi = vals[i]; vals[i] = vals[l]; vals[l] = vi; }
                }
            }
            k3_bar(T);
        }
        reg_pass(k, k);
    }
}

__global__ void __launch_bounds__(512) k3_merge_kernel(const K3Params p) {
    extern __shared__ __align__(16) uint8_t k3_smem[];
    __shared__ uint32_t s_diag[2][128];  // 64 rows x 2 halves of in-chunk match bits, double buffered (chunk c / c + 1)
    __shared__ int s_klist[64];
    __shared__ float4 s_kbox[64];
    __shared__ int s_kcat[64];
    __shared__ uint32_t s_kkey[64];
    __shared__ uint32_t s_rem[132];      // removed bits of the whole segment (4096) + slack for the 64-bit windows
    __shared__ unsigned long long s_keepmask, s_rembefore, s_remafter;
    __shared__ uint32_t s_rem_snapshot[2];
    __shared__ int s_kcount, s_ktotal, s_stop, s_flagged;

    const int s = blockIdx.x;
    const int tid = threadIdx.x;
    const int off = p.seg_offsets[s];
    int n = p.seg_counts ? p.seg_counts[s] : p.seg_cap;
    n = min(max(n, 0), p.seg_cap);
    if (n <= p.n_lo || n > p.n_hi) return;  // another tier's launch (or the cluster launch) owns this segment
    if (n == 0) {
        if (tid == 0) p.keep_count[s] = 0;
        return;
    }
    int P = 64;
    while (P < n) P <<= 1;
    // the launch is sized for the tier's capacity; the segment decides how many warps stay
    // (the phases are latency-bound — dependent shared-memory reads and min/max chains — so more warps per segment pay off:
    // profiles/r1_k3_greedynmm_n1024.summary.txt showed 2 warps per scheduler at 19 % issue utilisation)
    const int T = min((int)blockDim.x, P <= 64 ? 64 : (P <= 128 ? 128 : (P <= 256 ? 256 : 512)));
    if (tid >= T) return;
    const int lane = tid & 31;

    uint8_t* base = k3_smem;
    const size_t PP = (size_t)P;
    float4* sbox = reinterpret_cast<float4*>(base);
    uint64_t* keys = reinterpret_cast<uint64_t*>(base + 16 * PP);
    uint32_t* vals = reinterpret_cast<uint32_t*>(base + 24 * PP);
    int* parent = reinterpret_cast<int*>(base + 28 * PP);   // rank of the claiming keep, own rank for keeps, -1 none
    int* step = reinterpret_cast<int*>(base + 32 * PP);     // NMM: rank whose visit produced the claim; tie rule: backward claims
    int* scat = reinterpret_cast<int*>(base + 36 * PP);
    int* keepr = reinterpret_cast<int*>(base + 40 * PP);    // ranks of keeps in output order
    int* runs = reinterpret_cast<int*>(base + 44 * PP);     // NMM: first replay-list position of each keep rank

    const MatchCfg mc = match_cfg(p);

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
    k3_bar(T);
    bitonic_sort_hybrid(keys, vals, P, tid, T);
    int m = n;
    if (p.pre_cap > 0) m = min(m, p.pre_cap);
    for (int r = tid; r < n; r += T) {
        const int g = off + (int)vals[r];
        if (r < m) {
            const float* bp = p.boxes + (size_t)g * p.box_stride;
            sbox[r] = make_float4(bp[0], bp[1], bp[2], bp[3]);
            scat[r] = p.cats ? p.cats[(size_t)g * p.cat_stride] : 0;
            parent[r] = -1;
            step[r] = 0;
            runs[r] = -1;
        } else if (p.parent) {
            p.parent[g] = -1;  // cut by the pre-NMS cap
        }
    }
    for (int i = tid; i < 132; i += T) s_rem[i] = 0;
    if (tid == 0) { s_ktotal = 0; s_stop = 0; s_kcount = 0; s_flagged = 0; }
    k3_bar(T);
    auto keyhi = [&](int r) -> uint32_t { return mc.tie_rule ? (uint32_t)(keys[r] >> 32) : 0u; };

    const bool nmm = p.type == FSD_NMM;
    {
        // ---- 2. scan in chunks of 64 ranks -----------------------------------------------------------
        // greedy (NMS / GREEDYNMM): only the chunk's KEEPS act on lower ranks.  NMM (transitive, A.2.4 `nmm`): EVERY visited rank
        // acts — an unclaimed rank becomes a keep, a claimed one forwards its unclaimed matches to its keep — so the chunk's
        // "actors" are all its ranks, each carrying the label (keep rank) it stands for; the claim records the visiting rank
        // (`step`) for the append order of the replay.  Same two-barrier loop, same sweep.
        // match bits of one chunk: 64 threads per row (two warps = the two 32-bit halves), rows spread over the thread groups
        auto chunk_bits = [&](int c0, uint32_t* dg) {
            const int cn = min(64, m - c0);
            const int q = tid & 63;
            for (int r = tid >> 6; r < cn; r += T >> 6) {
                bool bit = false;
                if (q < cn && q > r)
                    bit = suppresses(sbox[c0 + r], sbox[c0 + q], scat[c0 + r], scat[c0 + q], keyhi(c0 + r), keyhi(c0 + q), mc);
                const uint32_t bal = __ballot_sync(0xffffffffu, bit);
                if (lane == 0) dg[r * 2 + (q >> 5)] = bal;
            }
        };
        chunk_bits(0, s_diag[0]);
        k3_bar(T);
        for (int c0 = 0, c = 0; c0 < m; c0 += 64, ++c) {
            const int cn = min(64, m - c0);
            const uint32_t* dg = s_diag[c & 1];
            if (tid < 32) {
                // resolve: lane l holds rows l and l + 32 in registers; the dependent chain is 64 shuffle + mask steps
                const uint64_t rowA = lane < cn ? ((uint64_t)dg[2 * lane] | ((uint64_t)dg[2 * lane + 1] << 32)) : 0ull;
                const uint64_t rowB = lane + 32 < cn ? ((uint64_t)dg[2 * lane + 64] | ((uint64_t)dg[2 * lane + 65] << 32)) : 0ull;
                const int w0 = c0 >> 5;
                uint64_t remw = (uint64_t)s_rem[w0] | ((uint64_t)s_rem[w0 + 1] << 32);
                const uint64_t before = remw;
                uint64_t keepmask = 0;
                const int kt = s_ktotal;
                int cnt = 0, stop = 0;
                const uint64_t live = ~remw & (cn == 64 ? ~0ull : ((1ull << cn) - 1ull));
                if (!nmm) {
                    for (int b = 0; b < cn && live != 0; ++b) {  // (a chunk whose ranks earlier keeps removed already has nothing to resolve)
                        const uint64_t row = __shfl_sync(0xffffffffu, b < 32 ? rowA : rowB, b & 31);
                        if (!((remw >> b) & 1ull)) {
                            if (p.max_keep > 0 && kt + cnt >= p.max_keep) { stop = 1; break; }
                            keepmask |= 1ull << b;
                            ++cnt;
                            remw |= row;
                        }
                    }
                    for (int pos = lane; pos < 64; pos += 32) {
                        if ((keepmask >> pos) & 1ull) {
                            const int idx = __popcll(keepmask & ((1ull << pos) - 1ull));
                            const int kr = c0 + pos;
                            s_klist[idx] = kr;
                            keepr[kt + idx] = kr;
                            s_kbox[idx] = sbox[kr];
                            s_kcat[idx] = scat[kr];
                            s_kkey[idx] = keyhi(kr);
                            parent[kr] = kr;
                        }
                    }
                } else {
                    // NMM: lane l carries the labels of ranks l and l + 32 (their keep, -1 while unclaimed) through the 64 visits
                    int labA = lane < cn && ((remw >> lane) & 1ull) ? parent[c0 + lane] : -1;
                    int labB = lane + 32 < cn && ((remw >> (lane + 32)) & 1ull) ? parent[c0 + lane + 32] : -1;
                    int stA = -1, stB = -1;
                    for (int b = 0; b < cn; ++b) {
                        const uint64_t row = __shfl_sync(0xffffffffu, b < 32 ? rowA : rowB, b & 31);
                        int lb = __shfl_sync(0xffffffffu, b < 32 ? labA : labB, b & 31);
                        if (!((remw >> b) & 1ull)) {  // unclaimed when visited: a new keep
                            lb = c0 + b;
                            keepmask |= 1ull << b;
                            ++cnt;
                            remw |= 1ull << b;
                            if (lane == (b & 31)) { if (b < 32) labA = lb; else labB = lb; }
                        }
                        const uint64_t fresh = row & ~remw;  // the unclaimed lower ranks of the chunk that match rank b
                        remw |= fresh;
                        if ((fresh >> lane) & 1ull) { labA = lb; stA = c0 + b; }
                        if ((fresh >> (lane + 32)) & 1ull) { labB = lb; stB = c0 + b; }
                    }
                    // every rank of the chunk is an actor of the sweep, standing for its label
                    if (lane < cn) { parent[c0 + lane] = labA; if (stA >= 0) step[c0 + lane] = stA; }
                    if (lane + 32 < cn) { parent[c0 + lane + 32] = labB; if (stB >= 0) step[c0 + lane + 32] = stB; }
                    for (int pos = lane; pos < cn; pos += 32) {
                        s_klist[pos] = pos < 32 ? labA : labB;
                        s_kbox[pos] = sbox[c0 + pos];
                        s_kcat[pos] = scat[c0 + pos];
                        s_kkey[pos] = 0u;
                        if ((keepmask >> pos) & 1ull) keepr[kt + __popcll(keepmask & ((1ull << pos) - 1ull))] = c0 + pos;
                    }
                }
                if (lane == 0) {
                    s_rem[w0] = (uint32_t)remw;
                    s_rem[w0 + 1] = (uint32_t)(remw >> 32);
                    s_keepmask = keepmask; s_rembefore = before; s_remafter = remw;
                    s_kcount = nmm ? cn : cnt; s_ktotal = kt + cnt; s_stop = stop;
                    s_rem_snapshot[0] = s_rem[w0 + 2]; s_rem_snapshot[1] = s_rem[w0 + 3];  // next chunk, before this chunk's sweep
                }
            }
            k3_bar(T);
            const int kc = s_kcount, stop = s_stop;
            if (tid < cn && !nmm) {
                // parents of the ranks this chunk's own keeps removed: the FIRST keep whose row holds the rank
                const uint64_t keepmask = s_keepmask;
                const int q = tid;
                if (!((s_rembefore >> q) & 1ull) && !((keepmask >> q) & 1ull) && ((s_remafter >> q) & 1ull)) {
                    uint64_t km = keepmask & ((1ull << q) - 1ull);
                    while (km) {
                        const int b = __ffsll((long long)km) - 1;
                        km &= km - 1;
                        if ((dg[2 * b + (q >> 5)] >> (q & 31)) & 1u) { parent[c0 + q] = c0 + b; break; }
                    }
                }
            }
            if (!stop) {
                // sweep: every not-yet-removed lower rank against this chunk's new keeps (first match claims it)
                // Most (keep, rank) pairs are disjoint boxes: four keeps are rejected at a time with independent min/max
                // chains (instruction-level parallelism on the common path); the exact test runs, in keep order, only for
                // a group that holds an overlapping keep.  thr <= 0 matches disjoint boxes too: no quick reject then.
                const bool quick = mc.thr > 0.0;
                for (int j = c0 + cn + tid; j < m; j += T) {
                    if ((s_rem[j >> 5] >> (j & 31)) & 1u) continue;
                    const float4 bj = sbox[j];
                    const int cj = scat[j];
                    const uint32_t kj = keyhi(j);
                    bool hit = false;
                    for (int k0 = 0; k0 < kc && !hit; k0 += 4) {
                        const int kn = min(4, kc - k0);
                        unsigned ov = 0xfu;
                        if (quick) {
                            ov = 0;
#pragma unroll
                            for (int u = 0; u < 4; ++u) {
                                const float4 a = s_kbox[min(k0 + u, kc - 1)];
                                if (fminf(a.z, bj.z) > fmaxf(a.x, bj.x) && fminf(a.w, bj.w) > fmaxf(a.y, bj.y)) ov |= 1u << u;
                            }
                        }
                        if (ov == 0) continue;
                        for (int u = 0; u < kn; ++u) {
                            if (!((ov >> u) & 1u)) continue;
                            const int k = k0 + u;
                            if (suppresses(s_kbox[k], bj, s_kcat[k], cj, s_kkey[k], kj, mc)) {
                                atomicOr(&s_rem[j >> 5], 1u << (j & 31));
                                parent[j] = s_klist[k];
                                if (nmm) step[j] = c0 + k;  // the visit that produced the claim
                                hit = true;
                                break;
                            }
                        }
                    }
                }
                // next chunk's match bits, off the critical path (same phase as the sweep) — unless the keeps so far have
                // already removed every rank of it (dense duplicates: most chunks after the first few), which is final
                if (c0 + 64 < m) {
                    const int w1 = (c0 + 64) >> 5, cn1 = min(64, m - c0 - 64);
                    const uint64_t r1 = (uint64_t)s_rem_snapshot[0] | ((uint64_t)s_rem_snapshot[1] << 32);
                    (void)w1;
                    if (nmm || (~r1 & (cn1 == 64 ? ~0ull : ((1ull << cn1) - 1ull))) != 0) chunk_bits(c0 + 64, s_diag[(c + 1) & 1]);
                }
            }
            k3_bar(T);
            if (stop) break;
        }
    }
    k3_bar(T);
    const int K = s_ktotal;
    const bool tie_merge = mc.tie_rule && p.type == FSD_GREEDYNMM;

    // ---- tie rule: a keep claims the earlier equal-score keeps it matches (each keep is claimed by the first such keep) ----
    if (tie_merge) {
        for (int i = tid; i < K; i += T) {
            const int kr = keepr[i];
            const uint32_t kk = keyhi(kr);
            for (int q = kr + 1; q < m && keyhi(q) == kk; ++q) {
                if (parent[q] == q && suppresses(sbox[q], sbox[kr], scat[q], scat[kr], kk, kk, mc)) {
                    atomicOr(&step[kr], q + 1);
                    if (!(atomicOr(&step[q], K3_HASB) & K3_HASB)) atomicAdd(&s_flagged, 1);
                    break;
                }
            }
        }
        k3_bar(T);
    }

    // ---- 3. outputs + merge replay ------------------------------------------------------------------
    if (p.parent) {
        for (int r = tid; r < m; r += T) {
            int pr = parent[r];
            if (tie_merge && pr == r && (step[r] & ~K3_HASB)) pr = (step[r] & ~K3_HASB) - 1;  // claimed by a later equal-score keep
            p.parent[off + (int)vals[r]] = pr < 0 ? -1 : off + (int)vals[pr];
        }
    }
    if (tid == 0) p.keep_count[s] = K;
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
    auto emit = [&](int i, int kr, const double (&kb)[4], float kscore, int kcat) {
        p.keep[off + i] = off + (int)vals[kr];
        float* mb = p.merged_boxes + (size_t)(off + i) * 4;
        mb[0] = (float)kb[0]; mb[1] = (float)kb[1]; mb[2] = (float)kb[2]; mb[3] = (float)kb[3];
        p.merged_scores[off + i] = kscore;
        if (p.merged_cats) p.merged_cats[off + i] = kcat;
    };
    auto fold_one = [&](double (&kb)[4], float kscore, int& kcat, int cr) {
        const float4 c = sbox[cr];
        if (has_match_f64(kb, c, p.metric, p.thr)) {
            kb[0] = fmin(kb[0], (double)c.x); kb[1] = fmin(kb[1], (double)c.y);
            kb[2] = fmax(kb[2], (double)c.z); kb[3] = fmax(kb[3], (double)c.w);
            // merged category: the keep's unless the candidate's score is not lower (sahi get_merged_category)
            const float cscore = p.scores[(size_t)(off + (int)vals[cr]) * p.score_stride];
            if (!(kscore > cscore)) kcat = scat[cr];
        }
    };
    const int flagged = tie_merge ? s_flagged : 0;  // keeps whose merge list holds an earlier equal-score keep (uniform over the CTA)
    if (p.type == FSD_GREEDYNMM && flagged == 0) {
        // candidates of a keep in append order = ascending rank: every thread owns one keep, all threads walk the claim array
        // together (uniform addresses: broadcast reads), a hit folds into the growing union box — no sort
        for (int i0 = 0; i0 < K; i0 += T) {
            const int i = i0 + tid;
            const int kr = i < K ? keepr[i] : -1;
            double kb[4] = {0, 0, 0, 0};
            float kscore = 0.f;
            int kcat = 0;
            if (kr >= 0) {
                const float4 b = sbox[kr];
                kb[0] = b.x; kb[1] = b.y; kb[2] = b.z; kb[3] = b.w;
                kscore = p.scores[(size_t)(off + (int)vals[kr]) * p.score_stride];
                kcat = scat[kr];
            }
            for (int r = keepr[i0] + 1; r < m; ++r)
                if (parent[r] == kr && r != kr) fold_one(kb, kscore, kcat, r);
            if (kr >= 0) emit(i, kr, kb, kscore, kcat);
        }
        return;
    }
    // ---- replay through sorted lists: NMM (step-major append order), and GREEDYNMM when the tie rule made a keep claim an
    // earlier keep (those keeps fold AFTER the keeps they claimed, reading their merged boxes: lists make that O(list)) ----
    // key = (keep rank : 15 bits | append sequence : 30 bits | candidate rank : 15 bits), ascending
    for (int r = tid; r < P; r += T) {
        uint64_t key = ~0ull;
        if (r < m) {
            int pr = parent[r];
            if (p.type == FSD_GREEDYNMM && pr == r && (step[r] & ~K3_HASB)) pr = (step[r] & ~K3_HASB) - 1;  // claimed keep
            if (pr >= 0 && pr != r) {
                const uint64_t seq = p.type == FSD_NMM ? (uint64_t)step[r] * 32768ull + (uint64_t)(32767 - r) : (uint64_t)r;
                key = ((uint64_t)pr << 45) | (seq << 15) | (uint64_t)r;
            }
        }
        keys[r] = key;
    }
    k3_bar(T);
    bitonic_sort_hybrid(keys, nullptr, P, tid, T);
    for (int q = tid; q < P; q += T) {
        const uint64_t key = keys[q];
        if (key == ~0ull) continue;
        const int pr = (int)(key >> 45);
        if (q == 0 || (int)(keys[q - 1] >> 45) != pr) runs[pr] = q;
    }
    k3_bar(T);
    auto fold_list = [&](int i) {
        const int kr = keepr[i];
        const float4 b = sbox[kr];
        double kb[4] = {(double)b.x, (double)b.y, (double)b.z, (double)b.w};
        const float kscore = p.scores[(size_t)(off + (int)vals[kr]) * p.score_stride];
        int kcat = scat[kr];
        int q = runs[kr];
        if (q >= 0) {
            for (; q < P; ++q) {
                const uint64_t key = keys[q];
                if (key == ~0ull || (int)(key >> 45) != kr) break;
                fold_one(kb, kscore, kcat, (int)(key & 32767ull));
            }
        }
        emit(i, kr, kb, kscore, kcat);
        if (flagged) {  // a later equal-score keep may fold THIS keep's merged box (only that keep reads it)
            sbox[kr] = make_float4((float)kb[0], (float)kb[1], (float)kb[2], (float)kb[3]);
            scat[kr] = kcat;
        }
    };
    for (int i = tid; i < K; i += T)
        if (!(flagged && (step[keepr[i]] & K3_HASB))) fold_list(i);
    if (flagged) {
        k3_bar(T);
        if (tid == 0)  // the keeps with backward claims, in rank order; every list is walked once: O(m) in total
            for (int i = 0; i < K; ++i)
                if (step[keepr[i]] & K3_HASB) fold_list(i);
    }
}

// ---- fallback: one CTA per segment with the arrays in the (L2-resident) workspace ---------------------------------------
// Round 1's kernel, kept as the cross-check of the cluster kernel (FSD_K3_SINGLE_CTA=1): plain rank order, no tie rule, NMM as
// a rank-by-rank walk.
__global__ void __launch_bounds__(512) k3_merge_global_kernel(const K3Params p) {
    __shared__ uint32_t s_diag[128];   // 64 rows x 2 halves of in-chunk match bits
    __shared__ int s_klist[64];
    // staged per chunk so that the inner loops never touch the (possibly L2-resident) workspace: the chunk's 64 boxes,
    // the boxes of its new keeps, and the removed-bit array of the whole segment (32768 bits)
    __shared__ float4 s_cbox[64];
    __shared__ int s_ccat[64];
    __shared__ float4 s_kbox[64];
    __shared__ int s_kcat[64];
    __shared__ uint32_t s_rem[1024];
    __shared__ int s_kcount, s_ktotal, s_stop;

    const int s = blockIdx.x;
    const int tid = threadIdx.x, T = blockDim.x;
    const int off = p.seg_offsets[s];
    int n = p.seg_counts ? p.seg_counts[s] : p.seg_cap;
    n = min(max(n, 0), p.seg_cap);
    if (n <= K3_SMEM_MAX_P && !p.use_global) return;  // the shared-memory launch owns this segment
    if (n == 0) {
        if (tid == 0) p.keep_count[s] = 0;
        return;
    }
    int P = 64;
    while (P < n) P <<= 1;

    uint8_t* base = p.workspace + (size_t)s * p.ws_per_segment;
    const size_t PP = (size_t)p.P;
    float4* sbox = reinterpret_cast<float4*>(base);
    uint64_t* keys = reinterpret_cast<uint64_t*>(base + 16 * PP);
    uint32_t* vals = reinterpret_cast<uint32_t*>(base + 24 * PP);
    int* parent = reinterpret_cast<int*>(base + 28 * PP);   // rank of the claiming keep, own rank for keeps, -1 none
    int* step = reinterpret_cast<int*>(base + 32 * PP);     // NMM: rank whose visit produced the claim
    int* scat = reinterpret_cast<int*>(base + 36 * PP);
    int* keepr = reinterpret_cast<int*>(base + 40 * PP);    // ranks of keeps in output order
    int* runs = reinterpret_cast<int*>(base + 44 * PP);     // first replay-list position of each keep rank

    MatchCfg mc = match_cfg(p);
    mc.tie_rule = 0;

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
            if (p.parent) p.parent[g] = -1;  // cut by the pre-NMS cap
        }
    }
    if (tid == 0) { s_ktotal = 0; s_stop = 0; }
    __syncthreads();

    if (p.type != FSD_NMM) {
        // ---- 2a. greedy scan in chunks of 64 ranks --------------------------------------------------
        uint32_t* rem = s_rem;
        for (int i = tid; i < (m + 31) / 32; i += T) rem[i] = 0;
        __syncthreads();
        for (int c0 = 0; c0 < m; c0 += 64) {
            const int cn = min(64, m - c0);
            if (tid < cn) { s_cbox[tid] = sbox[c0 + tid]; s_ccat[tid] = scat[c0 + tid]; }
            __syncthreads();
            // in-chunk match bits: 64 threads per row (two warps = two 32-bit halves)
            for (int r = tid >> 6; r < 64; r += T >> 6) {
                const int q = tid & 63;
                bool bit = false;
                if (r < cn && q < cn && q > r) bit = match_pair(s_cbox[r], s_cbox[q], s_ccat[r], s_ccat[q], mc);
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
            if (tid < kc) { s_kbox[tid] = s_cbox[s_klist[tid] - c0]; s_kcat[tid] = s_ccat[s_klist[tid] - c0]; }
            __syncthreads();
            // sweep: every not-yet-removed lower rank against this chunk's new keeps (first match claims it)
            for (int j = c0 + cn + tid; j < m; j += T) {
                if ((rem[j >> 5] >> (j & 31)) & 1u) continue;
                const float4 bj = sbox[j];
                const int cj = scat[j];
                for (int k = 0; k < kc; ++k) {
                    if (match_pair(s_kbox[k], bj, s_kcat[k], cj, mc)) {
                        atomicOr(&rem[j >> 5], 1u << (j & 31));
                        parent[j] = s_klist[k];
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

// ---- large segments (> 4096 boxes): one thread-block CLUSTER of 8 CTAs per segment -------------------------------------
// The single-CTA path is bound by one SM's instruction throughput once a segment no longer fits shared memory (N = 9900:
// ~27 M pair tests + two 16384-key sorts on one SM = 6.5-8.8 ms).  Here the same algorithm runs on 8 SMs: the arrays live
// in the L2-resident workspace, every phase that is parallel over ranks (rank sort, sweep, replay sort, fold) is split
// over the cluster's 4096 threads with cluster.sync() between dependent steps, and only the 64x64 in-chunk resolve stays
// on CTA 0, which broadcasts the chunk's keeps through a small scratch area.  Arrays written by one CTA and read by
// another are read with ld.global.cg (L2), never through a possibly stale L1 line.  All three merge types (NMM through the
// same chunked scan with every rank of a chunk acting for its keep).  CTA 0 resolves a chunk with one
// warp holding the 64x64 match bits in registers, and computes the NEXT chunk's match bits while the cluster sweeps.
namespace cg = cooperative_groups;

template <typename T> __device__ __forceinline__ T ld_cg(const T* p) { return __ldcg(p); }

struct K3Scratch {  // lives in global memory, one per segment
    int kcount, ktotal, stop, pad;
    int klist[64];
    int kcat[64];
    uint32_t kkey[64];
    float4 kbox[64];
};
static_assert(sizeof(K3Scratch) <= K3_SCRATCH_BYTES, "scratch area too small");

// Bitonic sort of P keys (P >= 8192, power of two) by the whole cluster.  Compare-exchange stages whose partner distance
// j is smaller than the per-CTA block B = P / 8 never leave a block, so each CTA runs them on its block in shared memory
// with block barriers only; just the log2(P/B) * (log2(P/B) + 1) / 2 = 6 stages with j >= B (for P = 16384) go through
// L2 with a cluster barrier each — about 10 cluster barriers per sort instead of 105.
__device__ void cluster_bitonic_sort(cg::cluster_group& cluster, uint64_t* keys, uint32_t* vals, int P, int crank, int tid, int T,
                                     uint64_t* lk, uint32_t* lv) {
    const int B = P / K3_CLUSTER, base = crank * B;
    const int gtid = crank * T + tid, TT = K3_CLUSTER * T;
    auto local_stages = [&](int k_first, int k_last) {  // all stages (k, j) with k_first <= k <= k_last and j < B
        for (int i = tid; i < B; i += T) { lk[i] = ld_cg(keys + base + i); if (vals) lv[i] = ld_cg(vals + base + i); }
        __syncthreads();
        for (int k = k_first; k <= k_last; k <<= 1) {
            for (int j = min(k >> 1, B >> 1); j > 0; j >>= 1) {
                for (int t = tid; t < (B >> 1); t += T) {
                    const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                    const int l = i | j;
                    const bool up = ((base + i) & k) == 0;
                    const uint64_t ki = lk[i], kl = lk[l];
                    if ((ki > kl) == up) {
                        lk[i] = kl; lk[l] = ki;
                        if (vals) { const uint32_t vi = lv[i]; lv[i] = lv[l]; lv[l] = vi; }
                    }
                }
                __syncthreads();
            }
        }
        for (int i = tid; i < B; i += T) { __stcg(keys + base + i, lk[i]); if (vals) __stcg(vals + base + i, lv[i]); }
    };
    local_stages(2, B);
    cluster.sync();
    for (int k = 2 * B; k <= P; k <<= 1) {
        for (int j = k >> 1; j >= B; j >>= 1) {
            for (int t = gtid; t < (P >> 1); t += TT) {
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const int l = i | j;
                const bool up = (i & k) == 0;
                const uint64_t ki = ld_cg(keys + i), kl = ld_cg(keys + l);
                if ((ki > kl) == up) {
                    __stcg(keys + i, kl); __stcg(keys + l, ki);
                    if (vals) { const uint32_t vi = ld_cg(vals + i), vl = ld_cg(vals + l); __stcg(vals + i, vl); __stcg(vals + l, vi); }
                }
            }
            cluster.sync();
        }
        local_stages(k, k);
        cluster.sync();
    }
}

__global__ void __cluster_dims__(K3_CLUSTER, 1, 1) __launch_bounds__(512) k3_merge_cluster_kernel(const K3Params p) {
    cg::cluster_group cluster = cg::this_cluster();
    __shared__ uint32_t s_diag[2][128];
    __shared__ float4 s_cbox[2][64];
    __shared__ int s_ccat[2][64];
    __shared__ uint32_t s_ckey[2][64];
    __shared__ float4 s_kbox[64];
    __shared__ int s_kcat[64];
    __shared__ uint32_t s_kkey[64];
    __shared__ int s_klist[64];
    __shared__ unsigned long long s_keepmask, s_rembefore, s_remafter;
    __shared__ int s_kcount, s_ktotal, s_stop;
    extern __shared__ __align__(16) uint8_t k3c_smem[];  // sort block: P/8 keys (8 B) + values (4 B); then the cache of this CTA's own ranks (24 B each)

    const int s = blockIdx.x / K3_CLUSTER;
    const int crank = (int)cluster.block_rank();
    const int tid = threadIdx.x, T = blockDim.x, gtid = crank * T + tid, TT = K3_CLUSTER * T;
    const int off = p.seg_offsets[s];
    int n = p.seg_counts ? p.seg_counts[s] : p.seg_cap;
    n = min(max(n, 0), p.seg_cap);
    if (n <= p.cluster_min) return;  // uniform over the cluster: the shared-memory launch owns this segment
    int P = 64;
    while (P < n) P <<= 1;
    uint8_t* base = p.workspace + (size_t)s * p.ws_per_segment;
    const size_t PP = (size_t)p.P;
    float4* sbox = reinterpret_cast<float4*>(base);
    uint64_t* keys = reinterpret_cast<uint64_t*>(base + 16 * PP);
    uint32_t* vals = reinterpret_cast<uint32_t*>(base + 24 * PP);
    int* parent = reinterpret_cast<int*>(base + 28 * PP);
    uint32_t* rem = reinterpret_cast<uint32_t*>(base + 32 * PP);  // removed bits: the first PP / 8 bytes of the NMM path's `step` slot
    int* scat = reinterpret_cast<int*>(base + 36 * PP);
    int* keepr = reinterpret_cast<int*>(base + 40 * PP);
    int* runs = reinterpret_cast<int*>(base + 44 * PP);
    int* stepc = reinterpret_cast<int*>(base + 48 * PP);          // NMM: rank whose visit produced the claim (the `step` slot holds the removed bits here)
    K3Scratch* sc = reinterpret_cast<K3Scratch*>(base + (size_t)K3_WS_BYTES_PER_BOX * PP);

    const MatchCfg mc = match_cfg(p);
    const int lane = tid & 31;

    // ---- 1. rank ----------------------------------------------------------------------------------------
    for (int i = gtid; i < P; i += TT) {
        uint64_t key = ~0ull;
        uint32_t v = 0xffffffffu;
        if (i < n) {
            const float scv = p.scores[(size_t)(off + i) * p.score_stride];
            const uint32_t tb = p.tie ? (uint32_t)p.tie[(size_t)(off + i) * p.tie_stride] : (uint32_t)i;
            key = ((uint64_t)score_key_desc(scv) << 32) | tb;
            v = (uint32_t)i;
        }
        __stcg(keys + i, key); __stcg(vals + i, v);
    }
    if (gtid == 0) sc->pad = 0;
    cluster.sync();
    uint64_t* lk = reinterpret_cast<uint64_t*>(k3c_smem);
    uint32_t* lv = reinterpret_cast<uint32_t*>(k3c_smem + (size_t)(p.P / K3_CLUSTER) * 8);
    cluster_bitonic_sort(cluster, keys, vals, P, crank, tid, T, lk, lv);
    int m = n;
    if (p.pre_cap > 0) m = min(m, p.pre_cap);
    for (int r = gtid; r < n; r += TT) {
        const int g = off + (int)ld_cg(vals + r);
        if (r < m) {
            const float* bp = p.boxes + (size_t)g * p.box_stride;
            __stcg(sbox + r, make_float4(bp[0], bp[1], bp[2], bp[3]));
            __stcg(scat + r, p.cats ? p.cats[(size_t)g * p.cat_stride] : 0);
            __stcg(parent + r, -1);
            __stcg(runs + r, -1);
        } else {
            if (p.parent) p.parent[g] = -1;  // cut by the pre-NMS cap
        }
    }
    for (int i = gtid; i < (m + 31) / 32 + 2; i += TT) __stcg(rem + i, 0u);
    cluster.sync();
    auto keyhi = [&](int r) -> uint32_t { return mc.tie_rule ? (uint32_t)(ld_cg(keys + r) >> 32) : 0u; };
    // Every CTA stages a chunk's boxes and computes its 64x64 match bits itself (double buffered: chunk c + 1 while chunk c is
    // swept): the chunk is then RESOLVED REDUNDANTLY by one warp of every CTA — identical inputs, identical keeps — so no
    // keep list has to travel between CTAs and ONE cluster barrier per chunk is enough (it publishes the removed bits).
    auto stage_chunk = [&](int c0, int buf) {
        const int cn = min(64, m - c0);
        if (tid < cn) { s_cbox[buf][tid] = ld_cg(sbox + c0 + tid); s_ccat[buf][tid] = ld_cg(scat + c0 + tid); s_ckey[buf][tid] = keyhi(c0 + tid); }
        __syncthreads();
        const int q = tid & 63;
        for (int r = tid >> 6; r < cn; r += T >> 6) {
            bool bit = false;
            if (q < cn && q > r)
                bit = suppresses(s_cbox[buf][r], s_cbox[buf][q], s_ccat[buf][r], s_ccat[buf][q], s_ckey[buf][r], s_ckey[buf][q], mc);
            const uint32_t bal = __ballot_sync(0xffffffffu, bit);
            if (lane == 0) s_diag[buf][r * 2 + (q >> 5)] = bal;
        }
        __syncthreads();
    };
    // The ranks a thread sweeps are fixed for the whole scan (rank = gtid + k * 4096): their boxes, categories and score keys
    // are cached in this CTA's shared memory once (the sort block is free now), so the sweep never goes back to L2 for them.
    float4* obox = reinterpret_cast<float4*>(k3c_smem);
    int* ocat = reinterpret_cast<int*>(k3c_smem + (size_t)(p.P / K3_CLUSTER) * 16);
    uint32_t* okey = reinterpret_cast<uint32_t*>(k3c_smem + (size_t)(p.P / K3_CLUSTER) * 20);
    uint32_t alive = 0;  // bit k: rank gtid + k * TT is not removed yet and not passed by the scan
    for (int k = 0; gtid + k * TT < m; ++k) {
        const int j = gtid + k * TT;
        obox[k * T + tid] = ld_cg(sbox + j);
        ocat[k * T + tid] = ld_cg(scat + j);
        okey[k * T + tid] = keyhi(j);
        alive |= 1u << k;
    }
    if (tid == 0) { s_ktotal = 0; s_stop = 0; s_kcount = 0; }
    stage_chunk(0, 0);

    // ---- 2. scan in chunks of 64 ranks (greedy, or NMM's transitive walk: see k3_merge_kernel) ----------------------------
    const bool nmm = p.type == FSD_NMM;
    for (int c0 = 0, c = 0; c0 < m; c0 += 64, ++c) {
        const int cn = min(64, m - c0);
        const int buf = c & 1;
        const uint32_t* dg = s_diag[buf];
        if (tid < 32) {
            const uint64_t rowA = lane < cn ? ((uint64_t)dg[2 * lane] | ((uint64_t)dg[2 * lane + 1] << 32)) : 0ull;
            const uint64_t rowB = lane + 32 < cn ? ((uint64_t)dg[2 * lane + 64] | ((uint64_t)dg[2 * lane + 65] << 32)) : 0ull;
            const int w0 = c0 >> 5;
            uint64_t remw = (uint64_t)ld_cg(rem + w0) | ((uint64_t)ld_cg(rem + w0 + 1) << 32);
            const uint64_t before = remw;
            uint64_t keepmask = 0;
            const int kt = s_ktotal;
            int cnt = 0, stop = 0;
            const uint64_t live = ~remw & (cn == 64 ? ~0ull : ((1ull << cn) - 1ull));
            if (!nmm) {
                for (int b = 0; b < cn && live != 0; ++b) {
                    const uint64_t row = __shfl_sync(0xffffffffu, b < 32 ? rowA : rowB, b & 31);
                    if (!((remw >> b) & 1ull)) {
                        if (p.max_keep > 0 && kt + cnt >= p.max_keep) { stop = 1; break; }
                        keepmask |= 1ull << b;
                        ++cnt;
                        remw |= row;
                    }
                }
                for (int pos = lane; pos < 64; pos += 32) {
                    if ((keepmask >> pos) & 1ull) {
                        const int idx = __popcll(keepmask & ((1ull << pos) - 1ull));
                        const int kr = c0 + pos;
                        s_klist[idx] = kr; s_kbox[idx] = s_cbox[buf][pos]; s_kcat[idx] = s_ccat[buf][pos]; s_kkey[idx] = s_ckey[buf][pos];
                        if (crank == 0) { __stcg(keepr + kt + idx, kr); __stcg(parent + kr, kr); }
                    }
                }
            } else {
                // NMM (see k3_merge_kernel): every rank of the chunk acts, standing for its label; lane l carries ranks l, l + 32
                int labA = lane < cn && ((remw >> lane) & 1ull) ? ld_cg(parent + c0 + lane) : -1;
                int labB = lane + 32 < cn && ((remw >> (lane + 32)) & 1ull) ? ld_cg(parent + c0 + lane + 32) : -1;
                int stA = -1, stB = -1;
                for (int b = 0; b < cn; ++b) {
                    const uint64_t row = __shfl_sync(0xffffffffu, b < 32 ? rowA : rowB, b & 31);
                    int lb = __shfl_sync(0xffffffffu, b < 32 ? labA : labB, b & 31);
                    if (!((remw >> b) & 1ull)) {
                        lb = c0 + b;
                        keepmask |= 1ull << b;
                        ++cnt;
                        remw |= 1ull << b;
                        if (lane == (b & 31)) { if (b < 32) labA = lb; else labB = lb; }
                    }
                    const uint64_t fresh = row & ~remw;
                    remw |= fresh;
                    if ((fresh >> lane) & 1ull) { labA = lb; stA = c0 + b; }
                    if ((fresh >> (lane + 32)) & 1ull) { labB = lb; stB = c0 + b; }
                }
                if (crank == 0) {
                    if (lane < cn) { __stcg(parent + c0 + lane, labA); if (stA >= 0) __stcg(stepc + c0 + lane, stA); }
                    if (lane + 32 < cn) { __stcg(parent + c0 + lane + 32, labB); if (stB >= 0) __stcg(stepc + c0 + lane + 32, stB); }
                }
                for (int pos = lane; pos < cn; pos += 32) {
                    s_klist[pos] = pos < 32 ? labA : labB;
                    s_kbox[pos] = s_cbox[buf][pos]; s_kcat[pos] = s_ccat[buf][pos]; s_kkey[pos] = 0u;
                    if (crank == 0 && ((keepmask >> pos) & 1ull)) __stcg(keepr + kt + __popcll(keepmask & ((1ull << pos) - 1ull)), c0 + pos);
                }
            }
            if (lane == 0) {
                s_keepmask = keepmask; s_rembefore = before; s_remafter = remw;
                s_kcount = nmm ? cn : cnt; s_ktotal = kt + cnt; s_stop = stop;
            }
        }
        __syncthreads();
        const int kc = s_kcount, stop = s_stop;
        if (crank == 0 && tid < cn && !nmm) {  // parents of the ranks this chunk's own keeps removed: the first keep whose row holds the rank
            const uint64_t keepmask = s_keepmask;
            const int q = tid;
            if (!((s_rembefore >> q) & 1ull) && !((keepmask >> q) & 1ull) && ((s_remafter >> q) & 1ull)) {
                uint64_t km = keepmask & ((1ull << q) - 1ull);
                while (km) {
                    const int b = __ffsll((long long)km) - 1;
                    km &= km - 1;
                    if ((dg[2 * b + (q >> 5)] >> (q & 31)) & 1u) { __stcg(parent + c0 + q, c0 + b); break; }
                }
            }
        }
        if (!stop) {
            const bool quick = mc.thr > 0.0;  // disjoint boxes cannot match a positive threshold: reject four keeps at a time
            for (int k = 0; (alive >> k) != 0; ++k) {
                if (!((alive >> k) & 1u)) continue;
                const int j = gtid + k * TT;
                if (j < c0 + cn) { alive &= ~(1u << k); continue; }  // the scan has passed this rank
                const float4 bj = obox[k * T + tid];
                bool hit = false;
                for (int k0 = 0; k0 < kc && !hit; k0 += 4) {
                    const int kn = min(4, kc - k0);
                    unsigned ov = 0xfu;
                    if (quick) {
                        ov = 0;
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const float4 a = s_kbox[min(k0 + u, kc - 1)];
                            if (fminf(a.z, bj.z) > fmaxf(a.x, bj.x) && fminf(a.w, bj.w) > fmaxf(a.y, bj.y)) ov |= 1u << u;
                        }
                    }
                    if (ov == 0) continue;
                    for (int u = 0; u < kn; ++u) {
                        if (!((ov >> u) & 1u)) continue;
                        const int kk = k0 + u;
                        if (suppresses(s_kbox[kk], bj, s_kcat[kk], ocat[k * T + tid], s_kkey[kk], okey[k * T + tid], mc)) {
                            atomicOr(rem + (j >> 5), 1u << (j & 31));
                            __stcg(parent + j, s_klist[kk]);
                            if (nmm) __stcg(stepc + j, c0 + kk);  // the visit that produced the claim
                            alive &= ~(1u << k);
                            hit = true;
                            break;
                        }
                    }
                }
            }
            if (c0 + 64 < m) stage_chunk(c0 + 64, buf ^ 1);  // next chunk's boxes + match bits, off the critical path
        }
        cluster.sync();  // the removed bits of this chunk's sweep are visible to every CTA's next resolve
        if (stop) break;
    }
    __syncthreads();
    const int K = s_ktotal;  // every CTA resolved every chunk: the same count everywhere
    const bool tie_merge = mc.tie_rule && p.type == FSD_GREEDYNMM;

    // ---- tie rule: a keep claims the earlier equal-score keeps it matches (see k3_merge_kernel).  Here the claim is written
    // into parent[] itself (the claimed keep then enters the claimer's replay list like any other candidate); what a rank IS
    // lives in runs[]: K3C_ISKEEP, K3_HASB (this keep's list holds an earlier keep), K3C_HASLIST | first list position.
    constexpr int K3C_ISKEEP = 1 << 29, K3C_HASLIST = 1 << 28;
    for (int r = gtid; r < m; r += TT) __stcg(runs + r, 0);
    cluster.sync();
    for (int i = gtid; i < K; i += TT) atomicOr(runs + ld_cg(keepr + i), K3C_ISKEEP);
    cluster.sync();
    if (tie_merge) {
        for (int i = gtid; i < K; i += TT) {
            const int kr = ld_cg(keepr + i);
            const uint32_t kk = keyhi(kr);
            const float4 bk = ld_cg(sbox + kr);
            const int ck = ld_cg(scat + kr);
            for (int q = kr + 1; q < m && keyhi(q) == kk; ++q) {
                if ((ld_cg(runs + q) & K3C_ISKEEP) && suppresses(ld_cg(sbox + q), bk, ld_cg(scat + q), ck, kk, kk, mc)) {
                    __stcg(parent + kr, q);
                    if (!(atomicOr(runs + q, K3_HASB) & K3_HASB)) atomicAdd(&sc->pad, 1);  // sc->pad counts the flagged keeps
                    break;
                }
            }
        }
        cluster.sync();
    }

    // ---- 3. outputs + merge replay --------------------------------------------------------------------------------
    if (p.parent) {
        for (int r = gtid; r < m; r += TT) {
            const int pr = ld_cg(parent + r);
            p.parent[off + (int)ld_cg(vals + r)] = pr < 0 ? -1 : off + (int)ld_cg(vals + pr);
        }
    }
    if (gtid == 0) p.keep_count[s] = K;
    if (p.type == FSD_NMS) {
        for (int i = gtid; i < K; i += TT) {
            const int kr = ld_cg(keepr + i);
            const int g = off + (int)ld_cg(vals + kr);
            p.keep[off + i] = g;
            const float4 b = ld_cg(sbox + kr);
            float* mb = p.merged_boxes + (size_t)(off + i) * 4;
            mb[0] = b.x; mb[1] = b.y; mb[2] = b.z; mb[3] = b.w;
            p.merged_scores[off + i] = p.scores[(size_t)g * p.score_stride];
            if (p.merged_cats) p.merged_cats[off + i] = ld_cg(scat + kr);
        }
        return;
    }
    // replay key = (keep rank : 15 bits | append sequence : 30 bits | candidate rank : 15 bits), ascending; a keep claimed
    // through the tie rule carries its claimer in parent[] and is listed like every other candidate
    for (int r = gtid; r < P; r += TT) {
        uint64_t key = ~0ull;
        if (r < m) {
            const int pr = ld_cg(parent + r);
            if (pr >= 0 && pr != r) {
                const uint64_t seq = nmm ? (uint64_t)ld_cg(stepc + r) * 32768ull + (uint64_t)(32767 - r) : (uint64_t)r;
                key = ((uint64_t)pr << 45) | (seq << 15) | (uint64_t)r;
            }
        }
        __stcg(keys + r, key);
    }
    cluster.sync();
    cluster_bitonic_sort(cluster, keys, nullptr, P, crank, tid, T, lk, lv);
    for (int q = gtid; q < P; q += TT) {
        const uint64_t key = ld_cg(keys + q);
        if (key == ~0ull) continue;
        const int pr = (int)(key >> 45);
        if (q == 0 || (int)(ld_cg(keys + q - 1) >> 45) != pr) atomicOr(runs + pr, K3C_HASLIST | q);
    }
    cluster.sync();
    auto fold_keep = [&](int i) {
        const int kr = ld_cg(keepr + i);
        const int g = off + (int)ld_cg(vals + kr);
        const float4 b = ld_cg(sbox + kr);
        double kb[4] = {(double)b.x, (double)b.y, (double)b.z, (double)b.w};
        const float kscore = p.scores[(size_t)g * p.score_stride];
        int kcat = ld_cg(scat + kr);
        const int w = ld_cg(runs + kr);
        if (w & K3C_HASLIST) {
            for (int q = w & 0xffff; q < P; ++q) {
                const uint64_t key = ld_cg(keys + q);
                if (key == ~0ull || (int)(key >> 45) != kr) break;
                const int cr = (int)(key & 32767ull);
                const float4 c = ld_cg(sbox + cr);
                if (has_match_f64(kb, c, p.metric, p.thr)) {
                    kb[0] = fmin(kb[0], (double)c.x); kb[1] = fmin(kb[1], (double)c.y);
                    kb[2] = fmax(kb[2], (double)c.z); kb[3] = fmax(kb[3], (double)c.w);
                    const float cscore = p.scores[(size_t)(off + (int)ld_cg(vals + cr)) * p.score_stride];
                    if (!(kscore > cscore)) kcat = ld_cg(scat + cr);
                }
            }
        }
        p.keep[off + i] = g;
        float* mb = p.merged_boxes + (size_t)(off + i) * 4;
        mb[0] = (float)kb[0]; mb[1] = (float)kb[1]; mb[2] = (float)kb[2]; mb[3] = (float)kb[3];
        p.merged_scores[off + i] = kscore;
        if (p.merged_cats) p.merged_cats[off + i] = kcat;
        if (tie_merge) {  // a later equal-score keep may fold THIS keep's merged box
            __stcg(sbox + kr, make_float4((float)kb[0], (float)kb[1], (float)kb[2], (float)kb[3]));
            __stcg(scat + kr, kcat);
        }
    };
    const int flagged = tie_merge ? ld_cg(&sc->pad) : 0;  // uniform over the cluster (read after the cluster barriers above)
    uint32_t* fbits = rem;  // the removed bits are dead after the scan: bitmap of the flagged keeps by KEEP INDEX
    if (flagged) {
        for (int i = gtid; i < (K + 31) / 32 + 1; i += TT) __stcg(fbits + i, 0u);
        cluster.sync();
    }
    for (int i = gtid; i < K; i += TT) {
        if (flagged && (ld_cg(runs + ld_cg(keepr + i)) & K3_HASB)) atomicOr(fbits + (i >> 5), 1u << (i & 31));
        else fold_keep(i);
    }
    if (flagged) {
        cluster.sync();
        if (gtid == 0) {  // keeps whose list holds an earlier keep, in rank order (rare: exact score ties between overlapping keeps)
            for (int w = 0; w < (K + 31) / 32; ++w) {
                uint32_t bits = ld_cg(fbits + w);
                while (bits) {
                    const int b = __ffs(bits) - 1;
                    bits &= bits - 1;
                    fold_keep(w * 32 + b);
                }
            }
        }
    }
}

static int pow2_at_least(int n) {
    int P = 64;
    while (P < n) P <<= 1;
    return P;
}

}  // namespace fsd

using namespace fsd;

// Segments above this many boxes take the 8-CTA cluster kernel.  Default 4096 = the shared-memory kernel's capacity; FSD_K3_CLUSTER_MIN
// = 2048 moves the hand-over down for measurements (smaller values are not supported by the cluster kernel).
static int k3_cluster_min() {
    const char* v = getenv("FSD_K3_CLUSTER_MIN");
    const int t = v ? atoi(v) : K3_SMEM_MAX_P;
    return t <= 2048 ? 2048 : K3_SMEM_MAX_P;
}

extern "C" int64_t fsd_merge_workspace_bytes(int64_t N, int S, int max_segment) {
    (void)N;
    if (max_segment <= k3_cluster_min()) return 256;  // unused, but keep the pointer non-null for callers
    return (int64_t)S * ((int64_t)pow2_at_least(max_segment) * K3_WS_BYTES_PER_BOX + K3_SCRATCH_BYTES) + 256;
}

extern "C" int fsd_merge(fsd_handle_t h, const float* boxes, int box_stride, const float* scores, int score_stride,
                         const int32_t* cats, int cat_stride, const int32_t* tie, int tie_stride,
                         const int32_t* seg_offsets, const int32_t* seg_counts, int S, int max_segment, int type,
                         int metric, double thr, int cmp_strict, int precision, int class_agnostic, int pre_cap,
                         int max_keep, int tie_rule, int32_t* keep, int32_t* keep_count, int32_t* parent, float* merged_boxes,
                         float* merged_scores, int32_t* merged_cats, void* workspace, int64_t workspace_bytes,
                         void* stream_) {
    FSD_CHECK_ARG(h && boxes && scores && seg_offsets && keep && keep_count && merged_boxes && merged_scores,
                  "fsd_merge: null argument");
    FSD_CHECK_ARG(type == FSD_NMS || type == FSD_GREEDYNMM || type == FSD_NMM, "fsd_merge: unknown merge type %d", type);
    FSD_CHECK_ARG(metric == FSD_IOU || metric == FSD_IOS, "fsd_merge: unknown match metric %d", metric);
    FSD_CHECK_ARG(S >= 0 && max_segment >= 0 && box_stride >= 4 && score_stride >= 1, "fsd_merge: bad sizes");
    FSD_CHECK_ARG(tie_rule == 0 || tie_rule == 1, "fsd_merge: unknown tie rule %d", tie_rule);
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
    p.class_agnostic = class_agnostic; p.pre_cap = pre_cap; p.max_keep = max_keep; p.thr = thr; p.tie_rule = tie_rule;
    p.keep = keep; p.keep_count = keep_count; p.parent = parent; p.merged_boxes = merged_boxes;
    p.merged_scores = merged_scores; p.merged_cats = merged_cats;
    p.P = pow2_at_least(max_segment);
    p.cluster_min = k3_cluster_min();
    p.use_global = p.P > p.cluster_min;
    p.workspace = reinterpret_cast<uint8_t*>(workspace);
    p.ws_per_segment = (size_t)p.P * K3_WS_BYTES_PER_BOX + K3_SCRATCH_BYTES;
    if (p.use_global) {
        FSD_CHECK_ARG(workspace && workspace_bytes >= fsd_merge_workspace_bytes(0, S, max_segment),
                      "fsd_merge: workspace too small (%lld bytes needed)", (long long)fsd_merge_workspace_bytes(0, S, max_segment));
        if (((uintptr_t)workspace & 15) != 0) { set_error("fsd_merge: workspace must be 16-byte aligned"); return FSD_ERR_ALIGN; }
    }
    cudaStream_t stream = (cudaStream_t)stream_;
    FSD_CUDA(cudaSetDevice(h->device));
    const bool force_single = getenv("FSD_K3_SINGLE_CTA") != nullptr;
    // the path is chosen per segment ON THE DEVICE from its actual count: both kernels are launched over all S segments when
    // the capacity allows large ones, and each returns at once for the segments that belong to the other
    if (p.use_global && force_single) {
        // cross-check mode (FSD_K3_SINGLE_CTA=1): one CTA per segment over the workspace (takes EVERY segment of the launch)
        TimedLaunch timed(h, FSD_KERNEL_MERGE, S, max_segment, stream);
        k3_merge_global_kernel<<<S, 512, 0, stream>>>(p);
        FSD_CUDA(cudaGetLastError());
        h->launches += 1;
        return FSD_OK;
    }
    {
        // shared-memory kernel in up to two tiers, so that a launch whose CAPACITY is large does not make every small segment
        // pay for 196 KB of shared memory (one CTA per SM): tier A takes the segments of at most 1024 boxes (48 KB, several CTAs
        // per SM), tier B those of 1025..4096; a CTA returns at once when its segment belongs to another launch
        FSD_CUDA(cudaFuncSetAttribute(k3_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, K3_SMEM_MAX_P * K3_BYTES_PER_BOX));
        p.use_global = 0;
        const int PA = p.P < 1024 ? p.P : 1024;
        p.n_lo = -1; p.n_hi = PA;
        {
            TimedLaunch timed(h, FSD_KERNEL_MERGE, S, max_segment, stream);
            k3_merge_kernel<<<S, PA <= 64 ? 64 : (PA <= 128 ? 128 : (PA <= 256 ? 256 : 512)), (size_t)PA * K3_BYTES_PER_BOX, stream>>>(p);
        }
        FSD_CUDA(cudaGetLastError());
        h->launches += 1;
        if (p.P > 1024 && p.cluster_min > 1024) {
            const int PB = p.P < p.cluster_min ? p.P : p.cluster_min;
            p.n_lo = 1024; p.n_hi = PB;
            {
                TimedLaunch timed(h, FSD_KERNEL_MERGE, S, -max_segment, stream);
                k3_merge_kernel<<<S, 512, (size_t)PB * K3_BYTES_PER_BOX, stream>>>(p);
            }
            FSD_CUDA(cudaGetLastError());
            h->launches += 1;
        }
    }
    if (p.P > p.cluster_min) {
        // segments above the hand-over size (4096 boxes by default): a cluster of 8 CTAs each (k3_merge_cluster_kernel)
        const size_t csmem = (size_t)(p.P / K3_CLUSTER) * 24;  // sort block (12 B / key) and, after it, the own-rank cache (24 B / rank)
        FSD_CUDA(cudaFuncSetAttribute(k3_merge_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)csmem));
        TimedLaunch timed(h, FSD_KERNEL_MERGE, S, -max_segment, stream);
        k3_merge_cluster_kernel<<<S * K3_CLUSTER, 512, csmem, stream>>>(p);
        FSD_CUDA(cudaGetLastError());
        h->launches += 1;
    }
    return FSD_OK;
}
