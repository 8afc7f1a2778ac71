// Kernel 10: 1x1 convolution (+ bias + activation [+ residual] -> concat slot [+ second destination]) on the 5th-generation
// tensor cores: TMA-fed, tcgen05.mma with the accumulator in tensor memory, warp-specialised and persistent.
//
// Same contract as Kernel 7 (fsd_pointwise_conv, include/fsd_b200.h) — it is reached through that entry point — for the ultralytics
// Conv(c1, c2, 1, 1) layers of the YOLO11 graph (C3k2.cv1/cv2, C3k.cv1-3, SPPF.cv1/cv2, C2PSA, head cv3; run from
// utils/yolo_wrapper.py:72).  A 1x1 convolution over channels-last activations is the GEMM
//     D[P pixels, N] = X[P, K] * W[N, K]^T            (both operands K-major in memory)
// with P in the millions and K, N <= 512 / 256: 2..8 flop per byte, i.e. bound by HBM, not by the tensor cores.  What the
// tensor-core path buys is everything around the math: no per-thread fragment loads, no register accumulators, one thread
// issuing the MMAs while TMA streams the activations and four warps drain finished tiles.
//
//   warp 0   : TMA producer.  The weights (<= 96 KB) are loaded once per CTA; activation tiles of 128 pixels stream through a ring
//              of KS-channel slabs (KS = 64 / 32 / 16 channels = one 128 / 64 / 32-byte swizzle span; UTMALDG.2D).
//   warp 1   : allocates tensor memory (2 accumulators x N fp32 columns) and issues tcgen05.mma M=128, N, K=16 per slab step (UTCHMMA);
//              tcgen05.commit releases ring slots and hands accumulators to the epilogue.
//   warps 2-5: epilogue.  tcgen05.ld (LDTM) 32 lanes x 16 columns, + bias, activation, fp16, staged through shared memory so that the
//              global stores (and the residual loads) are row-contiguous 16-byte vectors, exactly as in Kernel 7.
// The accumulator is double buffered, so the epilogue of tile t overlaps the loads and MMAs of tile t+1.
//
// The same kernel is the 3x3 convolution (stride 1, pad 1; fsd_conv3x3, ultralytics Conv(c1, c2, 3, 1) inside Bottleneck / C3k / the head
// branches) as an implicit GEMM: a tile is a 16 x 8 pixel patch of one image, and each of the nine taps is ONE 4-D TMA box
// [KS channels, 16, 8, 1] at the patch origin shifted by (kx - 1, ky - 1) — TMA's out-of-bounds zero fill is the padding — landing in
// the very same K-major swizzled slab layout as the 1x1 case.  The MMA loop runs over 9 x (C / KS) slabs against the tap-major weights
// [3][3][N][C] resident in shared memory; the epilogue maps accumulator rows back to (y, x).  The nine boxes of a tile overlap, so the
// input is fetched from HBM once and from L2 nine times.
#include "fsd_common.cuh"

namespace fsd {

// epilogue warps per CTA: EW / 4 warps per tensor-memory lane quarter, interleaving 16-column steps.  8 when several CTAs share an SM;
// 16 when the resident weights leave room for only one CTA (its epilogue must then hide its own latencies: SiLU alone is 16 N cycles
// of MUFU per tile and SM)
constexpr int K10_EPI_WARPS_MAX = 16;
constexpr int K10_TILE = 128;          // pixels per tile = MMA M
constexpr int K10_CHUNK = 16;          // accumulator columns per epilogue step (one tcgen05.ld.32x32b.x16)
constexpr int K10_STAGE_PITCH = 24;    // halves per staged output row (16 columns + 16 bytes: conflict-free 16-byte accesses)
constexpr int K10_MAX_STAGES = 16;

struct K10Params {
    const __half* bias;
    __half* out;
    const __half* res;
    __half* out2;
    long long P;
    int K, N, KS, n_slabs, stages, n_tiles, tmem_cols;
    int out_stride, res_stride, out2_stride, out2_c0;
    float slope;
    uint32_t b_bytes, b_region, slab_bytes, sbo_bytes, layout_type;  // b_region = b_bytes rounded up to 1024
    int taps, total_slabs, n_valid;           // 1 or 9 taps; total_slabs = taps * n_slabs; columns actually stored
    int H, W, tiles_x, tiles_per_image;       // 3x3 only: image size and the patch grid (16 x 8, halo mode: 8 x 16)
    int a_per_tile, baseoff_mode;             // ring slots one tile consumes; base-offset rule of the halo descriptors (measurement knob)
    int kh, kw, pad;                          // filter taps and padding: 3, 3, 1 or (halo mode only) 2, 2, 0
    int bias_in_mma;                          // 1: the bias enters through one extra K = 16 MMA step (ones column x bias row), 0: added in the epilogue
    uint32_t tx_bytes;                        // bytes one TMA box delivers into a ring slot (<= slab_bytes, the slot pitch)
    int halo_w;                               // halo mode: pixel columns of the patch (8 + kw - 1, or 16 with FSD_C3_HALO_W16=1)
    long long watchdog_cycles;                // > 0: a wait longer than this traps (debugging aid, FSD_K10_WATCHDOG_S); 0: wait for ever
};

// ---- PTX wrappers ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void k10_mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// mbarrier wait.  With FSD_K10_WATCHDOG_S=<seconds> a wait that long traps instead of hanging the GPU (bring-up aid; off by default: a
// kernel frozen by a profiler or by time-slicing with another context must not be killed for it).  try_wait carries a suspend-time hint, so a waiting
// thread mostly sleeps in hardware instead of spinning through issue slots that the epilogue warps need (the waits were 28 % of all
// instructions in profiles/r2_conv3_small_tc.summary.txt); `backoff_ns` adds a sleep between polls for waits that are not latency
// critical (the producer's free-slot wait).
__device__ __forceinline__ void k10_mbar_wait(uint64_t* bar, uint32_t parity, long long watchdog_cycles, unsigned backoff_ns = 0) {
    const uint32_t addr = smem_u32(bar);
    long long t0 = 0;
    for (unsigned polls = 1;; ++polls) {
        uint32_t done;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.b32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(addr), "r"(parity), "r"(20000u)
            : "memory");
        if (done) return;
        if (backoff_ns) __nanosleep(backoff_ns);
        if (watchdog_cycles > 0 && (polls & 0xfffu) == 0) {
            const long long now = clock64();
            if (t0 == 0) t0 = now;
            else if (now - t0 > watchdog_cycles) __trap();
        }
    }
}
__device__ __forceinline__ void k10_tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int x, int y) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
            smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(x), "r"(y)
        : "memory");
}
__device__ __forceinline__ void k10_tma_load_4d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c, int x, int y, int n) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
            smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c), "r"(x), "r"(y), "r"(n)
        : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_alloc(uint32_t* smem_result, uint32_t cols) {  // one full warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tc_dealloc(uint32_t taddr, uint32_t cols) {  // the allocating warp
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, fp16 inputs, fp32 accumulate; issued by ONE thread
__device__ __forceinline__ void tc_mma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// the same with the two 64-bit descriptors given as (lo, hi) halves: the issue loop below only ever adds to the low words
__device__ __forceinline__ void tc_mma_f16_lohi(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                                uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "setp.ne.b32 p, %6, 0;\n\t"
        "mov.b64 da, {%1, %2};\n\t"
        "mov.b64 db, {%3, %4};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}" ::"r"(d_tmem),
        "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
        : "memory");
}
// KSTEPS MMAs of K = 16 over one K-major slab pair; descriptor low words advance by 32 bytes (2 x 16-byte units) per step.
// The issuing thread is the serial resource of this kernel (one lane issues every MMA of the CTA): with the descriptors rebuilt
// from scratch the compiler spent ~37 instructions and ~190 cycles per MMA (profiles/r2_conv3_tc: 36 MMAs = 6900 of the 7000
// cycles per tile), so everything that does not change is hoisted and the steps are unrolled.
template <int KSTEPS>
__device__ __forceinline__ void tc_issue_slab(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                              bool clears) {
#pragma unroll
    for (int k = 0; k < KSTEPS; ++k) tc_mma_f16_lohi(d_tmem, a_lo + 2 * k, a_hi, b_lo + 2 * k, b_hi, idesc, (clears && k == 0) ? 0u : 1u);
}
// arrives on the mbarrier when every MMA issued so far by this thread has completed (implies fence::before_thread_sync)
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t (&v)[16]) {  // 32 lanes x 16 fp32 columns: lane = row
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// shared-memory matrix descriptor, K-major operand in a TMA-swizzled slab (rows of KS halves = one swizzle span, 8-row groups
// `sbo` bytes apart); layout: 2 = 128-byte swizzle, 4 = 64-byte, 6 = 32-byte (cute::UMMA::SmemDescriptor, version 1 = sm_100)
__device__ __forceinline__ uint64_t tc_smem_desc(uint32_t saddr, uint32_t sbo_bytes, uint32_t layout, uint32_t base_offset = 0) {
    uint64_t d = (uint64_t)((saddr & 0x3ffffu) >> 4);
    d |= (uint64_t)(base_offset & 7u) << 49;      // phase of the start address inside the swizzle pattern (0 when pattern-aligned)
    d |= (uint64_t)1 << 16;                       // leading byte offset (unused for swizzled K-major; CuTe writes 1)
    d |= (uint64_t)(sbo_bytes >> 4) << 32;        // stride byte offset between 8-row groups
    d |= (uint64_t)1 << 46;                       // descriptor version
    d |= (uint64_t)layout << 61;
    return d;
}

template <int ACT>
__device__ __forceinline__ float k10_act(float v, float slope) {
    if (ACT == 3) return tanh_silu(v);
    if (ACT == 1) return fast_silu(v);
    if (ACT == 2) return v > 0.f ? v : v * slope;
    return v;
}

// MODE 0: 1x1 (flat pixel tiles); 1: 3x3, one TMA box per tap; 2: 3x3, ONE halo box per tile, the taps are shifted descriptors
template <int ACT, int MINB, int MODE, int EW>
__global__ void __launch_bounds__(64 + 32 * EW, MINB)
k10_pointwise_tc_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w, const K10Params p) {
    constexpr bool CONV3 = MODE != 0;
    extern __shared__ uint8_t k10_raw[];
    __shared__ uint64_t full_bar[K10_MAX_STAGES], empty_bar[K10_MAX_STAGES], b_bar, acc_full[2], acc_empty[2];
    __shared__ uint32_t tmem_base_s;
    __shared__ __align__(16) float s_bias[256];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // dynamic smem (1024-byte aligned for the 128-byte swizzle atoms): [weights | activation ring | 4 x epilogue staging]
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(k10_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* smem_b = smem;
    uint8_t* smem_a = smem + p.b_region;
    __half* stage_base = reinterpret_cast<__half*>(smem_a + (size_t)p.stages * p.slab_bytes);

    constexpr int K10_THREADS = 64 + 32 * EW, SETS = EW / 4;
    for (int i = threadIdx.x; i < p.N; i += K10_THREADS) s_bias[i] = i < p.n_valid ? __half2float(__ldg(p.bias + i)) : 0.f;
    // Bias through the tensor cores: one extra K = 16 step whose A operand is a constant [128 x 16] tile with ones in column 0 and whose
    // B operand is [N x 16] with the bias in column 0 — the accumulator then already holds conv + bias and the epilogue saves one FADD per
    // element (it is bound by its instruction count).  Both tiles use the un-swizzled K-major canonical layout: 8-row x 16-byte core
    // matrices of 128 contiguous bytes, M/N-direction stride (SBO) 128 bytes, K-direction stride (LBO) = rows x 16 bytes.
    uint8_t* ones_tile = reinterpret_cast<uint8_t*>(stage_base + (size_t)EW * 32 * K10_STAGE_PITCH);
    uint8_t* bias_tile = ones_tile + 4096;
    if (p.bias_in_mma) {
        for (int i = threadIdx.x; i < 256 + 2 * p.N; i += K10_THREADS) {  // 16-byte pieces: 256 of the ones tile, 2 N of the bias tile
            uint4 v = make_uint4(0, 0, 0, 0);
            if (i < 128) v.x = 0x3c00u;                                              // row i (k-chunk 0): column 0 = 1.0
            else if (i >= 256 && i < 256 + p.N && i - 256 < p.n_valid) v.x = (uint32_t)__half_as_ushort(__ldg(p.bias + (i - 256)));
            if (i < 256) reinterpret_cast<uint4*>(ones_tile)[i] = v;
            else reinterpret_cast<uint4*>(bias_tile)[i - 256] = v;
        }
        fence_proxy_async();  // generic-proxy writes -> visible to the tensor core's (async proxy) operand reads
    }
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_x);
        tma_prefetch_desc(&map_w);
        for (int i = 0; i < p.stages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
        mbar_init(&b_bar, 1);
        for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 4 * (p.N / K10_CHUNK < SETS ? p.N / K10_CHUNK : SETS)); }
        fence_barrier_init();
    }
    if (warp == 1) tc_alloc(&tmem_base_s, (uint32_t)p.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;

    if (warp == 0) {
        // ================= TMA producer =================
        if (lane == 0) {
            mbar_expect_tx(&b_bar, p.b_bytes);
            for (int s = 0; s < p.total_slabs; ++s) {
                uint8_t* dst = smem_b + (size_t)s * p.N * p.KS * 2;
                if (CONV3) tma_load_3d(dst, &map_w, &b_bar, (s % p.n_slabs) * p.KS, 0, s / p.n_slabs);
                else k10_tma_load_2d(dst, &map_w, &b_bar, s * p.KS, 0);
            }
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
                int n = 0, ty = 0, tx = 0;
                if (CONV3) {
                    n = tile / p.tiles_per_image;
                    const int rem = tile - n * p.tiles_per_image;
                    ty = rem / p.tiles_x;
                    tx = rem - ty * p.tiles_x;
                }
                for (int s = 0; s < p.a_per_tile; ++s) {
                    k10_mbar_wait(&empty_bar[stage], phase ^ 1, p.watchdog_cycles, 64);
                    mbar_expect_tx(&full_bar[stage], p.tx_bytes);
                    uint8_t* dst = smem_a + (size_t)stage * p.slab_bytes;
                    if (MODE == 2) {  // the whole 18 x 16 pixel halo patch (rows y0-1.., columns x0-1..x0+14) of the 16 x 8 output tile
                        k10_tma_load_4d(dst, &map_x, &full_bar[stage], 0, tx * 8 - p.pad, ty * 16 - p.pad, n);
                    } else if (MODE == 1) {
                        const int tap = s / p.n_slabs, cs = s - tap * p.n_slabs;
                        const int ky = tap / 3, kx = tap - 3 * ky;
                        k10_tma_load_4d(dst, &map_x, &full_bar[stage], cs * p.KS, tx * 16 + kx - p.pad, ty * 8 + ky - p.pad, n);
                    } else {
                        k10_tma_load_2d(dst, &map_x, &full_bar[stage], s * p.KS, tile * K10_TILE);
                    }
                    if (++stage == p.stages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        // The whole warp walks the loop (uniform control flow lets the compiler keep the descriptors in uniform registers) and ONE
        // elected lane issues the MMAs and commits of a tile.
        {
            // instruction descriptor (cute::UMMA::InstrDescriptor): D = F32, A = B = F16, both K-major, N >> 3, M >> 4
            const uint32_t idesc = (1u << 4) | ((uint32_t)(p.N >> 3) << 17) | ((uint32_t)(K10_TILE >> 4) << 24);
            // shared-memory descriptors (cute::UMMA::SmemDescriptor, K-major): low word = address >> 4 | LBO (1, unused) << 16, high word =
            // SBO >> 4 | version 1 << 14 | swizzle mode << 29; base offset 0 — the swizzle phase follows the absolute address (measured)
            const uint32_t a_lo_base = ((smem_u32(smem_a) & 0x3ffffu) >> 4) | (1u << 16);
            const uint32_t b_lo_base = ((smem_u32(smem_b) & 0x3ffffu) >> 4) | (1u << 16);
            const uint32_t hi_common = (1u << 14) | (p.layout_type << 29);
            const uint32_t a_hi = (p.sbo_bytes >> 4) | hi_common, b_hi = a_hi;
            const uint32_t row16 = (uint32_t)p.KS * 2 >> 4;                  // one pixel row of a slab in 16-byte units
            const uint32_t a_hi_halo = ((uint32_t)p.halo_w * row16) | hi_common;  // halo patch: 8-row groups are halo_w pixel rows apart
            const uint32_t slab16 = p.slab_bytes >> 4, bslab16 = (uint32_t)(p.N * p.KS * 2) >> 4;
            const int ksteps = p.KS >> 4;
            // un-swizzled descriptors of the bias step: LBO (K direction) = rows x 16 bytes, SBO (M/N direction) = 128 bytes, layout 0
            const uint32_t ones_lo = ((smem_u32(ones_tile) & 0x3ffffu) >> 4) | ((2048u >> 4) << 16), ones_hi = (128u >> 4) | (1u << 14);
            const uint32_t biasd_lo = ((smem_u32(bias_tile) & 0x3ffffu) >> 4) | (((uint32_t)p.N * 16u >> 4) << 16), biasd_hi = ones_hi;
            k10_mbar_wait(&b_bar, 0, p.watchdog_cycles);
            int stage = 0, it = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++it) {
                const int acc = it & 1;
                k10_mbar_wait(&acc_empty[acc], ((it >> 1) & 1) ^ 1, p.watchdog_cycles);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * p.N);
                if (MODE == 2) {
                    // nine taps = nine descriptors into the one halo patch: pixel (yy, xx) sits at (yy * 16 + xx) * row bytes, so tap
                    // (ky, kx) starts (ky * halo_w + kx) rows in and the 8-pixel row groups are halo_w rows apart (halo_w = 10 for 3 x 3: the
                    // group stride need not be a multiple of the 1024-byte swizzle pattern, because the phase follows the absolute address)
                    k10_mbar_wait(&full_bar[stage], phase, p.watchdog_cycles);
                    tc_fence_after();
                    if (elect_one_sync()) {
                        const uint32_t a_lo0 = a_lo_base + (uint32_t)stage * slab16;
#pragma unroll
                        for (int ky = 0; ky < 3; ++ky) {
#pragma unroll
                            for (int kx = 0; kx < 3; ++kx) {
                                if (ky >= p.kh || kx >= p.kw) continue;  // (2 x 2 filters use four of the nine unrolled slots)
                                const uint32_t a_lo = a_lo0 + (uint32_t)(ky * p.halo_w + kx) * row16;
                                const uint32_t b_lo = b_lo_base + (uint32_t)(ky * p.kw + kx) * bslab16;
                                const bool first = ky == 0 && kx == 0;
                                // (measurement knob FSD_C3_BASEOFF=1: write the start address's pattern phase into the base-offset field,
                                // bits 49-51 — this is what BREAKS the results; the default leaves it 0)
                                const uint32_t a_hi_tap = p.baseoff_mode == 1 ? a_hi_halo | (((a_lo >> 3) & 7u) << 17) : a_hi_halo;
                                if (ksteps == 4) tc_issue_slab<4>(d_tmem, a_lo, a_hi_tap, b_lo, b_hi, idesc, first);
                                else if (ksteps == 2) tc_issue_slab<2>(d_tmem, a_lo, a_hi_tap, b_lo, b_hi, idesc, first);
                                else tc_issue_slab<1>(d_tmem, a_lo, a_hi_tap, b_lo, b_hi, idesc, first);
                            }
                        }
                        tc_commit(&empty_bar[stage]);
                        if (p.bias_in_mma) tc_mma_f16_lohi(d_tmem, ones_lo, ones_hi, biasd_lo, biasd_hi, idesc, 1u);
                        tc_commit(&acc_full[acc]);  // accumulator complete -> epilogue
                    }
                    __syncwarp();
                    if (++stage == p.stages) { stage = 0; phase ^= 1; }
                } else {
                    uint32_t b_lo = b_lo_base;
                    for (int s = 0; s < p.total_slabs; ++s, b_lo += bslab16) {
                        k10_mbar_wait(&full_bar[stage], phase, p.watchdog_cycles);
                        tc_fence_after();
                        if (elect_one_sync()) {
                            const uint32_t a_lo = a_lo_base + (uint32_t)stage * slab16;
                            if (ksteps == 4) tc_issue_slab<4>(d_tmem, a_lo, a_hi, b_lo, b_hi, idesc, s == 0);
                            else if (ksteps == 2) tc_issue_slab<2>(d_tmem, a_lo, a_hi, b_lo, b_hi, idesc, s == 0);
                            else tc_issue_slab<1>(d_tmem, a_lo, a_hi, b_lo, b_hi, idesc, s == 0);
                            tc_commit(&empty_bar[stage]);  // the slab may be overwritten once these MMAs have read it
                            if (s + 1 == p.total_slabs) {
                                if (p.bias_in_mma) tc_mma_f16_lohi(d_tmem, ones_lo, ones_hi, biasd_lo, biasd_hi, idesc, 1u);
                                tc_commit(&acc_full[acc]);  // accumulator complete -> epilogue
                            }
                        }
                        __syncwarp();
                        if (++stage == p.stages) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else {
        // ================= epilogue warps (TMEM lane quarter = warp % 4; warp set (warp - 2) / 4 takes every SETS-th 16-column step) ====
        const int q = warp & 3, half = (warp - 2) >> 2;
        __half* stg = stage_base + (size_t)(warp - 2) * 32 * K10_STAGE_PITCH;
        if (half * K10_CHUNK < p.N) {  // (a set beyond N / 16 has no step)
            int it = 0;
            for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++it) {
                const int acc = it & 1;
                k10_mbar_wait(&acc_full[acc], (it >> 1) & 1, p.watchdog_cycles);
                tc_fence_after();
                const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * p.N);
                // accumulator row m = 32 q + r  ->  pixel: 1x1: tile * 128 + m;  3x3: (y0 + m / 16, x0 + m % 16) of image n
                const long long pix0 = (long long)tile * K10_TILE + q * 32;
                int img = 0, y0 = 0, x0 = 0;
                if (CONV3) {
                    img = tile / p.tiles_per_image;
                    const int rem = tile - img * p.tiles_per_image;
                    const int ty = rem / p.tiles_x;
                    y0 = MODE == 2 ? ty * 16 + 4 * q : ty * 8 + 2 * q;
                    x0 = (rem - ty * p.tiles_x) * (MODE == 2 ? 8 : 16);
                }
                for (int c0 = half * K10_CHUNK; c0 < p.N; c0 += SETS * K10_CHUNK) {
                    uint32_t v[16];
                    tc_ld16(t_row + (uint32_t)c0, v);
                    tc_wait_ld();
                    if (c0 + SETS * K10_CHUNK >= p.N) {  // this warp's last read of the accumulator: hand it back before the stores
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) k10_mbar_arrive(&acc_empty[acc]);
                    }
                    uint32_t h[8];
                    const bool upper = c0 + 8 < p.n_valid;  // (N = 8 runs as a 16-column MMA: columns 8..15 are padding, skip their math)
#pragma unroll
                    for (int e4 = 0; e4 < 4; ++e4) {
                        if (e4 >= 2 && !upper) { h[2 * e4] = 0u; h[2 * e4 + 1] = 0u; continue; }
                        float4 bb = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (!p.bias_in_mma) bb = *reinterpret_cast<const float4*>(&s_bias[c0 + 4 * e4]);
                        const __half2 o0 = __floats2half2_rn(k10_act<ACT>(__uint_as_float(v[4 * e4]) + bb.x, p.slope),
                                                             k10_act<ACT>(__uint_as_float(v[4 * e4 + 1]) + bb.y, p.slope));
                        const __half2 o1 = __floats2half2_rn(k10_act<ACT>(__uint_as_float(v[4 * e4 + 2]) + bb.z, p.slope),
                                                             k10_act<ACT>(__uint_as_float(v[4 * e4 + 3]) + bb.w, p.slope));
                        h[2 * e4] = *reinterpret_cast<const uint32_t*>(&o0);
                        h[2 * e4 + 1] = *reinterpret_cast<const uint32_t*>(&o1);
                    }
                    uint4* d = reinterpret_cast<uint4*>(stg + (size_t)lane * K10_STAGE_PITCH);
                    d[0] = make_uint4(h[0], h[1], h[2], h[3]);
                    d[1] = make_uint4(h[4], h[5], h[6], h[7]);
                    __syncwarp();
                    // row-contiguous stores: two 16-byte pieces per row (+ residual, + second destination)
#pragma unroll
                    for (int i2 = 0; i2 < 2; ++i2) {
                        const int i = lane + 32 * i2;
                        const int r = i >> 1, c = i & 1;
                        long long pix = pix0 + r;
                        bool ok = pix < p.P;
                        if (CONV3) {
                            const int y = MODE == 2 ? y0 + (r >> 3) : y0 + (r >> 4), x = MODE == 2 ? x0 + (r & 7) : x0 + (r & 15);
                            ok = y < p.H && x < p.W;
                            pix = ((long long)img * p.H + y) * p.W + x;
                        }
                        const int col_first = c0 + c * 8;
                        if (ok && col_first < p.n_valid) {
                            uint4 o = *reinterpret_cast<const uint4*>(stg + (size_t)r * K10_STAGE_PITCH + c * 8);
                            const int col = c0 + c * 8;
                            if (p.res) {
                                const uint4 rr = __ldg(reinterpret_cast<const uint4*>(p.res + (size_t)pix * p.res_stride + col));
                                __half2* ho = reinterpret_cast<__half2*>(&o);
                                const __half2* hr = reinterpret_cast<const __half2*>(&rr);
#pragma unroll
                                for (int e = 0; e < 4; ++e) ho[e] = __hadd2(ho[e], hr[e]);
                            }
                            *reinterpret_cast<uint4*>(p.out + (size_t)pix * p.out_stride + col) = o;
                            if (p.out2 && col >= p.out2_c0) *reinterpret_cast<uint4*>(p.out2 + (size_t)pix * p.out2_stride + (col - p.out2_c0)) = o;
                        }
                    }
                    __syncwarp();
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        tc_fence_after();
        tc_dealloc(tmem_base, (uint32_t)p.tmem_cols);
    }
}

// Shapes the tensor-core path takes; everything else stays on Kernel 7 / the library convolution.
bool k10_supported(int K, int N) {
    return K % 16 == 0 && K >= 16 && K <= 512 && N % 16 == 0 && N >= 16 && N <= 256 && (size_t)K * N * 2 <= 96 * 1024;
}
// 3x3: N = 8 runs as a 16-column MMA whose upper half is never stored (the tap-major weights come zero-padded to 16 rows)
bool k10_conv3_supported(int K, int N) {
    const int n_mma = N == 8 ? 16 : N;
    return K % 16 == 0 && K >= 16 && K <= 256 && (N == 8 || (N % 16 == 0 && N >= 16 && N <= 256)) && (size_t)9 * K * n_mma * 2 <= 96 * 1024;
}

// 2 x 2, no padding (layer 1 of the backbone on the space-to-depth stem output): one channel slab only (halo mode)
bool k10_conv2_supported(int K, int N) {
    return (K == 16 || K == 32 || K == 64) && N % 16 == 0 && N >= 16 && N <= 256 && (size_t)4 * K * N * 2 <= 96 * 1024;
}

typedef CUresult (*k10_encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// fp16 tensor map of `rank` dimensions (dims / box innermost first; strides in bytes for dimensions 1..rank-1)
static bool k10_encode(fsd_context* h, CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides,
                       const uint32_t* box, CUtensorMapSwizzle swz) {
    cuuint64_t gdim[4], gstr[3];
    cuuint32_t bx[4], estr[4] = {1, 1, 1, 1};
    for (int i = 0; i < rank; ++i) { gdim[i] = dims[i]; bx[i] = box[i]; }
    for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides[i];
    return ((k10_encode_fn)h->encode_tiled)(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bx, estr,
                                            CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// Shared launcher.  taps = 1: x is [P, K] with row stride x_stride (elements), w is [N, K].  taps = 9: x is n_img channels-last images
// of H x W pixels (pixel stride x_stride), w is tap-major [3][3][n_mma][K]; P = n_img * H * W.
// -> FSD_OK with *taken = true when the launch was made; *taken = false: shape / resources not supported here
static int k10_launch(fsd_context* h, int taps, int kh, int pad, const void* x, int64_t x_stride, int n_img, int Hin, int Win, const void* w, const void* bias,
                      void* out, int64_t out_stride, const void* res, int64_t res_stride, void* out2, int64_t out2_stride, int out2_c0,
                      int64_t P, int K, int N, int act, float slope, cudaStream_t stream, bool* taken) {
    *taken = false;
    const int n_mma = (taps > 1 && N == 8) ? 16 : N;
    const int H = Hin - kh + 1 + 2 * pad, W = Win - kh + 1 + 2 * pad;  // output size (taps = 1: unused)
    if (!h->encode_tiled || P >= (1LL << 31) - K10_TILE) return FSD_OK;
    K10Params p;
    p.bias = (const __half*)bias; p.out = (__half*)out; p.res = (const __half*)res; p.out2 = (__half*)out2;
    p.P = P; p.K = K; p.N = n_mma; p.n_valid = N;
    p.KS = K % 64 == 0 ? 64 : (K % 32 == 0 ? 32 : 16);
    p.n_slabs = K / p.KS;
    p.taps = taps; p.total_slabs = taps * p.n_slabs;
    p.kh = kh; p.kw = kh; p.pad = pad;
    p.bias_in_mma = getenv("FSD_K10_BIAS_EPI") ? 0 : 1;
    p.watchdog_cycles = getenv("FSD_K10_WATCHDOG_S") ? (long long)(atof(getenv("FSD_K10_WATCHDOG_S")) * 2.0e9) : 0;
    p.out_stride = (int)out_stride; p.res_stride = (int)res_stride; p.out2_stride = (int)out2_stride; p.out2_c0 = out2_c0;
    p.slope = slope;
    p.b_bytes = (uint32_t)taps * K * n_mma * 2;
    p.b_region = (p.b_bytes + 1023u) & ~1023u;
    p.slab_bytes = (uint32_t)K10_TILE * p.KS * 2;
    p.sbo_bytes = 8u * p.KS * 2;
    p.layout_type = p.KS == 64 ? 2u : (p.KS == 32 ? 4u : 6u);
    const CUtensorMapSwizzle swz = p.KS == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (p.KS == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
    p.H = H; p.W = W; p.tiles_x = 1; p.tiles_per_image = 1;
    // 3x3 with one channel slab: ONE halo box per tile (FSD_C3_HALO=0 keeps one box per tap); FSD_C3_BASEOFF picks the descriptor rule
    const bool halo = taps > 1 && p.n_slabs == 1 && (taps != 9 || !(getenv("FSD_C3_HALO") && atoi(getenv("FSD_C3_HALO")) == 0));
    if (taps > 1 && taps != 9 && !halo) return FSD_OK;  // (2 x 2 filters only exist in halo mode)
    const uint32_t halo_rows = 16u + (uint32_t)kh - 1u;
    // measured (gpurun_out/r2_c3halo_*.log): the swizzle phase follows the ABSOLUTE shared-memory address, so a descriptor that starts
    // s pixel rows into a swizzle pattern needs base offset 0; writing (start >> 7) & 7 there breaks K = 32 and 64 (FSD_C3_BASEOFF=1 shows it)
    p.baseoff_mode = getenv("FSD_C3_BASEOFF") ? atoi(getenv("FSD_C3_BASEOFF")) : 0;
    p.a_per_tile = halo ? 1 : p.total_slabs;
    p.halo_w = getenv("FSD_C3_HALO_W16") ? 16 : 8 + kh - 1;
    p.tx_bytes = p.slab_bytes;
    if (halo) {
        p.tx_bytes = halo_rows * (uint32_t)p.halo_w * (uint32_t)p.KS * 2;
        p.slab_bytes = (p.tx_bytes + 1023u) & ~1023u;  // (stage bases stay pattern-aligned)
    }
    if (taps > 1) {
        p.tiles_x = halo ? (W + 7) / 8 : (W + 15) / 16;
        p.tiles_per_image = p.tiles_x * (halo ? (H + 15) / 16 : (H + 7) / 8);
        if ((int64_t)p.tiles_per_image * n_img >= (1LL << 31)) return FSD_OK;
        p.n_tiles = p.tiles_per_image * n_img;
    } else {
        p.n_tiles = (int)((P + K10_TILE - 1) / K10_TILE);
    }
    int cols = 32;
    while (cols < 2 * n_mma) cols <<= 1;
    p.tmem_cols = cols;
    // CTAs per SM: one CTA's epilogue warps cannot hide their own tensor-memory / shared-memory / store latencies (one tile per
    // ~2800 cycles measured with a single CTA of four epilogue warps per SM at K = N = 32, against 700 at the HBM roofline), so small
    // shapes run three (optionally four) CTAs per SM — bounded by tensor memory (512 columns per SM) and by a ring of >= 2 slabs per CTA
    // CTAs per SM: one CTA's epilogue warps cannot hide their own tensor-memory / shared-memory / store latencies (one tile per ~2800
    // cycles with a single CTA of four epilogue warps per SM at K = N = 32, against 700 at the HBM roofline), so small shapes run three
    // (optionally four) CTAs per SM — bounded by tensor memory (512 columns per SM) and by a ring of >= 2 slabs per CTA; a shape whose
    // resident weights leave room for one CTA only gets sixteen epilogue warps instead of eight
    int want = getenv("FSD_K10_CTAS") ? atoi(getenv("FSD_K10_CTAS")) : 3;  // 4 selects the 48-register build of the kernel
    if (want > 4) want = 4;
    int ctas = 1, stages = 0;
    size_t smem = 0;
    bool wide_epilogue = false;
    auto plan = [&](int epi_warps, int first_ctas) {
        const size_t staging = (size_t)epi_warps * 32 * K10_STAGE_PITCH * sizeof(__half) + 4096 + (size_t)n_mma * 32;  // + the bias step's tiles
        for (ctas = first_ctas; ctas >= 1; --ctas) {
            const size_t budget = (size_t)216 * 1024 / ctas - 3 * 1024;  // static shared memory + the driver's 1 KB per CTA
            const size_t fixed = 1024 + p.b_region + staging;
            if (budget < fixed + 2 * (size_t)p.slab_bytes) continue;
            stages = (int)((budget - fixed) / p.slab_bytes);
            if (stages > K10_MAX_STAGES) stages = K10_MAX_STAGES;
            smem = budget;  // (the padded request also keeps more CTAs than tensor memory allows from sharing an SM)
            return true;
        }
        stages = 0;
        return false;
    };
    const int tmem_ctas = 512 / p.tmem_cols < want ? 512 / p.tmem_cols : want;
    if (!plan(8, tmem_ctas < 1 ? 1 : tmem_ctas)) return FSD_OK;
    if (ctas == 1 && n_mma >= 64 && !(getenv("FSD_K10_EW") && atoi(getenv("FSD_K10_EW")) == 8)) {
        const int s8 = stages;
        const size_t m8 = smem;
        if (plan(16, 1)) wide_epilogue = true;
        else { stages = s8; smem = m8; ctas = 1; }
    }
    if (stages < 2) return FSD_OK;
    p.stages = stages;

    CUtensorMap mx, mw;
    if (taps > 1) {
        const uint64_t xd[4] = {(uint64_t)K, (uint64_t)Win, (uint64_t)Hin, (uint64_t)n_img};
        const uint64_t xs[3] = {(uint64_t)x_stride * 2, (uint64_t)Win * x_stride * 2, (uint64_t)Hin * Win * x_stride * 2};
        const uint32_t xb[4] = {(uint32_t)p.KS, halo ? (uint32_t)p.halo_w : 16u, halo ? halo_rows : 8u, 1};
        if (!k10_encode(h, &mx, x, 4, xd, xs, xb, swz)) return FSD_OK;
        const uint64_t wd[3] = {(uint64_t)K, (uint64_t)n_mma, (uint64_t)taps};
        const uint64_t ws[2] = {(uint64_t)K * 2, (uint64_t)n_mma * K * 2};
        const uint32_t wb[3] = {(uint32_t)p.KS, (uint32_t)n_mma, 1};
        if (!k10_encode(h, &mw, w, 3, wd, ws, wb, swz)) return FSD_OK;
    } else {
        const uint64_t xd[2] = {(uint64_t)K, (uint64_t)P};
        const uint64_t xs[1] = {(uint64_t)x_stride * 2};
        const uint32_t xb[2] = {(uint32_t)p.KS, K10_TILE};
        if (!k10_encode(h, &mx, x, 2, xd, xs, xb, swz)) return FSD_OK;
        const uint64_t wd[2] = {(uint64_t)K, (uint64_t)N};
        const uint64_t ws[1] = {(uint64_t)K * 2};
        const uint32_t wb[2] = {(uint32_t)p.KS, (uint32_t)N};
        if (!k10_encode(h, &mw, w, 2, wd, ws, wb, swz)) return FSD_OK;
    }

    const int grid = p.n_tiles < h->sm_count * ctas ? p.n_tiles : h->sm_count * ctas;
#define K10_GO2(ACT, C3)                                                                                                \
    {                                                                                                                   \
        auto kern = wide_epilogue ? k10_pointwise_tc_kernel<ACT, 1, C3, 16>                                            \
                                  : (ctas >= 4 ? k10_pointwise_tc_kernel<ACT, 4, C3, 8> : k10_pointwise_tc_kernel<ACT, 3, C3, 8>); \
        FSD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                  \
        kern<<<grid, wide_epilogue ? 64 + 32 * 16 : 64 + 32 * 8, smem, stream>>>(mx, mw, p);                            \
    }
#define K10_GO(ACT) { if (halo) K10_GO2(ACT, 2) else if (taps > 1) K10_GO2(ACT, 1) else K10_GO2(ACT, 0) }
    {
        // algorithmic bytes: input + output (+ residual, + second destination) once
        TimedLaunch timed(h, taps > 1 ? FSD_KERNEL_CONV3X3 : FSD_KERNEL_POINTWISE,
                          ((int64_t)(taps > 1 ? (int64_t)n_img * Hin * Win : P) * K + (int64_t)P * (N + (res ? N : 0) + (out2 ? N - out2_c0 : 0))) * 2, N, stream);
        if (act == 0) K10_GO(0) else if (act == 1 && silu_tanh_mode()) K10_GO(3) else if (act == 1) K10_GO(1) else K10_GO(2)
    }
#undef K10_GO
#undef K10_GO2
    FSD_CUDA(cudaGetLastError());
    h->launches += 1;
    *taken = true;
    return FSD_OK;
}

int launch_pointwise_tc(fsd_context* h, const void* x, int64_t x_stride, const void* w, const void* bias, void* out, int64_t out_stride,
                        const void* res, int64_t res_stride, void* out2, int64_t out2_stride, int out2_c0, int64_t P, int K, int N,
                        int act, float slope, cudaStream_t stream, bool* taken) {
    *taken = false;
    if (!k10_supported(K, N)) return FSD_OK;
    return k10_launch(h, 1, 1, 0, x, x_stride, 1, 1, 1, w, bias, out, out_stride, res, res_stride, out2, out2_stride, out2_c0, P, K, N, act, slope,
                      stream, taken);
}

}  // namespace fsd

using namespace fsd;

extern "C" int fsd_conv3x3_supported(int in_channels, int out_channels) { return k10_conv3_supported(in_channels, out_channels) ? 1 : 0; }

extern "C" int fsd_conv3x3(fsd_handle_t h, const void* x, int64_t x_pixel_stride, int n_images, int H, int W, const void* weight_taps,
                           const void* bias, void* out, int64_t out_pixel_stride, const void* residual, int64_t residual_pixel_stride,
                           int in_channels, int out_channels, int act, float slope, int dtype, void* stream_) {
    FSD_CHECK_ARG(h && x && weight_taps && bias && out, "fsd_conv3x3: null argument");
    FSD_CHECK_ARG(dtype == FSD_F16, "fsd_conv3x3: only fp16 is implemented");
    FSD_CHECK_ARG(n_images >= 0 && H > 0 && W > 0 && act >= 0 && act <= 2, "fsd_conv3x3: bad sizes / activation");
    FSD_CHECK_ARG(k10_conv3_supported(in_channels, out_channels), "fsd_conv3x3: unsupported shape %d -> %d (see fsd_conv3x3_supported)",
                  in_channels, out_channels);
    FSD_CHECK_ARG(x_pixel_stride >= in_channels && x_pixel_stride % 8 == 0, "fsd_conv3x3: bad input stride");
    FSD_CHECK_ARG(out_pixel_stride >= out_channels && out_pixel_stride % 8 == 0, "fsd_conv3x3: bad output stride");
    FSD_CHECK_ARG(!residual || (residual_pixel_stride >= out_channels && residual_pixel_stride % 8 == 0), "fsd_conv3x3: bad residual stride");
    if (((uintptr_t)x & 15) || ((uintptr_t)weight_taps & 15) || ((uintptr_t)out & 15) || ((uintptr_t)residual & 15)) {
        set_error("fsd_conv3x3: pointers must be 16-byte aligned");
        return FSD_ERR_ALIGN;
    }
    if (n_images == 0) return FSD_OK;
    FSD_CUDA(cudaSetDevice(h->device));
    bool taken = false;
    const int rc = k10_launch(h, 9, 3, 1, x, x_pixel_stride, n_images, H, W, weight_taps, bias, out, out_pixel_stride, residual, residual_pixel_stride,
                              nullptr, 0, 0, (int64_t)n_images * H * W, in_channels, out_channels, act, slope, (cudaStream_t)stream_, &taken);
    if (rc != FSD_OK) return rc;
    if (!taken) {
        set_error("fsd_conv3x3: the tensor-core path could not be set up for %d -> %d at %dx%d (tensor map / shared memory)", in_channels,
                  out_channels, H, W);
        return FSD_ERR_CAPACITY;
    }
    return FSD_OK;
}

extern "C" int fsd_conv2x2_supported(int in_channels, int out_channels) { return k10_conv2_supported(in_channels, out_channels) ? 1 : 0; }

extern "C" int fsd_conv2x2(fsd_handle_t h, const void* x, int64_t x_pixel_stride, int n_images, int H, int W, const void* weight_taps,
                           const void* bias, void* out, int64_t out_pixel_stride, int in_channels, int out_channels, int act, float slope,
                           int dtype, void* stream_) {
    FSD_CHECK_ARG(h && x && weight_taps && bias && out, "fsd_conv2x2: null argument");
    FSD_CHECK_ARG(dtype == FSD_F16, "fsd_conv2x2: only fp16 is implemented");
    FSD_CHECK_ARG(n_images >= 0 && H > 1 && W > 1 && act >= 0 && act <= 2, "fsd_conv2x2: bad sizes / activation");
    FSD_CHECK_ARG(k10_conv2_supported(in_channels, out_channels), "fsd_conv2x2: unsupported shape %d -> %d (see fsd_conv2x2_supported)",
                  in_channels, out_channels);
    FSD_CHECK_ARG(x_pixel_stride >= in_channels && x_pixel_stride % 8 == 0, "fsd_conv2x2: bad input stride");
    FSD_CHECK_ARG(out_pixel_stride >= out_channels && out_pixel_stride % 8 == 0, "fsd_conv2x2: bad output stride");
    if (((uintptr_t)x & 15) || ((uintptr_t)weight_taps & 15) || ((uintptr_t)out & 15)) {
        set_error("fsd_conv2x2: pointers must be 16-byte aligned");
        return FSD_ERR_ALIGN;
    }
    if (n_images == 0) return FSD_OK;
    FSD_CUDA(cudaSetDevice(h->device));
    bool taken = false;
    const int rc = k10_launch(h, 4, 2, 0, x, x_pixel_stride, n_images, H, W, weight_taps, bias, out, out_pixel_stride, nullptr, 0, nullptr, 0, 0,
                              (int64_t)n_images * (H - 1) * (W - 1), in_channels, out_channels, act, slope, (cudaStream_t)stream_, &taken);
    if (rc != FSD_OK) return rc;
    if (!taken) {
        set_error("fsd_conv2x2: the tensor-core path could not be set up for %d -> %d at %dx%d", in_channels, out_channels, H, W);
        return FSD_ERR_CAPACITY;
    }
    return FSD_OK;
}
