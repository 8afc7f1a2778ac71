// Shared host/device helpers for the sm_100a kernels behind include/fsd_b200.h.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <tuple>
#include <vector>

#include "../../include/fsd_b200.h"

namespace fsd {

void set_error(const char* fmt, ...);

#define FSD_CHECK_ARG(cond, ...)            \
    do {                                    \
        if (!(cond)) {                      \
            ::fsd::set_error(__VA_ARGS__);  \
            return FSD_ERR_ARG;             \
        }                                   \
    } while (0)

#define FSD_CUDA(call)                                                                          \
    do {                                                                                        \
        cudaError_t e__ = (call);                                                               \
        if (e__ != cudaSuccess) {                                                               \
            ::fsd::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, \
                             __LINE__);                                                         \
            return FSD_ERR_CUDA;                                                                \
        }                                                                                       \
    } while (0)

// Device-side table cached per (src,dst) resize pair (Kernel 1).
struct ResizeTable {
    int32_t* dev = nullptr;  // [dst] x int2-packed entries
    int n = 0;
};

}  // namespace fsd

// Handle: owns small device-side lookup tables and the TMA tensor-map cache.  No global state.
struct fsd_context {
    int device = 0;
    int sm_count = 0;
    int64_t launches = 0;
    std::mutex mu;
    // Kernel 1 coefficient tables keyed by (src, dst, axis)
    std::map<std::tuple<int, int, int>, fsd::ResizeTable> resize_tables;
    // Kernel 1 tensor maps keyed by (base, n, H, pitch/image_pitch, box_w, box_h)
    std::map<std::tuple<uintptr_t, int, int, int64_t, int64_t, int, int>, CUtensorMap> tensor_maps;
    std::map<std::tuple<int, int, int, int>, std::shared_ptr<void>> k1_plans;  // Kernel 1 geometry plans
    std::vector<void*> dev_allocs;  // device buffers owned by the handle (freed in fsd_destroy)
    // Kernel 2a: confidence (float bits) -> device pointer to the gate logit (computed and synchronised on first use)
    std::map<uint32_t, float*> decode_gates;
    void* encode_tiled = nullptr;  // cuTensorMapEncodeTiled, resolved through cudaGetDriverEntryPoint
    // optional device timing (fsd_kernel_timing_enable): CUDA events recorded on the launching stream right around
    // each kernel launch, inside the library, so no host-side preparation falls between the two records
    struct TimingSample { int kernel; int64_t units; int64_t tag; cudaEvent_t e0, e1; };
    unsigned timing = 0;  // bit mask of (1 << FSD_KERNEL_*)
    std::vector<TimingSample> timing_samples;
    std::vector<cudaEvent_t> event_pool;
};

namespace fsd {
// FSD_SILU=tanh selects the one-MUFU SiLU in the epilogue kernels (read per launch, so a test can toggle it)
inline bool silu_tanh_mode() {
    const char* v = getenv("FSD_SILU");
    return v && v[0] == 't';
}
// RAII bracket around one kernel launch; a no-op unless timing is enabled on the handle.
struct TimedLaunch {
    fsd_context* h; cudaStream_t stream; cudaEvent_t e1 = nullptr;
    TimedLaunch(fsd_context* h_, int kernel, int64_t units, int64_t tag, cudaStream_t s) : h(h_), stream(s) {
        if (!(h->timing >> kernel & 1u) || h->timing_samples.size() >= (1u << 20)) return;
        cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;  // a launch being captured into a CUDA graph cannot be bracketed
        if (cudaStreamIsCapturing(s, &cap) != cudaSuccess || cap != cudaStreamCaptureStatusNone) return;
        cudaEvent_t ev[2];
        for (int i = 0; i < 2; ++i) {
            if (!h->event_pool.empty()) { ev[i] = h->event_pool.back(); h->event_pool.pop_back(); }
            else if (cudaEventCreate(&ev[i]) != cudaSuccess) return;
        }
        e1 = ev[1];
        h->timing_samples.push_back({kernel, units, tag, ev[0], ev[1]});
        cudaEventRecord(ev[0], stream);
    }
    ~TimedLaunch() { if (e1) cudaEventRecord(e1, stream); }
};
}  // namespace fsd

namespace fsd {

// ---- PTX wrappers (sm_100a): mbarrier + TMA bulk tensor loads ------------------------------------
#ifdef __CUDACC__
// SiLU v * 1/(1 + 2^(-v*log2e)) on the two MUFU ops directly.  `__fdividef(v, 1 + __expf(-v))` computes the same thing but
// wraps each MUFU in denormal/range scaling (6 FMUL + 2 FSETP per value in SASS, profiles/r1_k6_stem.ncu-rep); the .ftz
// forms need none of it: a flushed denormal only turns a result of magnitude < 1e-36 into 0.
__device__ __forceinline__ float fast_silu(float v) {
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(v * -1.4426950408889634f));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
    return v * r;
}
// SiLU with ONE MUFU op: v * sigmoid(v) = h + h * tanh(h), h = v / 2.  tanh.approx is good to 2^-11 relative on tanh, i.e. the result
// is off by at most 2.4e-4 * |v| — half an fp16 rounding step for positive v, a few fp16 ulps of the (small) result for negative v.
// Opt-in (FSD_SILU=tanh): the epilogue kernels sit at 50-65 % XU-pipe utilisation with the two-MUFU form above.
__device__ __forceinline__ float tanh_silu(float v) {
    const float h = 0.5f * v;
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
    return fmaf(h, t, h);
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
// one lane of a fully converged warp (SASS: ELECT); use under a warp-uniform branch only
__device__ __forceinline__ bool elect_one_sync() {
    uint32_t pred;
    asm volatile(
        "{\n\t"
        ".reg .pred P;\n\t"
        "elect.sync _|P, 0xffffffff;\n\t"
        "selp.b32 %0, 1, 0, P;\n\t"
        "}"
        : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// 3-D tiled TMA load global -> shared, completion signalled on an mbarrier (SASS: UTMALDG).
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int x,
                                            int y, int z) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(x), "r"(y), "r"(z)
        : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
#endif

}  // namespace fsd
