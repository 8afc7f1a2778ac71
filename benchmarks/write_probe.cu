// Write-bandwidth probe for B200: what can a write-dominated kernel reach against the measured COPY peak?
// (Kernel 1 writes 2-8 bytes per byte it reads; Kernel 4's crop writes 2.4; a copy writes 1.)  Build + run:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o benchmarks/_bin/write_probe benchmarks/write_probe.cu && benchmarks/_bin/write_probe
// Prints one JSON line per variant: achieved GB/s over (bytes read + bytes written), CUDA events, 20 launches after 3 warm-ups.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

enum { ST_DEFAULT = 0, ST_CS = 1, ST_WT = 2 };

template <int MODE>
__device__ __forceinline__ void store16(uint4* p, uint4 v) {
    if (MODE == ST_CS) __stcs(p, v);
    else if (MODE == ST_WT) __stwt(p, v);
    else *p = v;
}

// pure fill, grid-stride over 16-byte words; UNROLL independent stores per thread per trip
template <int MODE, int UNROLL>
__global__ void __launch_bounds__(256) fill_kernel(uint4* out, size_t words) {
    const uint4 v = make_uint4(0x3c003c00u, 0x3c003c00u, 0x3c003c00u, 0x3c003c00u);
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + (UNROLL - 1) * stride < words; i += UNROLL * stride) {
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) store16<MODE>(out + i + u * stride, v);
    }
    for (; i < words; i += stride) store16<MODE>(out + i, v);
}

// contiguous 4 KB per warp trip (each lane 8 consecutive 16-byte words = one 128-byte line per lane... no: lane-interleaved)
template <int MODE>
__global__ void __launch_bounds__(256) fill_block_kernel(uint4* out, size_t words, int words_per_cta) {
    const uint4 v = make_uint4(1, 2, 3, 4);
    for (size_t base = (size_t)blockIdx.x * words_per_cta; base < words; base += (size_t)gridDim.x * words_per_cta) {
        for (int k = threadIdx.x; k < words_per_cta && base + k < words; k += blockDim.x) store16<MODE>(out + base + k, v);
    }
}

// bulk stores: the CTA fills CHUNK bytes of shared memory once and streams it out with cp.async.bulk (DEPTH in flight)
template <int CHUNK, int DEPTH>
__global__ void __launch_bounds__(128) fill_bulk_kernel(uint8_t* out, size_t bytes) {
    extern __shared__ __align__(128) uint8_t smem[];
    for (int k = threadIdx.x; k < CHUNK / 16; k += blockDim.x) reinterpret_cast<uint4*>(smem)[k] = make_uint4(1, 2, 3, 4);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
        int inflight = 0;
        for (size_t off = (size_t)blockIdx.x * CHUNK; off + CHUNK <= bytes; off += (size_t)gridDim.x * CHUNK) {
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(out + off), "r"(s), "n"(CHUNK) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            if (++inflight >= DEPTH) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(DEPTH - 1) : "memory");
        }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}

// the byte mix of Kernel 1's scale-1 path: read 16 uint8, write 16 fp16 (x/255)
template <int MODE>
__global__ void __launch_bounds__(256) convert_kernel(const uint4* in, uint4* out, size_t words_in) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < words_in; i += stride) {
        const uint4 q = __ldg(in + i);
        const uint32_t w[4] = {q.x, q.y, q.z, q.w};
        uint32_t o[8];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const __half2 a = __floats2half2_rn((float)(w[k] & 255u) / 255.f, (float)((w[k] >> 8) & 255u) / 255.f);
            const __half2 b = __floats2half2_rn((float)((w[k] >> 16) & 255u) / 255.f, (float)(w[k] >> 24) / 255.f);
            o[2 * k] = *reinterpret_cast<const uint32_t*>(&a);
            o[2 * k + 1] = *reinterpret_cast<const uint32_t*>(&b);
        }
        store16<MODE>(out + 2 * i, make_uint4(o[0], o[1], o[2], o[3]));
        store16<MODE>(out + 2 * i + 1, make_uint4(o[4], o[5], o[6], o[7]));
    }
}

__global__ void __launch_bounds__(256) copy_kernel(const uint4* in, uint4* out, size_t words) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 3 * stride < words; i += 4 * stride) {
        uint4 a = in[i], b = in[i + stride], c = in[i + 2 * stride], d = in[i + 3 * stride];
        out[i] = a; out[i + stride] = b; out[i + 2 * stride] = c; out[i + 3 * stride] = d;
    }
    for (; i < words; i += stride) out[i] = in[i];
}

template <typename F>
static void timeit(const char* name, double bytes, int grid, F launch) {
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    for (int i = 0; i < 3; ++i) launch();
    CK(cudaDeviceSynchronize());
    const int reps = 20;
    CK(cudaEventRecord(a));
    for (int i = 0; i < reps; ++i) launch();
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    CK(cudaGetLastError());
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, a, b));
    printf("{\"probe\": \"%s\", \"grid\": %d, \"mbytes\": %.1f, \"us\": %.2f, \"gbs\": %.1f}\n", name, grid, bytes / 1e6, ms * 1e3 / reps,
           bytes * reps / (ms * 1e-3) / 1e9);
    fflush(stdout);
}

int main() {
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    const size_t sizes[2] = {(size_t)453 << 20, (size_t)2400 << 20};
    uint8_t *out = nullptr, *in = nullptr;
    CK(cudaMalloc(&out, sizes[1] + 4096));
    CK(cudaMalloc(&in, sizes[1] + 4096));
    CK(cudaMemset(in, 7, sizes[1]));
    for (int s = 0; s < 2; ++s) {
        const size_t bytes = sizes[s], words = bytes / 16;
        char name[96];
        timeit("memset", (double)bytes, 0, [&] { CK(cudaMemsetAsync(out, 1, bytes)); });
        timeit("memcpy_d2d (r+w)", 2.0 * bytes, 0, [&] { CK(cudaMemcpyAsync(out, in, bytes, cudaMemcpyDeviceToDevice)); });
        for (int per_sm : {2, 4, 8, 16, 32}) {
            const int grid = sms * per_sm;
            snprintf(name, sizeof name, "fill v4 default u4 x%d", per_sm);
            timeit(name, (double)bytes, grid, [&] { fill_kernel<ST_DEFAULT, 4><<<grid, 256>>>((uint4*)out, words); });
            snprintf(name, sizeof name, "fill v4 cs u4 x%d", per_sm);
            timeit(name, (double)bytes, grid, [&] { fill_kernel<ST_CS, 4><<<grid, 256>>>((uint4*)out, words); });
        }
        timeit("fill v4 wt u4 x8", (double)bytes, sms * 8, [&] { fill_kernel<ST_WT, 4><<<sms * 8, 256>>>((uint4*)out, words); });
        timeit("fill v4 default u1 x8", (double)bytes, sms * 8, [&] { fill_kernel<ST_DEFAULT, 1><<<sms * 8, 256>>>((uint4*)out, words); });
        timeit("fill v4 default u8 x8", (double)bytes, sms * 8, [&] { fill_kernel<ST_DEFAULT, 8><<<sms * 8, 256>>>((uint4*)out, words); });
        for (int kb : {4, 16, 64}) {
            snprintf(name, sizeof name, "fill block %dKB/cta-trip x8", kb);
            timeit(name, (double)bytes, sms * 8, [&] { fill_block_kernel<ST_DEFAULT><<<sms * 8, 256>>>((uint4*)out, words, kb * 64); });
        }
        CK(cudaFuncSetAttribute(fill_bulk_kernel<32768, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768));
        for (int per_sm : {1, 2, 4}) {
            snprintf(name, sizeof name, "bulk store 4KB d8 x%d", per_sm);
            timeit(name, (double)bytes, sms * per_sm, [&] { fill_bulk_kernel<4096, 8><<<sms * per_sm, 128, 4096>>>(out, bytes); });
            snprintf(name, sizeof name, "bulk store 16KB d4 x%d", per_sm);
            timeit(name, (double)bytes, sms * per_sm, [&] { fill_bulk_kernel<16384, 4><<<sms * per_sm, 128, 16384>>>(out, bytes); });
            snprintf(name, sizeof name, "bulk store 32KB d4 x%d", per_sm);
            timeit(name, (double)bytes, sms * per_sm, [&] { fill_bulk_kernel<32768, 4><<<sms * per_sm, 128, 32768>>>(out, bytes); });
        }
        for (int per_sm : {4, 8, 16}) {
            const int grid = sms * per_sm;
            snprintf(name, sizeof name, "u8->f16 convert (r1+w2) default x%d", per_sm);
            timeit(name, 1.5 * bytes, grid, [&] { convert_kernel<ST_DEFAULT><<<grid, 256>>>((const uint4*)in, (uint4*)out, words / 2); });
            snprintf(name, sizeof name, "u8->f16 convert (r1+w2) cs x%d", per_sm);
            timeit(name, 1.5 * bytes, grid, [&] { convert_kernel<ST_CS><<<grid, 256>>>((const uint4*)in, (uint4*)out, words / 2); });
            snprintf(name, sizeof name, "copy v4 (r1+w1) x%d", per_sm);
            timeit(name, 2.0 * bytes, grid, [&] { copy_kernel<<<grid, 256>>>((const uint4*)in, (uint4*)out, words); });
        }
    }
    return 0;
}
