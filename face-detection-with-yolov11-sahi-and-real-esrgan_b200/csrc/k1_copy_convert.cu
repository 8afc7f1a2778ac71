// Kernel 1, scale-1 path: the slice (or the full image) already has the network-input size, so "letterbox" is a pure
// copy-convert  u8 HWC -> fp16 x/255  (channel order optionally reversed, planar or channels-last output).
//
// Reference: ultralytics LetterBox (no resize when the shape already matches) + `im[..., ::-1].transpose(2,0,1)` + `/255`, reached
// from docs sahi/predict.py:142-345 through utils/yolo_wrapper.py:72; C2's full-image pass (1024x768 at imgsz 1024) is this case.
//
// One thread = 16 pixels of one row: three aligned 16-byte loads (48 source bytes), byte permutes into exact-integer halves
// (0x6400 | b = 1024 + b), the two-term x/255 of the other Kernel 1 paths (bit-exact with float32 division + round-to-half for
// every byte value), six (channels-last) or 3 x 2 (planar) 16-byte stores.  Read 1 byte per 2 bytes written; the write-bandwidth
// probe (benchmarks/write_probe.cu) reaches 0.81-0.85 of the copy peak for this byte mix with small CTAs in several waves, which is
// how this grid is shaped (256 threads = 4096 pixels per CTA).  An entry whose x0 is not a multiple of 16 pixels (source bytes not
// 16-byte aligned) and the last partial group of a row take a byte-load path inside the same kernel.
#include "fsd_common.cuh"

namespace fsd {

constexpr int CC_THREADS = 256;

struct CopyParams {
    const uint8_t* images;
    int64_t row_pitch, image_pitch;
    const int32_t* entries;
    __half* out;
    int w, h, reverse, groups;  // groups = ceil(w / 16)
};

// fp16 pair (byte i, byte j of the 48-byte group) / 255; i, j are compile-time after unrolling
__device__ __forceinline__ uint32_t cc_pair(const uint32_t (&s)[12], int i, int j) {
    uint32_t m;
    if ((i >> 2) == (j >> 2)) {
        m = __byte_perm(s[i >> 2], 0x64646464u, (i & 3) | (4 << 4) | ((j & 3) << 8) | (4 << 12));
    } else {
        m = __byte_perm(s[i >> 2], s[j >> 2], (i & 3) | ((4 + (j & 3)) << 8));
        m = (m & 0x00ff00ffu) | 0x64006400u;
    }
    const __half2 k1024 = __halves2half2(__ushort_as_half(0x6400), __ushort_as_half(0x6400));
    const __half2 v = __hsub2(*reinterpret_cast<__half2*>(&m), k1024);                           // exact 0..255
    const __half2 c_hi = __halves2half2(__ushort_as_half(0x1C04), __ushort_as_half(0x1C04));   // fp16(1/255)
    const __half2 c_lo = __halves2half2(__ushort_as_half(0x0001), __ushort_as_half(0x0001));   // 2^-24
    const __half2 r = __hfma2(v, c_hi, __hmul2(v, c_lo));
    return *reinterpret_cast<const uint32_t*>(&r);
}

template <bool NHWC>
__global__ void __launch_bounds__(CC_THREADS)
k1_copy_convert_kernel(const CopyParams p) {
    const int b = blockIdx.y;
    const int item = blockIdx.x * CC_THREADS + threadIdx.x;
    if (item >= p.h * p.groups) return;
    const int y = item / p.groups, g = item - y * p.groups;
    const int img = __ldg(p.entries + 3 * b + 0), x0 = __ldg(p.entries + 3 * b + 1), y0 = __ldg(p.entries + 3 * b + 2);
    const uint8_t* src = p.images + (size_t)img * p.image_pitch + (size_t)(y0 + y) * p.row_pitch + (size_t)x0 * 3 + (size_t)g * 48;
    const int px = p.w - g * 16 < 16 ? p.w - g * 16 : 16;  // pixels of this group inside the row (16, or 8 at a ragged end)
    uint32_t s[12];
    if (px == 16 && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
        const uint4* q = reinterpret_cast<const uint4*>(src);
        const uint4 a = __ldg(q), c = __ldg(q + 1), d = __ldg(q + 2);
        s[0] = a.x; s[1] = a.y; s[2] = a.z; s[3] = a.w; s[4] = c.x; s[5] = c.y; s[6] = c.z; s[7] = c.w;
        s[8] = d.x; s[9] = d.y; s[10] = d.z; s[11] = d.w;
    } else {
        const int nbytes = px * 3;
#pragma unroll
        for (int wd = 0; wd < 12; ++wd) {
            uint32_t v = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (4 * wd + k < nbytes) v |= (uint32_t)__ldg(src + 4 * wd + k) << (8 * k);
            s[wd] = v;
        }
    }
    const int r0 = p.reverse ? 2 : 0, r2 = p.reverse ? 0 : 2;  // output channel 0 / 2 <- source byte r0 / r2 of the pixel
    const size_t plane = (size_t)p.h * p.w;
    if (NHWC) {
        // output halves 3x + c <- source byte 3x + (reverse ? 2 - c : c); 24 words = 6 x 16 bytes
        uint32_t o[24];
        if (p.reverse) {
#pragma unroll
            for (int q6 = 0; q6 < 8; ++q6) {  // 6 bytes = 2 pixels -> 3 words (b2,b1) (b0,b5) (b4,b3)
                const int e = 6 * q6;
                o[3 * q6 + 0] = cc_pair(s, e + 2, e + 1);
                o[3 * q6 + 1] = cc_pair(s, e + 0, e + 5);
                o[3 * q6 + 2] = cc_pair(s, e + 4, e + 3);
            }
        } else {
#pragma unroll
            for (int wd = 0; wd < 24; ++wd) o[wd] = cc_pair(s, 2 * wd, 2 * wd + 1);
        }
        uint4* dst = reinterpret_cast<uint4*>(p.out + (size_t)b * 3 * plane + ((size_t)y * p.w + (size_t)g * 16) * 3);
#pragma unroll
        for (int k = 0; k < 6; ++k)
            if (k < 3 || px == 16) dst[k] = make_uint4(o[4 * k], o[4 * k + 1], o[4 * k + 2], o[4 * k + 3]);
    } else {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const int sb = c == 0 ? r0 : (c == 1 ? 1 : r2);
            uint32_t o[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) o[k] = (sb == 0)   ? cc_pair(s, 6 * k + 0, 6 * k + 3)
                                               : (sb == 1) ? cc_pair(s, 6 * k + 1, 6 * k + 4)
                                                           : cc_pair(s, 6 * k + 2, 6 * k + 5);
            uint4* dst = reinterpret_cast<uint4*>(p.out + ((size_t)b * 3 + c) * plane + (size_t)y * p.w + (size_t)g * 16);
            dst[0] = make_uint4(o[0], o[1], o[2], o[3]);
            if (px == 16) dst[1] = make_uint4(o[4], o[5], o[6], o[7]);
        }
    }
}

// fp16 output, identity geometry (new == src == out, no border); w % 8 == 0 is guaranteed by the caller (out_w % 8 == 0)
int launch_copy_convert(fsd_context* h, const uint8_t* images, int64_t row_pitch, int64_t image_pitch, const int32_t* entries, int B,
                        int w, int hgt, int reverse, int nhwc, void* out, cudaStream_t stream) {
    CopyParams p;
    p.images = images; p.row_pitch = row_pitch; p.image_pitch = image_pitch; p.entries = entries;
    p.out = reinterpret_cast<__half*>(out); p.w = w; p.h = hgt; p.reverse = reverse; p.groups = (w + 15) / 16;
    dim3 grid(((unsigned)(hgt * p.groups) + CC_THREADS - 1) / CC_THREADS, B);
    {
        TimedLaunch timed(h, FSD_KERNEL_GATHER, B, w, stream);
        if (nhwc) k1_copy_convert_kernel<true><<<grid, CC_THREADS, 0, stream>>>(p);
        else k1_copy_convert_kernel<false><<<grid, CC_THREADS, 0, stream>>>(p);
    }
    FSD_CUDA(cudaGetLastError());
    h->launches += 1;
    return FSD_OK;
}

}  // namespace fsd
