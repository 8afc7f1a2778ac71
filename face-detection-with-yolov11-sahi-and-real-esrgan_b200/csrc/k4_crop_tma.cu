// Kernel 4a, TMA path — Real-ESRGAN tile crop as a bulk-copy pipeline (replaces the slicing in
// RealESRGANer.tile_process [EXT realesrgan 0.3.0], reached from utils/enhancer.py:214).
//
// A CTA owns 8 rows of one tile.  Its source bytes arrive through two 2-D TMA boxes (cp.async.bulk.tensor, SASS UTMALDG) of
// 8 rows x 640 B — the tile's 3-byte pixels start on an arbitrary byte, so the box starts at the 16-byte boundary below and
// the 0..15 byte offset is carried into the shared-memory reads — instead of seven 32-bit global loads + funnel shifts per
// 8 pixels (the LSU-queue stall that bounded the load/store version: lg_throttle 7.4, profiles/r1_k4_crop_stitch.summary.txt).
// The threads turn the staged bytes into planar RGB fp16 (byte permutes + the exact two-term /255 of Kernel 1) in shared memory,
// laid out so that each plane's 8 x pw halfs — ONE contiguous run of the packed [3, ph, pw] tile — has the same 16-byte
// phase as its destination; each run then leaves as one bulk store (cp.async.bulk.global.shared::cta, SASS UBLKCP), with at
// most 15 bytes of head / tail written by ordinary stores.
// Taken for fp16 tiles that lie inside the image (no reflect pre-/mod-pad) with an even width of at most 421 pixels — the
// reference's tile 400 / 256 with pad 10 on even-sized images; everything else stays on k4_crop_kernel.
#include "fsd_common.cuh"

namespace fsd {

constexpr int CR_ROWS = 8;      // tile rows per CTA
constexpr int CR_BOXW = 160;    // u32 per TMA box row (640 B); two boxes = 1280 B >= 421 * 3 + 15
constexpr int CR_THREADS = 256;
constexpr int CR_TT = 12;       // ints per tile-table row (k4_esrgan_tiles.cu)
constexpr int CR_MAX_PW = 421;

__device__ __forceinline__ uint32_t cr_norm255_pair(uint32_t w) {  // (vB << 16 | vA) -> half2(vA/255, vB/255), exact (Kernel 1)
    const __half2 magic = __halves2half2(__ushort_as_half(0x6400), __ushort_as_half(0x6400));
    uint32_t m = w | 0x64006400u;
    const __half2 v = __hsub2(*reinterpret_cast<__half2*>(&m), magic);
    const __half2 c_hi = __halves2half2(__ushort_as_half(0x1C04), __ushort_as_half(0x1C04));
    const __half2 c_lo = __halves2half2(__ushort_as_half(0x0001), __ushort_as_half(0x0001));
    const __half2 r = __hfma2(v, c_hi, __hmul2(v, c_lo));
    return *reinterpret_cast<const uint32_t*>(&r);
}

__device__ __forceinline__ void bulk_store(void* gdst, const void* ssrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes)
                 : "memory");
}

__global__ void __launch_bounds__(CR_THREADS)
k4_crop_tma_kernel(const __grid_constant__ CUtensorMap tmap, const int32_t* __restrict__ table, __half* __restrict__ tiles,
                   int64_t tiles_image_stride, int out_plane_bytes /* shared bytes reserved per output plane */) {
    extern __shared__ __align__(128) uint8_t cr_smem[];
    __shared__ __align__(8) uint64_t s_bar;
    const int32_t* t = table + (size_t)blockIdx.y * CR_TT;
    const int px0 = t[0], py0 = t[1], pw = t[2], ph = t[3];
    const int y0 = blockIdx.x * CR_ROWS;
    if (y0 >= ph) return;
    const int rows = min(CR_ROWS, ph - y0);
    const int img = blockIdx.z;
    __half* dst = tiles + (size_t)img * tiles_image_stride + ((int64_t)(uint32_t)t[8] | ((int64_t)t[9] << 32));
    uint32_t* s_in = reinterpret_cast<uint32_t*>(cr_smem);                       // [2][CR_ROWS][CR_BOXW]
    uint8_t* s_out = cr_smem + 2 * CR_ROWS * CR_BOXW * 4;                        // 3 planes, out_plane_bytes each
    const int xa = (px0 * 3) & ~15, off_b = px0 * 3 - xa;
    const int tid = threadIdx.x;
    if (tid == 0) {
        mbar_init(&s_bar, 1);
        fence_barrier_init();
    }
    __syncthreads();
    if (tid == 0) {
        mbar_expect_tx(&s_bar, 2 * CR_ROWS * CR_BOXW * 4);
        tma_load_3d(s_in, &tmap, &s_bar, xa >> 2, py0 + y0, img);
        tma_load_3d(s_in + CR_ROWS * CR_BOXW, &tmap, &s_bar, (xa >> 2) + CR_BOXW, py0 + y0, img);
    }
    // destination runs of the three planes and their 16-byte phase (while the boxes are in flight)
    const size_t run_halfs = (size_t)rows * pw;
    __half* grun[3];
    uint8_t* srun[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        grun[c] = dst + ((size_t)c * ph + y0) * pw;
        srun[c] = s_out + (size_t)c * out_plane_bytes + (reinterpret_cast<uintptr_t>(grun[c]) & 15);
    }
    mbar_wait(&s_bar, 0);
    // one warp per tile row, lanes over the row's 8-pixel groups: no division, and everything that depends on the row or on
    // the tile (word phase of the staged bytes, alignment class of the shared stores) is warp-uniform
    const int vecs = (pw + 7) >> 3;
    const int sh = (off_b & 3) * 8;
    const int wo = (off_b >> 2) & 1;  // the group's first byte lies in word `wo` of its 8-byte pair
    const int lane = tid & 31, r = tid >> 5;
    if (r < rows) {
        const uint32_t* in0 = s_in + r * CR_BOXW;
        for (int g = lane; g < vecs; g += 32) {
            const int x = g << 3;
            const int nx = min(8, pw - x);
            const int pi = (off_b + 24 * g) >> 3;  // first 8-byte pair of the group inside the staged 1280-byte row
            uint32_t w[8];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int p = min(pi + k, CR_BOXW - 1);  // 160 pairs per row: the clamp only touches bytes no pixel uses
                const uint2 v = *reinterpret_cast<const uint2*>(in0 + (p >= CR_BOXW / 2 ? CR_ROWS * CR_BOXW + (p - CR_BOXW / 2) * 2 : p * 2));
                w[2 * k] = v.x; w[2 * k + 1] = v.y;
            }
            uint32_t s[6];  // stream word k = bytes 4k .. 4k+3 of the group's 24 (B,G,R per pixel)
            if (wo) {
#pragma unroll
                for (int k = 0; k < 6; ++k) s[k] = __funnelshift_r(w[k + 1], w[k + 2], sh);
            } else {
#pragma unroll
                for (int k = 0; k < 6; ++k) s[k] = __funnelshift_r(w[k], w[k + 1], sh);
            }
#pragma unroll
            for (int c = 0; c < 3; ++c) {  // BGR channel c -> plane 2 - c (RGB)
                uint32_t o[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int ia = 6 * k + c, ib = ia + 3;
                    const uint32_t sel = (uint32_t)(ia & 3) | ((uint32_t)(4 * ((ib >> 2) - (ia >> 2)) + (ib & 3)) << 8);
                    o[k] = cr_norm255_pair(__byte_perm(s[ia >> 2], s[ib >> 2], sel) & 0x00ff00ffu);
                }
                uint8_t* q = srun[2 - c] + ((size_t)r * pw + x) * 2;
                if (nx == 8) {
                    const uint32_t a = smem_u32(q);
                    if ((a & 15) == 0) *reinterpret_cast<uint4*>(q) = make_uint4(o[0], o[1], o[2], o[3]);
                    else if ((a & 7) == 0) { reinterpret_cast<uint2*>(q)[0] = make_uint2(o[0], o[1]); reinterpret_cast<uint2*>(q)[1] = make_uint2(o[2], o[3]); }
                    else {
#pragma unroll
                        for (int k = 0; k < 4; ++k) reinterpret_cast<uint32_t*>(q)[k] = o[k];  // pw even -> always 4-byte aligned
                    }
                } else {
                    for (int e = 0; e < nx; ++e) reinterpret_cast<unsigned short*>(q)[e] = (unsigned short)(o[e >> 1] >> (16 * (e & 1)));
                }
            }
        }
    }
    fence_proxy_async();  // the generic-proxy writes above must be visible to the bulk (async-proxy) stores
    __syncthreads();
    if (tid < 3) {
        const int c = tid;
        uint8_t* g = reinterpret_cast<uint8_t*>(grun[c]);
        const uint8_t* sp = srun[c];
        const size_t total = run_halfs * 2;
        size_t head = (16 - (reinterpret_cast<uintptr_t>(g) & 15)) & 15;
        if (head > total) head = total;
        const size_t mid = (total - head) & ~(size_t)15;
        if (mid) bulk_store(g + head, sp + head, (uint32_t)mid);
        for (size_t i = 0; i < head; i += 2) *reinterpret_cast<unsigned short*>(g + i) = *reinterpret_cast<const unsigned short*>(sp + i);
        for (size_t i = head + mid; i < total; i += 2) *reinterpret_cast<unsigned short*>(g + i) = *reinterpret_cast<const unsigned short*>(sp + i);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // shared memory must outlive the reads of the bulk stores
    }
}

// Host side: true (and the launch is done) when every tile qualifies; false -> the caller uses the load/store kernel.
int launch_crop_tma(fsd_context* h, const uint8_t* bgr, int H, int W, int64_t row_pitch, int pre_h, int pre_w,
                    const int32_t* table_dev, const int32_t* table_host, int T, void* tiles, int n_images, int64_t image_pitch,
                    int64_t tiles_image_stride, int64_t units, cudaStream_t stream, bool* taken) {
    *taken = false;
    if (getenv("FSD_K4_NO_TMA") || pre_h != H || pre_w != W) return FSD_OK;
    if (((uintptr_t)bgr & 15) || (row_pitch & 15) || (image_pitch & 15) || ((uintptr_t)tiles & 15)) return FSD_OK;
    int max_ph = 0, max_pw = 0;
    for (int i = 0; i < T; ++i) {
        const int32_t* t = table_host + (size_t)i * CR_TT;
        if (t[0] + t[2] > W || t[1] + t[3] > H || (t[2] & 1) || t[2] > CR_MAX_PW || t[2] < 1) return FSD_OK;
        if (t[3] > max_ph) max_ph = t[3];
        if (t[2] > max_pw) max_pw = t[2];
    }
    if (n_images == 1) image_pitch = row_pitch * H;
    auto key = std::make_tuple((uintptr_t)bgr, n_images, H, row_pitch, image_pitch, CR_BOXW, CR_ROWS);
    CUtensorMap m;
    {
        std::lock_guard<std::mutex> lock(h->mu);
        auto it = h->tensor_maps.find(key);
        if (it == h->tensor_maps.end()) {
            typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                          const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                          CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
            cuuint64_t gdim[3] = {(cuuint64_t)(row_pitch / 4), (cuuint64_t)H, (cuuint64_t)n_images};
            cuuint64_t gstr[2] = {(cuuint64_t)row_pitch, (cuuint64_t)image_pitch};
            cuuint32_t box[3] = {(cuuint32_t)CR_BOXW, (cuuint32_t)CR_ROWS, 1};
            cuuint32_t estr[3] = {1, 1, 1};
            CUresult r = ((encode_fn)h->encode_tiled)(&m, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, (void*)bgr, gdim, gstr, box, estr,
                                                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                                      CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) return FSD_OK;  // e.g. a pitch the tensor-map rules reject: the load/store kernel takes over
            if (h->tensor_maps.size() > 256) h->tensor_maps.clear();
            it = h->tensor_maps.emplace(key, m).first;
        }
        m = it->second;
    }
    const int out_plane_bytes = ((CR_ROWS * max_pw * 2 + 16 + 15) / 16) * 16;
    const size_t smem = (size_t)2 * CR_ROWS * CR_BOXW * 4 + (size_t)3 * out_plane_bytes;
    dim3 grid((max_ph + CR_ROWS - 1) / CR_ROWS, T, n_images);
    {
        TimedLaunch timed(h, FSD_KERNEL_ESRGAN_CROP, units, n_images, stream);
        k4_crop_tma_kernel<<<grid, CR_THREADS, smem, stream>>>(m, table_dev, reinterpret_cast<__half*>(tiles), tiles_image_stride, out_plane_bytes);
    }
    FSD_CUDA(cudaGetLastError());
    h->launches += 1;
    *taken = true;
    return FSD_OK;
}

}  // namespace fsd
