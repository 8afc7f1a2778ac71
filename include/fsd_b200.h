/*
 * fsd_b200.h — C ABI of the B200-native sliced face-detection hot path.
 *
 * The reference (ihsanhadi57/Face-Detection-With-YOLOv11-SAHI-and-Real-ESRGAN) is pure Python and has
 * no FFI of its own: its plug-in boundary is the SAHI `DetectionModel` class protocol
 * (docs sahi/base.py:12-196) and `get_sliced_prediction` (docs sahi/predict.py:142-345).
 * This header is the boundary that sits UNDER that Python surface.  Each entry point names the
 * reference code it replaces (file:line relative to the reference root; "[EXT …]" marks arithmetic that
 * lives in an un-vendored dependency and is restated in SURVEY.md Appendix A).
 *
 * Conventions
 *   - every function returns 0 on success or a negative fsd_status; fsd_last_error() gives a
 *     thread-local message for the last failure on the calling thread;
 *   - all `dev` pointers are caller-allocated device memory, never retained or freed by the library;
 *   - kernels are enqueued on `stream` (a cudaStream_t passed as void*) and are asynchronous w.r.t. the host;
 *   - no torch types, no global state beyond the handle.
 */
#ifndef FSD_B200_H
#define FSD_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct fsd_context* fsd_handle_t;

enum fsd_status {
    FSD_OK = 0,
    FSD_ERR_ARG = -1,      /* bad argument (null pointer, negative size, unsupported enum) */
    FSD_ERR_ALIGN = -2,    /* pointer / pitch alignment not met (TMA needs 16-byte pitches) */
    FSD_ERR_CUDA = -3,     /* a CUDA runtime/driver call failed */
    FSD_ERR_CAPACITY = -4, /* problem larger than the compiled limits (see the function's doc) */
    FSD_ERR_NO_DEVICE = -5 /* no sm_100 device visible: there is NO CPU fallback */
};

enum fsd_dtype { FSD_F16 = 0, FSD_F32 = 1 };
enum fsd_merge_type { FSD_NMS = 0, FSD_GREEDYNMM = 1, FSD_NMM = 2 };
enum fsd_metric { FSD_IOU = 0, FSD_IOS = 1 };
enum fsd_layout { FSD_PLANAR = 0 /* [B,C,A] (NCHW) */, FSD_CHANNELS_LAST = 1 /* [B,A,C] (NHWC) */ };

/* ---- library / handle -------------------------------------------------------------------------- */
const char* fsd_version(void);
const char* fsd_last_error(void);
int fsd_create(int device, fsd_handle_t* out);
int fsd_destroy(fsd_handle_t h);
/* number of kernels this handle has launched since creation (bench.py's gpu_launches claim) */
int64_t fsd_launch_count(fsd_handle_t h);

/* Optional device timing for measurement (bench.py's roofline): while enabled, the instrumented kernels are bracketed
 * by CUDA events recorded on the launching stream INSIDE the library, right around the launch, so host-side preparation
 * never falls inside a sample.  enable() clears earlier samples; read() waits for each sample's end event and
 * returns rows [kernel id (FSD_KERNEL_*), units, tag, milliseconds]:
 *   FSD_KERNEL_GATHER         units = entries of the launch, tag = src_w
 *   FSD_KERNEL_DECODE         units = bytes of the head tensors (80 values per anchor), tag = anchors per entry
 *   FSD_KERNEL_MERGE          units = segments, tag = max_segment
 *   FSD_KERNEL_ESRGAN_CROP / _STITCH   units = algorithmic bytes, tag = images of the launch
 *   FSD_KERNEL_BIAS_ACT / _STEM / _POINTWISE / _SPPF / _CONV3X3 / _DWCONV   units = bytes the launch moves, tag = channels
 *   FSD_KERNEL_FINALIZE / _ATTACH / _PACK   units = entries / images
 * Launches captured into a CUDA graph cannot be bracketed (bench.py times the backbone's kernels in an eager step).
 * (the reference has no counterpart; its timing is wall-clock around model.predict, utils/yolo_wrapper.py:67-80) */
enum { FSD_KERNEL_GATHER = 1, FSD_KERNEL_DECODE = 2, FSD_KERNEL_MERGE = 3, FSD_KERNEL_ESRGAN_CROP = 4, FSD_KERNEL_BIAS_ACT = 5,
       FSD_KERNEL_STEM = 6, FSD_KERNEL_POINTWISE = 7, FSD_KERNEL_FINALIZE = 8, FSD_KERNEL_ATTACH = 9, FSD_KERNEL_PACK = 10,
       FSD_KERNEL_ESRGAN_STITCH = 11, FSD_KERNEL_SPPF = 12, FSD_KERNEL_CONV3X3 = 13, FSD_KERNEL_DWCONV = 14 };
int fsd_kernel_timing_enable(fsd_handle_t h, unsigned kernel_mask /* OR of (1u << FSD_KERNEL_*); 0 = off */);
int fsd_kernel_timing_read(fsd_handle_t h, double* samples /* [cap,4] or NULL */, int cap, int* n);

/* ---- (a2) slice plan — replaces sahi.slicing.get_slice_bboxes [EXT sahi 0.11.34], called at
 *      docs sahi/predict.py:229-238.  Host-side, integer exact.  boxes_xyxy holds up to `cap` rows of
 *      [x0,y0,x1,y1]; *n receives the number of slices (also when it exceeds cap -> FSD_ERR_CAPACITY). */
int fsd_slice_plan(int image_h, int image_w, int slice_h, int slice_w, double overlap_h_ratio,
                   double overlap_w_ratio, int32_t* boxes_xyxy, int cap, int* n);

/* ---- (a4) letterbox geometry — replaces ultralytics LetterBox(auto=True, scaleup=True, center=True)
 *      [EXT ultralytics], reached from utils/yolo_wrapper.py:74-80.  Host-side.  geom[0..7] =
 *      new_w, new_h (resized, un-padded), pad_left, pad_top, out_w, out_h, mode, reserved;
 *      mode: 0 copy (no resize), 1 cv2 INTER_LINEAR fixed point, 2 cv2 exact-2x area average.
 *      *gain = min(out_h/src_h, out_w/src_w) as ultralytics scale_boxes recomputes it. */
int fsd_letterbox_geometry(int src_h, int src_w, int imgsz, int stride, int32_t geom[8], double* gain);

/* ---- Kernel 1 (a4) fused slice gather + letterbox + u8 -> fp16/fp32 normalise.
 *      Replaces numpy slice views (docs sahi/predict.py:270-276), np.ascontiguousarray (:106) and
 *      ultralytics preprocess (cv2.resize INTER_LINEAR, copyMakeBorder 114, [..., ::-1], HWC->CHW, /255).
 *      images: dev, n_images x [H, row_pitch] bytes of HWC uint8 (3 channels), image i at i*image_pitch;
 *              base pointer, row_pitch and image_pitch must be multiples of 16 (TMA tensor-map rule).
 *      entries: dev int32 [B,3] = (image_index, x0, y0) of each source box; all boxes are src_w x src_h.
 *      out: dev [B,3,out_h,out_w] of dtype (out_h/out_w from fsd_letterbox_geometry), stored planar
 *           (FSD_PLANAR, NCHW) or channels-last (FSD_CHANNELS_LAST, [B,out_h,out_w,3]) per out_layout. */
int fsd_gather_letterbox(fsd_handle_t h, const uint8_t* images, int n_images, int H, int W,
                         int64_t row_pitch, int64_t image_pitch, const int32_t* entries, int B, int src_w,
                         int src_h, int imgsz, int stride, int reverse_channels, int dtype, int out_layout,
                         void* out, void* stream);

/* ---- Kernel 2a (a6,a7) fused pose-head decode + confidence gate + compaction.
 *      Replaces ultralytics Detect/Pose._inference, DFL, dist2bbox, kpts_decode and the `xc` candidate
 *      filter of non_max_suppression [EXT ultralytics], reached from utils/yolo_wrapper.py:74-80.
 *      Per level l (strides 8,16,32): box[l] [B,64,h_l*w_l], cls[l] [B,1,...], kpt[l] [B,15,...] in
 *      `layout`/`dtype`.  A candidate row is 24 floats:
 *        [x1,y1,x2,y2 (letterboxed px), score, anchor_index, 15 kpt values (x,y,conf)x5, 3 pad].
 *      Candidates of batch entry b are written to cand[b*cap_per_entry ...] in anchor order is NOT
 *      guaranteed; `anchor_index` makes any later ordering deterministic.  count[b] receives the number of
 *      candidates found (may exceed cap_per_entry: the excess is dropped and reported by the caller). */
int fsd_pose_decode(fsd_handle_t h, const void* const box[3], const void* const cls[3],
                    const void* const kpt[3], const int32_t level_hw[6], int B, int layout, int dtype,
                    float conf, float* cand, int cap_per_entry, int32_t* count, void* stream);

/* ---- Kernel 3 (a7,a12) batched segment NMS / GREEDYNMM / NMM.
 *      Replaces torchvision.ops.nms inside ultralytics non_max_suppression (stage 1: per slice, IOU,
 *      thr 0.7, strict >, fp32, max_keep 300, pre_cap 30000) and sahi.postprocess.combine.{NMS,GreedyNMM,NMM}
 *      Postprocess with has_match + merge_object_prediction_pair [EXT sahi 0.11.34] (stage 2: per image),
 *      selected at docs sahi/predict.py:44-49,250-259 and run at :297,:319.
 *      boxes [.,4] f32 at row stride box_stride floats, scores at stride score_stride, cats int32 at stride
 *      cat_stride (NULL: one class), tie int32 tie-break keys at stride tie_stride (NULL: index in segment);
 *      segment s covers rows seg_offsets[s] .. +min(seg_counts[s], max_segment)-1 (seg_counts NULL: all full).
 *      cmp_strict: 0 -> match if metric >= thr (sahi), 1 -> metric > thr (torchvision).
 *      precision:  0 -> fp64 metric (sahi 0.11.34), 1 -> fp32 metric (torchvision).
 *      Rank (visiting) order inside a segment: score descending, ties by ascending tie key.
 *      tie_rule: 0 -> plain greedy loop (torchvision; sahi's older pure-torch generations);
 *                1 -> sahi 0.11.34's rule for EQUAL scores (SURVEY A.2.4 variant N): the current box does not test an
 *                     equal-score candidate whose (x1,y1,x2,y2) tuple is lexicographically larger — both may be kept — and a
 *                     keep visited later does test, and for GREEDYNMM claims into its merge list, the earlier equal-score
 *                     keeps it matches (it then folds their MERGED boxes).  Ignored for NMM (variant N's nmm has no such rule).
 *      Outputs (dev): keep [.] global row ids in rank order, written from seg_offsets[s]; keep_count [S];
 *      parent [.] (may be NULL) = global row id of the keep that claimed a row (its own id for keeps — or, with tie_rule 1,
 *      the later equal-score keep that claimed this keep; -1 for rows cut by pre_cap/max_keep); merged_boxes [.,4],
 *      merged_scores [.], merged_cats [.] (may be NULL) are indexed like keep: the union box after the has_match replay
 *      (NMS: the kept row itself).
 *      The path is chosen per segment on the device from its actual count (no host round trip): segments up to 4096 boxes
 *      run in one CTA's shared memory, larger ones on a cluster of 8 CTAs over `workspace`
 *      (fsd_merge_workspace_bytes() bytes; only touched when max_segment > 4096).
 *      Limit: max_segment <= 32768 (FSD_ERR_CAPACITY beyond). */
int64_t fsd_merge_workspace_bytes(int64_t N, int S, int max_segment);
int fsd_merge(fsd_handle_t h, const float* boxes, int box_stride, const float* scores, int score_stride,
              const int32_t* cats, int cat_stride, const int32_t* tie, int tie_stride,
              const int32_t* seg_offsets, const int32_t* seg_counts, int S, int max_segment, int type, int metric,
              double thr, int cmp_strict, int precision, int class_agnostic, int pre_cap, int max_keep, int tie_rule,
              int32_t* keep, int32_t* keep_count, int32_t* parent, float* merged_boxes, float* merged_scores,
              int32_t* merged_cats, void* workspace, int64_t workspace_bytes, void* stream);

/* ---- Kernel 2b (a8-a11) finalize kept detections: un-letterbox (scale_boxes / scale_coords / clip),
 *      int() truncation, + slice shift.  Replaces ultralytics scale_boxes/scale_coords/clip_boxes [EXT],
 *      utils/yolo_wrapper.py:137-159 (astype(int), +shift, keypoints +shift), the sahi ObjectAnnotation clamp
 *      [EXT] and ObjectPrediction.get_shifted_object_prediction (docs sahi/prediction.py:94-120).
 *      cand: Kernel 2a rows; keep/keep_count: Kernel 3 stage-1 output, one segment per batch entry (segment b
 *      starts at row b*cap_per_entry); entry_geom [B,8] int32 = (shift_x, shift_y, src_w, src_h, box_pad_x,
 *      box_pad_y, full_w, full_h); entry_fgeom [B,4] f32 = (gain, kpt_pad_x, kpt_pad_y, 0).
 *      Rows are packed per group (image): group_range [G,2] = first / one-past-last entry of the group (entries
 *      of a group are contiguous), group_offsets [G] = first det row of the group; out_count [G] is IN/OUT: rows
 *      already present (so a second call can append the full-image pass) -> rows after this call.
 *      det rows are 24 floats: [x1,y1,x2,y2 (full-image ints as float), score, entry (int bits), 15 kpts,
 *      source cand row (int bits), 2 pad]; rows of a group are in (entry, stage-1 rank) order — the order the
 *      reference appends ObjectPredictions in. */
int fsd_finalize_dets(fsd_handle_t h, const float* cand, int cap_per_entry, const int32_t* keep,
                      const int32_t* keep_count, int B, const int32_t* entry_geom, const float* entry_fgeom,
                      const int32_t* group_range, const int32_t* group_offsets, int G, int truncate, float* det,
                      int det_cap_per_group, int32_t* out_count, void* stream);

/* ---- (a1) result packing — the tail of get_sliced_prediction (docs sahi/predict.py:317-345): the merged boxes of
 *      all G images (Kernel 3 stage-2 outputs, indexed from group_offsets[g]) plus the key-points picked by
 *      fsd_attach_keypoints are packed into ONE contiguous list of 24-float rows so a single D2H copy returns a
 *      whole batch: [x1,y1,x2,y2, score, source det row (int bits, -1: no key-points), 15 kpts, stage-2 keep id
 *      (int bits), image index (int bits), pad].  out_offsets [G+1] receives the first row of each image. */
int fsd_pack_results(fsd_handle_t h, const float* det, const int32_t* group_offsets, const int32_t* keep,
                     const int32_t* keep_count, const float* merged_boxes, const float* merged_scores,
                     const int32_t* src_index, int G, float* out, int32_t* out_offsets, void* stream);

/* ---- (a5) backbone conv epilogue: x = act(x + bias[c]) in place over a dense channels-last tensor [n_pixels, channels]
 *      (act: 0 none, 1 SiLU, 2 LeakyReLU(slope)).  The conv backbones stay PyTorch/cuDNN (BASELINE.json north_star);
 *      this replaces the two eager elementwise passes torch runs after every convolution of ultralytics' Conv
 *      (conv -> +bias -> SiLU) and basicsr's RRDB blocks (conv -> +bias -> LeakyReLU 0.2) with one read+write. */
int fsd_bias_act_inplace(fsd_handle_t h, void* x, const void* bias, int64_t n_pixels, int channels, int act,
                         float slope, int dtype, void* stream);

/* ---- (a5) general conv epilogue (fp16, channels-last): out[pix,c] = act(x[pix,c] + bias[c]) (+ residual[pix,c]).
 *      `out` / `residual` / `out2` may be channel slots of wider channels-last buffers: each has its own pixel stride
 *      (elements).  Channels >= out2_first_channel are also stored to out2 (NULL = none).  With it the epilogue of
 *      ultralytics' Conv writes straight into the concat buffer of C3k2 / SPPF / C2PSA (`torch.cat((...), 1)` in
 *      ultralytics/nn/modules/block.py, reached from utils/yolo_wrapper.py:72) and folds Bottleneck's `x + cv2(cv1(x))`.
 *      up2x (NULL = none): a [N, 2*height, 2*width] pixel grid (own pixel stride) that receives every result pixel four times
 *      = `nn.Upsample(2, "nearest")` of the output, stored straight into the FPN's next concat buffer. */
int fsd_bias_act(fsd_handle_t h, const void* x, const void* bias, void* out, int64_t out_pixel_stride,
                 const void* residual, int64_t residual_pixel_stride, void* out2, int64_t out2_pixel_stride,
                 int out2_first_channel, void* up2x, int64_t up2x_pixel_stride, int height, int width,
                 int64_t n_pixels, int channels, int act, float slope, int dtype, void* stream);

/* ---- (a5) SPPF pooling: buf is the [N,H,W,4c] fp16 channels-last concat buffer whose channel slot 0 holds y; fills
 *      slots 1..3 with m(y), m(m(y)), m(m(m(y))), m = MaxPool2d(kernel 5, stride 1, padding 2) (ultralytics SPPF). */
int fsd_sppf_pool(fsd_handle_t h, void* buf, int N, int H, int W, int c, int dtype, void* stream);

/* ---- (a5) YOLO stem: out = SiLU(conv3x3_stride2_pad1(x) + bias) for the 3-channel network input, fp16 channels-last:
 *      x [E,H,W,3] (Kernel 1's channels-last output), weight [16,3,3,3] (o,c,ky,kx dense), bias [16],
 *      out [E,(H-1)/2+1,(W-1)/2+1,16].  Replaces layer 0 of yolo11-pose.yaml (ultralytics Conv(3,16,3,2), run inside
 *      model.predict at utils/yolo_wrapper.py:72) as cuDNN convolution + epilogue pass; tensor cores via mma.sync.
 *      space_to_depth != 0: out is [E, OH/2 + 1, OW/2 + 1, 64] with out[e][Y+1][X+1][(dy*2+dx)*16 + c] = result(e, 2Y+dy, 2X+dx, c)
 *      (row 0 / column 0 are NOT written: the caller keeps them zero) — the layout in which layer 1 (Conv(16,32,3,2)) is a
 *      2x2 stride-1 convolution over 64 channels, for which cuDNN has a fast kernel (benchmarks/b1_s2d_probe.py). */
int fsd_stem_conv(fsd_handle_t h, const void* x, int E, int H, int W, const void* weight, const void* bias,
                  int out_channels, int dtype, int space_to_depth, void* out, void* stream);

/* ---- (a5) 1x1 convolution + bias + activation (+ residual) for the low-intensity layers, fp16 channels-last:
 *      out[pix, :N] = act(x[pix, :K] . weight[N, K]^T + bias) (+ residual[pix, :N]); x / out / residual / out2 are channel slots
 *      with their own pixel strides (elements), exactly as in fsd_bias_act.  K % 16 == 0, K <= 128, N in {16,32,64,128}
 *      (the weight matrix and two staging tiles per warp live in shared memory).  Replaces cuDNN convolution + epilogue pass for
 *      ultralytics Conv(c1, c2, 1, 1) layers (C3k2.cv1/cv2, C3k.cv1-3, head cv3; run from utils/yolo_wrapper.py:72). */
/* 0: shape not supported; 1: the mma.sync kernel only; 2: the tcgen05 kernel takes it (channels in multiples of 16, K <= 512, N <= 256,
 * K * N * 2 bytes <= 96 KB) */
int fsd_pointwise_conv_supported(int in_channels, int out_channels);
int fsd_pointwise_conv(fsd_handle_t h, const void* x, int64_t x_pixel_stride, const void* weight, const void* bias,
                       void* out, int64_t out_pixel_stride, const void* residual, int64_t residual_pixel_stride,
                       void* out2, int64_t out2_pixel_stride, int out2_first_channel, int64_t n_pixels,
                       int in_channels, int out_channels, int act, float slope, int dtype, void* stream);

/* ---- (a5) 3x3 convolution, stride 1, pad 1, + bias + activation (+ residual) -> channel slot, on the tensor cores
 *      (csrc/k10_pointwise_tc.cu: implicit GEMM, nine shifted 4-D TMA boxes per 16x8-pixel tile, tcgen05.mma into tensor memory).
 *      x: n_images channels-last images [H, W, in_channels] fp16, x_pixel_stride elements between pixels (a channel slot of a wider
 *      buffer is fine); weight_taps: TAP-MAJOR [3][3][n][in_channels] fp16 with n = out_channels (16 when out_channels == 8: rows
 *      8..15 zero); out / residual as in fsd_pointwise_conv.  fsd_conv3x3_supported: in_channels % 16 == 0 (<= 256), out_channels 8
 *      or a multiple of 16 (<= 256), 9 * in * n * 2 bytes <= 96 KB (the weights stay resident in shared memory).  Replaces cuDNN
 *      convolution + fsd_bias_act for ultralytics Conv(c1, c2, 3, 1) (Bottleneck.cv1/cv2, C3k, head cv2/cv3/cv4 branches; run from
 *      utils/yolo_wrapper.py:72).  fp32 accumulation; results agree with the library path to fp16 rounding. */
int fsd_conv3x3_supported(int in_channels, int out_channels);
int fsd_conv3x3(fsd_handle_t h, const void* x, int64_t x_pixel_stride, int n_images, int H, int W, const void* weight_taps,
                const void* bias, void* out, int64_t out_pixel_stride, const void* residual, int64_t residual_pixel_stride,
                int in_channels, int out_channels, int act, float slope, int dtype, void* stream);

/* ---- (a5) 2x2 convolution, stride 1, no padding, + bias + activation on the same tensor-core pipeline (halo mode): layer 1 of
 *      yolo11-pose (Conv(16, 32, 3, 2)) evaluated on the space-to-depth output of fsd_stem_conv, where it is algebraically a 2x2
 *      stride-1 convolution over 64 channels.  x: [n, H, W, in_channels]; out: [n, H-1, W-1, out_channels]; weight_taps TAP-MAJOR
 *      [2][2][out_channels][in_channels].  in_channels 16 / 32 / 64, out_channels a multiple of 16. */
int fsd_conv2x2_supported(int in_channels, int out_channels);
int fsd_conv2x2(fsd_handle_t h, const void* x, int64_t x_pixel_stride, int n_images, int H, int W, const void* weight_taps,
                const void* bias, void* out, int64_t out_pixel_stride, int in_channels, int out_channels, int act, float slope,
                int dtype, void* stream);

/* ---- (a5) depth-wise 3x3 convolution, stride 1, pad 1, + bias + activation -> channel slot (csrc/k11_dwconv3x3.cu).
 *      x: n_images channels-last images [H, W, channels] fp16; weight_taps: TAP-MAJOR [9][channels] fp16 (tap = 3 ky + kx).
 *      Replaces cuDNN's grouped convolution + fsd_bias_act for ultralytics DWConv(c, c, 3) (head cv3 branches) and C2PSA's positional
 *      convolution (run from utils/yolo_wrapper.py:72).  fp32 accumulation in tap order. */
int fsd_dwconv3x3(fsd_handle_t h, const void* x, int64_t x_pixel_stride, int n_images, int H, int W, const void* weight_taps,
                  const void* bias, void* out, int64_t out_pixel_stride, int channels, int act, float slope, int dtype, void* stream);

/* ---- (a5) YOLO neck: out = concat(nearest_upsample_2x(a), b) along channels, channels-last, in ONE pass.
 *      a [N,ah,aw,ca], b [N,2ah,2aw,cb], out [N,2ah,2aw,ca+cb]; replaces torch's upsample kernel + concat kernel. */
int fsd_upsample2x_concat(fsd_handle_t h, const void* a, const void* b, void* out, int N, int ah, int aw, int ca,
                          int cb, int dtype, void* stream);

/* ---- Kernel 4 (a15) Real-ESRGAN tile crop / stitch.  Replaces RealESRGANer.enhance/pre_process/
 *      tile_process/post_process [EXT realesrgan 0.3.0], reached from utils/enhancer.py:214.
 *      fsd_esrgan_tile_table: host-side tile table; each row = 12 int32:
 *        [px0,py0,pw,ph (padded input rect), in_x0,in_y0,in_w,in_h (interior), tile_elem_offset lo,hi,
 *         out_elem_offset lo,hi]  (offsets in elements into the packed buffers: input tiles 3*ph*pw each, output
 *        tiles 3*(ph*scale)*(pw*scale) each, both rounded up to 8 elements).  padded_hw = H,W after pre-/mod-pad. */
int fsd_esrgan_tile_table(int H, int W, int scale, int tile, int tile_pad, int pre_pad, int32_t* table,
                          int cap, int* n_tiles, int32_t padded_hw[2]);
/* crop: u8 HWC BGR image -> packed [3,ph,pw] RGB tiles of dtype, value/255; pre_h/pre_w = H,W + pre_pad (the
 *       right/bottom reflect pre-pad and mod-pad are folded into the index map); table given on device and host.
 *       n_images same-sized images (image_pitch bytes apart) are cropped by ONE launch with the same table; image i's
 *       tiles start tiles_image_stride ELEMENTS after image i-1's (a 1080p frame is only ~20 MB: batching frames is
 *       what lets the launch reach HBM speed).  n_images = 1 reproduces the reference's per-image call. */
int fsd_esrgan_crop(fsd_handle_t h, const uint8_t* bgr, int H, int W, int64_t row_pitch, int pre_h, int pre_w,
                    const int32_t* table_dev, const int32_t* table_host, int T, int dtype, void* tiles,
                    int n_images, int64_t image_pitch, int64_t tiles_image_stride, void* stream);
/* stitch: packed [3,ph*s,pw*s] RGB network outputs -> u8 HWC BGR [out_h,out_w] (= H*s, W*s);
 *         clamp(0,1)*255, round-half-even; halos, mod-pad and pre-pad are dropped.  n_images outputs
 *         (out_image_pitch bytes apart) are stitched by one launch from tile sets tiles_image_stride elements apart. */
int fsd_esrgan_stitch(fsd_handle_t h, const void* tiles_out, const int32_t* table_dev,
                      const int32_t* table_host, int T, int scale, int dtype, uint8_t* out_bgr, int out_h,
                      int out_w, int64_t out_pitch, int n_images, int64_t tiles_image_stride, int64_t out_image_pitch,
                      void* stream);

/* ---- (f1) WIDER-FACE evaluation IoU — replaces the Cython bbox_overlaps of WiderFace-Evaluation,
 *      imported at eval/eval_official_widerface.py:20-33 and called at :330.  boxes [N,4], query [K,4]
 *      xyxy f64 (dev); overlaps [N,K] f64 (dev), "+1" pixel convention. */
int fsd_bbox_overlaps_p1(fsd_handle_t h, const double* boxes, int N, const double* query, int K,
                         double* overlaps, void* stream);

/* ---- (f1) WIDER-FACE official-protocol PR curve in ONE launch — replaces _image_eval, _img_pr_info and the
 *      `pr_curve += ...` accumulation of _evaluate_setting (eval/eval_official_widerface.py:302-377, 398-445).
 *      pred [n_pred,5] f64 rows (x,y,w,h,score) of all G images, image g = rows pred_off[g] .. pred_off[g+1]-1 in the
 *      order the evaluator stores them; gt [n_gt,4] f64 (x,y,w,h) with gt_off likewise; evaluate [n_gt] int32 =
 *      1 for ground-truth boxes that count in this setting (the `ignore` array of :432-434), 0 for ignored ones;
 *      thresh [T] f64 = 1 - (t+1)/T as the host computes them.  Outputs: pred_recall / proposal [n_pred] f64 (what
 *      _image_eval returns per image: cumulative matched count, and 1 / -1 flags) and pr_curve [T,2] f64 =
 *      sum over images of _img_pr_info (zeroed by this call).  scratch: fsd_widerface_scratch_bytes() bytes. */
int64_t fsd_widerface_scratch_bytes(int64_t n_pred, int64_t n_gt);
int fsd_widerface_pr_curve(fsd_handle_t h, const double* pred, const int32_t* pred_off, const double* gt,
                           const int32_t* gt_off, const int32_t* evaluate, int G, int64_t n_pred, int64_t n_gt,
                           double iou_thresh, const double* thresh, int T, double* pred_recall, double* proposal,
                           void* scratch, int64_t scratch_bytes, double* pr_curve, void* stream);

/* ---- (f4) JPEG ingest — replaces the host-side PIL decode of read_image_as_pil ([EXT sahi.utils.cv], reached from
 *      docs sahi/predict.py:229 and docs sahi/prediction.py:173) and the temporary-JPEG hand-offs between pipeline stages
 *      (pipeline_v4_yolo/1_Inference.py:328-330): a baseline JPEG (host bytes) is decoded by nvJPEG on `stream` straight into
 *      dst [H, row_pitch] (interleaved RGB; bgr != 0: BGR as cv2.imread gives).  libnvjpeg is loaded on first use; without it
 *      these two entry points fail with FSD_ERR_ARG.  NOT bit-exact with PIL's libjpeg (a few LSB): opt-in, outside the
 *      bit-exact parity claims. */
int fsd_jpeg_info(const uint8_t* jpeg, int64_t len, int* width, int* height, int* channels);
int fsd_jpeg_decode(fsd_handle_t h, const uint8_t* jpeg, int64_t len, int bgr, uint8_t* dst, int64_t row_pitch, int H, int W,
                    void* stream);

/* ---- (f2) keypoint attach — replaces YOLOv11PoseDetectionModel.attach_keypoints_to_predictions
 *      (utils/yolo_wrapper.py:168-217), batched over S images: for each merged box pick the LAST stage-1
 *      detection of the same image with the identical box, else the detection whose box has the largest IoU
 *      (first maximum, then the last detection sharing that box) if that IoU > 0.5.
 *      merged rows of image s: m_off[s] .. +m_cnt[s]; detections: d_off[s] .. +d_cnt[s] (strides in floats);
 *      src_index [.] (indexed like merged) receives the global detection row or -1. */
int fsd_attach_keypoints(fsd_handle_t h, const float* merged, int merged_stride, const int32_t* m_off,
                         const int32_t* m_cnt, const float* dets, int det_stride, const int32_t* d_off,
                         const int32_t* d_cnt, int S, int32_t* src_index, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FSD_B200_H */
