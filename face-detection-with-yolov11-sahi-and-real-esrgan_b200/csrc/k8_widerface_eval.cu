// (f1) WIDER-FACE official-protocol evaluation on the device.
//
// Replaces the evaluator maths of eval/eval_official_widerface.py:302-377,398-445:
//   bbox_overlaps  — the Cython module of the external WiderFace-Evaluation repo imported at :20-33 ("+1" pixel convention);
//   _image_eval    — per prediction (in stored order) the best ground-truth box (first maximum, numpy argmax), then the
//                    order-dependent greedy matching: ignored ground truth marks the prediction -1, a free one becomes a hit;
//   _img_pr_info   — for each of thresh_num thresholds the LAST prediction with score >= thresh (np.where(...)[0][-1]), the
//                    number of valid proposals up to it and the recall count at it;
//   the `pr_curve += _img_pr_info` accumulation over the data set (:441-442).
// One launch per (data set, setting): one CTA per image.  Every accumulated quantity is an integer count held in a double, so
// the atomic accumulation order cannot change a bit of the result; the IoU uses explicitly rounded fp64 operations in the
// Cython source's order (the library is built -fmad=false), so matches at exactly IoU == 0.5 fall as they do on the CPU.
#include "fsd_common.cuh"

namespace fsd {

// xyxy boxes, "+1" convention, operation order of bbox.pyx
__device__ __forceinline__ double overlap_p1(double b0, double b1, double b2, double b3, double q0, double q1, double q2,
                                             double q3) {
    const double qa = __dmul_rn(__dadd_rn(__dsub_rn(q2, q0), 1.0), __dadd_rn(__dsub_rn(q3, q1), 1.0));
    const double iw = __dadd_rn(__dsub_rn(fmin(b2, q2), fmax(b0, q0)), 1.0);
    if (iw > 0) {
        const double ih = __dadd_rn(__dsub_rn(fmin(b3, q3), fmax(b1, q1)), 1.0);
        if (ih > 0) {
            const double ba = __dmul_rn(__dadd_rn(__dsub_rn(b2, b0), 1.0), __dadd_rn(__dsub_rn(b3, b1), 1.0));
            const double inter = __dmul_rn(iw, ih);
            return __ddiv_rn(inter, __dsub_rn(__dadd_rn(ba, qa), inter));
        }
    }
    return 0.0;
}

__global__ void bbox_overlaps_p1_kernel(const double* __restrict__ boxes, int N, const double* __restrict__ query,
                                        int K, double* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)N * K) return;
    const double* b = boxes + 4 * (size_t)(i / K);
    const double* q = query + 4 * (size_t)(i % K);
    out[i] = overlap_p1(b[0], b[1], b[2], b[3], q[0], q[1], q[2], q[3]);
}

constexpr int K8_THREADS = 256;

struct K8Params {
    const double* pred; const int32_t* pred_off;   // [.,5] x,y,w,h,score ; [G+1]
    const double* gt; const int32_t* gt_off;       // [.,4] x,y,w,h       ; [G+1]
    const int32_t* evaluate;                        // [.] per ground-truth box: 1 = counts in this setting, 0 = ignored
    const double* thresh; int T;                    // score thresholds 1 - (t+1)/T, computed by the host in float64
    double iou_thresh;
    double* pred_recall; double* proposal;          // [.] outputs of _image_eval, indexed like pred
    int32_t* best_idx; double* best_val; double* cum_valid;  // [.] scratch, indexed like pred
    int32_t* recall_flag;                           // [.] scratch, indexed like gt
    double* pr_curve;                               // [T,2] accumulated (caller zeroes it)
};

__global__ void __launch_bounds__(K8_THREADS) k8_widerface_pr_kernel(const K8Params p) {
    __shared__ int s_idx[K8_THREADS];
    __shared__ double s_val[K8_THREADS];
    __shared__ int s_matched, s_valid;
    const int g = blockIdx.x, tid = threadIdx.x;
    const int p0 = p.pred_off[g], n = p.pred_off[g + 1] - p0;
    const int g0 = p.gt_off[g], k = p.gt_off[g + 1] - g0;
    if (n <= 0 || k <= 0) return;  // eval_official_widerface.py:428-429: such images add nothing to the curve
    const double* pred = p.pred + 5 * (size_t)p0;
    const double* gt = p.gt + 4 * (size_t)g0;
    // ---- best ground truth per prediction (xywh -> xyxy in fp64, as `_pred[:, 2] + _pred[:, 0]`) ---------------
    for (int h = tid; h < n; h += K8_THREADS) {
        const double b0 = pred[5 * h], b1 = pred[5 * h + 1];
        const double b2 = __dadd_rn(pred[5 * h + 2], b0), b3 = __dadd_rn(pred[5 * h + 3], b1);
        double best = -1.0;
        int bi = 0;
        for (int j = 0; j < k; ++j) {
            const double q0 = gt[4 * j], q1 = gt[4 * j + 1];
            const double v = overlap_p1(b0, b1, b2, b3, q0, q1, __dadd_rn(gt[4 * j + 2], q0), __dadd_rn(gt[4 * j + 3], q1));
            if (v > best) { best = v; bi = j; }  // strict: the FIRST maximum, like numpy argmax
        }
        p.best_idx[p0 + h] = bi;
        p.best_val[p0 + h] = best;
    }
    for (int j = tid; j < k; j += K8_THREADS) p.recall_flag[g0 + j] = 0;
    if (tid == 0) { s_matched = 0; s_valid = 0; }
    __syncthreads();
    // ---- greedy matching in prediction order: chunks staged in shared memory, one thread walks them ----------------
    for (int c0 = 0; c0 < n; c0 += K8_THREADS) {
        const int cn = min(K8_THREADS, n - c0);
        if (tid < cn) { s_idx[tid] = p.best_idx[p0 + c0 + tid]; s_val[tid] = p.best_val[p0 + c0 + tid]; }
        __syncthreads();
        if (tid == 0) {
            int matched = s_matched, valid = s_valid;
            for (int i = 0; i < cn; ++i) {
                double prop = 1.0;
                if (s_val[i] >= p.iou_thresh) {
                    const int j = g0 + s_idx[i];
                    if (p.evaluate[j] == 0) { p.recall_flag[j] = -1; prop = -1.0; }
                    else if (p.recall_flag[j] == 0) { p.recall_flag[j] = 1; ++matched; }
                }
                if (prop == 1.0) ++valid;
                p.proposal[p0 + c0 + i] = prop;
                p.pred_recall[p0 + c0 + i] = (double)matched;
                p.cum_valid[p0 + c0 + i] = (double)valid;
            }
            s_matched = matched; s_valid = valid;
        }
        __syncthreads();
    }
    // ---- PR info per threshold, accumulated over images ---------------------------------------------------------
    for (int t = tid; t < p.T; t += K8_THREADS) {
        const double th = p.thresh[t];
        int last = n - 1;
        while (last >= 0 && !(pred[5 * last + 4] >= th)) --last;
        if (last < 0) continue;
        atomicAdd(p.pr_curve + 2 * t, p.cum_valid[p0 + last]);
        atomicAdd(p.pr_curve + 2 * t + 1, p.pred_recall[p0 + last]);
    }
}

}  // namespace fsd

using namespace fsd;

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

extern "C" int64_t fsd_widerface_scratch_bytes(int64_t n_pred, int64_t n_gt) {
    return n_pred * (int64_t)(sizeof(int32_t) + 2 * sizeof(double)) + n_gt * (int64_t)sizeof(int32_t) + 64;
}

extern "C" int fsd_widerface_pr_curve(fsd_handle_t h, const double* pred, const int32_t* pred_off, const double* gt,
                                      const int32_t* gt_off, const int32_t* evaluate, int G, int64_t n_pred, int64_t n_gt,
                                      double iou_thresh, const double* thresh, int T, double* pred_recall,
                                      double* proposal, void* scratch, int64_t scratch_bytes, double* pr_curve,
                                      void* stream_) {
    FSD_CHECK_ARG(h && G >= 0 && T > 0 && n_pred >= 0 && n_gt >= 0, "fsd_widerface_pr_curve: bad sizes");
    FSD_CHECK_ARG(pred_off && gt_off && thresh && pr_curve, "fsd_widerface_pr_curve: null argument");
    cudaStream_t stream = (cudaStream_t)stream_;
    FSD_CUDA(cudaSetDevice(h->device));
    FSD_CUDA(cudaMemsetAsync(pr_curve, 0, sizeof(double) * 2 * (size_t)T, stream));
    if (G == 0 || n_pred == 0 || n_gt == 0) return FSD_OK;
    FSD_CHECK_ARG(pred && gt && evaluate && pred_recall && proposal && scratch, "fsd_widerface_pr_curve: null argument");
    FSD_CHECK_ARG(scratch_bytes >= fsd_widerface_scratch_bytes(n_pred, n_gt) && ((uintptr_t)scratch & 7) == 0,
                  "fsd_widerface_pr_curve: scratch must hold fsd_widerface_scratch_bytes() bytes, 8-byte aligned");
    K8Params p;
    p.pred = pred; p.pred_off = pred_off; p.gt = gt; p.gt_off = gt_off; p.evaluate = evaluate;
    p.thresh = thresh; p.T = T; p.iou_thresh = iou_thresh;
    p.pred_recall = pred_recall; p.proposal = proposal;
    uint8_t* s = reinterpret_cast<uint8_t*>(scratch);
    p.best_val = reinterpret_cast<double*>(s);
    p.cum_valid = p.best_val + n_pred;
    p.best_idx = reinterpret_cast<int32_t*>(p.cum_valid + n_pred);
    p.recall_flag = p.best_idx + n_pred;
    p.pr_curve = pr_curve;
    k8_widerface_pr_kernel<<<G, K8_THREADS, 0, stream>>>(p);
    FSD_CUDA(cudaGetLastError());
    h->launches += 1;
    return FSD_OK;
}
