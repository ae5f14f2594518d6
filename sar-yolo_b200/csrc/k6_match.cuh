// k6_match.cuh — validator matching on the GPU.
//
// Replaces utils/metrics.py:55-75 (box_iou) + engine/validator.py:222-262 (match_predictions, use_scipy=False;
// JDE variant models/yolo/jde/val.py:683-736) for a whole batch in one launch.  The reference builds the
// (labels x detections) IoU matrix on the device, copies it to the host per image and runs numpy:
//     matches = nonzero(iou >= t); sort by iou descending;
//     keep the first match of every detection  (np.unique on column 1 -> its best label)
//     keep the first match of every label      (np.unique on column 0, list now ordered by detection index
//                                               -> the lowest-index detection among those that chose the label)
// The best label of a detection does not depend on t (only whether it clears t does), so one pass over the labels
// per detection plus, per threshold, an atomicMin per detection on its label reproduces the result exactly.
// One CTA per image.
#pragma once
#include "common.cuh"

namespace sarpost {

constexpr int kMatchThreads = 256;
constexpr int kMaxThr = 16;

struct MatchParams {
    const float *dets;
    const int32_t *det_counts;
    int32_t max_det, row_len;
    const float *gt_boxes, *gt_cls;
    const int32_t *gt_counts;
    int32_t max_gt;
    float iouv[kMaxThr];
    int32_t n_thr;
    uint8_t *correct;
    int32_t *matched_gt;
    int32_t tag_thr;
};

// utils/metrics.py:71-75 in torch's fp32 operation order
__device__ __forceinline__ float box_iou_ref(const float4 a, const float4 b) {
    const float w = fmaxf(__fsub_rn(fminf(a.z, b.z), fmaxf(a.x, b.x)), 0.0f);
    const float h = fmaxf(__fsub_rn(fminf(a.w, b.w), fmaxf(a.y, b.y)), 0.0f);
    const float inter = __fmul_rn(w, h);
    const float area_a = __fmul_rn(__fsub_rn(a.z, a.x), __fsub_rn(a.w, a.y));
    const float area_b = __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y));
    return __fdiv_rn(inter, __fadd_rn(__fsub_rn(__fadd_rn(area_a, area_b), inter), 1e-7f));
}

// dynamic smem: best_iou[max_det] f32 | best_l[max_det] i32 | dmin[max_gt] i32
__global__ void __launch_bounds__(kMatchThreads) k6_match(const __grid_constant__ MatchParams p) {
    extern __shared__ __align__(16) unsigned char match_smem[];
    float *best_iou = reinterpret_cast<float *>(match_smem);
    int32_t *best_l = reinterpret_cast<int32_t *>(best_iou + p.max_det);
    int32_t *dmin = best_l + p.max_det;
    const int b = blockIdx.x, tid = threadIdx.x;
    const int nd = min(max(p.det_counts[b], 0), p.max_det), ng = min(max(p.gt_counts[b], 0), p.max_gt);
    const float *dets = p.dets + static_cast<int64_t>(b) * p.max_det * p.row_len;
    const float4 *gts = reinterpret_cast<const float4 *>(p.gt_boxes) + static_cast<int64_t>(b) * p.max_gt;
    const float *gcls = p.gt_cls + static_cast<int64_t>(b) * p.max_gt;
    // best label of every detection: highest IoU among same-class labels (lower label index on ties)
    for (int d = tid; d < nd; d += kMatchThreads) {
        const float *row = dets + static_cast<int64_t>(d) * p.row_len;
        const float4 db = make_float4(row[0], row[1], row[2], row[3]);
        const float dc = row[5];
        float bi = 0.0f;
        int bl = -1;
        for (int l = 0; l < ng; ++l) {
            if (gcls[l] != dc) continue;  // iou * correct_class (validator.py:240-241)
            const float v = box_iou_ref(gts[l], db);
            if (v > bi) { bi = v; bl = l; }
        }
        best_iou[d] = bi;
        best_l[d] = bl;
    }
    __syncthreads();
    uint8_t *corr = p.correct + static_cast<int64_t>(b) * p.max_det * p.n_thr;
    for (int t = 0; t < p.n_thr; ++t) {
        const float thr = p.iouv[t];
        for (int l = tid; l < ng; l += kMatchThreads) dmin[l] = 0x7fffffff;
        __syncthreads();
        for (int d = tid; d < nd; d += kMatchThreads)
            if (best_l[d] >= 0 && best_iou[d] >= thr) atomicMin(&dmin[best_l[d]], d);
        __syncthreads();
        for (int d = tid; d < p.max_det; d += kMatchThreads) {
            const bool ok = d < nd && best_l[d] >= 0 && best_iou[d] >= thr && dmin[best_l[d]] == d;
            corr[static_cast<int64_t>(d) * p.n_thr + t] = ok ? 1 : 0;
            if (p.matched_gt && t == p.tag_thr) p.matched_gt[static_cast<int64_t>(b) * p.max_det + d] = ok ? best_l[d] : -1;
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// Same matching from a PRECOMPUTED IoU matrix — the boundary of BaseValidator.match_predictions(pred_classes,
// true_classes, iou) (engine/validator.py:222; JDE: jde/val.py:683), whose callers build `iou (L, D)` themselves
// (box_iou, mask IoU, OKS, probiou).  One image per launch, one CTA.
// ---------------------------------------------------------------------------------------------
struct MatchIouParams {
    const float *iou;         // (n_gt, n_det), row stride `ld`
    int64_t ld;
    const float *pred_cls;    // (n_det)
    const float *true_cls;    // (n_gt)
    int32_t n_det, n_gt;
    float iouv[kMaxThr];
    int32_t n_thr;
    uint8_t *correct;         // (n_det, n_thr)
    int32_t *matched_gt;      // (n_det) or nullptr
    int32_t tag_thr;
};

__global__ void __launch_bounds__(kMatchThreads) k6_match_iou(const __grid_constant__ MatchIouParams p) {
    extern __shared__ __align__(16) unsigned char match_smem[];
    float *best_iou = reinterpret_cast<float *>(match_smem);
    int32_t *best_l = reinterpret_cast<int32_t *>(best_iou + p.n_det);
    int32_t *dmin = best_l + p.n_det;
    const int tid = threadIdx.x;
    for (int d = tid; d < p.n_det; d += kMatchThreads) {  // consecutive threads read consecutive columns of a row: coalesced
        const float dc = p.pred_cls[d];
        float bi = 0.0f;
        int bl = -1;
        for (int l = 0; l < p.n_gt; ++l) {
            const float v = p.true_cls[l] == dc ? p.iou[static_cast<int64_t>(l) * p.ld + d] : 0.0f;  // iou * correct_class
            if (v > bi) { bi = v; bl = l; }
        }
        best_iou[d] = bi;
        best_l[d] = bl;
    }
    __syncthreads();
    for (int t = 0; t < p.n_thr; ++t) {
        const float thr = p.iouv[t];
        for (int l = tid; l < p.n_gt; l += kMatchThreads) dmin[l] = 0x7fffffff;
        __syncthreads();
        for (int d = tid; d < p.n_det; d += kMatchThreads)
            if (best_l[d] >= 0 && best_iou[d] >= thr) atomicMin(&dmin[best_l[d]], d);
        __syncthreads();
        for (int d = tid; d < p.n_det; d += kMatchThreads) {
            const bool ok = best_l[d] >= 0 && best_iou[d] >= thr && dmin[best_l[d]] == d;
            p.correct[static_cast<int64_t>(d) * p.n_thr + t] = ok ? 1 : 0;
            if (p.matched_gt && t == p.tag_thr) p.matched_gt[d] = ok ? best_l[d] : -1;
        }
        __syncthreads();
    }
}

}  // namespace sarpost
