/*
 * oracle/nms_greedy.c — TEST INFRASTRUCTURE (see oracle/__init__.py).
 *
 * Plain-C restatement of the greedy IoU suppression that the reference reaches through
 * `torchvision.ops.nms(boxes, scores, iou_thres)` at ultralytics/utils/ops.py:296.
 * torchvision is a third-party dependency whose source is not under /root/reference
 * (requirements.txt:2 pins torchvision==0.17.2); this file restates its published CPU
 * algorithm (torchvision/csrc/ops/cpu/nms_kernel.cpp) from its behaviour:
 *
 *   - candidates are visited in STABLE descending-score order (equal scores: lower index first);
 *   - a visited, not-yet-suppressed candidate is kept, then suppresses every later candidate j
 *     with  inter / ((area_i + area_j) - inter) > iou_threshold ;
 *   - every arithmetic step is an individually rounded fp32 operation (no fused multiply-add),
 *     the comparison promotes the fp32 ratio to double against the double threshold;
 *   - NaN ratios (0/0 for degenerate boxes) compare false, i.e. never suppress;
 *   - result = kept indices in visiting order.
 *
 * Build:  gcc -O2 -ffp-contract=off -fno-fast-math -shared -fPIC nms_greedy.c -o _build/liboracle_nms.so
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* stable merge sort of indices by descending score */
static void msort_desc(const float *s, int64_t *idx, int64_t *tmp, int64_t n) {
    for (int64_t width = 1; width < n; width *= 2) {
        for (int64_t lo = 0; lo < n; lo += 2 * width) {
            int64_t mid = lo + width < n ? lo + width : n;
            int64_t hi = lo + 2 * width < n ? lo + 2 * width : n;
            int64_t a = lo, b = mid, k = lo;
            while (a < mid && b < hi) {
                /* take from the right run only when strictly greater: keeps ties stable */
                if (s[idx[b]] > s[idx[a]]) tmp[k++] = idx[b++];
                else tmp[k++] = idx[a++];
            }
            while (a < mid) tmp[k++] = idx[a++];
            while (b < hi) tmp[k++] = idx[b++];
        }
        memcpy(idx, tmp, (size_t)n * sizeof(int64_t));
    }
}

static inline float fmaxf_(float a, float b) { return a < b ? b : a; } /* std::max(a,b) */
static inline float fminf_(float a, float b) { return b < a ? b : a; } /* std::min(a,b) */

/*
 * boxes: n x 4 (x1,y1,x2,y2) row-major fp32; scores: n fp32; keep: capacity n (int64).
 * max_keep <= 0 means unlimited; otherwise the visit stops after max_keep keeps (equivalent to
 * truncating the full result, ops.py:297 `i = i[:max_det]`).
 * Returns the number of kept indices, or -1 on allocation failure.
 */
int64_t oracle_nms_greedy(const float *boxes, const float *scores, int64_t n, double iou_threshold,
                          int64_t max_keep, int64_t *keep) {
    if (n <= 0) return 0;
    int64_t *order = (int64_t *)malloc((size_t)n * sizeof(int64_t));
    int64_t *tmp = (int64_t *)malloc((size_t)n * sizeof(int64_t));
    float *area = (float *)malloc((size_t)n * sizeof(float));
    unsigned char *dead = (unsigned char *)calloc((size_t)n, 1);
    if (!order || !tmp || !area || !dead) { free(order); free(tmp); free(area); free(dead); return -1; }
    for (int64_t i = 0; i < n; ++i) {
        order[i] = i;
        float w = boxes[4 * i + 2] - boxes[4 * i + 0];
        float h = boxes[4 * i + 3] - boxes[4 * i + 1];
        area[i] = w * h;
    }
    msort_desc(scores, order, tmp, n);
    int64_t nk = 0;
    for (int64_t a = 0; a < n; ++a) {
        int64_t i = order[a];
        if (dead[i]) continue;
        keep[nk++] = i;
        if (max_keep > 0 && nk >= max_keep) break;
        const float ix1 = boxes[4 * i], iy1 = boxes[4 * i + 1], ix2 = boxes[4 * i + 2], iy2 = boxes[4 * i + 3];
        const float iarea = area[i];
        for (int64_t b = a + 1; b < n; ++b) {
            int64_t j = order[b];
            if (dead[j]) continue;
            float xx1 = fmaxf_(ix1, boxes[4 * j]);
            float yy1 = fmaxf_(iy1, boxes[4 * j + 1]);
            float xx2 = fminf_(ix2, boxes[4 * j + 2]);
            float yy2 = fminf_(iy2, boxes[4 * j + 3]);
            float w = fmaxf_(0.0f, xx2 - xx1);
            float h = fmaxf_(0.0f, yy2 - yy1);
            float inter = w * h;
            float sum = iarea + area[j];
            float uni = sum - inter;
            float ovr = inter / uni;
            if ((double)ovr > iou_threshold) dead[j] = 1;
        }
    }
    free(order); free(tmp); free(area); free(dead);
    return nk;
}
