/*
 * sarpost.h — C ABI of libsarpost.so: B200 (sm_100a) detection post-processing for SAR-YOLO.
 *
 * The reference (HaoqianSong/SAR-YOLO, an Ultralytics 8.3.63 fork) has no FFI on this path: the
 * boundary is two Python callables.  Each entry point below names the reference interface it
 * replaces (paths relative to /root/reference/ultralytics).  INTEGRATION.md shows the ctypes stub a
 * reference maintainer would add.
 *
 * Conventions
 *   - plain C types only; device buffers are raw `void*` / typed pointers, streams are `void*`
 *     holding a cudaStream_t (NULL = legacy default stream);
 *   - every function returns 0 on success, a negative SARPOST_E* code on failure; the message is
 *     available (per thread) from sarpost_last_error();
 *   - no function synchronises the device except the *_host entry points; all work is enqueued on
 *     the given stream; the library keeps no global mutable state (re-entrant, one workspace per call);
 *   - tensors are contiguous, row-major, fp32 unless a dtype field says otherwise.
 */
#ifndef SARPOST_H_
#define SARPOST_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SARPOST_VERSION 100 /* 0.1.0 */
#define SARPOST_MAX_LEVELS 8
#define SARPOST_MAX_CLASSES 2048 /* upper bound for nc (class-filter bitmap lives in kernel params) */

#define SARPOST_F32 0
#define SARPOST_F16 1

#define SARPOST_LAYOUT_CAT 0   /* one (B, no, H, W) tensor per level: the branches concatenated (head.py:204-206) */
#define SARPOST_LAYOUT_SPLIT 1 /* one tensor per branch and level: the convolution outputs as they are, no torch.cat */

#define SARPOST_OK 0
#define SARPOST_EINVAL (-1)    /* bad argument */
#define SARPOST_ECUDA (-2)     /* CUDA runtime error (message holds cudaGetErrorString) */
#define SARPOST_EWORKSPACE (-3)/* workspace too small */
#define SARPOST_EUNSUPPORTED (-4)

/*
 * Raw Detect/JDE head output: the list `x` handed to Detect._inference (nn/modules/head.py:100) /
 * JDE._inference (:214) — one (B, no, H_l, W_l) tensor per level, channel order
 * box[4*reg_max] | cls[nc] | extras_raw[n_extra_raw] | extras_sigmoid[n_extra_sigmoid]
 * (head.py:204-206, :232-235; JDE: 256 raw embedding + 6 state logits).
 */
typedef struct sarpost_head {
    int32_t nl;              /* number of levels, 1..SARPOST_MAX_LEVELS */
    int32_t batch;           /* B */
    int32_t no;              /* channels per anchor in memory, >= 4*reg_max + nc + n_extra_raw + n_extra_sigmoid
                                (trailing channels beyond the declared extras are ignored: declare 0 extras to
                                get 6-column rows from a JDE head) */
    int32_t nc;              /* number of classes */
    int32_t reg_max;         /* DFL bins per side; only 16 is supported (head.py:39) */
    int32_t n_extra_raw;     /* extras copied through unchanged (JDE embedding, head.py:247) */
    int32_t n_extra_sigmoid; /* extras passed through sigmoid (JDE state, head.py:247) */
    int32_t dtype;           /* element type of the level tensors: SARPOST_F32 (0) or SARPOST_F16 (1, `half=True`
                                pipelines: engine/validator.py:115-117).  fp16 logits are upcast exactly and all
                                arithmetic stays fp32; sarpost_decode then writes y in the same type as the input */
    int32_t h[SARPOST_MAX_LEVELS];
    int32_t w[SARPOST_MAX_LEVELS];
    float stride[SARPOST_MAX_LEVELS]; /* Detect.stride (head.py:41) */
    const void *data[SARPOST_MAX_LEVELS]; /* device (or, for *_host calls, host) pointers.  SPLIT layout: the box branch,
                                             (B, 4*reg_max, H_l, W_l) — the output of cv2[l] (head.py:204) */
    /* ---- SARPOST_LAYOUT_SPLIT only (device entry points; ignored for layout CAT) ----
     * JDE.forward concatenates cv2 | cv3 | cv4 [| state] per level (head.py:204-206) only so that _inference can split
     * them again (:232-235); at 1280x1280 P2, batch 16, that cat writes 2.85 GB of which the path reads 0.57 GB.  With the
     * split layout the caller hands over the branch outputs directly.  The embedding may be channels-last
     * ((B, H, W, E) in memory, what a channels_last cv4 branch writes): the <= max_det kept rows are then read as one
     * contiguous E*4-byte run each instead of E isolated 4-byte reads 128-byte DRAM lines apart. */
    int32_t layout;            /* SARPOST_LAYOUT_CAT (0, default) or SARPOST_LAYOUT_SPLIT */
    int32_t emb_channels_last; /* 1: emb[l] is (B, H_l, W_l, n_extra_raw) in memory; 0: (B, n_extra_raw, H_l, W_l) */
    const void *cls[SARPOST_MAX_LEVELS];   /* (B, nc, H_l, W_l): output of cv3[l] */
    const void *emb[SARPOST_MAX_LEVELS];   /* output of cv4[l] (n_extra_raw channels), NULL when n_extra_raw == 0 */
    const void *state[SARPOST_MAX_LEVELS]; /* (B, n_extra_sigmoid, H_l, W_l) state logits, NULL when n_extra_sigmoid == 0 */
} sarpost_head_t;

/* Keyword arguments of ops.non_max_suppression (utils/ops.py:167-182). */
typedef struct sarpost_nms_params {
    float conf_thres;      /* compared in fp32: score > (float)conf_thres (ops.py:234,271,275) */
    double iou_thres;      /* compared in double against the fp32 IoU (torchvision CPU nms) */
    int32_t agnostic;      /* ops.py:289 */
    int32_t multi_label;   /* ops.py:239,270-272 (only effective when nc > 1) */
    int32_t max_det;       /* ops.py:297 */
    int32_t max_nms;       /* ops.py:285-286 */
    float max_wh;          /* ops.py:289,295 class offset = cls * max_wh */
    const int32_t *classes;/* HOST pointer to the `classes` filter (ops.py:278-279) or NULL */
    int32_t n_classes;
    const float *labels;   /* sarpost_nms_decoded only: DEVICE pointer (B, max_labels, 5) rows cls,x,y,w,h — the apriori labels of
                              `save_hybrid` (ops.py:256-261): appended to every image's candidates with probability 1.0 for
                              their class and zero extras; NULL = none */
    const int32_t *label_counts; /* DEVICE (B): valid label rows per image */
    int32_t max_labels;
    const float *rescale;  /* DEVICE pointer (B, 5) = pad_x, pad_y, gain, w0, h0 per image, or NULL.  When set the gather
                              kernel applies ops.scale_boxes + clip_boxes (utils/ops.py:92-127, :319-338) to the output
                              boxes: x = clamp((x - pad_x) / gain, 0, w0), y likewise with pad_y, h0 — the per-image
                              loop of models/yolo/jde/predict.py:49 (ignored by sarpost_merge_tiles). */
    /* Fused gather + exchange over NVLink peer memory (multi-GPU, one process per GPU): when n_peers > 0 the gather
     * kernel stores every output row — and every image's count — into the buffers of ALL n_peers ranks (P2P-mapped
     * device pointers, e.g. torch symmetric memory), at image slot peer_slot_offset + b, instead of `out`/`counts`:
     * the all-gather of SURVEY §8e happens inside K5.  The caller runs a cross-rank barrier afterwards. */
    float *peer_out[8];      /* each (total_images, max_det, 6 + nm); 16-byte aligned (6-column rows leave as 16-byte stores) */
    int32_t *peer_counts[8]; /* each (total_images) */
    int32_t n_peers;
    int32_t peer_slot_offset;
    int32_t prediction_dtype;/* sarpost_nms_decoded only: element type of `prediction`, SARPOST_F32 or SARPOST_F16 (the `y` of a
                              `half=True` model); fp16 values are upcast exactly, arithmetic and output rows stay fp32 */
    int32_t workspace_clean;/* 0: the call zeroes the score-histogram head of the workspace itself (one memset node).
                              1: the caller guarantees the first sarpost_workspace_clean_bytes(batch) bytes are zero —
                              true after sarpost_workspace_prepare and after every successful call that used the
                              workspace with the SAME batch (each call leaves that region zeroed again). */
    int32_t out_tail_cols;  /* columns reserved at the END of every output row: rows are (6 + nm + out_tail_cols) floats
                              wide and the call leaves the tail untouched (sarpost_state_head fills it).  0 = none. */
    int64_t *stats;         /* instrumentation, DEVICE (B, 4) int64 or NULL: per image, what the NMS kernel did — sorted
                              candidates consumed before max_det keeps were found (or the candidates ran out), IoU pair
                              tests executed, NMS sub-chunks, selection passes.  bench.py's `clustered` leg reports them. */
    /* Results layout (models/yolo/jde/predict.py:52-66, engine/results.py:1011-1014): when
     * res_boxes != NULL the gather kernel writes, instead of the (6 + nm)-column rows of `out` (which may then be NULL),
     * what JDEPredictor.postprocess builds per image with split / argmax / cat:
     *   res_boxes   DEVICE (B, max_det, 7)  x1,y1,x2,y2, state_id, conf, cls — state_id = argmax over the
     *               n_extra_sigmoid state probabilities (first maximum, like torch.argmax) as a float; -1 when the
     *               head has no state channels or they are deferred (sarpost_state_ids fills the column then)
     *   res_embeds  DEVICE (B, max_det, n_extra_raw) the raw embedding of every kept row, contiguous (NULL allowed
     *               when n_extra_raw == 0)
     * Rows beyond counts[b] are left untouched.  Device entry points sarpost_fused / sarpost_nms_decoded only (for
     * a decoded prediction the first nm - res_state_cols extras columns are the embedding); not combinable with the
     * peer_out exchange or out_tail_cols. */
    float *res_boxes;
    float *res_embeds;
    int32_t res_state_cols; /* sarpost_nms_decoded only: how many of the trailing extras columns are state probabilities */
    int32_t nms_cluster;    /* CTAs (of 512 threads, one per SM) the NMS kernel spends per image: 0 = automatic (the largest of
                              4 / 2 / 1 with batch * CTAs <= #SMs: lowest latency for one call), or 1 / 2 / 4 / 8.  Callers that
                              keep several calls in flight on different streams pass 1, so that the NMS kernels of
                              consecutive calls fit next to each other.  Results do not depend on it. */
} sarpost_nms_params_t;

/* Last error message of the calling thread ("" if none). */
const char *sarpost_last_error(void);
int32_t sarpost_version(void);

/*
 * Bytes of device scratch sarpost_nms_decoded / sarpost_fused need (an upper bound valid for any
 * split of `anchors` = total anchors A per image into <= SARPOST_MAX_LEVELS levels), and the same
 * for sarpost_merge_tiles.  Negative = error code.
 */
int64_t sarpost_workspace_bytes(int32_t batch, int64_t anchors, int32_t nc, int32_t multi_label,
                                int32_t max_det);
int64_t sarpost_merge_workspace_bytes(int32_t n_frames, int32_t tiles_per_frame, int32_t dets_per_tile,
                                      int32_t max_det);
/* Persistent workspaces (see sarpost_nms_params_t.workspace_clean): size of the region that must be zero on
 * entry for `batch` images (frames for the merge), and a helper that zeroes it on `stream`. */
int64_t sarpost_workspace_clean_bytes(int32_t batch);
int32_t sarpost_workspace_prepare(void *workspace, int64_t workspace_bytes, int32_t batch, void *stream);

/*
 * Replaces Detect._inference / JDE._inference (nn/modules/head.py:100-131, :214-249) including
 * DFL (nn/modules/block.py:77-80), make_anchors and dist2bbox (utils/tal.py:366-390).
 * y: device (B, 4 + nc + n_extra_raw + n_extra_sigmoid, A), A = sum_l H_l*W_l, same element type as head->dtype.
 */
int32_t sarpost_decode(const sarpost_head_t *head, void *y, void *stream);

/*
 * Replaces ops.non_max_suppression (utils/ops.py:167-316, non-rotated branch) on an already
 * decoded prediction (B, C, A), C = 4 + nc + nm, rows cx,cy,w,h | nc probabilities | nm extras.
 * With apriori labels (params->labels) size the workspace for `anchors + max_labels` anchors; kept_index of a
 * label row is (anchors + label_row)*nc + class.
 *   out        device (B, max_det, 6 + nm): x1,y1,x2,y2,conf,cls,extras — rows [0, counts[b]) valid,
 *              in descending confidence; rows beyond are left untouched
 *   counts     device (B) int32
 *   kept_index device (B, max_det) int32 or NULL: anchor*nc + class of every output row
 * The input is not modified (the reference rewrites prediction[..., :4] in place, ops.py:243-244).
 */
int32_t sarpost_nms_decoded(const void *prediction, int32_t batch, int32_t channels, int64_t anchors,
                            int32_t nc, const sarpost_nms_params_t *params, float *out, int32_t *counts,
                            int32_t *kept_index, void *workspace, int64_t workspace_bytes, void *stream);

/*
 * Decode + non_max_suppression in one pass from the raw level logits (head.py:214-249 followed by
 * ops.py:167-316) without materialising y.  Output as sarpost_nms_decoded with
 * nm = n_extra_raw + n_extra_sigmoid (extras are gathered only for the kept rows).
 */
int32_t sarpost_fused(const sarpost_head_t *head, const sarpost_nms_params_t *params, float *out,
                      int32_t *counts, int32_t *kept_index, void *workspace, int64_t workspace_bytes,
                      void *stream);


/*
 * Plans.  A serving loop calls sarpost_fused with the same geometry and thresholds thousands of times; only the
 * addresses change.  A plan freezes everything else — validated geometry, candidate filter (the `classes` list is
 * folded in at creation), workspace binding, launch configuration, the TMA tensor maps (re-encoded only for a level
 * whose address changed) — so a run is three kernel launches and nothing more on the host: the cached form SURVEY §7
 * (hard part 7) asks for.  Results are those of sarpost_fused on the same inputs, bit for bit.
 *   create   `head` gives geometry, dtype and layout (its pointers are only checked for alignment class); `params` as for
 *            sarpost_fused with workspace_clean = 1: `workspace` (>= sarpost_workspace_bytes) must have been prepared with
 *            sarpost_workspace_prepare and belongs to the plan until it is destroyed.  Peer buffers (n_peers > 0) are fixed
 *            at creation; rescale / stats / res_* pointers of `params` are ignored — they travel in the io block.
 *   run      all work on `stream`.  Runs of one plan must be stream-ordered (one plan per stream); the plan is updated in
 *            place, so it must not be run from two threads at once.
 */
typedef struct sarpost_plan sarpost_plan_t;
typedef struct sarpost_plan_io {
    const void *data[SARPOST_MAX_LEVELS];  /* as sarpost_head_t.data / cls / emb / state */
    const void *cls[SARPOST_MAX_LEVELS];
    const void *emb[SARPOST_MAX_LEVELS];
    const void *state[SARPOST_MAX_LEVELS];
    float *out;            /* as sarpost_fused */
    int32_t *counts;
    int32_t *kept_index;   /* or NULL */
    const float *rescale;  /* as sarpost_nms_params_t.rescale, or NULL */
    int64_t *stats;        /* as sarpost_nms_params_t.stats, or NULL */
    float *res_boxes;      /* as sarpost_nms_params_t.res_boxes / res_embeds, or NULL */
    float *res_embeds;
} sarpost_plan_io_t;
int32_t sarpost_plan_create(const sarpost_head_t *head, const sarpost_nms_params_t *params, void *workspace,
                            int64_t workspace_bytes, sarpost_plan_t **plan);
int32_t sarpost_plan_run(sarpost_plan_t *plan, const sarpost_plan_io_t *io, void *stream);
void sarpost_plan_destroy(sarpost_plan_t *plan);

/*
 * Extras (raw embedding + sigmoid state, head.py:247) of an explicit list of n (image, anchor) pairs,
 * for callers that learn which rows need them only later (e.g. after sarpost_merge_tiles).
 *   image_index, anchor_index  device (n) int32;  out  device (n, n_extra_raw + n_extra_sigmoid)
 */
int32_t sarpost_gather_extras(const sarpost_head_t *head, const int32_t *image_index, const int32_t *anchor_index,
                              int32_t n, float *out, void *stream);

/*
 * Cross-tile merge for sliced (SAHI-style) inference: per frame, shift every tile's detections by
 * the tile origin and run the same class-offset NMS (ops.py:289-297 semantics).
 *   dets        device (n_frames*tiles_per_frame, dets_per_tile, row_len) rows x1,y1,x2,y2,conf,cls,...
 *   det_counts  device (n_frames*tiles_per_frame) int32 valid rows per tile
 *   origins     device (n_frames*tiles_per_frame, 2) fp32 (x0, y0) of each tile in frame pixels
 *   out         device (n_frames, max_det, row_len); counts device (n_frames)
 *   kept_index  device (n_frames, max_det) int32 or NULL: tile_in_frame*dets_per_tile + row
 * conf_thres / multi_label / classes of `params` are ignored (rows were filtered per tile).
 */
int32_t sarpost_merge_tiles(const float *dets, const int32_t *det_counts, const float *origins,
                            int32_t n_frames, int32_t tiles_per_frame, int32_t dets_per_tile,
                            int32_t row_len, const sarpost_nms_params_t *params, float *out,
                            int32_t *counts, int32_t *kept_index, void *workspace,
                            int64_t workspace_bytes, void *stream);

/*
 * Validator matching on the GPU: replaces, for a whole batch at once and without the per-image
 * device->host copy, utils/metrics.py:55-75 (box_iou, eps 1e-7) and BaseValidator.match_predictions
 * (engine/validator.py:222-262, use_scipy=False; JDE variant models/yolo/jde/val.py:683-736):
 *   iou[l][d] = box_iou(gt_l, det_d) zeroed where the classes differ; for every threshold t: the pairs with
 *   iou >= t are ordered by descending iou, each detection keeps its best label, each label then keeps the
 *   lowest-index detection among those that chose it; correct[d][t] = 1 for the surviving pairs.
 * (Equal IoUs: the reference's order is numpy's unstable argsort; here the lower label index wins.)
 *   dets        device (B, max_det, row_len) rows x1,y1,x2,y2,conf,cls,...;  det_counts device (B)
 *   gt_boxes    device (B, max_gt, 4) xyxy;  gt_cls device (B, max_gt) fp32;  gt_counts device (B)
 *   iouv        HOST (n_thr <= 16) thresholds, e.g. linspace(0.5, 0.95, 10)
 *   correct     device (B, max_det, n_thr) uint8
 *   matched_gt  device (B, max_det) int32 or NULL: label index matched at threshold index `tag_thr` or -1
 *               (the JDE validator reads true_tags[matched_gt], jde/val.py:731-735)
 */
int32_t sarpost_match_predictions(const float *dets, const int32_t *det_counts, int32_t batch, int32_t max_det,
                                  int32_t row_len, const float *gt_boxes, const float *gt_cls, const int32_t *gt_counts,
                                  int32_t max_gt, const float *iouv, int32_t n_thr, uint8_t *correct,
                                  int32_t *matched_gt, int32_t tag_thr, void *stream);

/*
 * The same matching at the boundary of BaseValidator.match_predictions(pred_classes, true_classes, iou)
 * (engine/validator.py:222-262; JDE variant with tags: models/yolo/jde/val.py:683-736) for one image: the caller has
 * already built the IoU matrix (box_iou, mask IoU, OKS ...); the reference copies it to the host and loops in numpy.
 *   iou        device (n_gt, n_det) fp32, rows `iou_row_stride` elements apart
 *   pred_cls   device (n_det) fp32;  true_cls device (n_gt) fp32
 *   iouv       HOST (n_thr <= 16);  correct device (n_det, n_thr) uint8
 *   matched_gt device (n_det) int32 or NULL: label index matched at threshold index tag_thr, else -1
 */
int32_t sarpost_match_from_iou(const float *iou, int32_t n_gt, int32_t n_det, int64_t iou_row_stride, const float *pred_cls,
                               const float *true_cls, const float *iouv, int32_t n_thr, uint8_t *correct, int32_t *matched_gt,
                               int32_t tag_thr, void *stream);

/*
 * Deferred JDE state head.  Replaces the per-anchor evaluation of JDE.state_predictor inside
 * JDE.forward (nn/modules/head.py:189-190 = Linear(E, E/2), ReLU, Dropout (identity in eval), Linear(E/2, S);
 * applied at :198-204, sigmoid at :247) by one evaluation per KEPT row after the gather: for every image b and row
 * r < counts[b]
 *     rows[b][r][state_col .. +n_state) = sigmoid(W2 · relu(W1 · rows[b][r][emb_col .. +embed_dim) + b1) + b2)
 * in place, fp32 FMA accumulation (results within 1e-5 of the reference's per-anchor matmul; the embedding columns
 * are bit-identical, so the only difference is summation order).  Run sarpost_fused on levels WITHOUT the state
 * channels (head.n_extra_sigmoid = 0) with params.out_tail_cols = n_state, then this call with emb_col = 6,
 * state_col = 6 + embed_dim: the output rows equal the reference's [xyxy, conf, cls, emb, state].
 *   rows   device (B, max_det, row_len);  counts device (B)
 *   w1     device (hidden, embed_dim) row-major (nn.Linear.weight), b1 (hidden);  w2 (n_state, hidden), b2 (n_state)
 *   embed_dim, hidden <= 1024, n_state <= 64
 */
int32_t sarpost_state_head(float *rows, const int32_t *counts, int32_t batch, int32_t max_det, int32_t row_len,
                           int32_t emb_col, int32_t embed_dim, int32_t state_col, int32_t n_state, int32_t hidden,
                           const float *w1, const float *b1, const float *w2, const float *b2, void *stream);

/*
 * Deferred state head for the results layout (params.res_boxes / res_embeds of a sarpost_fused call on levels WITHOUT
 * state channels): evaluates the same MLP on embeds (B, max_det, embed_dim) and writes only what
 * JDEPredictor.postprocess keeps of it (models/yolo/jde/predict.py:61-64): boxes7[b][r][4] = argmax_s sigmoid(...)
 * (first maximum) as a float, for r < counts[b].  boxes7 device (B, max_det, 7).
 */
int32_t sarpost_state_ids(const float *embeds, const int32_t *counts, float *boxes7, int32_t batch, int32_t max_det,
                          int32_t embed_dim, int32_t n_state, int32_t hidden, const float *w1, const float *b1,
                          const float *w2, const float *b2, void *stream);

/* sizeof(sarpost_head_t) / sizeof(sarpost_nms_params_t) as this library was compiled: lets a binding written in
 * another language (ctypes, cgo, JNI ...) verify its struct mirrors before the first call. */
int32_t sarpost_abi_sizes(int32_t *head_bytes, int32_t *params_bytes);

/*
 * End-to-end entry with HOST buffers (what a caller holding CPU tensors uses; timed as `e2e` by
 * bench.py).  A context owns pinned staging, device buffers and streams for one head geometry.
 * sarpost_fused_host copies only the channels the path reads (box + cls) host->device, runs the
 * fused pipeline, copies counts/boxes/indices back, gathers the extras of the kept rows from the
 * host tensors, and returns after the stream is idle.
 *   head->data[l]  HOST pointers (pinned or pageable)
 *   out            HOST (B, max_det, 6 + nm); counts HOST (B); kept_index HOST (B, max_det) or NULL
 */
typedef struct sarpost_host_ctx sarpost_host_ctx_t;
int32_t sarpost_host_ctx_create(int32_t device, sarpost_host_ctx_t **ctx);
void sarpost_host_ctx_destroy(sarpost_host_ctx_t *ctx);
int32_t sarpost_fused_host(sarpost_host_ctx_t *ctx, const sarpost_head_t *head,
                           const sarpost_nms_params_t *params, float *out, int32_t *counts,
                           int32_t *kept_index);
/* bytes moved by the last sarpost_fused_host call */
int32_t sarpost_host_ctx_last_traffic(const sarpost_host_ctx_t *ctx, int64_t *h2d_bytes, int64_t *d2h_bytes);

/*
 * Software pipeline over successive batches held in DEVICE memory (throughput mode; bench.py's `value`).  Same arguments
 * and results as sarpost_fused, but the work is enqueued on the pipeline's own streams: every decode kernel on one
 * low-priority stream, back to back; the NMS + gather kernels of batch i on a high-priority stream of their own, released
 * when the decode kernel of batch i ends — by then the decode kernel of batch i+1 holds every SM, so they stay pending and
 * take the first SMs that fall free when it drains, ahead of the decode kernel of batch i+2, which streams on what is left
 * (tiles are handed out dynamically).  NMS + gather of every batch are hidden under a later batch's decode; results lag one
 * decode kernel behind.  `depth` (1..8) = batches whose NMS + gather may be pending or running behind the decode stream
 * (depth + 1 workspaces rotate; depth 1 = plain stream order, no overlap; 3 is a good default).  No host synchronisation
 * anywhere.
 *   submit  waits (device side) for the work enqueued so far on `in_stream` — the stream that produced the level
 *           tensors — then enqueues the batch.  The level tensors and the outputs must stay valid until a later
 *           sarpost_pipeline_wait has been passed.
 *   wait    makes `stream` wait (device side) until every batch submitted so far is complete.
 */
typedef struct sarpost_pipeline sarpost_pipeline_t;
int32_t sarpost_pipeline_create(int32_t device, int32_t depth, sarpost_pipeline_t **pl);
void sarpost_pipeline_destroy(sarpost_pipeline_t *pl);
int32_t sarpost_pipeline_submit(sarpost_pipeline_t *pl, const sarpost_head_t *head, const sarpost_nms_params_t *params,
                                float *out, int32_t *counts, int32_t *kept_index, void *in_stream);
int32_t sarpost_pipeline_wait(sarpost_pipeline_t *pl, void *stream);

/*
 * Introspection for benchmarks/tests: number of kernels the last call on this thread launched, and
 * optional per-stage CUDA-event timing.  When stage timing is enabled the calls record events around
 * every stage on the caller's stream; sarpost_stage_times() synchronises those events and returns
 * milliseconds for {K1 candidates (incl. histogram memset), K2-K4 select+sort+NMS, K5 gather, whole call}.
 */
int32_t sarpost_last_launch_count(void);
int32_t sarpost_set_stage_timing(int32_t enabled); /* 0 off, 1 time the last call, 2 accumulate: sarpost_stage_times
                                                      then returns the MEAN over all calls since it was enabled / last read */
int32_t sarpost_stage_times(float *ms4);

#ifdef __cplusplus
}
#endif
#endif /* SARPOST_H_ */
