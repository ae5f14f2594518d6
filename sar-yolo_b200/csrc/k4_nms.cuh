// k4_nms.cuh — stage K3/K4: lazy stable sort + class-offset greedy NMS with early exit; K5: gather.
//
// Replaces the stable descending sort and greedy suppression of torchvision.ops.nms called at
// ops.py:296 on the class-offset boxes of ops.py:289,295, the `[:max_det]` cut (ops.py:297) and the
// row gather (ops.py:311).  Bit-exact to torchvision's CPU kernel (oracle/nms_greedy.c).
//
// Greedy NMS compares a candidate only with boxes KEPT before it, and stops at max_det keeps — so at
// most n*max_det pair tests are needed, not n^2/2, and usually only the head of the sorted order is
// ever looked at.  One CTA (or one cluster of CTAs, below) per image; selection scheme in k2_select_sort.cuh:
//   select    the per-image SAMPLED score histogram written by K1 is scanned from the top to size the next
//             run of whole score buckets (estimates only steer the size; membership is exact); the run is
//             collected by streaming the image's candidate scores once (tiles whose best score is below
//             the run are skipped);
//   phase 1   the collected candidates are tested against the kept list as it stood when the chunk began — this
//             needs no order, so it happens BEFORE sorting, and in the suppression-heavy regime (clustered
//             detections) a chunk spans thousands of candidates of which a handful survive;
//   sort      the survivors are sorted in shared memory (bitonic network on the 64-bit composite score|~slot =
//             descending score, source order on ties); a single bucket larger than that falls back to a stable
//             LSD radix sort in global memory;
//   phase 2   sub-chunks of <= 256 sorted survivors: test against boxes kept since the chunk began + upper-triangular
//             suppression bitmask among themselves (columns in registers, rows broadcast from shared memory, IoU only
//             for pairs that intersect with comparable areas); big sub-chunks are shared with the idle CTAs of the
//             cluster (candidates pushed through distributed shared memory, results stored back: 2 cluster barriers);
//   sweep     one warp resolves the sub-chunk on the bitmask, 32 candidates per step when no two live
//             candidates of the group overlap, else one step per KEPT box; all threads append to the kept list.
// Small batches (B*CL <= #SMs) run a thread-block CLUSTER of CL CTAs per image: collection and phase 1 are split
// across the CTAs (interleaved tiles), survivors travel to the master CTA through distributed shared memory, the
// master sorts / sweeps and replicates the new kept boxes: 2 cluster barriers per chunk.
#pragma once
#include <cooperative_groups.h>
#include <cuda_fp16.h>

#include "common.cuh"
#include "k2_select_sort.cuh"

namespace sarpost {

#ifdef SARPOST_PHASE_PROF
__device__ unsigned long long g_phase_n[16], g_phase_max[16];  // visits / longest single visit per phase
__device__ unsigned long long g_phase[16];  // cycles of block 0 per phase: 0 prologue, 1 collect, 2 share phase 1, 3 deliver + barrier 1,
// 4 sort, 5 radix fallback, 6 zoom histogram, 9 master tail + barrier 2, 8 publish; inside process_sorted: 10 load, 13 pair round (phase 1 +
// bitmask, incl. its cluster barriers), 11 sweep (warp 0), 14 append to the kept list + barriers
#define PROF_MARK(i)                                                         \
    do {                                                                     \
        if (blockIdx.x == 0 && threadIdx.x == 0) {                           \
            const long long _t = clock64();                                  \
            g_phase[i] += static_cast<unsigned long long>(_t - prof_t);      \
            g_phase_n[i] += 1ull;                                            \
            if (static_cast<unsigned long long>(_t - prof_t) > g_phase_max[i]) g_phase_max[i] = static_cast<unsigned long long>(_t - prof_t); \
            prof_t = _t;                                                     \
        }                                                                    \
    } while (0)
#else
#define PROF_MARK(i) do {} while (0)
#endif

constexpr int kNmsThreads = 512;
constexpr int kNmsWarps = kNmsThreads / 32;
constexpr int kSortCap = 1024;   // candidates sorted in shared memory at once
constexpr int kSub = 256;        // candidates per NMS sub-chunk
constexpr int kSubWords = kSub / 32;

// Where the extras columns of an output row live (gathered by K5 for the kept rows only).
struct ExtrasSrc {
    int32_t mode;               // 0 decoded prediction, 1 raw level tensors, 2 source detection rows (merge)
    int32_t nm, nc;
    // mode 0
    const void *pred;
    int32_t channels;
    int64_t anchors;
    // mode 1
    int32_t nl, no, n_extra_raw;
    int32_t lvl_aoff[kMaxLevels + 1];
    int32_t lvl_hw[kMaxLevels];
    const void *lvl_ptr[kMaxLevels];
    int32_t is_half;            // element type of the level tensors (mode 1) / of the prediction (mode 0)
    // mode 1, split layout (SARPOST_LAYOUT_SPLIT): the extras come from their own branch tensors
    int32_t split, emb_cl;
    const void *lvl_emb[kMaxLevels];    // (B, E, H, W), or channels-last (B, H, W, E) when emb_cl
    const void *lvl_state[kMaxLevels];  // (B, S, H, W) logits
    // mode 2
    const float *dets;
    int32_t dets_per_tile, row_len;
};

__device__ __forceinline__ float load_elem(const void *base, int64_t idx, bool is_half) {
    return is_half ? __half2float(__ldg(static_cast<const __half *>(base) + idx)) : __ldg(static_cast<const float *>(base) + idx);
}

// Extras of (image b, anchor) for modes 0 and 1: store(c, value) for c = lane, lane + 32, ... < nm.  Raw embedding
// columns are copied, state columns pass through the sigmoid (head.py:247); a decoded prediction is copied verbatim.
// Every source layout reduces to at most two strided column segments, walked by ONE copy loop (the kernel is a few
// hundred instructions per warp: code size — instruction-cache misses — is what it pays for).  Up to eight independent
// loads per lane are in flight before the first store (a plain load/store loop would serialise on the possible
// aliasing of source and destination).
template <class Store>
__device__ __forceinline__ void gather_extras_row(const ExtrasSrc &e, int b, uint32_t anchor, int lane, const Store &store) {
    const bool hf = e.is_half != 0;
    const void *base[2];
    int64_t at[2], cstride[2];
    int n[2];
    if (e.mode == 0) {
        base[0] = e.pred;
        at[0] = (static_cast<int64_t>(b) * e.channels + 4 + e.nc) * e.anchors + anchor;
        cstride[0] = e.anchors;
        n[0] = e.nm;
        base[1] = nullptr; at[1] = 0; cstride[1] = 0; n[1] = 0;
    } else {
        int l = 0;
#pragma unroll
        for (int i = 1; i < kMaxLevels; ++i) l += (i < e.nl && anchor >= static_cast<uint32_t>(e.lvl_aoff[i])) ? 1 : 0;
        const int64_t hw = e.lvl_hw[l];
        const int64_t pos = anchor - e.lvl_aoff[l];
        n[0] = e.n_extra_raw;
        n[1] = e.nm - e.n_extra_raw;
        cstride[0] = cstride[1] = hw;
        if (!e.split) {
            base[0] = base[1] = e.lvl_ptr[l];
            at[0] = (static_cast<int64_t>(b) * e.no + 4 * kRegMax + e.nc) * hw + pos;
            at[1] = at[0] + n[0] * hw;
        } else {
            base[0] = e.lvl_emb[l];
            base[1] = e.lvl_state[l];
            if (e.emb_cl) {  // one contiguous run of n_raw elements per kept row
                at[0] = (static_cast<int64_t>(b) * hw + pos) * n[0];
                cstride[0] = 1;
            } else {
                at[0] = static_cast<int64_t>(b) * n[0] * hw + pos;
            }
            at[1] = static_cast<int64_t>(b) * n[1] * hw + pos;
        }
    }
    // segment 1 (state columns: a handful) is fetched first so that its latency overlaps segment 0's
    const bool sig = e.mode != 0;
    float s1 = 0.0f;
    if (lane < n[1]) s1 = load_elem(base[1], at[1] + lane * cstride[1], hf);
#pragma unroll 1
    for (int c0 = lane; c0 < n[0]; c0 += 256) {
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u)
            if (c0 + 32 * u < n[0]) v[u] = load_elem(base[0], at[0] + (c0 + 32 * u) * cstride[0], hf);
#pragma unroll
        for (int u = 0; u < 8; ++u)
            if (c0 + 32 * u < n[0]) store(c0 + 32 * u, v[u]);
    }
#pragma unroll 1
    for (int c = lane; c < n[1]; c += 32) {
        const float v = c < 32 ? s1 : load_elem(base[1], at[1] + c * cstride[1], hf);
        store(n[0] + c, sig ? sigmoid_rn(v) : v);
    }
}

// ---------------------------------------------------------------------------------------------
// K5: one warp per output row.  Row = x1,y1,x2,y2,conf,cls,extras (ops.py:272/275, :311).
// Extras come from the decoded prediction (nms_decoded), from the raw level tensors (fused: raw
// embedding, sigmoid state — head.py:247), or from the source detection rows (merge).
// ---------------------------------------------------------------------------------------------
struct GatherParams {
    CandStore st;
    ExtrasSrc ex;
    const uint32_t *kept_slot;  // [B*max_det]
    const int32_t *counts;      // [B]
    float *out;                 // [B, max_det, 6+nm]
    int32_t *kept_index;        // [B*max_det] or nullptr
    const float *rescale;       // [B,5] pad_x, pad_y, gain, w0, h0 or nullptr (ops.scale_boxes + clip_boxes)
    int32_t max_det;
    int32_t tail_cols;          // columns reserved (unwritten) at the end of every output row
    // fused gather + exchange: rows/counts are stored into every rank's buffer (P2P-mapped pointers over NVLink)
    float *peer_out[8];
    int32_t *peer_counts[8];
    int32_t n_peers, peer_slot_offset;
    // results layout (sarpost_nms_params_t.res_boxes): 7-column boxes with the state id + contiguous embeddings
    float *res_boxes;    // [B, max_det, 7] x1,y1,x2,y2,state_id,conf,cls or nullptr
    float *res_embeds;   // [B, max_det, res_n_raw]
    int32_t res_n_raw;   // leading extras columns that are the embedding; the remaining nm - res_n_raw are state probabilities
};

// ops.scale_boxes (utils/ops.py:92-127, padding=True, xyxy) followed by clip_boxes (:319-338), in torch's fp32
// operation order: subtract the pad, true division by the gain, clamp to the original image.
__device__ __forceinline__ float4 rescale_box(float4 b, const float *rs) {
    const float px = rs[0], py = rs[1], gain = rs[2], w0 = rs[3], h0 = rs[4];
    b.x = fminf(fmaxf(__fdiv_rn(__fsub_rn(b.x, px), gain), 0.0f), w0);
    b.y = fminf(fmaxf(__fdiv_rn(__fsub_rn(b.y, py), gain), 0.0f), h0);
    b.z = fminf(fmaxf(__fdiv_rn(__fsub_rn(b.z, px), gain), 0.0f), w0);
    b.w = fminf(fmaxf(__fdiv_rn(__fsub_rn(b.w, py), gain), 0.0f), h0);
    return b;
}

constexpr int kGatherWarps = 8;

// One output row (image b, row r) by one warp: candidate `slot` with key `key`.
__device__ __forceinline__ void gather_row(const GatherParams &p, int b, int r, int lane, uint32_t slot, uint32_t key) {
    const int row_len = 6 + p.ex.nm + p.tail_cols;
    // destinations of this row: the local output, or the same slot in every rank's buffer (peer stores)
    const int n_dst = p.n_peers > 0 ? p.n_peers : 1;
    const int64_t img = p.n_peers > 0 ? p.peer_slot_offset + b : b;
    auto dst = [&](int q) { return (p.n_peers > 0 ? p.peer_out[q] : p.out) + (img * p.max_det + r) * row_len; };
    const int64_t seg = static_cast<int64_t>(b) * p.st.cap;
    if (p.ex.mode == 2) {
        const float *src = p.ex.dets + (static_cast<int64_t>(b) * p.st.tpi * p.ex.dets_per_tile + key) * p.ex.row_len;
        const float4 bx = p.st.box[seg + slot];
        // (destination, column) pairs spread over the lanes: with peer stores every lane talks to its own peer instead of
        // one lane writing to eight of them in turn
        const int ncols = 6 + p.ex.nm;
        for (int idx = lane; idx < ncols * n_dst; idx += 32) {
            const int q = idx / ncols, c = idx - q * ncols;
            const float v = c == 0 ? bx.x : c == 1 ? bx.y : c == 2 ? bx.z : c == 3 ? bx.w : src[c];
            dst(q)[c] = v;
        }
        if (lane == 0 && p.kept_index) p.kept_index[static_cast<int64_t>(b) * p.max_det + r] = static_cast<int32_t>(key);
        return;
    }
    const uint32_t anchor = key / static_cast<uint32_t>(p.ex.nc), cls = key - anchor * static_cast<uint32_t>(p.ex.nc);
    // Two output forms share one pass over the extras: the (6 + nm)-column row, or — results layout,
    // models/yolo/jde/predict.py:52-66 — boxes = cat(xyxy, argmax(states), conf, cls) + the raw embedding columns
    const bool res = p.res_boxes != nullptr;
    const int64_t row = static_cast<int64_t>(b) * p.max_det + r;
    const int n_raw = p.res_n_raw, n_sig = p.ex.nm - n_raw;
    float *emb = res ? p.res_embeds + row * n_raw : nullptr;
    float best = -1.0f;  // probabilities are >= 0
    int best_i = 0x7fffffff;
    const bool label_row = p.ex.mode == 0 && anchor >= static_cast<uint32_t>(p.ex.anchors);  // apriori label: zero extras (ops.py:258)
    if (p.ex.nm > 0 && label_row) {
        for (int c = lane; c < (res ? n_raw : p.ex.nm); c += 32) {
            if (res) emb[c] = 0.0f;
            else
                for (int q = 0; q < n_dst; ++q) dst(q)[6 + c] = 0.0f;
        }
        if (res && lane == 0 && n_sig > 0) { best = 0.0f; best_i = 0; }  // -> state id 0
    } else if (p.ex.nm > 0) {
        gather_extras_row(p.ex, b, anchor, lane, [&](int c, float v) {
            if (!res) {
                for (int q = 0; q < n_dst; ++q) dst(q)[6 + c] = v;
            } else if (c < n_raw) {
                emb[c] = v;
            } else if (v > best) {  // ascending c per lane + strict > : first maximum
                best = v;
                best_i = c - n_raw;
            }
        });
    }
    if (res) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, best, off);
            const int oi = __shfl_xor_sync(0xffffffffu, best_i, off);
            if (ov > best || (ov == best && oi < best_i)) { best = ov; best_i = oi; }
        }
    }
    if (res ? lane == 0 : lane < 6 * n_dst) {
        float4 bx = p.st.box[seg + slot];
        if (p.rescale) bx = rescale_box(bx, p.rescale + 5 * b);
        const float sc = p.st.score[seg + slot];
        if (res) {
            float *o = p.res_boxes + row * 7;
            o[0] = bx.x;
            o[1] = bx.y;
            o[2] = bx.z;
            o[3] = bx.w;
            o[4] = n_sig > 0 ? static_cast<float>(best_i) : -1.0f;
            o[5] = sc;
            o[6] = static_cast<float>(cls);
        } else {
            // the six leading columns, one (destination, column) pair per lane (peer stores: every lane its own peer)
            for (int idx = lane; idx < 6 * n_dst; idx += 32) {
                const int q = idx / 6, c = idx - q * 6;
                dst(q)[c] = c == 0 ? bx.x : c == 1 ? bx.y : c == 2 ? bx.z : c == 3 ? bx.w : c == 4 ? sc : static_cast<float>(cls);
            }
        }
        if (lane == 0 && p.kept_index) p.kept_index[row] = static_cast<int32_t>(key);
    }
}

struct NmsParams {
    CandStore st;
    uint32_t *tmp_key_a, *tmp_val_a; // [B*cap] scratch for the oversized-bucket fallback sort
    uint32_t *tmp_key_b, *tmp_val_b; // [B*cap]
    uint32_t *kept_slot;             // [B*max_det]
    int32_t *counts;                 // [B]
    int32_t max_det, max_nms;
    int32_t nc;                      // key % nc = class (merge: class comes from cls_override)
    const float *cls_override;       // merge path: class id per slot (float) or nullptr
    float max_wh;                    // 0 when agnostic
    float thr;                       // largest float <= iou_thres
    long long *stats;                // [B*4] or nullptr: candidates consumed, pair tests, sub-chunks, collections (instrumentation)
    int32_t *tile_counter;           // the decode kernel's tile counter: left zeroed for the next call (workspace_clean)
    unsigned int *resident_counter;  // pipeline only (else nullptr): every CTA checks in here as soon as it runs (k_gate)
};

// IoU(a,b) > thr, bit-exact to the fp32 division of the reference but without paying for it on every
// pair: a 2-instruction approximate quotient decides unless it lands within a few ulp of the threshold
// (`border`), in which case the caller re-evaluates the pair with the correctly rounded division
// (common.cuh iou_gt).  Branch-free so that callers can keep several pairs in flight.
__device__ __forceinline__ bool iou_gt_approx(const float4 a, const float area_a, const float4 b, const float area_b,
                                              const float thr, const float band, bool &border) {
    const float xx1 = fmaxf(a.x, b.x);
    const float yy1 = fmaxf(a.y, b.y);
    const float xx2 = fminf(a.z, b.z);
    const float yy2 = fminf(a.w, b.w);
    const float w = fmaxf(0.0f, __fsub_rn(xx2, xx1));
    const float h = fmaxf(0.0f, __fsub_rn(yy2, yy1));
    const float inter = __fmul_rn(w, h);
    const float uni = __fsub_rn(__fadd_rn(area_a, area_b), inter);
    const float q = __fdividef(inter, uni);  // <= 2 ulp for |uni| < 2^126
    border = fabsf(__fsub_rn(q, thr)) <= band;  // false for NaN
    return q > thr;
}

// One stable LSD radix pass (8-bit digit) over n (key,val) pairs in global memory, kNmsThreads threads.
// Warp w owns the contiguous share [w*per_warp, (w+1)*per_warp) so relative order is preserved.
template <class DigitFn>
__device__ __forceinline__ void radix_pass_global(const uint32_t *in_key, const uint32_t *in_val, uint32_t *out_key,
                                                  uint32_t *out_val, int n, const DigitFn &digit, int *cnt, int *warp_tot) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int kStride = kNmsWarps + 1;
    for (int i = threadIdx.x; i < 256 * kStride; i += kNmsThreads) cnt[i] = 0;
    __syncthreads();
    const int per_warp = (n + kNmsWarps - 1) / kNmsWarps;
    const int iters = (per_warp + 31) / 32;
    for (int it = 0; it < iters; ++it) {
        const int o = it * 32 + lane, i = warp * per_warp + o;
        const bool ok = o < per_warp && i < n;
        const uint32_t d = ok ? digit(in_key[i], in_val[i]) : 256u;
        const uint32_t peers = __match_any_sync(0xffffffffu, d);
        if (ok && (peers & lanemask_lt()) == 0) cnt[d * kStride + warp] += __popc(peers);
        __syncwarp();
    }
    __syncthreads();
    // exclusive scan over (digit-major, warp-minor): 256*16 = 4096 entries, 8 per thread
    int local[8], sum = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int e = threadIdx.x * 8 + i;
        local[i] = cnt[(e / kNmsWarps) * kStride + (e % kNmsWarps)];
        sum += local[i];
    }
    int inc = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += v;
    }
    if (lane == 31) warp_tot[warp] = inc;
    __syncthreads();
    int run = inc - sum;
    for (int w = 0; w < warp; ++w) run += warp_tot[w];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int e = threadIdx.x * 8 + i;
        cnt[(e / kNmsWarps) * kStride + (e % kNmsWarps)] = run;
        run += local[i];
    }
    __syncthreads();
    for (int it = 0; it < iters; ++it) {
        const int o = it * 32 + lane, i = warp * per_warp + o;
        const bool ok = o < per_warp && i < n;
        uint32_t key = 0, val = 0;
        if (ok) {
            key = in_key[i];
            val = in_val[i];
        }
        const uint32_t d = ok ? digit(key, val) : 256u;
        const uint32_t peers = __match_any_sync(0xffffffffu, d);
        int base = 0;
        if (ok) base = cnt[d * kStride + warp];
        __syncwarp();
        if (ok) {
            const int rank = __popc(peers & lanemask_lt());
            if (rank == 0) cnt[d * kStride + warp] = base + __popc(peers);
            out_key[base + rank] = key;
            out_val[base + rank] = val;
        }
        __syncwarp();
    }
    __syncthreads();
}

__device__ __forceinline__ unsigned long long shfl_xor_u64(unsigned long long v, int lane_mask) {
    const uint32_t lo = __shfl_xor_sync(0xffffffffu, static_cast<uint32_t>(v), lane_mask);
    const uint32_t hi = __shfl_xor_sync(0xffffffffu, static_cast<uint32_t>(v >> 32), lane_mask);
    return (static_cast<unsigned long long>(hi) << 32) | lo;
}

// Bitonic sorting network (descending) over 2^lpw <= 1024 keys in shared memory, blockDim = kNmsThreads.
// Thread t owns the adjacent pair (2t, 2t+1): compare-exchange strides 1..32 stay inside a warp (registers +
// shuffles, no barrier); only strides >= 64 go through shared memory.  15 barriers for 1024 keys instead of 55.
__device__ __forceinline__ void bitonic_sort_desc(unsigned long long *key, int lpw) {
    const int t = threadIdx.x, half_n = 1 << (lpw - 1);
    auto reg_steps = [&](unsigned long long &a, unsigned long long &b, int lk, int lj_hi) {
        const bool desc = ((t >> (lk - 1)) & 1) == 0;  // = ((2t >> lk) & 1) == 0, identical for 2t+1
        for (int lj = lj_hi; lj >= 1; --lj) {
            const int h = 1 << (lj - 1);  // partner thread = t ^ h
            const unsigned long long pa = shfl_xor_u64(a, h), pb = shfl_xor_u64(b, h);
            const bool keep_max = ((t & h) == 0) == desc;
            a = keep_max ? (a > pa ? a : pa) : (a < pa ? a : pa);
            b = keep_max ? (b > pb ? b : pb) : (b < pb ? b : pb);
        }
        if ((a < b) == desc) {
            const unsigned long long x = a;
            a = b;
            b = x;
        }
    };
    if (t < half_n) {  // warp-uniform: half_n is a multiple of 32
        unsigned long long a = key[2 * t], b = key[2 * t + 1];
        const int first = lpw < 6 ? lpw : 6;
        for (int lk = 1; lk <= first; ++lk) reg_steps(a, b, lk, lk - 1);
        key[2 * t] = a;
        key[2 * t + 1] = b;
    }
    __syncthreads();
    for (int lk = 7; lk <= lpw; ++lk) {
        for (int lj = lk - 1; lj >= 6; --lj) {
            if (t < half_n) {
                const int i = ((t >> lj) << (lj + 1)) | (t & ((1 << lj) - 1)), q = i | (1 << lj);
                const bool desc = ((i >> lk) & 1) == 0;
                const unsigned long long x = key[i], y = key[q];
                if ((x < y) == desc) {
                    key[i] = y;
                    key[q] = x;
                }
            }
            __syncthreads();
        }
        if (t < half_n) {
            unsigned long long a = key[2 * t], b = key[2 * t + 1];
            reg_steps(a, b, lk, 5);
            key[2 * t] = a;
            key[2 * t + 1] = b;
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// Shared memory of the NMS kernel: one dynamic allocation, fixed-size arrays first (compile-time offsets), the
// kept list (max_det entries) last.
// ---------------------------------------------------------------------------------------------
constexpr int kShareCap = 1024;  // members one CTA collects (and phase-1 tests) per chunk
constexpr int kMaxNmsCluster = 8;

struct NmsSmemLayout {
    static constexpr int plist = 0;                                  // u64[kShareCap]  this CTA's members of the chunk
    static constexpr int surv = plist + kShareCap * 8;               // u64[kShareCap]  its phase-1 survivors
    static constexpr int skey = surv + kShareCap * 8;                // u64[kSortCap]   master: every CTA's survivors, sorted
    static constexpr int radix_end = skey + kSortCap * 8;            // (the radix fallback aliases plist|surv|skey)
    static constexpr int bstart = radix_end;                         // i32[kBuckets + 4]
    static constexpr int mask = bstart + (kBuckets + 4) * 4;         // u32[kSubWords][kSub]  word-major bitmask
    static constexpr int a_box = mask + kSub * kSubWords * 4;        // float4[kSub]  (a_box|c_box double as the tile list)
    static constexpr int c_box = a_box + kSub * 16;                  // float4[kSub]
    static constexpr int a_area = c_box + kSub * 16;                 // f32[kSub]
    static constexpr int c_area = a_area + kSub * 4;
    static constexpr int a_slot = c_area + kSub * 4;                 // u32[kSub]
    static constexpr int c_slot = a_slot + kSub * 4;
    static constexpr int dead = c_slot + kSub * 4;                   // i32[kSub]  master: verdicts of the incremental phase 1
    static constexpr int deadw = dead + kSub * 4;                    // u32[kShareCap / 32]  share phase 1: dead bits
    static constexpr int misc = deadw + kShareCap / 32 * 4;          // i32[96]
    static constexpr int zstart = misc + 96 * 4;                     // i32[kBuckets + 4]  zoom: exact sub-bucket ranks of one bucket
    static constexpr int kept = zstart + (kBuckets + 4) * 4;         // float4[max_det] | f32[max_det] | u32[max_det]
};
static_assert(NmsSmemLayout::radix_end >= 256 * (kNmsWarps + 1) * 4, "radix counters must fit the plist|surv|skey region");
static_assert(kTileListCap * 4 <= 2 * kSub * 16, "tile list must fit the a_box|c_box region");
static_assert(NmsSmemLayout::kept % 16 == 0, "kept boxes must be 16-byte aligned");

__host__ __device__ inline size_t nms_smem_bytes(int max_det) { return NmsSmemLayout::kept + static_cast<size_t>(max_det) * 24; }

// misc[] slots
enum : int {
    kMWarp = 0,      // [0, 16) per-warp partial sums
    kMTot = 16,      // [16, 32) per-warp totals of the histogram prologue
    kMNAll = 32,     // exact candidate count of the image
    kMRunEnd = 33,   // chosen end of the bucket run
    kMShareCnt = 34, // members collected by this CTA (exact, may exceed kShareCap)
    kMSurvCnt = 35,  // phase-1 survivors of this CTA
    kMOff = 36,      // offset of this CTA's survivors in the master's list
    kMKept = 37,     // kept count after the sweep (master)
    kMListN = 38,    // tile list length
    kMCtl = 40,      // [40, 44) control block written by the master into every CTA: action, kept, members, survivors
    kMCtr = 48,      // [48, 56) master only: two banks (chunk parity) of {members, survivors, overflow}
    kMKm = 56,       // [56, 64) master only: kept bits per group of the sub-chunk just swept
    kMOrd = 64,      // [64, 68) helper CTAs: work order from the master {kOrdHelp | 0, candidates, k_from, kept}
};
constexpr int kOrdHelp = 1;
// a sub-chunk's pair work (bitmask + incremental phase 1) is shared with the other CTAs of the cluster when it exceeds this
// many IoU pairs: below, two cluster barriers cost more than the helpers save
constexpr long long kDistPairs = 10000;
enum : int { kActOk = 0, kActHalve = 1, kActCareful = 2, kActRadix = 3, kActZoom = 4 };
constexpr uint32_t kZoomSpan = (1u << kBucketShift) / kBuckets;  // float values per sub-bucket of a zoomed bucket (8)

// One CTA, or a thread-block cluster of CL CTAs, per image.
//   every CTA   collects its share (interleaved tiles) of the next run of score buckets and tests those candidates
//               against the kept list as of the start of the chunk (phase 1 needs no order, so it runs BEFORE any
//               sorting, on up to CL*kShareCap candidates at once); only the survivors travel to the master
//               (distributed shared memory, offsets from a remote atomic) — cluster barrier #1;
//   the master  sorts the survivors (bitonic network on score|~slot), resolves them in order (sub-chunks of <= 256:
//               incremental phase 1 against boxes kept since the chunk began, bitmask, one-warp sweep), replicates the
//               new kept boxes and a control block into the peers — cluster barrier #2.
// A chunk that would cross the max_nms rank cut, overflow the master's list or a CTA's share is redone (all members to
// the master / half the bucket run / radix fallback): phase 1 is idempotent and nothing is committed before barrier #2.
// MINB = 2 (one CTA per image only): the register budget is halved so that two CTAs share an SM — for batches with more
// images than SMs, where what the kernel costs is SM-time (it is a latency chain at an IPC of ~0.4), not its own duration.
template <int CL, int MINB = 1>
__global__ void __launch_bounds__(kNmsThreads, MINB) k4_nms(const __grid_constant__ NmsParams p) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    const int crank = CL > 1 ? static_cast<int>(cluster.block_rank()) : 0;
    auto cluster_sync = [&]() {
        if constexpr (CL > 1) cluster.sync(); else __syncthreads();
    };
    // Shared arrays are reached through offsets from the one dynamic symbol, so every local access compiles to
    // LDS/STS; only the explicitly remote ones (map_shared_rank) are generic.
    extern __shared__ __align__(16) unsigned char dyn[];
    using L = NmsSmemLayout;
#define PLIST (reinterpret_cast<unsigned long long *>(dyn + L::plist))
#define SURV (reinterpret_cast<unsigned long long *>(dyn + L::surv))
#define SKEY (reinterpret_cast<unsigned long long *>(dyn + L::skey))
#define S_BSTART (reinterpret_cast<int32_t *>(dyn + L::bstart))
#define S_MASK (reinterpret_cast<uint32_t *>(dyn + L::mask))
#define A_BOX (reinterpret_cast<float4 *>(dyn + L::a_box))
#define C_BOX (reinterpret_cast<float4 *>(dyn + L::c_box))
#define TILE_LIST (reinterpret_cast<uint32_t *>(dyn + L::a_box)) /* aliases A_BOX|C_BOX, idle while collecting */
#define S_A_AREA (reinterpret_cast<float *>(dyn + L::a_area))
#define S_C_AREA (reinterpret_cast<float *>(dyn + L::c_area))
#define S_A_SLOT (reinterpret_cast<uint32_t *>(dyn + L::a_slot))
#define S_C_SLOT (reinterpret_cast<uint32_t *>(dyn + L::c_slot))
#define S_DEAD (reinterpret_cast<int32_t *>(dyn + L::dead))
#define S_DEADW (reinterpret_cast<uint32_t *>(dyn + L::deadw))
#define S_MISC (reinterpret_cast<int32_t *>(dyn + L::misc))
#define S_ZSTART (reinterpret_cast<int32_t *>(dyn + L::zstart))
#define KEPT_BOX (reinterpret_cast<float4 *>(dyn + L::kept))
#define KEPT_AREA (reinterpret_cast<float *>(dyn + L::kept + static_cast<size_t>(p.max_det) * 16))
#define KEPT_SLOT (reinterpret_cast<uint32_t *>(dyn + L::kept + static_cast<size_t>(p.max_det) * 20))

    const int b = blockIdx.x / CL, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t seg = static_cast<int64_t>(b) * p.st.cap;
    const int32_t *tcount = p.st.tile_count + static_cast<int64_t>(b) * p.st.tpi;
    const uint32_t *tmaxv = p.st.tile_max + static_cast<int64_t>(b) * p.st.tpi;
    const float *score = p.st.score + seg;
    const float band = fmaxf(p.thr * 2e-6f, 1e-37f);
    // IoU <= min(area)/max(area): the fp32 intersection never exceeds either fp32 area (rounding is monotone) and the
    // union is >= the larger area up to an ulp, so a pair whose area ratio is <= thr (less a 1e-5 relative margin, far
    // above those ulps) cannot exceed the threshold: no IoU arithmetic needed for boxes of clearly different size.
    const float ratio_cut = p.thr * (1.0f - 1e-5f);
    const bool need_cls = p.max_wh != 0.0f && (p.cls_override != nullptr || p.nc > 1);
#ifdef SARPOST_PHASE_PROF
    long long prof_t = clock64();
#endif
    if (tid < kSub) S_DEAD[tid] = 0;
    if (tid >= 32 && tid < 96) S_MISC[tid] = 0;  // control block, counters (both banks), work order
    if (p.resident_counter != nullptr && tid == 0) atomicAdd(p.resident_counter, 1u);
    // The kernel is launched as a programmatic dependent of the candidate kernel (cudaLaunchAttributeProgrammaticStream-
    // Serialization): its CTAs — each needs a whole SM — are placed while that kernel drains and wait here until all of its
    // results are visible.  Nothing above touches global memory the candidate kernel uses (a plain launch passes at once).
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (blockIdx.x == 0 && tid == 0) *p.tile_counter = 0;

    // class-offset box of a candidate slot (ops.py:289,295)
    auto offset_box = [&](uint32_t slot) {
        const float4 bx = p.st.box[seg + slot];
        float cls = 0.0f;
        if (need_cls)
            cls = p.cls_override ? p.cls_override[seg + slot] : static_cast<float>(p.st.key[seg + slot] % static_cast<uint32_t>(p.nc));
        const float off = __fmul_rn(cls, p.max_wh);
        return make_float4(__fadd_rn(bx.x, off), __fadd_rn(bx.y, off), __fadd_rn(bx.z, off), __fadd_rn(bx.w, off));
    };
    // does any kept box k in [k_lo, k_hi), k = k_lo + part, + nparts, ... suppress (ob, oa)?  `dead` counts only verdicts
    // the approximate quotient can be trusted with (not within the band around the threshold); if none of those fired
    // and some pair was borderline, every pair is re-evaluated with the correctly rounded division
    auto killed_by_kept = [&](const float4 ob, const float oa, int k_lo, int k_hi, int part, int nparts) {
        bool dead = false, any_border = false;
#pragma unroll 4
        for (int k = k_lo + part; k < k_hi; k += nparts) {
            bool bd;
            const bool gt = iou_gt_approx(KEPT_BOX[k], KEPT_AREA[k], ob, oa, p.thr, band, bd);
            dead |= gt && !bd;
            any_border |= bd;
        }
        if (any_border && !dead)  // rare: a quotient within a few ulp of the threshold
            for (int k = k_lo + part; k < k_hi; k += nparts) dead |= iou_gt(KEPT_BOX[k], KEPT_AREA[k], ob, oa, p.thr);
        return dead;
    };
    // The same verdict for the master's incremental phase 1, where boxes kept a moment ago rarely overlap the next
    // candidates: four kept boxes per step, and the IoU arithmetic runs only when some lane of the warp sees its
    // candidate intersect one of them.  Called by every thread of the block (`valid` = the thread owns a candidate);
    // `part` / `nparts` must be warp-uniform.  Disjoint boxes have inter = 0 -> IoU 0 or NaN, never > thr >= 0.
    auto killed_by_kept_sparse = [&](bool valid, const float4 ob, const float oa, int k_lo, int k_hi, int part, int nparts) {
        bool dead = false;
        for (int k = k_lo + part; k < k_hi; k += 4 * nparts) {
            float4 kb[4];
            float ka4[4];
            bool ov[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int ku = k + u * nparts;
                kb[u] = KEPT_BOX[ku < k_hi ? ku : k];
                ka4[u] = KEPT_AREA[ku < k_hi ? ku : k];
                ov[u] = valid && ku < k_hi && kb[u].z > ob.x && ob.z > kb[u].x && kb[u].w > ob.y && ob.w > kb[u].y &&
                        !(fminf(ka4[u], oa) <= ratio_cut * fmaxf(ka4[u], oa));
            }
            if (!__any_sync(0xffffffffu, ov[0] || ov[1] || ov[2] || ov[3])) continue;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (!ov[u]) continue;
                const float ka = ka4[u];
                bool bd;
                bool gt = iou_gt_approx(kb[u], ka, ob, oa, p.thr, band, bd);
                if (bd) gt = iou_gt(kb[u], ka, ob, oa, p.thr);  // rare: a quotient within a few ulp of the threshold
                dead |= gt;
            }
        }
        return dead;
    };

    // ---- descending exclusive scan of the sampled score histogram: S_BSTART[d] ~ estimated rank of the first
    //      candidate of bucket 4095-d in the sorted order (x kHistSample).  Estimates only steer how many
    //      buckets a chunk spans; membership, ranks and results are exact.  n_all = exact candidate count. ----
    {
        int32_t *hist = p.st.hist + static_cast<int64_t>(b) * kBuckets;
        constexpr int kPer = kBuckets / kNmsThreads;  // 8 consecutive d per thread
        int loc[kPer], sum = 0;
#pragma unroll
        for (int i = 0; i < kPer; ++i) {
            loc[i] = hist[kBuckets - 1 - (tid * kPer + i)] * kHistSample;
            if constexpr (CL == 1) hist[kBuckets - 1 - (tid * kPer + i)] = 0;  // leave it zeroed for the next call (workspace_clean)
            sum += loc[i];
        }
        int tot = 0;
        for (int t = tid; t < p.st.tpi; t += kNmsThreads) tot += tcount[t];
        int inc = sum;
#pragma unroll
        for (int dd = 1; dd < 32; dd <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, inc, dd);
            if (lane >= dd) inc += v;
        }
        tot = __reduce_add_sync(0xffffffffu, tot);
        if (lane == 31) S_MISC[kMWarp + warp] = inc;
        if (lane == 0) S_MISC[kMTot + warp] = tot;
        __syncthreads();
        int run = inc - sum;
        for (int w = 0; w < warp; ++w) run += S_MISC[kMWarp + w];
#pragma unroll
        for (int i = 0; i < kPer; ++i) {
            S_BSTART[tid * kPer + i] = run;
            run += loc[i];
        }
        if (tid == kNmsThreads - 1) S_BSTART[kBuckets] = run;
        tot = 0;
#pragma unroll
        for (int w = 0; w < kNmsWarps; ++w) tot += S_MISC[kMTot + w];
        __syncthreads();
        if (tid == 0) S_MISC[kMNAll] = tot;
    }
    __syncthreads();
    // Distributed shared memory may only be touched once every CTA of the cluster runs and has initialised its control
    // block: each CTA arrives here and waits right before its first remote access (by then the barrier is long complete)
    if constexpr (CL > 1) asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    bool cluster_ready = CL == 1;
    PROF_MARK(0);
    const int n_all = S_MISC[kMNAll];
    const int n_limit = min(n_all, p.max_nms);  // ops.py:285-286: only the top max_nms ranks are eligible
    int kept = 0;
    int pos = 0, d = 0;  // pos = exact number of candidates consumed (all buckets above descending index d)
    long long st_walk = 0, st_pairs = 0;  // instrumentation (NmsParams::stats); st_walk/st_sub/st_coll uniform, st_pairs per CTA
    int st_sub = 0, st_coll = 0;

    // The pair work of one sub-chunk for this CTA's slice (`cpart` of `cparts` CTAs of the cluster).  A_BOX / S_A_AREA hold
    // `sub` sorted candidates that were already tested against kept[0, k_from); boxes [k_from, k_now) have been kept since.
    //   (a) incremental phase 1: candidates x those boxes (this CTA's share of the boxes) -> S_DEAD verdicts
    //   (b) suppression bitmask among the candidates themselves (this CTA's share of the rows), word-major:
    //       S_MASK[w * kSub + r] = word w (columns 32w .. 32w+31) of row r, bits j > r only
    // Both land in the MASTER's shared memory: local stores on the master, distributed-shared-memory stores from a helper.
    auto pair_round = [&](const int sub, const int k_from, const int k_now, const int cpart, const int cparts) {
        int32_t *dead_out = S_DEAD;
        uint32_t *mask_out = S_MASK;
        if constexpr (CL > 1) {
            if (cpart != 0) {
                dead_out = cluster.map_shared_rank(S_DEAD, 0);
                mask_out = cluster.map_shared_rank(S_MASK, 0);
            }
        }
        if (k_now > k_from) {
            int sub_p2 = 64;
            while (sub_p2 < sub) sub_p2 <<= 1;
            const int cand = tid & (sub_p2 - 1), part = tid / sub_p2, nparts = kNmsThreads / sub_p2;
            const int cc = cand < sub ? cand : 0;
            if (killed_by_kept_sparse(cand < sub, A_BOX[cc], S_A_AREA[cc], k_from, k_now, cpart * nparts + part, cparts * nparts)) dead_out[cand] = 1;
            st_pairs += static_cast<long long>(sub) * (k_now - k_from) / cparts;
        }
        // Work items = (word w, row r < min(sub, 32(w+1))), flattened word-major and cut into equal slices: one per CTA, then
        // one per warp.  A warp keeps the 32 column boxes of its current word in registers (lane = column) and streams rows
        // past them, four at a time: one broadcast shared load and a few compares per (row, word); the IoU arithmetic runs
        // only for rows that intersect some column of the word with a comparable area (disjoint boxes: inter = 0 -> IoU 0
        // or NaN, never > thr; area ratio <= thr: see ratio_cut).
        const int words = (sub + 31) >> 5;
        const int n_items = 16 * words * (words - 1) + sub;  // full words contribute 32(w+1) rows each, the last one `sub`
        const int c_lo = static_cast<int>(static_cast<long long>(n_items) * cpart / cparts);
        const int c_hi = static_cast<int>(static_cast<long long>(n_items) * (cpart + 1) / cparts);
        const int it_lo = c_lo + static_cast<int>(static_cast<long long>(c_hi - c_lo) * warp / kNmsWarps);
        const int it_hi = c_lo + static_cast<int>(static_cast<long long>(c_hi - c_lo) * (warp + 1) / kNmsWarps);
        int it = it_lo, w = 0;
        while (w + 1 < words && 16 * (w + 1) * (w + 2) <= it) ++w;
        while (it < it_hi) {
            const int r_begin = it - 16 * w * (w + 1);
            const int r_end = min(w == words - 1 ? sub : 32 * (w + 1), r_begin + (it_hi - it));
            const int j = (w << 5) + lane;
            const bool jv = j < sub;
            const float4 cb = A_BOX[jv ? j : 0];
            const float ca = S_A_AREA[jv ? j : 0];
            for (int r0 = r_begin; r0 < r_end; r0 += 4) {
                float4 rb[4];
                float ra4[4];
                bool ov[4], any[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    rb[u] = A_BOX[min(r0 + u, r_end - 1)];
                    ra4[u] = S_A_AREA[min(r0 + u, r_end - 1)];
                }
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    ov[u] = jv && j > r0 + u && rb[u].z > cb.x && cb.z > rb[u].x && rb[u].w > cb.y && cb.w > rb[u].y &&
                            !(fminf(ra4[u], ca) <= ratio_cut * fmaxf(ra4[u], ca));
#pragma unroll
                for (int u = 0; u < 4; ++u) any[u] = __any_sync(0xffffffffu, ov[u]);
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    if (r0 + u >= r_end) break;
                    uint32_t bits = 0u;
                    if (any[u]) {
                        bool g = false;
                        if (ov[u]) {
                            bool bd;
                            g = iou_gt_approx(rb[u], ra4[u], cb, ca, p.thr, band, bd);
                            if (bd) g = iou_gt(rb[u], ra4[u], cb, ca, p.thr);  // rare: quotient within a few ulp of thr
                        }
                        bits = __ballot_sync(0xffffffffu, g);
                    }
                    if (lane == 0) mask_out[w * kSub + r0 + u] = bits;
                }
            }
            it += r_end - r_begin;
            ++w;
        }
        if (cpart == 0) st_pairs += static_cast<long long>(sub) * (sub - 1) / 2;
    };

    // MASTER ONLY.  Walks `cnt` sorted candidates whose slots are produced by slot_at(i); they have already been tested
    // against kept[0, k_from).  Sub-chunks of <= 256: load, pair work (shared with the other CTAs of the cluster when it is
    // worth two cluster barriers), one-warp sweep over the bitmask, append to the kept list (replicated into the peers).
    auto process_sorted = [&](auto slot_at, int cnt, int k_from) {
        int pdone = 0;
        while (pdone < cnt && kept < p.max_det) {
            const int need = p.max_det - kept;
            const int sub = min(cnt - pdone, min(kSub, max(64, 2 * need)));
            // ---- load sub-chunk ----
            if (tid < sub) {
                const uint32_t slot = slot_at(pdone + tid);
                const float4 ob = offset_box(slot);
                A_BOX[tid] = ob;
                S_A_AREA[tid] = box_area_rn(ob);
                S_A_SLOT[tid] = slot;
            }
            __syncthreads();
            PROF_MARK(10);
            // ---- pair work: alone, or with the helpers (candidates + a work order pushed into their shared memory) ----
            bool dist = false;
            if constexpr (CL > 1) {
                dist = static_cast<long long>(sub) * (sub / 2 + (kept - k_from)) >= kDistPairs;
                if (dist) {
                    for (int r2 = 1; r2 < CL; ++r2) {
                        if (tid < sub) {
                            cluster.map_shared_rank(A_BOX, r2)[tid] = A_BOX[tid];
                            cluster.map_shared_rank(S_A_AREA, r2)[tid] = S_A_AREA[tid];
                        }
                        if (tid == 0) {
                            int32_t *o = cluster.map_shared_rank(&S_MISC[kMOrd], r2);
                            o[1] = sub;
                            o[2] = k_from;
                            o[3] = kept;
                            o[0] = kOrdHelp;
                        }
                    }
                    cluster.sync();  // helpers start
                }
            }
            pair_round(sub, k_from, kept, 0, dist ? CL : 1);
            if (dist) cluster_sync(); else __syncthreads();  // every verdict and mask word is in
            PROF_MARK(13);
            // ---- sweep (warp 0): which candidates are kept, 32 per step ----
            const int words = (sub + 31) >> 5;
            if (warp == 0) {
                uint32_t km[kSubWords], deadw[kSubWords];  // kept bits of the groups resolved so far / phase-1 verdicts (warp-uniform)
#pragma unroll
                for (int g = 0; g < kSubWords; ++g) {
                    km[g] = 0u;
                    deadw[g] = __ballot_sync(0xffffffffu, S_DEAD[(g << 5) + lane] != 0);
                }
                int kl = kept;
                // column g of the bitmask (word g of the rows g2*32+lane, g2 <= g; word-major layout: conflict-free) is loaded
                // one group ahead, before the dependent chain of group g-1, so the sweep never waits on shared memory
                uint32_t cur[kSubWords], nxt[kSubWords];
                cur[0] = S_MASK[lane];
#pragma unroll
                for (int g = 0; g < kSubWords; ++g) {
                    if (g + 1 < kSubWords) {
#pragma unroll
                        for (int g2 = 0; g2 <= g + 1; ++g2) nxt[g2] = S_MASK[(g + 1) * kSub + (g2 << 5) + lane];
                    }
                    if (g < words && kl < p.max_det) {
                        // removal word of group g = OR over the rows kept in earlier groups (lane = row)
                        uint32_t rem = 0u;
#pragma unroll
                        for (int g2 = 0; g2 < g; ++g2) rem |= ((km[g2] >> lane) & 1u) ? cur[g2] : 0u;
                        rem = __reduce_or_sync(0xffffffffu, rem) | deadw[g];
                        const int nvalid = min(32, sub - (g << 5));
                        const uint32_t live = ~rem & (nvalid == 32 ? 0xffffffffu : ((1u << nvalid) - 1u));
                        const uint32_t diag = lane < nvalid ? cur[g] : 0u;
                        // rows that overlap a later live row of the group are the only ones whose fate matters to
                        // others: walk just those in order (usually none or a handful of the 32)
                        uint32_t pending = __ballot_sync(0xffffffffu, ((live >> lane) & 1u) && (diag & live));
                        uint32_t keptm = live;
                        while (pending) {
                            const int j = __ffs(static_cast<int>(pending)) - 1;
                            pending &= pending - 1u;
                            const uint32_t dj = __shfl_sync(0xffffffffu, diag, j);
                            if ((keptm >> j) & 1u) keptm &= ~dj;  // j is still alive: it suppresses its overlaps
                        }
                        int c = __popc(keptm);
                        if (kl + c > p.max_det) {  // keep only the first (max_det - kl) of them
                            const int allow = p.max_det - kl;
                            uint32_t t = keptm, res = 0u;
                            for (int q = 0; q < allow; ++q) {
                                const uint32_t low = t & (0u - t);
                                res |= low;
                                t ^= low;
                            }
                            keptm = res;
                            c = allow;
                        }
                        kl += c;
                        km[g] = keptm;
                    }
                    if (g + 1 < kSubWords) {
#pragma unroll
                        for (int g2 = 0; g2 <= g + 1; ++g2) cur[g2] = nxt[g2];
                    }
                }
                if (lane == 0) {
                    S_MISC[kMKept] = kl;
#pragma unroll
                    for (int g = 0; g < kSubWords; ++g) S_MISC[kMKm + g] = static_cast<int32_t>(km[g]);
                }
                PROF_MARK(11);
            }
            __syncthreads();
            // ---- append the kept candidates (in order) to the kept list, here and in every peer; clear the verdicts ----
            if (tid < kSub) {
                const uint32_t kmg = static_cast<uint32_t>(S_MISC[kMKm + warp]);  // tid < 256: warp == group
                if (tid < sub && ((kmg >> lane) & 1u)) {
                    int idx = kept + __popc(kmg & lanemask_lt());
                    for (int g2 = 0; g2 < warp; ++g2) idx += __popc(static_cast<uint32_t>(S_MISC[kMKm + g2]));
                    const float4 bx = A_BOX[tid];
                    const float ar = S_A_AREA[tid];
                    KEPT_BOX[idx] = bx;
                    KEPT_AREA[idx] = ar;
                    KEPT_SLOT[idx] = S_A_SLOT[tid];
                    if constexpr (CL > 1) {
                        for (int r2 = 1; r2 < CL; ++r2) {  // visible to the peer at the next cluster barrier
                            cluster.map_shared_rank(KEPT_BOX, r2)[idx] = bx;
                            cluster.map_shared_rank(KEPT_AREA, r2)[idx] = ar;
                        }
                    }
                }
                S_DEAD[tid] = 0;
            }
            ++st_sub;
            st_walk += sub;
            kept = S_MISC[kMKept];
            pdone += sub;
            __syncthreads();
            PROF_MARK(14);
        }
    };

    // this CTA's share of the image's tiles: crank, crank + CL, ... (everything when CL == 1)
    const int n_share = crank < p.st.tpi ? (p.st.tpi - crank + CL - 1) / CL : 0;
    int par = 0;          // chunk parity: which bank of the master's counters this round uses
    bool careful = false; // redo of a chunk that crosses the max_nms cut: every member goes to the master (no phase 1 first)
    int last_m = 1, last_s = 1;  // members / survivors of the last committed chunk (survival rate steers the chunk size)
    int m = 0, s_all = 0;        // members / survivors of the last attempt (cluster-uniform after barrier #2)
    bool last_p1 = false;        // the last attempt ran phase 1 on the shares

    // One attempt at the candidates whose score bit patterns lie in [lo_bits, hi_bits]: collect, phase 1, deliver, and —
    // on the master — sort + resolve.  Returns the cluster-uniform verdict: kActOk (committed: `kept` advanced, the range
    // is consumed), kActCareful (the range crosses the max_nms rank cut: come back with careful = true), or `on_ovf`
    // when the range does not fit the shared-memory lists — kActHalve / kActZoom: nothing was committed, the caller
    // narrows the range; kActRadix: the range cannot be narrowed and was resolved here through the global radix sort.
    auto attempt = [&](const uint32_t lo_bits, const uint32_t hi_bits, const int on_ovf) -> int {
        int action = kActOk;
        const int k0 = kept;
        ++st_coll;
        __syncthreads();
        if (tid == 0) { S_MISC[kMShareCnt] = 0; S_MISC[kMSurvCnt] = 0; }
        if (tid < kShareCap / 32) S_DEADW[tid] = 0u;
        __syncthreads();
        // ---- collect this CTA's members of the range ----
        for_each_candidate_in(p.st, tcount, tmaxv, score, lo_bits, hi_bits, crank, CL, n_share, TILE_LIST, &S_MISC[kMListN],
                              [&](uint32_t slot, uint32_t bits) {
                                  const int at = atomicAdd(&S_MISC[kMShareCnt], 1);  // members are rare (a few hundred per image)
                                  if (at < kShareCap) PLIST[at] = (static_cast<unsigned long long>(bits) << 32) | (0xffffffffu - slot);
                              });
        __syncthreads();
        PROF_MARK(1);
        const int my_cnt = S_MISC[kMShareCnt];
        const int my_n = min(my_cnt, kShareCap);
        // ---- phase 1 on the share: members x kept[0, k0) ----
        const bool do_p1 = k0 > 0 && !careful;
        last_p1 = do_p1;
        if (do_p1) {
            for (int base = 0; base < my_n; base += kNmsThreads) {
                const int nb = min(kNmsThreads, my_n - base);
                int sub_p2 = 64;
                while (sub_p2 < nb) sub_p2 <<= 1;
                const int ci = tid & (sub_p2 - 1), part = tid / sub_p2, nparts = kNmsThreads / sub_p2;
                bool dead = false;
                if (ci < nb) {
                    const uint32_t slot = 0xffffffffu - static_cast<uint32_t>(PLIST[base + ci]);
                    const float4 ob = offset_box(slot);
                    dead = killed_by_kept(ob, box_area_rn(ob), 0, k0, part, nparts);
                }
                const uint32_t bal = __ballot_sync(0xffffffffu, dead);  // a warp = 32 consecutive members, one part
                if (lane == 0 && bal) atomicOr(&S_DEADW[(base + ci) >> 5], bal);
            }
            st_pairs += static_cast<long long>(my_n) * k0;
            __syncthreads();
            for (int i0 = warp * 32; i0 < my_n; i0 += kNmsThreads) {  // unordered compaction (the survivors get sorted anyway)
                const int i = i0 + lane;
                const bool alive = i < my_n && !((S_DEADW[i >> 5] >> (i & 31)) & 1u);
                const uint32_t bal = __ballot_sync(0xffffffffu, alive);
                int at = 0;
                if (lane == 0 && bal) at = atomicAdd(&S_MISC[kMSurvCnt], __popc(bal));
                at = __shfl_sync(0xffffffffu, at, 0);
                if (alive) SURV[at + __popc(bal & lanemask_lt())] = PLIST[i];
            }
            __syncthreads();
        }
        PROF_MARK(2);
        const unsigned long long *mine = do_p1 ? SURV : PLIST;
        const int s_cnt = do_p1 ? S_MISC[kMSurvCnt] : my_n;
        // ---- survivors -> master (offset from a remote atomic on the master's counters) ----
        if constexpr (CL > 1) {
            if (!cluster_ready) {
                asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
                cluster_ready = true;
            }
        }
        int32_t *ctr = &S_MISC[kMCtr + par * 4];
        if constexpr (CL > 1) ctr = cluster.map_shared_rank(&S_MISC[kMCtr + par * 4], 0);
        if (tid == 0) {
            atomicAdd(ctr + 0, my_cnt);
            S_MISC[kMOff] = atomicAdd(ctr + 1, s_cnt);
            if (my_cnt > kShareCap) atomicOr(ctr + 2, 1);
        }
        __syncthreads();
        {
            const int off = S_MISC[kMOff];
            unsigned long long *dst = SKEY;
            if constexpr (CL > 1) dst = cluster.map_shared_rank(SKEY, 0);
            for (int i = tid; i < s_cnt; i += kNmsThreads)
                if (off + i < kSortCap) dst[off + i] = mine[i];
        }
        cluster_sync();  // barrier #1
        PROF_MARK(3);
        if (crank == 0) {
            const int mm = S_MISC[kMCtr + par * 4 + 0];
            const int ss = S_MISC[kMCtr + par * 4 + 1];
            const bool share_ovf = S_MISC[kMCtr + par * 4 + 2] != 0;
            if (share_ovf || ss > kSortCap) action = on_ovf;
            else if (!careful && k0 > 0 && pos + mm > n_limit) action = kActCareful;  // crosses the rank cut: ranks need every member
            else action = kActOk;
            __syncthreads();
            if (tid < 4) S_MISC[kMCtr + (par ^ 1) * 4 + tid] = 0;  // the other bank is idle until the next round
            if (action == kActOk) {
                if (ss > 0) {
                    int lpw = 6;
                    while ((1 << lpw) < ss) ++lpw;
                    for (int i = ss + tid; i < (1 << lpw); i += kNmsThreads) SKEY[i] = 0ull;
                    __syncthreads();
                    bitonic_sort_desc(SKEY, lpw);
                    PROF_MARK(4);
                    // without a preceding phase 1 (first chunk, careful redo) the sorted list holds every member, so
                    // the rank cut applies directly; after phase 1 the whole chunk is known to lie above the cut
                    const int cnt = do_p1 ? ss : min(ss, n_limit - pos);
                    process_sorted([&](int i) { return 0xffffffffu - static_cast<uint32_t>(SKEY[i]); }, cnt, do_p1 ? k0 : 0);
                }
            } else if (action == kActRadix) {
                // ---- more equal (or nearly equal) scores than shared memory holds: collect the range to global scratch,
                //      stable LSD radix sort on (score bits desc, slot asc), then stream it through the suppression phases
                //      (rare path: thousands of candidates within 8 adjacent float values) ----
                uint32_t *ka = p.tmp_key_a + seg, *va = p.tmp_val_a + seg, *kb = p.tmp_key_b + seg, *vb = p.tmp_val_b + seg;
                const uint32_t *ik = ka, *iv = va;
                uint32_t *ok = kb, *ov = vb;
                int slot_bits = 1;
                while ((static_cast<int64_t>(1) << slot_bits) < p.st.cap) ++slot_bits;
                const int passes_slot = (slot_bits + 7) / 8;
                // the members share every score bit above the range's span: only the bits that can differ need a pass
                const int passes_key = (32 - __clz(static_cast<int>(lo_bits ^ hi_bits)) + 7) / 8;
                __syncthreads();
                if (tid == 0) S_MISC[kMShareCnt] = 0;
                __syncthreads();
                for_each_candidate_in(p.st, tcount, tmaxv, score, lo_bits, hi_bits, 0, 1, p.st.tpi, TILE_LIST, &S_MISC[kMListN],
                                      [&](uint32_t slot, uint32_t bits) {
                                          const int at = atomicAdd(&S_MISC[kMShareCnt], 1);
                                          ka[at] = bits;
                                          va[at] = slot;
                                      });
                __syncthreads();
                int *cnt = reinterpret_cast<int *>(dyn + L::plist);  // aliases plist|surv|skey
                int *wt = S_MISC + kMWarp;
                for (int ps = 0; ps < passes_slot + passes_key; ++ps) {
                    const int sh = ps < passes_slot ? ps * 8 : (ps - passes_slot) * 8;
                    if (ps < passes_slot)
                        radix_pass_global(ik, iv, ok, ov, mm, [sh](uint32_t, uint32_t v) { return (v >> sh) & 255u; }, cnt, wt);
                    else
                        radix_pass_global(ik, iv, ok, ov, mm, [sh](uint32_t k, uint32_t) { return ((~k) >> sh) & 255u; }, cnt, wt);
                    const uint32_t *tk = ik, *tv = iv;
                    ik = ok; iv = ov;
                    ok = const_cast<uint32_t *>(tk); ov = const_cast<uint32_t *>(tv);
                }
                const uint32_t *sorted = iv;
                const int lim = min(mm, n_limit - pos);
                for (int piece = 0; piece < lim && kept < p.max_det; piece += kSortCap)
                    process_sorted([&](int i) { return sorted[piece + i]; }, min(kSortCap, lim - piece), 0);
                PROF_MARK(5);
            }
            // ---- the verdict into every CTA (the kept boxes were replicated sub-chunk by sub-chunk); helpers are released ----
            if constexpr (CL > 1) {
                __syncthreads();
                if (tid < CL) {
                    int32_t *ctl = cluster.map_shared_rank(&S_MISC[kMCtl], tid);
                    ctl[0] = action;
                    ctl[1] = kept;
                    ctl[2] = mm;
                    ctl[3] = ss;
                    *cluster.map_shared_rank(&S_MISC[kMOrd], tid) = 0;
                }
            } else {
                if (tid == 0) {
                    S_MISC[kMCtl + 0] = action;
                    S_MISC[kMCtl + 1] = kept;
                    S_MISC[kMCtl + 2] = mm;
                    S_MISC[kMCtl + 3] = ss;
                }
            }
        }
        if constexpr (CL > 1) {
            if (crank != 0) {
                // helper: every cluster barrier here is either the start of a pair round (work order present) or barrier #2
                for (;;) {
                    cluster.sync();
                    if (S_MISC[kMOrd] != kOrdHelp) break;
                    pair_round(S_MISC[kMOrd + 1], S_MISC[kMOrd + 2], S_MISC[kMOrd + 3], crank, CL);
                    cluster.sync();  // results delivered
                }
            } else {
                cluster.sync();  // barrier #2
            }
        } else {
            __syncthreads();  // barrier #2
        }
        action = S_MISC[kMCtl + 0];
        kept = S_MISC[kMCtl + 1];
        m = S_MISC[kMCtl + 2];
        s_all = S_MISC[kMCtl + 3];
        par ^= 1;
        PROF_MARK(9);
        return action;
    };

    // Two levels of positions, both walked in descending score order with the same loop: the 4096 score buckets
    // (S_BSTART: ESTIMATED ranks from K1's sampled histogram), and — when a single bucket holds more candidates than
    // shared memory (clustered detections whose scores saturate in one 0.4 % wide bucket) — the 4096 sub-buckets of 8
    // adjacent float values inside it (S_ZSTART: EXACT counts from one scan of the bucket).  Only a sub-bucket that still
    // overflows (thousands of practically equal scores) takes the global radix sort.
    bool zoom_on = false;
    int zd = 0, z = 0;  // zoomed bucket (descending index) and position inside it
    uint32_t z_lo = 0;  // smallest score bit pattern of the zoomed bucket
    while (pos < n_limit && kept < p.max_det) {
        if (zoom_on) {
            if (z >= kBuckets || S_ZSTART[kBuckets] == S_ZSTART[z]) {  // bucket exhausted: back to the coarse walk
                zoom_on = false;
                d = zd + 1;
                continue;
            }
        } else if (d >= kBuckets) {
            break;
        }
        const int32_t *start = zoom_on ? S_ZSTART : S_BSTART;
        const int from = zoom_on ? z : d;
        // ---- next chunk: a run of whole (sub-)buckets [from, to) whose population (estimated / exact) fits the target ----
        if (tid == 0) {
            int target;
            if (pos == 0) {
                // the first chunk is smaller: NMS usually finishes inside it and sorting cost grows with the chunk
                target = min(kSortCap / 2, max(256, 2 * p.max_det));
            } else {
                // only survivors of the kept-list test reach the master: when few survive (clustered detections) a chunk
                // can span many more candidates than the master's list holds
                const long long want = 640ll * last_m / max(last_s, 1);
                target = static_cast<int>(min(static_cast<long long>(CL) * (kShareCap * 3 / 4), max(768ll, want)));
            }
            if (careful) target = min(target, (kSortCap * 3) / 4);
            const int base = start[from];
            const int rest = zoom_on ? start[kBuckets] - base : n_all - pos;  // exact in both cases
            int lo = from + 1, hi = kBuckets;
            if (rest <= kSortCap) {
                lo = kBuckets;  // everything that is left fits shared memory: one chunk, no estimate needed
            } else if (start[lo] - base <= target) {
                while (lo < hi) {
                    const int mid = (lo + hi + 1) >> 1;
                    if (start[mid] - base <= target) lo = mid; else hi = mid - 1;
                }
            }
            S_MISC[kMRunEnd] = lo;
        }
        __syncthreads();
        int to = S_MISC[kMRunEnd];
        int action;
        for (;;) {  // attempt; a failed attempt (overflow / rank cut) narrows [from, to) or sets `careful` and comes back here
            uint32_t lo_bits, hi_bits;
            int on_ovf = kActHalve;
            if (zoom_on) {
                lo_bits = z_lo + static_cast<uint32_t>(kBuckets - to) * kZoomSpan;
                hi_bits = z_lo + static_cast<uint32_t>(kBuckets - from) * kZoomSpan - 1u;
                if (to - from == 1) on_ovf = kActRadix;
            } else {
                const int b_lo = kBuckets - to, b_hi = kBuckets - 1 - from;
                lo_bits = bucket_floor_bits(b_lo);
                hi_bits = b_hi >= kBuckets - 1 ? 0xffffffffu : bucket_floor_bits(b_hi + 1) - 1u;
                // the clamped end buckets (everything below 2^-16, everything >= 1.0) have no fixed span to subdivide
                if (to - from == 1) on_ovf = (b_hi >= 1 && b_hi < kBuckets - 1) ? kActZoom : kActRadix;
            }
            action = attempt(lo_bits, hi_bits, on_ovf);
            if (action == kActHalve) {
                to = from + (to - from) / 2;
                continue;
            }
            if (action == kActCareful) {
                careful = true;
                if (m > (kSortCap * 3) / 4 && to - from > 1) to = from + max(1, (to - from) * ((kSortCap * 3) / 4) / m);
                continue;
            }
            if (action == kActOk && last_p1) {
                last_m = max(m, 1);
                last_s = max(s_all, 1);
            }
            break;
        }
        careful = false;
        if (action == kActZoom) {
            // ---- exact histogram of the bucket's sub-buckets (descending), every CTA of the cluster for itself ----
            zd = from;
            z_lo = bucket_floor_bits(kBuckets - 1 - from);
            for (int i = tid; i <= kBuckets; i += kNmsThreads) S_ZSTART[i] = 0;
            __syncthreads();
            for_each_candidate_in(p.st, tcount, tmaxv, score, z_lo, z_lo + (static_cast<uint32_t>(kBuckets) * kZoomSpan - 1u), 0, 1, p.st.tpi,
                                  TILE_LIST, &S_MISC[kMListN], [&](uint32_t, uint32_t bits) {
                                      atomicAdd(&S_ZSTART[kBuckets - 1 - static_cast<int>((bits - z_lo) / kZoomSpan)], 1);
                                  });
            __syncthreads();
            {
                constexpr int kPer = kBuckets / kNmsThreads;
                int loc[kPer], sum = 0;
#pragma unroll
                for (int i = 0; i < kPer; ++i) {
                    loc[i] = S_ZSTART[tid * kPer + i];
                    sum += loc[i];
                }
                int inc = sum;
#pragma unroll
                for (int dd = 1; dd < 32; dd <<= 1) {
                    const int v = __shfl_up_sync(0xffffffffu, inc, dd);
                    if (lane >= dd) inc += v;
                }
                if (lane == 31) S_MISC[kMWarp + warp] = inc;
                __syncthreads();
                int run = inc - sum;
                for (int w = 0; w < warp; ++w) run += S_MISC[kMWarp + w];
#pragma unroll
                for (int i = 0; i < kPer; ++i) {
                    S_ZSTART[tid * kPer + i] = run;
                    run += loc[i];
                }
                if (tid == kNmsThreads - 1) S_ZSTART[kBuckets] = run;
                __syncthreads();
            }
            zoom_on = true;
            z = 0;
            PROF_MARK(6);
            continue;
        }
        pos += m;
        if (zoom_on) z = to; else d = to;
    }
    // ---- publish (master CTA) ----
    __syncthreads();
    if (crank == 0) {
        for (int k = tid; k < kept; k += kNmsThreads) p.kept_slot[static_cast<int64_t>(b) * p.max_det + k] = KEPT_SLOT[k];
        if (tid == 0) p.counts[b] = kept;
    }
    if (p.stats && tid == 0) {  // accumulated with atomics (pair tests are counted per CTA): the caller zeroes the buffer
        unsigned long long *o = reinterpret_cast<unsigned long long *>(p.stats) + static_cast<int64_t>(b) * 4;
        if (crank == 0) {
            atomicAdd(o + 0, static_cast<unsigned long long>(st_walk));
            atomicAdd(o + 2, static_cast<unsigned long long>(st_sub));
            atomicAdd(o + 3, static_cast<unsigned long long>(st_coll));
        }
        atomicAdd(o + 1, static_cast<unsigned long long>(st_pairs));
    }
    if constexpr (CL > 1) {
        if (!cluster_ready) asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");  // no attempt ran (nothing to do)
        cluster.sync();  // no CTA may exit while a peer can still address its shared memory
        // every CTA read the score histogram in its prologue: leave it zeroed for the next call (workspace_clean)
        if (crank == 0)
            for (int i = tid; i < kBuckets; i += kNmsThreads) p.st.hist[static_cast<int64_t>(b) * kBuckets + i] = 0;
    }
    PROF_MARK(8);

#undef PLIST
#undef SURV
#undef SKEY
#undef S_BSTART
#undef S_MASK
#undef A_BOX
#undef C_BOX
#undef TILE_LIST
#undef S_A_AREA
#undef S_C_AREA
#undef S_A_SLOT
#undef S_C_SLOT
#undef S_DEAD
#undef S_DEADW
#undef S_MISC
#undef S_ZSTART
#undef KEPT_BOX
#undef KEPT_AREA
#undef KEPT_SLOT
}

// Pipeline gate (sarpost_pipeline_*): holds the stream — i.e. the next batch's decode kernel — until the NMS kernel of the
// previous batch is resident (its CTAs need whole SMs: 512 threads x 124 registers; once the decode kernel's CTAs have
// spread over every SM there is no room for them until it ends), or until `timeout_ns` has passed.  `expected` = CTAs of
// all NMS kernels launched through the pipeline so far (the counter only ever grows; the comparison is wrap-safe).
__global__ void k_gate(const unsigned int *counter, unsigned int expected, unsigned int timeout_ns) {
    if (threadIdx.x != 0) return;
    unsigned long long t0, t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (;;) {
        const unsigned int v = *reinterpret_cast<const volatile unsigned int *>(counter);
        if (static_cast<int>(v - expected) >= 0) break;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        if (t - t0 > timeout_ns) break;
        __nanosleep(100);
    }
}

// K5.  Launched as a programmatic dependent of the NMS kernel (cudaLaunchAttributeProgrammaticStreamSerialization): its
// CTAs are scheduled while that kernel is still finishing and wait here until its results are visible, so the launch
// latency is off the critical path (griddepcontrol.wait returns at once for a plain launch).
// (5 CTAs per SM: 16 images x 300 rows = 608 CTAs then fit one wave; at 4 per SM the last 16 CTAs ran as a second wave
// that doubled the kernel's tail)
__global__ void __launch_bounds__(kGatherWarps * 32, 5) k5_gather(const __grid_constant__ GatherParams p) {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const int b = blockIdx.y;
    const int r = blockIdx.x * kGatherWarps + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    const int n_rows = p.counts[b];
    if (p.n_peers > 0 && r == 0 && lane < p.n_peers) p.peer_counts[lane][p.peer_slot_offset + b] = n_rows;  // the counts travel with the rows
    if (p.n_peers > 0 && p.ex.nm == 0 && p.tail_cols == 0 && p.res_boxes == nullptr && (p.max_det & 1) == 0) {
        // Exchange of plain 6-column rows (sliced inference: tile detections / merged frames): the block's rows are
        // assembled in shared memory and go out as 16-byte stores of one contiguous run per peer.  Written row by row, every
        // 24-byte row is its own NVLink write to each of the peers, and the exchange becomes bound by the packet rate
        // (8 GPUs: 134 k packets per rank and step) instead of by anything the rows weigh.
        __shared__ __align__(16) float srow[kGatherWarps * 6];
        const int warp = threadIdx.x >> 5;
        if (r < n_rows) {
            const int64_t seg = static_cast<int64_t>(b) * p.st.cap;
            const uint32_t slot = p.kept_slot[static_cast<int64_t>(b) * p.max_det + r];
            const uint32_t key = p.st.key[seg + slot];
            float4 bx = p.st.box[seg + slot];
            float c4, c5;
            if (p.ex.mode == 2) {
                const float *src = p.ex.dets + (static_cast<int64_t>(b) * p.st.tpi * p.ex.dets_per_tile + key) * p.ex.row_len;
                c4 = src[4];
                c5 = src[5];
            } else {
                if (p.rescale) bx = rescale_box(bx, p.rescale + 5 * b);
                c4 = p.st.score[seg + slot];
                c5 = static_cast<float>(key % static_cast<uint32_t>(p.ex.nc));
            }
            if (lane < 6) srow[warp * 6 + lane] = lane == 0 ? bx.x : lane == 1 ? bx.y : lane == 2 ? bx.z : lane == 3 ? bx.w : lane == 4 ? c4 : c5;
            if (lane == 0 && p.kept_index) p.kept_index[static_cast<int64_t>(b) * p.max_det + r] = static_cast<int32_t>(key);
        }
        __syncthreads();
        const int r0 = blockIdx.x * kGatherWarps;
        const int nf = 6 * min(max(n_rows - r0, 0), kGatherWarps);  // valid floats of this block
        const int64_t base = ((static_cast<int64_t>(p.peer_slot_offset) + b) * p.max_det + r0) * 6;  // 16-byte aligned: max_det and r0 are even
        constexpr int kChunks = kGatherWarps * 6 / 4;
        for (int t = threadIdx.x; t < p.n_peers * kChunks; t += kGatherWarps * 32) {
            const int q = t / kChunks, f0 = (t - q * kChunks) * 4;
            float *d = p.peer_out[q] + base;
            if (f0 + 4 <= nf) *reinterpret_cast<float4 *>(d + f0) = *reinterpret_cast<const float4 *>(srow + f0);
            else
                for (int f = f0; f < nf; ++f) d[f] = srow[f];  // rows beyond counts[b] stay untouched
        }
        return;
    }
    if (r >= n_rows) return;
    const uint32_t slot = p.kept_slot[static_cast<int64_t>(b) * p.max_det + r];
    gather_row(p, b, r, lane, slot, p.st.key[static_cast<int64_t>(b) * p.st.cap + slot]);
}

// Extras (raw embedding, sigmoid state; head.py:247) of an explicit list of (image, anchor) pairs — used when
// the rows that need them are only known after a later stage (cross-tile merge).  One warp per pair.
struct GatherExtrasParams {
    const int32_t *image_index, *anchor_index;
    int32_t n, batch;
    float *out;  // [n, nm]
    ExtrasSrc ex;  // mode 1
};

__global__ void __launch_bounds__(kGatherWarps * 32) k_gather_extras(const __grid_constant__ GatherExtrasParams p) {
    const int q = blockIdx.x * kGatherWarps + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (q >= p.n) return;
    const int b = p.image_index[q], a = p.anchor_index[q];
    float *o = p.out + static_cast<int64_t>(q) * p.ex.nm;
    if (b < 0 || b >= p.batch || a < 0 || a >= p.ex.lvl_aoff[p.ex.nl]) {  // out-of-range pair: zero row
        for (int c = lane; c < p.ex.nm; c += 32) o[c] = 0.0f;
        return;
    }
    gather_extras_row(p.ex, b, static_cast<uint32_t>(a), lane, [&](int c, float v) { o[c] = v; });
}

}  // namespace sarpost
