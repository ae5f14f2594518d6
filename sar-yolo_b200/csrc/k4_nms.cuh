// k4_nms.cuh — stage K4: class-offset greedy NMS with early exit, and K5: max_det gather.
//
// Replaces ops.py:289-297 (`c = cls*max_wh; boxes = box + c; i = torchvision.ops.nms(...)[:max_det]`)
// and ops.py:311 (`output[xi] = x[i]`).  Bit-exact to torchvision's CPU kernel (oracle/nms_greedy.c).
//
// Greedy NMS only ever compares a candidate with boxes that were KEPT before it, and the result is
// cut at max_det (ops.py:297), so at most n*max_det pair tests are needed instead of n^2/2.
// One CTA per image walks the sorted candidates in chunks of kChunk:
//   phase 1  every candidate of the chunk is tested against the kept list (<= max_det boxes, smem);
//   phase 2  the survivors are compacted and an upper-triangular 64-bit suppression mask is built
//            among them (tiled IoU bitmask, all threads);
//   sweep    one warp resolves the chunk sequentially on the bitmask — one step per KEPT box, not
//            per candidate — appending to the kept list; stops at max_det.
#pragma once
#include "common.cuh"

namespace sarpost {

constexpr int kNmsThreads = 512;
constexpr int kChunk = 512;
constexpr int kChunkWords = kChunk / 64;

struct NmsParams {
    CandStore st;
    const uint32_t *sorted;   // [B*cap] candidate slots in descending-score order
    const int32_t *n_sorted;  // [B]
    uint32_t *kept_slot;      // [B*max_det]
    int32_t *counts;          // [B]
    int32_t max_det;
    int32_t nc;               // key % nc = class (merge: class comes from cls_override)
    const float *cls_override;// merge path: class id per slot (float) or nullptr
    float max_wh;             // 0 when agnostic
    float thr;                // largest float <= iou_thres
};

// dynamic smem layout: kept_box[max_det] float4 | kept_area[max_det] | kept_slot[max_det]
__global__ void __launch_bounds__(kNmsThreads, 1) k4_nms(const __grid_constant__ NmsParams p) {
    extern __shared__ __align__(16) unsigned char nms_smem[];
    float4 *kept_box = reinterpret_cast<float4 *>(nms_smem);
    float *kept_area = reinterpret_cast<float *>(kept_box + p.max_det);
    uint32_t *kept_slot = reinterpret_cast<uint32_t *>(kept_area + p.max_det);
    __shared__ float4 ch_box[kChunk];
    __shared__ float ch_area[kChunk];
    __shared__ uint32_t ch_slot[kChunk];
    __shared__ unsigned long long mask[kChunk * kChunkWords];
    __shared__ int warp_tot[kNmsThreads / 32];
    __shared__ int s_kept, s_m;

    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = p.n_sorted[b];
    const int64_t seg = static_cast<int64_t>(b) * p.st.cap;
    const uint32_t *sorted = p.sorted + seg;
    int kept = 0;

    for (int start = 0; start < n && kept < p.max_det; start += kChunk) {
        // ---- load chunk, build class-offset boxes (ops.py:289,295), phase 1 ----
        const int i = start + tid;
        bool alive = i < n;
        float4 ob = make_float4(0.f, 0.f, 0.f, 0.f);
        float area = 0.f;
        uint32_t slot = 0;
        if (alive) {
            slot = sorted[i];
            const float4 bx = p.st.box[seg + slot];
            const float cls = p.cls_override ? p.cls_override[seg + slot]
                                             : static_cast<float>(p.st.key[seg + slot] % static_cast<uint32_t>(p.nc));
            const float off = __fmul_rn(cls, p.max_wh);
            ob = make_float4(__fadd_rn(bx.x, off), __fadd_rn(bx.y, off), __fadd_rn(bx.z, off), __fadd_rn(bx.w, off));
            area = box_area_rn(ob);
        }
        for (int k = 0; k < kept; ++k) {
            if (!__any_sync(0xffffffffu, alive)) break;
            if (alive && iou_gt(kept_box[k], kept_area[k], ob, area, p.thr)) alive = false;
        }
        // ---- ordered compaction of survivors ----
        const uint32_t bal = __ballot_sync(0xffffffffu, alive);
        if (lane == 0) warp_tot[warp] = __popc(bal);
        __syncthreads();
        int base = 0, m = 0;
#pragma unroll
        for (int w = 0; w < kNmsThreads / 32; ++w) {
            const int c = warp_tot[w];
            base += (w < warp) ? c : 0;
            m += c;
        }
        if (alive) {
            const int pos = base + __popc(bal & lanemask_lt());
            ch_box[pos] = ob;
            ch_area[pos] = area;
            ch_slot[pos] = slot;
        }
        __syncthreads();
        // ---- phase 2: suppression bitmask among the m survivors (row i, bits j > i) ----
        const int words = (m + 63) >> 6;
        for (int item = tid; item < m * words; item += kNmsThreads) {
            const int r = item / words, w = item - r * words;
            if (w < (r >> 6)) continue;
            const float4 rb = ch_box[r];
            const float ra = ch_area[r];
            unsigned long long bits = 0ull;
            const int j0 = w << 6;
            const int jn = min(64, m - j0);
            for (int jj = max(0, r + 1 - j0); jj < jn; ++jj)
                if (iou_gt(rb, ra, ch_box[j0 + jj], ch_area[j0 + jj], p.thr)) bits |= 1ull << jj;
            mask[r * kChunkWords + w] = bits;
        }
        __syncthreads();
        // ---- sweep (warp 0): lane w accumulates removal word w ----
        if (warp == 0) {
            unsigned long long remv = 0ull;
            for (int g = 0; g < words && kept < p.max_det; ++g) {
                const unsigned long long cur = __shfl_sync(0xffffffffu, remv, g);
                const int nvalid = min(64, m - (g << 6));
                unsigned long long live = ~cur & (nvalid == 64 ? ~0ull : ((1ull << nvalid) - 1ull));
                while (live) {
                    const int j = __ffsll(static_cast<long long>(live)) - 1;
                    const int r = (g << 6) + j;
                    if (lane == 0) {
                        kept_box[kept] = ch_box[r];
                        kept_area[kept] = ch_area[r];
                        kept_slot[kept] = ch_slot[r];
                    }
                    ++kept;
                    if (kept >= p.max_det) break;
                    const unsigned long long row = (lane >= g && lane < words) ? mask[r * kChunkWords + lane] : 0ull;
                    remv |= row;
                    const unsigned long long diag = __shfl_sync(0xffffffffu, row, g);
                    live &= ~diag;
                    live &= live - 1ull;  // drop bit j itself (lowest set bit; diag never has bits <= j)
                }
            }
            if (lane == 0) s_kept = kept;
        }
        __syncthreads();
        kept = s_kept;
    }
    // ---- publish ----
    for (int k = tid; k < kept; k += kNmsThreads) p.kept_slot[static_cast<int64_t>(b) * p.max_det + k] = kept_slot[k];
    if (tid == 0) p.counts[b] = kept;
}

// ---------------------------------------------------------------------------------------------
// K5: one warp per output row.  Row = x1,y1,x2,y2,conf,cls,extras (ops.py:272/275, :311).
// Extras come from the decoded prediction (nms_decoded), from the raw level tensors (fused: raw
// embedding, sigmoid state — head.py:247), or from the source detection rows (merge).
// ---------------------------------------------------------------------------------------------
struct GatherParams {
    CandStore st;
    const uint32_t *kept_slot;  // [B*max_det]
    const int32_t *counts;      // [B]
    float *out;                 // [B, max_det, 6+nm]
    int32_t *kept_index;        // [B*max_det] or nullptr
    int32_t max_det, nc, nm;
    int32_t mode;               // 0 decoded, 1 fused, 2 merge, 3 none (boxes only)
    // mode 0
    const float *pred;
    int32_t channels;
    int64_t anchors;
    // mode 1
    int32_t nl, no, n_extra_raw;
    int32_t lvl_aoff[kMaxLevels + 1];
    int32_t lvl_hw[kMaxLevels];
    const float *lvl_ptr[kMaxLevels];
    // mode 2
    const float *dets;
    int32_t dets_per_tile, row_len;
};

constexpr int kGatherWarps = 8;

__global__ void __launch_bounds__(kGatherWarps * 32) k5_gather(const __grid_constant__ GatherParams p) {
    const int b = blockIdx.y;
    const int r = blockIdx.x * kGatherWarps + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (r >= p.counts[b]) return;
    const int64_t seg = static_cast<int64_t>(b) * p.st.cap;
    const uint32_t slot = p.kept_slot[static_cast<int64_t>(b) * p.max_det + r];
    const uint32_t key = p.st.key[seg + slot];
    const int row_len = 6 + p.nm;
    float *o = p.out + (static_cast<int64_t>(b) * p.max_det + r) * row_len;
    if (p.mode == 2) {
        const float *src = p.dets + (static_cast<int64_t>(b) * p.st.tpi * p.dets_per_tile + key) * p.row_len;
        const float4 bx = p.st.box[seg + slot];
        for (int c = lane; c < row_len; c += 32)
            o[c] = c == 0 ? bx.x : c == 1 ? bx.y : c == 2 ? bx.z : c == 3 ? bx.w : src[c];
        if (lane == 0 && p.kept_index) p.kept_index[static_cast<int64_t>(b) * p.max_det + r] = static_cast<int32_t>(key);
        return;
    }
    const uint32_t anchor = key / static_cast<uint32_t>(p.nc), cls = key - anchor * static_cast<uint32_t>(p.nc);
    if (lane == 0) {
        const float4 bx = p.st.box[seg + slot];
        o[0] = bx.x;
        o[1] = bx.y;
        o[2] = bx.z;
        o[3] = bx.w;
        o[4] = p.st.score[seg + slot];
        o[5] = static_cast<float>(cls);
        if (p.kept_index) p.kept_index[static_cast<int64_t>(b) * p.max_det + r] = static_cast<int32_t>(key);
    }
    if (p.mode == 0) {
        const float *src = p.pred + (static_cast<int64_t>(b) * p.channels + 4 + p.nc) * p.anchors + anchor;
        for (int c = lane; c < p.nm; c += 32) o[6 + c] = __ldg(src + static_cast<int64_t>(c) * p.anchors);
    } else if (p.mode == 1) {
        int l = 0;
#pragma unroll
        for (int i = 1; i < kMaxLevels; ++i) l += (i < p.nl && anchor >= static_cast<uint32_t>(p.lvl_aoff[i])) ? 1 : 0;
        const int hw = p.lvl_hw[l];
        const float *src = p.lvl_ptr[l] + (static_cast<int64_t>(b) * p.no + 4 * kRegMax + p.nc) * hw + (anchor - p.lvl_aoff[l]);
        for (int c = lane; c < p.nm; c += 32) {
            const float v = __ldg(src + static_cast<int64_t>(c) * hw);
            o[6 + c] = c < p.n_extra_raw ? v : sigmoid_rn(v);
        }
    }
}

}  // namespace sarpost
