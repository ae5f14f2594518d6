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
//             the run are skipped), and sorted in shared memory (bitonic network on the 64-bit composite
//             score|~slot = descending score, source order on ties); a single bucket larger than that
//             falls back to a stable LSD radix sort in global memory;
//   phase 1   a sub-chunk of <= 256 sorted candidates is tested against the kept list (smem);
//   phase 2   an upper-triangular suppression bitmask is built among the survivors (tiled IoU
//             bitmask, all threads);
//   sweep     one warp resolves the sub-chunk on the bitmask, 32 candidates per step when no two live
//             candidates of the group overlap, else one step per KEPT box; appends to the kept list.
// Small batches (B*CL <= #SMs) run a thread-block CLUSTER of CL CTAs per image: the selection scan, phase 1
// and the bitmask are split across the CTAs, the sort/compaction are replicated (same data, same result),
// the sweep stays on the master CTA; candidates, survivors' flags, mask rows and new kept boxes travel through
// distributed shared memory, with 3 cluster barriers per sub-chunk.
#pragma once
#include <cooperative_groups.h>
#include <cuda_fp16.h>

#include "common.cuh"
#include "k2_select_sort.cuh"

namespace sarpost {

#ifdef SARPOST_PHASE_PROF
__device__ unsigned long long g_phase[16];
#define PROF_MARK(i)                                                         \
    do {                                                                     \
        if (blockIdx.x == 0 && threadIdx.x == 0) {                           \
            const long long _t = clock64();                                  \
            g_phase[i] += static_cast<unsigned long long>(_t - prof_t);      \
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
    // mode 2
    const float *dets;
    int32_t dets_per_tile, row_len;
};

// element index (from the start of its tensor) of extras channel 0 of (image b, anchor), the element stride
// between channels, and the tensor base pointer
__device__ __forceinline__ int64_t extras_base(const ExtrasSrc &e, int b, uint32_t anchor, int64_t *stride, const void **base) {
    if (e.mode == 0) {
        *stride = e.anchors;
        *base = e.pred;
        return (static_cast<int64_t>(b) * e.channels + 4 + e.nc) * e.anchors + anchor;
    }
    int l = 0;
#pragma unroll
    for (int i = 1; i < kMaxLevels; ++i) l += (i < e.nl && anchor >= static_cast<uint32_t>(e.lvl_aoff[i])) ? 1 : 0;
    const int hw = e.lvl_hw[l];
    *stride = hw;
    *base = e.lvl_ptr[l];
    return (static_cast<int64_t>(b) * e.no + 4 * kRegMax + e.nc) * hw + (anchor - e.lvl_aoff[l]);
}

__device__ __forceinline__ float load_elem(const void *base, int64_t idx, bool is_half) {
    return is_half ? __half2float(__ldg(static_cast<const __half *>(base) + idx)) : __ldg(static_cast<const float *>(base) + idx);
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

// dynamic shared memory: kept_box[max_det] float4 | kept_area[max_det] | kept_slot[max_det] | (pad to 16) |
// plist[kSortCap] u64 (this CTA's share of a collected chunk, cluster variant)
__host__ __device__ inline size_t nms_kept_bytes(int max_det) { return (static_cast<size_t>(max_det) * 24 + 15) / 16 * 16; }
__host__ __device__ inline size_t nms_smem_bytes(int max_det) { return nms_kept_bytes(max_det) + kSortCap * 8; }

constexpr int kMaxNmsCluster = 4;

template <int CL>
__global__ void __launch_bounds__(kNmsThreads, 1) k4_nms(const __grid_constant__ NmsParams p) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    const int crank = CL > 1 ? static_cast<int>(cluster.block_rank()) : 0;
    auto cluster_sync = [&]() {
        if constexpr (CL > 1) cluster.sync(); else __syncthreads();
    };
    // All shared arrays are referenced through their symbols (never through pointers kept in
    // structs) so every access compiles to LDS/STS rather than a generic LD/ST.
    extern __shared__ float4 dyn_kept[];
    __shared__ int32_t s_bstart[kBuckets + 4];
    __shared__ uint32_t s_mask[kSub * kSubWords];
    __shared__ float s_a_area[kSub], s_c_area[kSub];
    __shared__ uint32_t s_a_slot[kSub], s_c_slot[kSub];
    __shared__ int32_t s_dead[kSub];  // phase-1 verdicts; written by every CTA of the cluster, cleared by its owner
    __shared__ int32_t s_misc[32];
    // union region: skey[kSortCap] u64 | a_box[kSub] | c_box[kSub]; the fallback radix sort aliases all of it
    __shared__ __align__(16) unsigned char s_uni[256 * (kNmsWarps + 1) * 4];
    static_assert(sizeof(s_uni) >= kSortCap * 8 + 2 * kSub * 16, "union region too small");
#define KEPT_BOX (dyn_kept)
#define KEPT_AREA (reinterpret_cast<float *>(dyn_kept + p.max_det))
#define KEPT_SLOT (reinterpret_cast<uint32_t *>(dyn_kept + p.max_det) + p.max_det)
#define SKEY (reinterpret_cast<unsigned long long *>(s_uni))
#define A_BOX (reinterpret_cast<float4 *>(s_uni + kSortCap * 8))
#define C_BOX (reinterpret_cast<float4 *>(s_uni + kSortCap * 8 + kSub * 16))
#define TILE_LIST (reinterpret_cast<uint32_t *>(s_uni + kSortCap * 8))  /* aliases A_BOX|C_BOX, idle while collecting */
#define PLIST (reinterpret_cast<unsigned long long *>(reinterpret_cast<unsigned char *>(dyn_kept) + nms_kept_bytes(p.max_det)))
    static_assert(kTileListCap * 4 <= 2 * kSub * 16, "tile list must fit the a_box|c_box region");

    const int b = blockIdx.x / CL, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t seg = static_cast<int64_t>(b) * p.st.cap;
    const int32_t *tcount = p.st.tile_count + static_cast<int64_t>(b) * p.st.tpi;
    const uint32_t *tmaxv = p.st.tile_max + static_cast<int64_t>(b) * p.st.tpi;
    const float *score = p.st.score + seg;
    const float band = fmaxf(p.thr * 2e-6f, 1e-37f);
#ifdef SARPOST_PHASE_PROF
    long long prof_t = clock64();
#endif
    if (tid < kSub) s_dead[tid] = 0;

    // ---- descending exclusive scan of the sampled score histogram: s_bstart[d] ~ estimated rank of the first
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
        if (lane == 31) s_misc[warp] = inc;
        if (lane == 0) s_misc[16 + warp] = tot;
        __syncthreads();
        int run = inc - sum;
        for (int w = 0; w < warp; ++w) run += s_misc[w];
#pragma unroll
        for (int i = 0; i < kPer; ++i) {
            s_bstart[tid * kPer + i] = run;
            run += loc[i];
        }
        if (tid == kNmsThreads - 1) s_bstart[kBuckets] = run;
        tot = 0;
#pragma unroll
        for (int w = 0; w < kNmsWarps; ++w) tot += s_misc[16 + w];
        __syncthreads();
        if (tid == 0) s_misc[19] = tot;
    }
    __syncthreads();
    if constexpr (CL > 1) {  // every CTA of the cluster has read the histogram: the master zeroes it
        cluster.sync();
        if (crank == 0)
            for (int i = tid; i < kBuckets; i += kNmsThreads) p.st.hist[static_cast<int64_t>(b) * kBuckets + i] = 0;
    }
    PROF_MARK(0);
    const int n_all = s_misc[19];
    const int n_limit = min(n_all, p.max_nms);  // ops.py:285-286: only the top max_nms ranks are eligible
    int kept = 0;
    int pos = 0, d = 0;  // pos = exact number of candidates consumed (all buckets above descending index d)
    long long st_walk = 0, st_pairs = 0;  // instrumentation (NmsParams::stats), uniform across threads
    int st_sub = 0, st_coll = 0;

    // Walks `cnt` sorted candidates whose slots are produced by slot_at(i), i in [0,cnt).
    auto process_sorted = [&](auto slot_at, int cnt) {
        int pdone = 0;
        while (pdone < cnt && kept < p.max_det) {
            const int need = p.max_det - kept;
            const int sub = min(cnt - pdone, min(kSub, max(64, 2 * need)));
            // ---- load sub-chunk: class-offset boxes (ops.py:289,295) ----
            if (tid < sub) {
                const uint32_t slot = slot_at(pdone + tid);
                const float4 bx = p.st.box[seg + slot];
                const float cls = p.cls_override ? p.cls_override[seg + slot]
                                                 : static_cast<float>(p.st.key[seg + slot] % static_cast<uint32_t>(p.nc));
                const float off = __fmul_rn(cls, p.max_wh);
                const float4 ob = make_float4(__fadd_rn(bx.x, off), __fadd_rn(bx.y, off), __fadd_rn(bx.z, off), __fadd_rn(bx.w, off));
                A_BOX[tid] = ob;
                s_a_area[tid] = box_area_rn(ob);
                s_a_slot[tid] = slot;
            }
            __syncthreads();
            PROF_MARK(3);
            // ---- phase 1: candidates x kept list, kNmsThreads/sub_p2 threads per candidate ----
            if (kept > 0) {
                // this CTA's share of the candidates (the whole sub-chunk when CL == 1)
                const int per = (sub + CL - 1) / CL, c_lo = crank * per, c_n = max(0, min(sub, c_lo + per) - c_lo);
                int sub_p2 = 64;
                while (sub_p2 < c_n) sub_p2 <<= 1;
                const int cand = c_lo + (tid & (sub_p2 - 1)), part = tid / sub_p2, nparts = kNmsThreads / sub_p2;
                if (cand < c_lo + c_n) {
                    const float4 ob = A_BOX[cand];
                    const float oa = s_a_area[cand];
                    // `dead` counts only verdicts the approximate quotient can be trusted with (not within the
                    // band around the threshold); if none of those fired and some pair was borderline, every pair
                    // is re-evaluated with the correctly rounded division
                    bool dead = false, any_border = false;
#pragma unroll 4
                    for (int k = part; k < kept; k += nparts) {
                        bool bd;
                        const bool gt = iou_gt_approx(KEPT_BOX[k], KEPT_AREA[k], ob, oa, p.thr, band, bd);
                        dead |= gt && !bd;
                        any_border |= bd;
                    }
                    if (any_border && !dead)  // rare: a quotient within a few ulp of the threshold
                        for (int k = part; k < kept; k += nparts) dead |= iou_gt(KEPT_BOX[k], KEPT_AREA[k], ob, oa, p.thr);
                    if (dead) {
                        s_dead[cand] = 1;
                        if constexpr (CL > 1)
                            for (int r2 = 0; r2 < CL; ++r2)
                                if (r2 != crank) *cluster.map_shared_rank(&s_dead[cand], r2) = 1;
                    }
                }
                cluster_sync();
            }
            PROF_MARK(4);
            // ---- ordered compaction of survivors (first kSub threads = 8 warps) ----
            bool alive = false;
            uint32_t bal = 0;
            if (tid < kSub) {
                alive = tid < sub && !s_dead[tid];
                // cleared by its owner here: peers write it again only after the two cluster barriers that follow
                s_dead[tid] = 0;
                bal = __ballot_sync(0xffffffffu, alive);
                if (lane == 0) s_misc[warp] = __popc(bal);
            }
            __syncthreads();
            int m = 0;
#pragma unroll
            for (int w = 0; w < kSubWords; ++w) m += s_misc[w];
            if (alive) {
                int base = 0;
                for (int w = 0; w < warp; ++w) base += s_misc[w];
                const int at = base + __popc(bal & lanemask_lt());
                C_BOX[at] = A_BOX[tid];
                s_c_area[at] = s_a_area[tid];
                s_c_slot[at] = s_a_slot[tid];
            }
            __syncthreads();
            PROF_MARK(5);
            // ---- phase 2: suppression bitmask among the m survivors (row r, bits j > r) ----
            // warp per row r, lane j tests the pair (r, 32w + j) for every word w >= r/32; a ballot packs the word
            const int words = (m + 31) >> 5;
            uint32_t *mask_out = s_mask;  // the master's bitmask (it runs the sweep)
            if constexpr (CL > 1) mask_out = cluster.map_shared_rank(s_mask, 0);
            for (int r = warp * CL + crank; r < m; r += kNmsWarps * CL) {
                const float4 rb = C_BOX[r];
                const float ra = s_c_area[r];
                for (int w = r >> 5; w < words; w += 2) {
                    const int j0 = (w << 5) + lane, j1 = j0 + 32;
                    const bool v0 = j0 > r && j0 < m, v1 = (w + 1 < words) && j1 < m;
                    bool bd0 = false, bd1 = false;
                    const int jc0 = v0 ? j0 : r, jc1 = v1 ? j1 : r;
                    const float4 cb0 = C_BOX[jc0], cb1 = C_BOX[jc1];
                    // disjoint boxes have inter = 0 -> IoU 0 (or NaN), never > thr >= 0: skip the arithmetic when no
                    // lane of the warp sees an overlap (the usual case for boxes spread over the image)
                    const bool ov0 = v0 && fminf(rb.z, cb0.z) > fmaxf(rb.x, cb0.x) && fminf(rb.w, cb0.w) > fmaxf(rb.y, cb0.y);
                    const bool ov1 = v1 && fminf(rb.z, cb1.z) > fmaxf(rb.x, cb1.x) && fminf(rb.w, cb1.w) > fmaxf(rb.y, cb1.y);
                    if (!__any_sync(0xffffffffu, ov0 || ov1)) {
                        if (lane == 0) {
                            mask_out[r * kSubWords + w] = 0u;
                            if (w + 1 < words) mask_out[r * kSubWords + w + 1] = 0u;
                        }
                        continue;
                    }
                    bool g0 = iou_gt_approx(rb, ra, cb0, s_c_area[jc0], p.thr, band, bd0);
                    bool g1 = iou_gt_approx(rb, ra, cb1, s_c_area[jc1], p.thr, band, bd1);
                    if (__any_sync(0xffffffffu, (bd0 && v0) || (bd1 && v1))) {  // rare: quotient within a few ulp of thr
                        if (bd0) g0 = iou_gt(rb, ra, cb0, s_c_area[jc0], p.thr);
                        if (bd1) g1 = iou_gt(rb, ra, cb1, s_c_area[jc1], p.thr);
                    }
                    const uint32_t bits0 = __ballot_sync(0xffffffffu, g0 && v0);
                    const uint32_t bits1 = __ballot_sync(0xffffffffu, g1 && v1);
                    if (lane == 0) {
                        mask_out[r * kSubWords + w] = bits0;
                        if (w + 1 < words) mask_out[r * kSubWords + w + 1] = bits1;
                    }
                }
            }
            cluster_sync();
            PROF_MARK(6);
            // ---- sweep (warp 0 of the master CTA) ----
            if (warp == 0 && crank == 0) {
                uint32_t km[kSubWords];  // kept bits of the groups resolved so far (warp-uniform)
#pragma unroll
                for (int g = 0; g < kSubWords; ++g) km[g] = 0u;
                int kl = kept;
                // column g of the bitmask (word g of the rows g2*32+lane, g2 <= g) is loaded one group ahead, before
                // the dependent chain of group g-1, so the sweep never waits on shared memory
                uint32_t cur[kSubWords], nxt[kSubWords];
                cur[0] = s_mask[lane * kSubWords];
#pragma unroll
                for (int g = 0; g < kSubWords; ++g) {
                    if (g + 1 < kSubWords) {
#pragma unroll
                        for (int g2 = 0; g2 <= g + 1; ++g2) nxt[g2] = s_mask[((g2 << 5) + lane) * kSubWords + g + 1];
                    }
                    if (g < words && kl < p.max_det) {
                        // removal word of group g = OR over the rows kept in earlier groups (lane = row)
                        uint32_t rem = 0u;
#pragma unroll
                        for (int g2 = 0; g2 < g; ++g2) rem |= ((km[g2] >> lane) & 1u) ? cur[g2] : 0u;
                        rem = __reduce_or_sync(0xffffffffu, rem);
                        const int nvalid = min(32, m - (g << 5));
                        const uint32_t live = ~rem & (nvalid == 32 ? 0xffffffffu : ((1u << nvalid) - 1u));
                        const int r = (g << 5) + lane;
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
                        if ((keptm >> lane) & 1u) {
                            const int idx = kl + __popc(keptm & lanemask_lt());
                            KEPT_BOX[idx] = C_BOX[r];
                            KEPT_AREA[idx] = s_c_area[r];
                            KEPT_SLOT[idx] = s_c_slot[r];
                        }
                        kl += c;
                        km[g] = keptm;
                    }
                    if (g + 1 < kSubWords) {
#pragma unroll
                        for (int g2 = 0; g2 <= g + 1; ++g2) cur[g2] = nxt[g2];
                    }
                }
                if (lane == 0) s_misc[16] = kl;
            }
            if constexpr (CL > 1) {
                if (crank == 0) {  // replicate the new kept boxes and the new count in the other CTAs of the cluster
                    __syncthreads();
                    const int kl = s_misc[16];
                    for (int i = kept + tid; i < kl; i += kNmsThreads)
                        for (int r2 = 1; r2 < CL; ++r2) {
                            *cluster.map_shared_rank(&KEPT_BOX[i], r2) = KEPT_BOX[i];
                            *cluster.map_shared_rank(&KEPT_AREA[i], r2) = KEPT_AREA[i];
                        }
                    if (tid == 0)
                        for (int r2 = 1; r2 < CL; ++r2) *cluster.map_shared_rank(&s_misc[16], r2) = kl;
                }
                cluster.sync();
            } else {
                __syncthreads();
            }
            st_pairs += static_cast<long long>(sub) * kept + static_cast<long long>(m) * (m - 1) / 2;
            ++st_sub;
            st_walk += sub;
            kept = s_misc[16];
            pdone += sub;
            __syncthreads();
            PROF_MARK(7);
        }
    };

    // this CTA's share of the image's tiles (everything when CL == 1)
    const int tiles_per_cta = (p.st.tpi + CL - 1) / CL;
    const int tile_lo = min(p.st.tpi, crank * tiles_per_cta), tile_hi = min(p.st.tpi, tile_lo + tiles_per_cta);
    // members of score buckets [b_lo, b_hi] among this CTA's tiles -> `list` (first kSortCap of them), exact
    // count -> s_misc[18]
    auto collect_smem = [&](int b_lo, int b_hi, unsigned long long *list) {
        const uint32_t lo_bits = bucket_floor_bits(b_lo);
        const uint32_t hi_bits = b_hi >= kBuckets - 1 ? 0xffffffffu : bucket_floor_bits(b_hi + 1) - 1u;
        for_each_candidate_in(p.st, tcount, tmaxv, score, lo_bits, hi_bits, tile_lo, tile_hi, TILE_LIST, &s_misc[20],
                              [&](uint32_t slot, uint32_t bits) {
                                  const int at = atomicAdd(&s_misc[18], 1);  // members are rare (a few hundred per image)
                                  if (at < kSortCap) list[at] = (static_cast<unsigned long long>(bits) << 32) | (0xffffffffu - slot);
                              });
    };

    while (d < kBuckets && pos < n_limit && kept < p.max_det) {
        // ---- next chunk: a run of whole buckets [d, d1) whose ESTIMATED population fits the target ----
        if (tid == 0) {
            // the first chunk is smaller: NMS usually finishes inside it and sorting cost grows with the chunk
            const int target = pos == 0 ? min(kSortCap / 2, max(256, 2 * p.max_det)) : (kSortCap * 3) / 4;
            const int base = s_bstart[d];
            int lo = d + 1, hi = kBuckets;
            if (n_all - pos <= kSortCap) {
                lo = kBuckets;  // everything that is left fits shared memory: one chunk, no estimate needed
            } else if (s_bstart[lo] - base <= target) {
                while (lo < hi) {
                    const int mid = (lo + hi + 1) >> 1;
                    if (s_bstart[mid] - base <= target) lo = mid; else hi = mid - 1;
                }
            }
            s_misc[17] = lo;
        }
        __syncthreads();
        int d1 = s_misc[17];
        int m = 0;
        int my_off = 0, my_cnt = 0;
        ++st_coll;
        for (;;) {  // collect; if the real population overflows shared memory, halve the bucket run and retry
            __syncthreads();
            if (tid == 0) s_misc[18] = 0;
            __syncthreads();
            if constexpr (CL == 1) {
                collect_smem(kBuckets - d1, kBuckets - 1 - d, SKEY);
                __syncthreads();
                m = s_misc[18];
            } else {
                collect_smem(kBuckets - d1, kBuckets - 1 - d, PLIST);
                __syncthreads();
                my_cnt = s_misc[18];
                if (tid < CL) *cluster.map_shared_rank(&s_misc[24 + crank], tid) = my_cnt;  // my count -> every CTA
                cluster.sync();
                m = 0;
                my_off = 0;
#pragma unroll
                for (int r2 = 0; r2 < CL; ++r2) {
                    my_off += r2 < crank ? s_misc[24 + r2] : 0;
                    m += s_misc[24 + r2];
                }
                cluster.sync();  // everyone has read the counts before a retry may overwrite them
            }
            if (m <= kSortCap || d1 - d == 1) break;
            d1 = d + (d1 - d) / 2;
        }
        if constexpr (CL > 1) {
            if (m <= kSortCap) {  // all-gather the shares: every CTA ends up with the same chunk in SKEY
                for (int i = tid; i < my_cnt; i += kNmsThreads) {
                    const unsigned long long v = PLIST[i];
#pragma unroll
                    for (int r2 = 0; r2 < CL; ++r2) *cluster.map_shared_rank(&SKEY[my_off + i], r2) = v;
                }
            }
            cluster.sync();
        }
        PROF_MARK(1);
        if (m <= kSortCap) {
            if (m > 0) {
                int lpw = 6;
                while ((1 << lpw) < m) ++lpw;
                const int pw = 1 << lpw;
                for (int i = m + tid; i < pw; i += kNmsThreads) SKEY[i] = 0ull;
                __syncthreads();
                bitonic_sort_desc(SKEY, lpw);
                PROF_MARK(2);
                process_sorted([&](int i) { return 0xffffffffu - static_cast<uint32_t>(SKEY[i]); }, min(m, n_limit - pos));
            }
        } else {
            // ---- a single bucket larger than shared memory: collect to global scratch, stable LSD radix sort
            //      on (score bits desc, slot asc), then stream it through the suppression phases.  With a
            //      cluster the master CTA does this alone (rare path); the others wait and read the result. ----
            uint32_t *ka = p.tmp_key_a + seg, *va = p.tmp_val_a + seg, *kb = p.tmp_key_b + seg, *vb = p.tmp_val_b + seg;
            const int b_one = kBuckets - 1 - d;
            const uint32_t *ik = ka, *iv = va;
            uint32_t *ok = kb, *ov = vb;
            int slot_bits = 1;
            while ((static_cast<int64_t>(1) << slot_bits) < p.st.cap) ++slot_bits;
            const int passes_slot = (slot_bits + 7) / 8;
            if (crank == 0) {
                __syncthreads();
                if (tid == 0) s_misc[18] = 0;
                __syncthreads();
                for_each_candidate_in(p.st, tcount, tmaxv, score, bucket_floor_bits(b_one),
                                      b_one >= kBuckets - 1 ? 0xffffffffu : bucket_floor_bits(b_one + 1) - 1u, 0, p.st.tpi,
                                      TILE_LIST, &s_misc[20], [&](uint32_t slot, uint32_t bits) {
                                          const int at = atomicAdd(&s_misc[18], 1);
                                          ka[at] = bits;
                                          va[at] = slot;
                                      });
                __syncthreads();
                int *cnt = reinterpret_cast<int *>(s_uni);
                int *wt = s_misc;
                for (int ps = 0; ps < passes_slot + 4; ++ps) {
                    const int sh = ps < passes_slot ? ps * 8 : (ps - passes_slot) * 8;
                    if (ps < passes_slot)
                        radix_pass_global(ik, iv, ok, ov, m, [sh](uint32_t, uint32_t v) { return (v >> sh) & 255u; }, cnt, wt);
                    else
                        radix_pass_global(ik, iv, ok, ov, m, [sh](uint32_t k, uint32_t) { return ((~k) >> sh) & 255u; }, cnt, wt);
                    const uint32_t *tk = ik, *tv = iv;
                    ik = ok; iv = ov;
                    ok = const_cast<uint32_t *>(tk); ov = const_cast<uint32_t *>(tv);
                }
                __threadfence();
            } else if ((passes_slot + 4) & 1) {  // same ping-pong parity as the master: where the sorted values end up
                iv = vb;
            }
            if constexpr (CL > 1) cluster.sync();
            const uint32_t *sorted = iv;
            const int lim = min(m, n_limit - pos);
            for (int piece = 0; piece < lim && kept < p.max_det; piece += kSortCap) {
                process_sorted([&](int i) { return sorted[piece + i]; }, min(kSortCap, lim - piece));
            }
        }
        pos += m;
        d = d1;
    }
    // ---- publish (master CTA) ----
    __syncthreads();
    if (crank == 0) {
        for (int k = tid; k < kept; k += kNmsThreads) p.kept_slot[static_cast<int64_t>(b) * p.max_det + k] = KEPT_SLOT[k];
        if (tid == 0) p.counts[b] = kept;
        if (tid == 0 && p.stats) {
            long long *o = p.stats + static_cast<int64_t>(b) * 4;
            o[0] = st_walk;
            o[1] = st_pairs;
            o[2] = st_sub;
            o[3] = st_coll;
        }
    }
    if constexpr (CL > 1) cluster.sync();  // no CTA may exit while a peer can still address its shared memory
    PROF_MARK(8);
#undef KEPT_BOX
#undef KEPT_AREA
#undef KEPT_SLOT
#undef SKEY
#undef A_BOX
#undef C_BOX
#undef TILE_LIST
#undef PLIST
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

__global__ void __launch_bounds__(kGatherWarps * 32) k5_gather(const __grid_constant__ GatherParams p) {
    const int b = blockIdx.y;
    const int r = blockIdx.x * kGatherWarps + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    const int n_rows = p.counts[b];
    const int row_len = 6 + p.ex.nm + p.tail_cols;
    // destinations of this row: the local output, or the same slot in every rank's buffer (peer stores)
    const int n_dst = p.n_peers > 0 ? p.n_peers : 1;
    const int64_t img = p.n_peers > 0 ? p.peer_slot_offset + b : b;
    if (p.n_peers > 0 && r == 0 && lane < p.n_peers) p.peer_counts[lane][img] = n_rows;  // the counts travel with the rows
    if (r >= n_rows) return;
    auto dst = [&](int q) { return (p.n_peers > 0 ? p.peer_out[q] : p.out) + (img * p.max_det + r) * row_len; };
    const int64_t seg = static_cast<int64_t>(b) * p.st.cap;
    const uint32_t slot = p.kept_slot[static_cast<int64_t>(b) * p.max_det + r];
    const uint32_t key = p.st.key[seg + slot];
    if (p.ex.mode == 2) {
        const float *src = p.ex.dets + (static_cast<int64_t>(b) * p.st.tpi * p.ex.dets_per_tile + key) * p.ex.row_len;
        const float4 bx = p.st.box[seg + slot];
        for (int c = lane; c < 6 + p.ex.nm; c += 32) {
            const float v = c == 0 ? bx.x : c == 1 ? bx.y : c == 2 ? bx.z : c == 3 ? bx.w : src[c];
            for (int q = 0; q < n_dst; ++q) dst(q)[c] = v;
        }
        if (lane == 0 && p.kept_index) p.kept_index[static_cast<int64_t>(b) * p.max_det + r] = static_cast<int32_t>(key);
        return;
    }
    const uint32_t anchor = key / static_cast<uint32_t>(p.ex.nc), cls = key - anchor * static_cast<uint32_t>(p.ex.nc);
    if (lane == 0) {
        float4 bx = p.st.box[seg + slot];
        if (p.rescale) bx = rescale_box(bx, p.rescale + 5 * b);
        const float sc = p.st.score[seg + slot];
        for (int q = 0; q < n_dst; ++q) {
            float *o = dst(q);
            o[0] = bx.x;
            o[1] = bx.y;
            o[2] = bx.z;
            o[3] = bx.w;
            o[4] = sc;
            o[5] = static_cast<float>(cls);
        }
        if (p.kept_index) p.kept_index[static_cast<int64_t>(b) * p.max_det + r] = static_cast<int32_t>(key);
    }
    if (p.ex.nm > 0 && p.ex.mode == 0 && anchor >= static_cast<uint32_t>(p.ex.anchors)) {
        for (int c = lane; c < p.ex.nm; c += 32)
            for (int q = 0; q < n_dst; ++q) dst(q)[6 + c] = 0.0f;  // apriori label row: no extras (ops.py:258)
    } else if (p.ex.nm > 0) {
        int64_t stride;
        const void *base;
        const int64_t at = extras_base(p.ex, b, anchor, &stride, &base);
        const bool hf = p.ex.is_half != 0;
        const int n_raw = p.ex.mode == 0 ? p.ex.nm : p.ex.n_extra_raw;  // a decoded prediction is copied verbatim
        for (int c = lane; c < p.ex.nm; c += 32) {
            float v = load_elem(base, at + c * stride, hf);
            v = c < n_raw ? v : sigmoid_rn(v);
            for (int q = 0; q < n_dst; ++q) dst(q)[6 + c] = v;
        }
    }
}

// Extras (raw embedding, sigmoid state; head.py:247) of an explicit list of (image, anchor) pairs — used when
// the rows that need them are only known after a later stage (cross-tile merge).  One warp per pair.
struct GatherExtrasParams {
    const int32_t *image_index, *anchor_index;
    int32_t n;
    float *out;  // [n, nm]
    int32_t nl, no, nc, batch, n_extra_raw, nm;
    int32_t lvl_aoff[kMaxLevels + 1];
    int32_t lvl_hw[kMaxLevels];
    const void *lvl_ptr[kMaxLevels];
    int32_t is_half;
};

__global__ void __launch_bounds__(kGatherWarps * 32) k_gather_extras(const __grid_constant__ GatherExtrasParams p) {
    const int q = blockIdx.x * kGatherWarps + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (q >= p.n) return;
    const int b = p.image_index[q], a = p.anchor_index[q];
    float *o = p.out + static_cast<int64_t>(q) * p.nm;
    if (b < 0 || b >= p.batch || a < 0 || a >= p.lvl_aoff[p.nl]) {  // out-of-range pair: zero row
        for (int c = lane; c < p.nm; c += 32) o[c] = 0.0f;
        return;
    }
    int l = 0;
#pragma unroll
    for (int i = 1; i < kMaxLevels; ++i) l += (i < p.nl && a >= p.lvl_aoff[i]) ? 1 : 0;
    const int hw = p.lvl_hw[l];
    const int64_t at = (static_cast<int64_t>(b) * p.no + 4 * kRegMax + p.nc) * hw + (a - p.lvl_aoff[l]);
    for (int c = lane; c < p.nm; c += 32) {
        const float v = load_elem(p.lvl_ptr[l], at + static_cast<int64_t>(c) * hw, p.is_half != 0);
        o[c] = c < p.n_extra_raw ? v : sigmoid_rn(v);
    }
}

}  // namespace sarpost
