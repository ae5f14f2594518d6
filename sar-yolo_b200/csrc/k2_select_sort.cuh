// k2_select_sort.cuh — stage K2: per-image top-`max_nms` selection and score-bucket partition.
//
// Replaces ops.py:285-286 (`x[x[:,4].argsort(descending=True)[:max_nms]]`) and prepares the stable
// descending order that torchvision.ops.nms establishes internally (ops.py:296).  Tie rule
// everywhere: equal scores keep source order (lower anchor, then lower class, first) = the rule of
// torchvision's stable sort; the reference's own unstable argsort at the max_nms cut is
// torch-version defined.
//
// NMS stops after max_det keeps (ops.py:297), so it usually consumes only the first few hundred
// candidates of the sorted order.  A full sort is therefore wasted work; instead:
//   K2 (this file)  one 4096-bin histogram over a monotone function of the score (16 octaves below
//                   1.0 at 8 mantissa bits), boundary bucket of the top max_nms, and an unordered
//                   scatter of the selected candidates into their bucket's range (descending buckets).
//   K4 (k4_nms.cuh) sorts bucket runs lazily, chunk by chunk, only as far as NMS actually walks.
// Inside a bucket order is restored by sorting on the composite (score bits, slot): the slot index
// of the tile-segmented store IS the source order, so the result is exact and deterministic even
// though the scatter uses shared-memory atomics.
//
// One CTA (1024 threads) per image.
#pragma once
#include "common.cuh"

namespace sarpost {

constexpr int kPartThreads = 1024;
constexpr int kPartWarps = kPartThreads / 32;
constexpr int kBuckets = 4096;
// bucket(bits) = clamp((bits >> 15) - kBucketBase, 0, 4095); 1.0f >> 15 = 32512 -> bucket 4095.
constexpr int kBucketShift = 15;
constexpr int kBucketBase = (0x3F800000 >> kBucketShift) - (kBuckets - 1);

__device__ __forceinline__ int score_bucket(uint32_t bits) {
    const int b = static_cast<int>(bits >> kBucketShift) - kBucketBase;
    return min(max(b, 0), kBuckets - 1);
}

struct PartParams {
    CandStore st;
    uint32_t *part_key;   // [B*cap] score bits, partitioned by descending bucket
    uint32_t *part_val;   // [B*cap] candidate slot
    int32_t *bstart;      // [B*(kBuckets+1)] start of descending bucket d = 4095 - bucket; [4096] = n_sel
    int32_t max_nms;
};

// block-wide exclusive scan over one int per thread (blockDim = kPartThreads)
__device__ __forceinline__ int block_excl_scan_1024(int v, int *warp_tot /*[kPartWarps+1]*/, int *total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int n = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += n;
    }
    if (lane == 31) warp_tot[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        const int t = warp_tot[lane];
        int ti = t;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int n = __shfl_up_sync(0xffffffffu, ti, d);
            if (lane >= d) ti += n;
        }
        warp_tot[lane] = ti - t;
        if (lane == 31) warp_tot[kPartWarps] = ti;
    }
    __syncthreads();
    const int res = warp_tot[warp] + inc - v;
    *total = warp_tot[kPartWarps];
    __syncthreads();
    return res;
}

// Visit every candidate of the image: fn(slot, score_bits).  Each warp owns a contiguous run of tiles;
// the tile counts of the run are fetched with one coalesced load and, when the tile region is exactly
// kTileA slots, the scores of kUnroll tiles are fetched with independent 128-bit loads before any of
// them is consumed (the loop is latency-bound: one CTA per image, data in L2).
template <class Fn>
__device__ __forceinline__ void for_each_candidate(const CandStore &st, const int32_t *tcount, const float *score,
                                                   const Fn &fn) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int tpw = (st.tpi + nwarps - 1) / nwarps;
    const int t_begin = warp * tpw, t_end = min(st.tpi, t_begin + tpw);
    if (st.region == kTileA) {
        constexpr int kUnroll = 8;
        for (int tb = t_begin; tb < t_end; tb += 32) {
            const int my_c = (tb + lane < t_end) ? tcount[tb + lane] : 0;
            const int nt = min(32, t_end - tb);
            for (int u0 = 0; u0 < nt; u0 += kUnroll) {
                uint4 v[kUnroll];
                int c[kUnroll];
#pragma unroll
                for (int u = 0; u < kUnroll; ++u) {
                    c[u] = __shfl_sync(0xffffffffu, my_c, (u0 + u) & 31);
                    if (u0 + u >= nt) c[u] = 0;
                    v[u] = make_uint4(0u, 0u, 0u, 0u);
                    if (lane * 4 < c[u])
                        v[u] = *reinterpret_cast<const uint4 *>(score + static_cast<size_t>(tb + u0 + u) * kTileA + lane * 4);
                }
#pragma unroll
                for (int u = 0; u < kUnroll; ++u) {
                    const int i0 = lane * 4;
                    const uint32_t slot = static_cast<uint32_t>(tb + u0 + u) * kTileA + i0;
                    if (i0 < c[u]) fn(slot, v[u].x);
                    if (i0 + 1 < c[u]) fn(slot + 1, v[u].y);
                    if (i0 + 2 < c[u]) fn(slot + 2, v[u].z);
                    if (i0 + 3 < c[u]) fn(slot + 3, v[u].w);
                }
            }
        }
    } else {
        for (int t = t_begin; t < t_end; ++t) {
            const int c = tcount[t];
            for (int i = lane; i < c; i += 32) {
                const uint32_t slot = static_cast<uint32_t>(t) * st.region + i;
                fn(slot, __float_as_uint(score[slot]));
            }
        }
    }
}

__global__ void __launch_bounds__(kPartThreads, 1) k2_select_partition(const __grid_constant__ PartParams p) {
    __shared__ int hist[kBuckets];
    __shared__ int warp_tot[kPartWarps + 1];
    __shared__ int s_db;
    const int b = blockIdx.x, tid = threadIdx.x;
    const int32_t *tcount = p.st.tile_count + static_cast<int64_t>(b) * p.st.tpi;
    const float *score = p.st.score + static_cast<int64_t>(b) * p.st.cap;
    uint32_t *pkey = p.part_key + static_cast<int64_t>(b) * p.st.cap;
    uint32_t *pval = p.part_val + static_cast<int64_t>(b) * p.st.cap;
    int32_t *bstart = p.bstart + static_cast<int64_t>(b) * (kBuckets + 1);

    for (int i = tid; i < kBuckets; i += kPartThreads) hist[i] = 0;
    if (tid == 0) s_db = kBuckets - 1;
    __syncthreads();
    for_each_candidate(p.st, tcount, score, [&](uint32_t, uint32_t bits) { atomicAdd(&hist[score_bucket(bits)], 1); });
    __syncthreads();

    // descending exclusive scan: thread t owns d = 4t..4t+3 (d = 4095 - bucket)
    int loc[4], sum = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        loc[i] = hist[kBuckets - 1 - (tid * 4 + i)];
        sum += loc[i];
    }
    int total;
    const int ex = block_excl_scan_1024(sum, warp_tot, &total);
    int run = ex;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        if (run < p.max_nms && run + loc[i] >= p.max_nms) s_db = tid * 4 + i;  // boundary bucket of the top max_nms
        run += loc[i];
    }
    __syncthreads();
    const int db = s_db;  // total < max_nms: stays 4095 (everything selected)
    run = ex;
    int n_sel_local = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int d = tid * 4 + i;
        const bool sel = d <= db;
        bstart[d] = sel ? run : -1;                       // fixed up below for d > db
        hist[kBuckets - 1 - d] = sel ? run : -1;          // becomes the scatter cursor; -1 = not selected
        if (d == db) n_sel_local = run + loc[i];
        run += loc[i];
    }
    if (tid * 4 <= db && db < tid * 4 + 4) warp_tot[0] = n_sel_local;  // exactly one thread owns db
    __syncthreads();
    const int n_sel = warp_tot[0];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int d = tid * 4 + i;
        if (d > db) bstart[d] = n_sel;
    }
    if (tid == 0) bstart[kBuckets] = n_sel;

    // unordered scatter of the selected candidates into their bucket range
    for_each_candidate(p.st, tcount, score, [&](uint32_t slot, uint32_t bits) {
        const int bk = score_bucket(bits);
        if (hist[bk] >= 0) {  // cursors only grow, so the sign test is race-free
            const int pos = atomicAdd(&hist[bk], 1);
            pkey[pos] = bits;
            pval[pos] = slot;
        }
    });
}

}  // namespace sarpost
