// k2_select_sort.cuh — stage K2/K3: per-image top-`max_nms` selection and stable descending sort.
//
// Replaces ops.py:285-286 (`x[x[:,4].argsort(descending=True)[:max_nms]]`) and the stable
// descending score sort inside torchvision.ops.nms (ops.py:296).  Tie rule everywhere: equal scores
// keep source order (lower anchor, then lower class, first) — the rule of torchvision's stable sort;
// the reference's own unstable argsort at the max_nms cut is torch-version defined.
//
// One CTA (1024 threads) per image.
//   1. exclusive scan of the image's tile counts -> tile_off (source rank of every candidate), n.
//   2. n <= kSmallN : rank-by-counting in shared memory (stable by construction).
//      otherwise    : if n > max_nms, one 4096-bin histogram over the high score bits picks the
//                     boundary bucket (everything in a higher bucket is certainly in the top max_nms,
//                     everything lower certainly is not); then a 4-pass LSD radix sort (8-bit digits)
//                     of the selected candidates.  Pass 1 reads the tile-segmented store in source
//                     order, so the selection needs no separate compaction pass; the boundary bucket
//                     is sorted whole and the list is cut at max_nms afterwards (exact, stable).
// Output: sorted[b][0..n_sorted[b]) = candidate slots (index inside the image's segment).
#pragma once
#include "common.cuh"

namespace sarpost {

constexpr int kSortThreads = 1024;
constexpr int kSortWarps = kSortThreads / 32;
constexpr int kSmallN = 2048;
constexpr int kHistBits = 12;  // bucket = score_bits >> 19 (sign bit is 0 for every candidate)
constexpr int kHistShift = 31 - kHistBits;

struct SortParams {
    CandStore st;
    uint32_t *key_a, *val_a, *key_b, *val_b;  // [B*cap] ping-pong; result ends in *_a
    int32_t *tile_off;                         // [B*(tpi+1)]
    int32_t *n_sorted;                         // [B]
    int32_t max_nms;
};

__device__ __forceinline__ int cnt_index(int digit, int warp) { return digit * (kSortWarps + 1) + warp; }

// block-wide exclusive scan helper over one int per thread; returns exclusive prefix, total via *total
__device__ __forceinline__ int block_excl_scan(int v, int *warp_tot /*[kSortWarps+1]*/, int *total) {
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
        int t = warp_tot[lane];
        int ti = t;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int n = __shfl_up_sync(0xffffffffu, ti, d);
            if (lane >= d) ti += n;
        }
        warp_tot[lane] = ti - t;
        if (lane == 31) warp_tot[kSortWarps] = ti;
    }
    __syncthreads();
    const int res = warp_tot[warp] + inc - v;
    *total = warp_tot[kSortWarps];
    __syncthreads();
    return res;
}

// One stable LSD pass.  Source elements are produced by `load(w, it, lane, key, val)` which must
// enumerate warp w's share of the input in order (it = 0,1,...; returns false when exhausted for the
// whole warp); each warp's share precedes the next warp's share in input order.
template <class Loader>
__device__ __forceinline__ int radix_pass(const Loader &load, int shift, uint32_t *out_key, uint32_t *out_val,
                                          int *cnt /*[256*(kSortWarps+1)]*/, int *warp_tot) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 256 * (kSortWarps + 1); i += kSortThreads) cnt[i] = 0;
    __syncthreads();
    // (a) per-warp digit histogram
    for (int it = 0;; ++it) {
        uint32_t key, val;
        bool ok;
        if (!load(warp, it, lane, key, val, ok)) break;
        const uint32_t d = ok ? ((key >> shift) & 255u) : 256u;
        const uint32_t peers = __match_any_sync(0xffffffffu, d);
        if (ok && (peers & lanemask_lt()) == 0) cnt[cnt_index(d, warp)] += __popc(peers);
        __syncwarp();
    }
    __syncthreads();
    // (b) exclusive scan in (digit-major, warp-minor) order: 8192 entries, 8 per thread
    int local[8];
    int sum = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int e = threadIdx.x * 8 + i;
        local[i] = cnt[cnt_index(e >> 5, e & 31)];
        sum += local[i];
    }
    int total;
    int run = block_excl_scan(sum, warp_tot, &total);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int e = threadIdx.x * 8 + i;
        cnt[cnt_index(e >> 5, e & 31)] = run;
        run += local[i];
    }
    __syncthreads();
    // (c) stable scatter
    for (int it = 0;; ++it) {
        uint32_t key, val;
        bool ok;
        if (!load(warp, it, lane, key, val, ok)) break;
        const uint32_t d = ok ? ((key >> shift) & 255u) : 256u;
        const uint32_t peers = __match_any_sync(0xffffffffu, d);
        int base = 0;
        if (ok) base = cnt[cnt_index(d, warp)];
        __syncwarp();
        if (ok) {
            const int rank = __popc(peers & lanemask_lt());
            if (rank == 0) cnt[cnt_index(d, warp)] = base + __popc(peers);
            out_key[base + rank] = key;
            out_val[base + rank] = val;
        }
        __syncwarp();
    }
    __syncthreads();
    return total;
}

__global__ void __launch_bounds__(kSortThreads, 1) k2_select_sort(const __grid_constant__ SortParams p) {
    __shared__ int cnt[256 * (kSortWarps + 1)];  // 33,792 B; also the 4096-bin histogram / small-n scratch
    __shared__ int warp_tot[kSortWarps + 1];
    __shared__ int s_bucket;
    const int b = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tpi = p.st.tpi;
    const int32_t *tcount = p.st.tile_count + static_cast<int64_t>(b) * tpi;
    int32_t *toff = p.tile_off + static_cast<int64_t>(b) * (tpi + 1);
    const float *score = p.st.score + static_cast<int64_t>(b) * p.st.cap;
    uint32_t *key_a = p.key_a + static_cast<int64_t>(b) * p.st.cap, *val_a = p.val_a + static_cast<int64_t>(b) * p.st.cap;
    uint32_t *key_b = p.key_b + static_cast<int64_t>(b) * p.st.cap, *val_b = p.val_b + static_cast<int64_t>(b) * p.st.cap;

    // 1. tile offsets
    int carry = 0;
    for (int t0 = 0; t0 < tpi; t0 += kSortThreads) {
        const int t = t0 + tid;
        const int c = t < tpi ? tcount[t] : 0;
        int total;
        const int ex = block_excl_scan(c, warp_tot, &total);
        if (t < tpi) toff[t] = carry + ex;
        carry += total;
    }
    const int n = carry;
    if (tid == 0) toff[tpi] = n;
    if (n == 0) {
        if (tid == 0) p.n_sorted[b] = 0;
        return;
    }
    __syncthreads();  // toff visible to the block (global writes by this block, read below)

    // 2a. small n: rank by counting
    if (n <= kSmallN) {
        float *s_sc = reinterpret_cast<float *>(cnt);
        uint32_t *s_slot = reinterpret_cast<uint32_t *>(cnt) + kSmallN;
        for (int t = warp; t < tpi; t += kSortWarps) {
            const int c = tcount[t], o = toff[t];
            for (int i = lane; i < c; i += 32) {
                const uint32_t slot = static_cast<uint32_t>(t) * p.st.region + i;
                s_sc[o + i] = score[slot];
                s_slot[o + i] = slot;
            }
        }
        __syncthreads();
        for (int e = tid; e < n; e += kSortThreads) {
            const float se = s_sc[e];
            int rank = 0;
            for (int f = 0; f < n; ++f) {
                const float sf = s_sc[f];
                rank += (sf > se) || (sf == se && f < e);
            }
            if (rank < p.max_nms) val_a[rank] = s_slot[e];
        }
        if (tid == 0) p.n_sorted[b] = n < p.max_nms ? n : p.max_nms;
        return;
    }

    // 2b. boundary bucket of the top-max_nms (only when something must be dropped)
    uint32_t min_bucket = 0;
    if (n > p.max_nms) {
        constexpr int kBins = 1 << kHistBits;
        for (int i = tid; i < kBins; i += kSortThreads) cnt[i] = 0;
        __syncthreads();
        for (int t = warp; t < tpi; t += kSortWarps) {
            const int c = tcount[t];
            for (int i = lane; i < c; i += 32) {
                const uint32_t bits = __float_as_uint(score[static_cast<int64_t>(t) * p.st.region + i]);
                atomicAdd(&cnt[min(bits >> kHistShift, static_cast<uint32_t>(kBins - 1))], 1);
            }
        }
        __syncthreads();
        // thread t owns buckets [kBins-1-4t-3, kBins-1-4t] taken from the top down
        int loc[4];
        int sum = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            loc[i] = cnt[kBins - 1 - (tid * 4 + i)];
            sum += loc[i];
        }
        int total;
        int run = block_excl_scan(sum, warp_tot, &total);
        if (run < p.max_nms && run + sum >= p.max_nms) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                if (run < p.max_nms && run + loc[i] >= p.max_nms) s_bucket = kBins - 1 - (tid * 4 + i);
                run += loc[i];
            }
        }
        __syncthreads();
        min_bucket = static_cast<uint32_t>(s_bucket);
        __syncthreads();
    }

    // 3. LSD radix sort, ascending on ~score_bits (= descending score), stable.
    const int tpw = (tpi + kSortWarps - 1) / kSortWarps;  // contiguous tiles per warp
    const int groups_per_tile = (p.st.region + 31) / 32;
    auto load_tiles = [&](int w, int it, int ln, uint32_t &key, uint32_t &val, bool &ok) -> bool {
        const int t = w * tpw + it / groups_per_tile;
        if (it >= tpw * groups_per_tile || t >= tpi) return false;
        const int i = (it % groups_per_tile) * 32 + ln;
        ok = i < tcount[t];
        if (ok) {
            const uint32_t slot = static_cast<uint32_t>(t) * p.st.region + i;
            const uint32_t bits = __float_as_uint(score[slot]);
            key = ~bits;
            val = slot;
            ok = min(bits >> kHistShift, (1u << kHistBits) - 1u) >= min_bucket;
        }
        return true;
    };
    const int n_sel = radix_pass(load_tiles, 0, key_b, val_b, cnt, warp_tot);
    const int per_warp = (n_sel + kSortWarps - 1) / kSortWarps;
    const int iters = (per_warp + 31) / 32;
    auto make_linear = [&](const uint32_t *k_in, const uint32_t *v_in) {
        return [=](int w, int it, int ln, uint32_t &key, uint32_t &val, bool &ok) -> bool {
            if (it >= iters) return false;
            const int o = it * 32 + ln;
            const int i = w * per_warp + o;
            ok = o < per_warp && i < n_sel;
            if (ok) {
                key = k_in[i];
                val = v_in[i];
            }
            return true;
        };
    };
    radix_pass(make_linear(key_b, val_b), 8, key_a, val_a, cnt, warp_tot);
    radix_pass(make_linear(key_a, val_a), 16, key_b, val_b, cnt, warp_tot);
    radix_pass(make_linear(key_b, val_b), 24, key_a, val_a, cnt, warp_tot);
    if (tid == 0) p.n_sorted[b] = n_sel < p.max_nms ? n_sel : p.max_nms;
}

}  // namespace sarpost
