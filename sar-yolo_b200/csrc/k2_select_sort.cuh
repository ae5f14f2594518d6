// k2_select_sort.cuh — stage K2/K3 helpers: per-image top-`max_nms` selection and lazy stable sort.
//
// Replaces ops.py:285-286 (`x[x[:,4].argsort(descending=True)[:max_nms]]`) and prepares the stable
// descending order that torchvision.ops.nms establishes internally (ops.py:296).  Tie rule
// everywhere: equal scores keep source order (lower anchor, then lower class, first) = the rule of
// torchvision's stable sort; the reference's own unstable argsort at the max_nms cut is
// torch-version defined.
//
// NMS stops after max_det keeps (ops.py:297), so it usually consumes only the first few hundred
// candidates of the sorted order: a full sort (or even a full partition) is wasted work.  Instead
//   * K1 accumulates, per image, a 4096-bin histogram of score_bucket(score) while it emits the
//     candidates (one RED per candidate, common.cuh);
//   * the NMS kernel (k4_nms.cuh) scans that histogram from the top: the cumulative counts give the
//     rank range of every bucket, the boundary bucket of the top max_nms, and the next run of whole
//     buckets that fits shared memory; it then streams the image's candidate scores once, collects
//     the members of that bucket run, sorts them in shared memory on the 64-bit composite
//     (score bits, ~slot) and hands them to the suppression phases.  More runs are extracted only if
//     NMS has not reached max_det yet.
// The slot index of the tile-segmented store IS the source order, so sorting on the composite key is
// exact and deterministic although candidates are collected with atomics in arbitrary order.
//
// This header holds the device helpers; there is no separate K2 launch.
#pragma once
#include "common.cuh"

namespace sarpost {

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
                    // fn(slot, bits, valid) is called by all 32 lanes (it may vote)
                    fn(slot, v[u].x, i0 < c[u]);
                    fn(slot + 1, v[u].y, i0 + 1 < c[u]);
                    fn(slot + 2, v[u].z, i0 + 2 < c[u]);
                    fn(slot + 3, v[u].w, i0 + 3 < c[u]);
                }
            }
        }
    } else {
        for (int t = t_begin; t < t_end; ++t) {
            const int c = tcount[t];
            for (int i0 = 0; i0 < c; i0 += 32) {
                const int i = i0 + lane;
                const uint32_t slot = static_cast<uint32_t>(t) * st.region + i;
                const bool valid = i < c;
                fn(slot, valid ? __float_as_uint(score[slot]) : 0u, valid);
            }
        }
    }
}

}  // namespace sarpost
