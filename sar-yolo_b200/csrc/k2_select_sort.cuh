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

// Visit every candidate of the image whose score bit pattern lies in [lo_bits, hi_bits]: fn(slot, bits).
// The scan is instruction-bound (one CTA streams every candidate score of its image out of L2), so the
// common case — no member among a lane's four scores — costs one 128-bit load, four range tests and a
// single branch.  Each warp owns a contiguous run of tiles; the tile counts of the run are fetched with
// one coalesced load and, when the tile region is exactly kTileA slots, kUnroll tiles are in flight.
// Tiles whose best score (tile_max, written by K1) is below the range are skipped without touching them.
template <class Fn>
__device__ __forceinline__ void for_each_candidate_in(const CandStore &st, const int32_t *tcount, const uint32_t *tmax,
                                                      const float *score,
                                                      uint32_t lo_bits, uint32_t hi_bits, const Fn &fn) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int tpw = (st.tpi + nwarps - 1) / nwarps;
    const int t_begin = warp * tpw, t_end = min(st.tpi, t_begin + tpw);
    const uint32_t span = hi_bits - lo_bits;  // in range  <=>  (bits - lo_bits) <= span  (unsigned)
    if (st.region == kTileA) {
        constexpr int kUnroll = 8;
        for (int tb = t_begin; tb < t_end; tb += 32) {
            // a tile whose best score is below the range has no member: treat it as empty
            const int my_c = (tb + lane < t_end && tmax[tb + lane] >= lo_bits) ? tcount[tb + lane] : 0;
            const int nt = min(32, t_end - tb);
            for (int u0 = 0; u0 < nt; u0 += kUnroll) {
                uint4 v[kUnroll];
                int c[kUnroll];
#pragma unroll
                for (int u = 0; u < kUnroll; ++u) {
                    c[u] = __shfl_sync(0xffffffffu, my_c, (u0 + u) & 31) - lane * 4;  // valid entries of this lane
                    if (u0 + u >= nt) c[u] = 0;
                    v[u] = make_uint4(0u, 0u, 0u, 0u);
                    if (c[u] > 0)
                        v[u] = *reinterpret_cast<const uint4 *>(score + static_cast<size_t>(tb + u0 + u) * kTileA + lane * 4);
                }
#pragma unroll
                for (int u = 0; u < kUnroll; ++u) {
                    const bool t0 = (v[u].x - lo_bits) <= span && c[u] > 0, t1 = (v[u].y - lo_bits) <= span && c[u] > 1;
                    const bool t2 = (v[u].z - lo_bits) <= span && c[u] > 2, t3 = (v[u].w - lo_bits) <= span && c[u] > 3;
                    if (t0 | t1 | t2 | t3) {
                        const uint32_t slot = static_cast<uint32_t>(tb + u0 + u) * kTileA + lane * 4;
                        if (t0) fn(slot, v[u].x);
                        if (t1) fn(slot + 1, v[u].y);
                        if (t2) fn(slot + 2, v[u].z);
                        if (t3) fn(slot + 3, v[u].w);
                    }
                }
            }
        }
    } else {
        for (int t = t_begin; t < t_end; ++t) {
            const int c = tmax[t] >= lo_bits ? tcount[t] : 0;
            for (int i = lane; i < c; i += 32) {
                const uint32_t slot = static_cast<uint32_t>(t) * st.region + i;
                const uint32_t bits = __float_as_uint(score[slot]);
                if ((bits - lo_bits) <= span) fn(slot, bits);
            }
        }
    }
}

}  // namespace sarpost
