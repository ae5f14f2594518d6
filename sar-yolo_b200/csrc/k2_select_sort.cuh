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
// Block-wide (contains __syncthreads).  The scan is instruction-bound (one CTA streams candidate scores of
// its image out of L2), so
//   1. tiles whose best score (tile_max, written by K1) is below the range, or that are empty, are dropped
//      up front: the surviving tile ids are compacted into `tile_list` (shared memory, kTileListCap entries
//      per round);
//   2. warps then walk that list with kUnroll tiles in flight; when the tile region is exactly kTileA slots a
//      lane fetches its four scores with one 128-bit load, and the common case — no member among them —
//      costs four range tests and a single branch.
constexpr int kTileListCap = 2048;

template <class Fn>
__device__ __forceinline__ void for_each_candidate_in(const CandStore &st, const int32_t *tcount, const uint32_t *tmax,
                                                      const float *score, uint32_t lo_bits, uint32_t hi_bits,
                                                      int tile_first, int tile_step, int n_share /* this CTA's share of the image's
                                                      tiles: tile_first + i*tile_step, i in [0, n_share) — interleaved across the
                                                      CTAs of a cluster so that dense regions of the image spread over all of them */,
                                                      uint32_t *tile_list /*[kTileListCap] smem*/, int *list_n /*smem*/,
                                                      const Fn &fn) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const uint32_t span = hi_bits - lo_bits;  // in range  <=>  (bits - lo_bits) <= span  (unsigned)
    for (int t0 = 0; t0 < n_share; t0 += kTileListCap) {
        __syncthreads();
        if (tid == 0) *list_n = 0;
        __syncthreads();
        const int t1 = min(n_share, t0 + kTileListCap);
        for (int tb = t0 + warp * 32; tb < t1; tb += nwarps * 32) {
            const int v = tb + lane;
            const int t = tile_first + v * tile_step;
            // both loads are issued together (one round trip to L2, not two)
            const uint32_t tm = v < t1 ? tmax[t] : 0u;
            const int tc = v < t1 ? tcount[t] : 0;
            const int c = tm >= lo_bits ? tc : 0;
            const uint32_t bal = __ballot_sync(0xffffffffu, c > 0);
            if (bal) {
                int base = 0;
                if (lane == 0) base = atomicAdd(list_n, __popc(bal));
                base = __shfl_sync(0xffffffffu, base, 0);
                // the tile's population travels with its id (regions of kTileA slots: count <= 128), so the scan below
                // does not have to fetch it again
                if (c > 0) tile_list[base + __popc(bal & lanemask_lt())] = static_cast<uint32_t>(t) | (st.region == kTileA ? static_cast<uint32_t>(c) << 24 : 0u);
            }
        }
        __syncthreads();
        const int n_act = *list_n;
        if (st.region == kTileA) {
            constexpr int kUnroll = 8;
            for (int e0 = warp * kUnroll; e0 < n_act; e0 += nwarps * kUnroll) {
                uint4 v[kUnroll];
                int c[kUnroll];
                uint32_t tt[kUnroll];
#pragma unroll
                for (int u = 0; u < kUnroll; ++u) {
                    const bool on = e0 + u < n_act;
                    const uint32_t ent = on ? tile_list[e0 + u] : 0u;
                    tt[u] = ent & 0xffffffu;
                    c[u] = static_cast<int>(ent >> 24) - lane * 4;  // valid entries of this lane (0 - ... when the entry is off)
                    v[u] = make_uint4(0u, 0u, 0u, 0u);
                    if (c[u] > 0) v[u] = *reinterpret_cast<const uint4 *>(score + static_cast<size_t>(tt[u]) * kTileA + lane * 4);
                }
#pragma unroll
                for (int u = 0; u < kUnroll; ++u) {
                    const bool h0 = (v[u].x - lo_bits) <= span && c[u] > 0, h1 = (v[u].y - lo_bits) <= span && c[u] > 1;
                    const bool h2 = (v[u].z - lo_bits) <= span && c[u] > 2, h3 = (v[u].w - lo_bits) <= span && c[u] > 3;
                    if (h0 | h1 | h2 | h3) {
                        const uint32_t slot = tt[u] * kTileA + lane * 4;
                        if (h0) fn(slot, v[u].x);
                        if (h1) fn(slot + 1, v[u].y);
                        if (h2) fn(slot + 2, v[u].z);
                        if (h3) fn(slot + 3, v[u].w);
                    }
                }
            }
        } else {
            for (int e = warp; e < n_act; e += nwarps) {
                const uint32_t t = tile_list[e];
                const int c = tcount[t];
                for (int i0 = 0; i0 < c; i0 += 128) {
                    uint32_t bits[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int i = i0 + u * 32 + lane;
                        bits[u] = i < c ? __float_as_uint(score[static_cast<size_t>(t) * st.region + i]) : 0u;
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int i = i0 + u * 32 + lane;
                        if (i < c && (bits[u] - lo_bits) <= span) fn(t * static_cast<uint32_t>(st.region) + i, bits[u]);
                    }
                }
            }
        }
    }
    __syncthreads();
}

}  // namespace sarpost
