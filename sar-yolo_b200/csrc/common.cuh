// common.cuh — shared device helpers for libsarpost (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/sarpost.h"

namespace sarpost {

constexpr int kTileA = 128;        // anchors per K1 tile (= candidate region granularity)
constexpr int kMaxLevels = SARPOST_MAX_LEVELS;
constexpr int kClsWords = SARPOST_MAX_CLASSES / 32;
constexpr int kRegMax = 16;

// ---------------------------------------------------------------------------------------------
// Candidate store ("tile-segmented"): every tile of kTileA anchors (or every SAHI tile) owns a
// fixed region of `region` slots inside its image's segment of `cap` slots; a tile's candidates
// are compacted to the front of its region in source order and tile_count[] says how many.
// Source order of an image = tiles in order, then position in tile.  No atomics, deterministic.
// ---------------------------------------------------------------------------------------------
struct CandStore {
    float4 *box;        // [B*cap] x1,y1,x2,y2 (un-offset)
    float *score;       // [B*cap]
    uint32_t *key;      // [B*cap] anchor*nc + cls  (merge: tile*dets_per_tile + row)
    int32_t *tile_count;// [B*tpi]
    uint32_t *tile_max; // [B*tpi] largest score bit pattern among the tile's candidates (0 if none)
    int32_t *hist;      // [B*kBuckets] per-image SAMPLED histogram of score_bucket(): candidates of every
                        // kHistSample-th anchor (zeroed before K1); only steers chunk sizes, never results
    int64_t cap;        // slots per image
    int32_t tpi;        // tiles per image
    int32_t region;     // slots per tile region
};

// Monotone score -> bucket map used for top-k selection and lazy sorting: 16 octaves below 1.0 at
// 8 mantissa bits (relative width 0.4 %); everything lower shares bucket 0, everything >= 1.0 bucket 4095.
constexpr int kBuckets = 4096;
constexpr int kHistSample = 8;  // K1 histograms the candidates of anchors with (anchor % kHistSample) == 0
constexpr int kBucketShift = 15;
constexpr int kBucketBase = (0x3F800000 >> kBucketShift) - (kBuckets - 1);
__device__ __forceinline__ int score_bucket(uint32_t bits) {
    const int b = static_cast<int>(bits >> kBucketShift) - kBucketBase;
    return min(max(b, 0), kBuckets - 1);
}

// Smallest score bit pattern that maps to bucket >= b (score_bucket is monotone in the bit pattern).
__host__ __device__ __forceinline__ uint32_t bucket_floor_bits(int b) {
    return b <= 0 ? 0u : static_cast<uint32_t>(b + kBucketBase) << kBucketShift;
}

// fp32 sigmoid as torch computes it: 1 / (1 + exp(-x)), each step rounded (head.py:131,249).  The correctly rounded
// reciprocal IS the correctly rounded quotient 1 / d (same real number, same rounding), without the general division's
// scaling and slow-path checks.
__device__ __forceinline__ float sigmoid_rn(float x) {
    return __frcp_rn(__fadd_rn(1.0f, expf(-x)));
}

// IoU(a,b) > thr with torchvision CPU nms semantics (see oracle/nms_greedy.c): every op an
// individually rounded fp32 op (no FMA), `thr` = largest float <= the double threshold.
// NaN (0/0) compares false.
__device__ __forceinline__ bool iou_gt(const float4 a, const float area_a, const float4 b, const float area_b,
                                       const float thr) {
    const float xx1 = fmaxf(a.x, b.x);
    const float yy1 = fmaxf(a.y, b.y);
    const float xx2 = fminf(a.z, b.z);
    const float yy2 = fminf(a.w, b.w);
    const float w = fmaxf(0.0f, __fsub_rn(xx2, xx1));
    const float h = fmaxf(0.0f, __fsub_rn(yy2, yy1));
    const float inter = __fmul_rn(w, h);
    const float uni = __fsub_rn(__fadd_rn(area_a, area_b), inter);
    return __fdiv_rn(inter, uni) > thr;
}

__device__ __forceinline__ float box_area_rn(const float4 b) {
    return __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y));
}

// ---------------------------------------------------------------------------------------------
// mbarrier / TMA PTX wrappers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// 3-D tiled TMA load global -> shared, completion on an mbarrier.
__device__ __forceinline__ void tma_load_3d(void *smem_dst, const CUtensorMap *map, int c0, int c1, int c2,
                                            uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2),
        "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap *map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

__device__ __forceinline__ uint32_t lanemask_lt() {
    uint32_t m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

}  // namespace sarpost
