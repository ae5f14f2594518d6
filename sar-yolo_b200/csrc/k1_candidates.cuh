// k1_candidates.cuh — stage K1: raw head logits (or a decoded prediction) -> candidate store.
//
// Replaces, in one pass over the logits:
//   head.py:218-249   level concat, DFL softmax-expectation (block.py:77-80), make_anchors
//                     (tal.py:366-378), dist2bbox (tal.py:381-390), * stride, cls.sigmoid()
//   ops.py:234        candidate mask  amax(cls) > conf
//   ops.py:241-246    xywh -> xyxy (ops.py:416-433)
//   ops.py:268-279    best-class (first max) or multi-label expansion, `classes` filter
// Output = tile-segmented candidate store (common.cuh).
#pragma once
#include <cuda_fp16.h>

#include "common.cuh"

namespace sarpost {

struct HeadGeom {
    int32_t nl, batch, no, nc, n_extra_raw, n_extra_sig;
    int32_t tpi;                           // tiles per image (levels never share a tile)
    int32_t lvl_tile_begin[kMaxLevels + 1];
    int32_t lvl_hw[kMaxLevels];
    int32_t lvl_w[kMaxLevels];
    int32_t lvl_aoff[kMaxLevels + 1];      // first global anchor index of the level; [nl] = A
    float lvl_stride[kMaxLevels];
    const void *lvl_ptr[kMaxLevels];  // elements of type float or __half (kernel template parameter); split layout: box branch
    int32_t is_half;
    uint32_t lvl_w_magic[kMaxLevels];  // ceil(2^32 / W_l): y = pos / W_l without an integer division
    // split layout (SARPOST_LAYOUT_SPLIT): one tensor per branch, see include/sarpost.h
    int32_t split, emb_cl;
    const void *lvl_cls[kMaxLevels];
    const void *lvl_emb[kMaxLevels];
    const void *lvl_state[kMaxLevels];
};

// Element pointer of (image b, logical channel c, position pos) of level l, c in [0, 4*reg_max + nc): box then class rows.
template <typename T>
__device__ __forceinline__ const T *hot_channel_ptr(const HeadGeom &g, int l, int b, int c, int pos) {
    const int64_t hw = g.lvl_hw[l];
    if (!g.split) return static_cast<const T *>(g.lvl_ptr[l]) + (static_cast<int64_t>(b) * g.no + c) * hw + pos;
    if (c < 4 * kRegMax) return static_cast<const T *>(g.lvl_ptr[l]) + (static_cast<int64_t>(b) * (4 * kRegMax) + c) * hw + pos;
    return static_cast<const T *>(g.lvl_cls[l]) + (static_cast<int64_t>(b) * g.nc + (c - 4 * kRegMax)) * hw + pos;
}

// floor(n / d) for n*d < 2^32-ish operands via a precomputed magic = ceil(2^32 / d) and one correction step
__device__ __forceinline__ int fast_div(int n, int d, uint32_t magic) {
    int q = static_cast<int>(__umulhi(static_cast<uint32_t>(n), magic));
    if (q * d > n) --q;               // magic rounds up: the estimate can be one too large
    else if ((q + 1) * d <= n) ++q;   // (only when magic was clamped, d == 1)
    return q;
}

__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ float to_f32(__half v) { return __half2float(v); }
__device__ __forceinline__ void from_f32(float v, float *o) { *o = v; }
__device__ __forceinline__ void from_f32(float v, __half *o) { *o = __float2half_rn(v); }

struct CandFilter {
    float conf;            // (float)conf_thres
    int32_t multi_label;   // already AND-ed with nc > 1 (ops.py:239)
    int32_t has_cls_filter;
    uint32_t cls_allow[kClsWords];
};

__device__ __forceinline__ bool cls_allowed(const CandFilter &f, int j) {
    return !f.has_cls_filter || ((f.cls_allow[j >> 5] >> (j & 31)) & 1u);
}

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// DFL expectation of one box side from 16 logits (block.py:77-80): sum_k k*softmax_k.
// exp(v-m) is evaluated as 2^(v*L - m*L) with one FFMA + one MUFU.EX2; the common factor
// 2^(rounding of m*L) cancels in the ratio.
template <class Acc>
__device__ __forceinline__ float dfl_side(const Acc &acc, int c0) {
    float v[kRegMax];
#pragma unroll
    for (int k = 0; k < kRegMax; ++k) v[k] = acc(c0 + k);
    float m = v[0];
#pragma unroll
    for (int k = 1; k < kRegMax; ++k) m = fmaxf(m, v[k]);
    constexpr float kLog2e = 1.4426950408889634f;
    const float nml = -m * kLog2e;
    float sum = 0.0f, wsum = 0.0f;
#pragma unroll
    for (int k = 0; k < kRegMax; ++k) {
        const float e = ex2_approx(fmaf(v[k], kLog2e, nml));
        sum += e;
        wsum = fmaf(static_cast<float>(k), e, wsum);
    }
    return __fdiv_rn(wsum, sum);
}

// Anchor/stride decode in the reference's operation order (tal.py:381-390, head.py:245), every
// step an individually rounded fp32 op.  Returns xywh in pixels.
template <class Acc>
__device__ __forceinline__ float4 decode_xywh(const Acc &acc, int x, int y, float stride) {
    const float dl = dfl_side(acc, 0), dt = dfl_side(acc, kRegMax), dr = dfl_side(acc, 2 * kRegMax),
                db = dfl_side(acc, 3 * kRegMax);
    const float ax = __fadd_rn(static_cast<float>(x), 0.5f), ay = __fadd_rn(static_cast<float>(y), 0.5f);
    const float x1 = __fsub_rn(ax, dl), y1 = __fsub_rn(ay, dt);
    const float x2 = __fadd_rn(ax, dr), y2 = __fadd_rn(ay, db);
    float4 o;
    o.x = __fmul_rn(__fmul_rn(__fadd_rn(x1, x2), 0.5f), stride);
    o.y = __fmul_rn(__fmul_rn(__fadd_rn(y1, y2), 0.5f), stride);
    o.z = __fmul_rn(__fsub_rn(x2, x1), stride);
    o.w = __fmul_rn(__fsub_rn(y2, y1), stride);
    return o;
}

// ops.py:416-433
__device__ __forceinline__ float4 xywh2xyxy_rn(const float4 c) {
    const float hw = __fmul_rn(c.z, 0.5f), hh = __fmul_rn(c.w, 0.5f);
    return make_float4(__fsub_rn(c.x, hw), __fsub_rn(c.y, hh), __fadd_rn(c.x, hw), __fadd_rn(c.y, hh));
}

// Staging area of the multi-label emission (below): the tile's boxes by thread, and per compacted slot the score and
// its source (thread << 3 | class).
constexpr int kClsCache = 8;
constexpr int kStageEmitBytes = kTileA * 16 + kTileA * kClsCache * (4 + 2);

// Block-wide (kTileA threads) ordered emission of this tile's candidates.
//   score(j): class probability j of this thread's anchor.
//   anchor0:  anchor index of thread 0 of the tile (thread t of a valid tile holds anchor0 + t).
//   stg:      kStageEmitBytes of 16-byte aligned shared memory that no thread reads through score() any more once every
//             thread has evaluated its classes (the TMA kernel passes the tile's own stage buffer), or nullptr.
// scratch: int[8] shared.  Contains two __syncthreads() (three on the staged multi-label path).
//
// Multi-label rows (ops.py:268-270) of one anchor are adjacent slots, so thread t's stores start at a slot that depends
// on every earlier thread: written straight from the registers a warp's 32 stores hit 32 different sectors for every one
// of the <= nc rounds (3 arrays each).  With `stg` the compacted tile is assembled in shared memory first and then
// copied out slot-per-thread: fully coalesced 512 B / 128 B / 128 B warp stores.
template <class ScoreFn>
__device__ __forceinline__ void emit_candidates(bool valid, const float4 xyxy, uint32_t anchor, uint32_t anchor0, int nc,
                                                const CandFilter &f, const ScoreFn &score, const CandStore &st,
                                                int b, int tile_in_image, int *scratch, unsigned char *stg) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int cnt = 0;
    float best = 0.0f;
    int bj = 0;
    // multi-label with few classes (the SAR posture heads have nc = 6): every probability is evaluated once and
    // kept in registers across the block scan; larger heads re-evaluate in the write loop
    float pc[kClsCache];
    uint32_t pass = 0u;
    const bool cached = f.multi_label && nc <= kClsCache;
    float tmax = 0.0f;  // largest candidate score of the tile (lets the selection scan skip whole tiles)
    if (valid) {
        if (cached) {
#pragma unroll
            for (int j = 0; j < kClsCache; ++j) {
                pc[j] = j < nc ? score(j) : 0.0f;
                if (j < nc && (pc[j] > f.conf) && cls_allowed(f, j)) {
                    pass |= 1u << j;
                    tmax = fmaxf(tmax, pc[j]);
                }
            }
            cnt = __popc(pass);
        } else if (f.multi_label) {
            for (int j = 0; j < nc; ++j) {
                const float p = score(j);
                if ((p > f.conf) && cls_allowed(f, j)) {
                    ++cnt;
                    tmax = fmaxf(tmax, p);
                }
            }
        } else {
            best = score(0);
            for (int j = 1; j < nc; ++j) {
                const float p = score(j);
                if (p > best) { best = p; bj = j; }  // strict >: first max wins (ops.py:274)
            }
            cnt = (best > f.conf) && cls_allowed(f, bj);
            tmax = cnt ? best : 0.0f;
        }
    }
    const uint32_t wmax = __reduce_max_sync(0xffffffffu, __float_as_uint(fmaxf(tmax, 0.0f)));
    // exclusive scan of cnt over the block
    int inc = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int n = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += n;
    }
    if (lane == 31) scratch[warp] = inc;
    if (lane == 0) scratch[4 + warp] = static_cast<int>(wmax);
    __syncthreads();
    int base = 0;
#pragma unroll
    for (int w = 0; w < kTileA / 32; ++w) base += (w < warp) ? scratch[w] : 0;
    if (tid == 0) {
        int tot = 0;
#pragma unroll
        for (int w = 0; w < kTileA / 32; ++w) tot += scratch[w];
        st.tile_count[static_cast<int64_t>(b) * st.tpi + tile_in_image] = tot;
        uint32_t mx = 0u;
#pragma unroll
        for (int w = 0; w < kTileA / 32; ++w) mx = max(mx, static_cast<uint32_t>(scratch[4 + w]));
        st.tile_max[static_cast<int64_t>(b) * st.tpi + tile_in_image] = mx;
    }
    int64_t pos = static_cast<int64_t>(b) * st.cap + static_cast<int64_t>(tile_in_image) * st.region + base + (inc - cnt);
    int32_t *hist = st.hist + static_cast<int64_t>(b) * kBuckets;
    const bool sampled = (anchor % kHistSample) == 0;  // 1-in-8 sample keeps the RED traffic negligible
    if (cached && stg != nullptr) {
        // every thread is past its last score() read (the barrier above): the staging area may be overwritten
        float4 *sx = reinterpret_cast<float4 *>(stg);
        float *sscore = reinterpret_cast<float *>(stg + kTileA * 16);
        uint16_t *ssrc = reinterpret_cast<uint16_t *>(stg + kTileA * 16 + kTileA * kClsCache * 4);
        int tot = 0;
#pragma unroll
        for (int w = 0; w < kTileA / 32; ++w) tot += scratch[w];
        if (cnt) {
            sx[tid] = xyxy;
            int o = base + (inc - cnt);
#pragma unroll
            for (int j = 0; j < kClsCache; ++j) {
                if ((pass >> j) & 1u) {
                    sscore[o] = pc[j];
                    ssrc[o] = static_cast<uint16_t>((tid << 3) | j);
                    ++o;
                }
            }
        }
        __syncthreads();
        const int64_t g0 = static_cast<int64_t>(b) * st.cap + static_cast<int64_t>(tile_in_image) * st.region;
        for (int k = tid; k < tot; k += kTileA) {
            const uint32_t src = ssrc[k];
            const uint32_t t = src >> 3, j = src & 7u;
            const float sc = sscore[k];
            st.box[g0 + k] = sx[t];
            st.score[g0 + k] = sc;
            st.key[g0 + k] = (anchor0 + t) * static_cast<uint32_t>(nc) + j;
            if (((anchor0 + t) % kHistSample) == 0) atomicAdd(hist + score_bucket(__float_as_uint(sc)), 1);
        }
        fence_proxy_async();  // TMA kernel: the staging area is the stage buffer the next bulk load lands in
    } else if (cnt) {
        if (cached) {
#pragma unroll
            for (int j = 0; j < kClsCache; ++j) {
                if ((pass >> j) & 1u) {
                    st.box[pos] = xyxy;
                    st.score[pos] = pc[j];
                    st.key[pos] = anchor * static_cast<uint32_t>(nc) + static_cast<uint32_t>(j);
                    if (sampled) atomicAdd(hist + score_bucket(__float_as_uint(pc[j])), 1);
                    ++pos;
                }
            }
        } else if (f.multi_label) {
            for (int j = 0; j < nc; ++j) {
                const float p = score(j);
                if ((p > f.conf) && cls_allowed(f, j)) {
                    st.box[pos] = xyxy;
                    st.score[pos] = p;
                    st.key[pos] = anchor * static_cast<uint32_t>(nc) + static_cast<uint32_t>(j);
                    if (sampled) atomicAdd(hist + score_bucket(__float_as_uint(p)), 1);  // RED.ADD, result unused
                    ++pos;
                }
            }
        } else {
            st.box[pos] = xyxy;
            st.score[pos] = best;
            st.key[pos] = anchor * static_cast<uint32_t>(nc) + static_cast<uint32_t>(bj);
            if (sampled) atomicAdd(hist + score_bucket(__float_as_uint(best)), 1);
        }
    }
    __syncthreads();  // scratch reusable, and (TMA kernel) every read of the stage buffer is done
}

// Which level does tile r (index within an image) belong to.
__device__ __forceinline__ int tile_level(const HeadGeom &g, int r) {
    int l = 0;
#pragma unroll
    for (int i = 1; i < kMaxLevels; ++i) l += (r >= g.lvl_tile_begin[i]) ? 1 : 0;  // entries >= nl hold tpi (never reached)
    return l;
}

// ---------------------------------------------------------------------------------------------
// K1 (TMA): persistent CTAs, kTileA threads, multi-stage mbarrier ring.  One TMA box per tile =
// {kTileA anchors, 4*reg_max + nc channels, 1 image} landing as smem[channel][anchor]: thread t
// reads column t -> conflict-free, and the extras channels are never fetched.
// ---------------------------------------------------------------------------------------------
struct K1TmaParams {
    CUtensorMap maps[kMaxLevels];      // cat layout: box + cls rows of the level tensor; split layout: the box branch
    CUtensorMap maps_cls[kMaxLevels];  // split layout: the class branch
    HeadGeom g;
    CandFilter f;
    CandStore st;
    int32_t stages;
    int32_t n_tiles;  // B * tpi
    int32_t *tile_counter;  // zero at launch: next tile to hand out (the NMS kernel zeroes it again)
};

constexpr int kMaxStages = 8;

__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// Tiles are handed out through an atomic counter instead of a fixed stride: CTAs that become resident late — SMs held by
// the NMS kernel of the previous batch when batches are pipelined (sarpost_pipeline_*), or simply a partial last wave —
// take what is left instead of owning a fixed share, so the kernel ends when the work does.  Which CTA decodes a tile
// does not affect the result: every tile writes its own region of the candidate store.
template <typename T>
__global__ void __launch_bounds__(kTileA, sizeof(T) == 2 ? 6 : 3) k1_fused_tma(const __grid_constant__ K1TmaParams p) {
    extern __shared__ unsigned char dyn_smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[kMaxStages];
    __shared__ int ring_b[kMaxStages], ring_r[kMaxStages];  // image / tile-in-image of the tile in flight in each stage; b < 0: no more work
    __shared__ int scratch[8];
    const int tid = threadIdx.x;
    const int nch = 4 * kRegMax + p.g.nc;
    const uint32_t stage_bytes = static_cast<uint32_t>(nch) * kTileA * sizeof(T);
    unsigned char *dyn = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(dyn_smem_raw) + 127) & ~uintptr_t(127));

    // thread 0: tile index fetched ahead of its use, so the atomic's round trip is off the issue path.  The first tile of a
    // CTA is its block index (no round trip before the first load is in flight); the counter hands out the tiles behind
    // the first wave: counter value v = tile gridDim.x + v.
    int next_t = blockIdx.x;
    if (tid == 0) {
        for (int s = 0; s < p.stages; ++s) mbar_init(&full_bar[s], 1);
        fence_barrier_init();
        for (int l = 0; l < p.g.nl; ++l) {
            tma_prefetch_desc(&p.maps[l]);
            if (p.g.split) tma_prefetch_desc(&p.maps_cls[l]);
        }
    }
    __syncthreads();

    // thread 0: put the next tile into stage s (or mark the stage "no more work"), then fetch the index after it
    auto issue = [&](int s) {
        const int t = next_t;
        if (t >= p.n_tiles) {
            ring_b[s] = -1;
            mbar_arrive(&full_bar[s]);  // completes the phase without a transfer
            // out of tiles: at most stages-1 are left in this CTA's ring.  Once every CTA has said so the NMS kernel (a
            // programmatic dependent) may start placing its CTAs on the SMs that fall free; it still waits for this grid
            // to complete before it reads anything (griddepcontrol.wait in k4_nms).
            asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
            return;
        }
        next_t = static_cast<int>(gridDim.x) + atomicAdd(p.tile_counter, 1);
        const int b = t / p.g.tpi, r = t - b * p.g.tpi;
        ring_b[s] = b;
        ring_r[s] = r;
        const int l = tile_level(p.g, r);
        const int a0 = (r - p.g.lvl_tile_begin[l]) * kTileA;
        mbar_arrive_expect_tx(&full_bar[s], stage_bytes);  // release: the ring entries are visible to whoever sees the phase flip
        unsigned char *dst = dyn + static_cast<size_t>(s) * stage_bytes;
        tma_load_3d(dst, &p.maps[l], a0, 0, b, &full_bar[s]);
        if (p.g.split)  // the class rows land right behind the 4*reg_max box rows: same smem[channel][anchor] tile as the cat layout
            tma_load_3d(dst + static_cast<size_t>(4 * kRegMax) * kTileA * sizeof(T), &p.maps_cls[l], a0, 0, b, &full_bar[s]);
    };
    if (tid == 0)
        for (int s = 0; s < p.stages; ++s) issue(s);

    int s = 0;
    uint32_t phase = 0;
    for (;;) {
        mbar_wait(&full_bar[s], phase);
        const int b = ring_b[s];
        if (b < 0) break;  // tiles are handed out in order: every later stage is empty too, nothing is in flight
        const int r = ring_r[s];
        const int l = tile_level(p.g, r);
        const int pos = (r - p.g.lvl_tile_begin[l]) * kTileA + tid;  // position inside the level
        const bool valid = pos < p.g.lvl_hw[l];
        const T *buf = reinterpret_cast<const T *>(dyn + static_cast<size_t>(s) * stage_bytes) + tid;
        auto acc = [&](int c) { return to_f32(buf[c * kTileA]); };
        const int w = p.g.lvl_w[l];
        const int yy = fast_div(pos, w, p.g.lvl_w_magic[l]), xx = pos - yy * w;
        const float4 xyxy = xywh2xyxy_rn(decode_xywh(acc, xx, yy, p.g.lvl_stride[l]));
        auto score = [&](int j) { return sigmoid_rn(to_f32(buf[(4 * kRegMax + j) * kTileA])); };
        emit_candidates(valid, xyxy, static_cast<uint32_t>(p.g.lvl_aoff[l] + pos), static_cast<uint32_t>(p.g.lvl_aoff[l] + pos - tid), p.g.nc, p.f,
                        score, p.st, b, r, scratch, dyn + static_cast<size_t>(s) * stage_bytes);
        // emit_candidates ended with __syncthreads(): stage s (buffer and ring entry) is free again
        if (tid == 0) issue(s);
        if (++s == p.stages) { s = 0; phase ^= 1u; }
    }
}

// ---------------------------------------------------------------------------------------------
// K1 (LDG): same math with direct coalesced global loads; used when a level cannot be described
// by a tensor map (H*W*4 not a multiple of 16 bytes — odd rect-inference levels — unaligned base,
// or 4*reg_max + nc > 256 box rows).  grid = (tpi, B).
// ---------------------------------------------------------------------------------------------
struct K1LdgParams {
    HeadGeom g;
    CandFilter f;
    CandStore st;
};

template <typename T>
__global__ void __launch_bounds__(kTileA) k1_fused_ldg(const __grid_constant__ K1LdgParams p) {
    __shared__ int scratch[8];
    __shared__ __align__(16) unsigned char stg[kStageEmitBytes];
    const int r = blockIdx.x, b = blockIdx.y;
    const int l = tile_level(p.g, r);
    const int hw = p.g.lvl_hw[l];
    const int pos = (r - p.g.lvl_tile_begin[l]) * kTileA + threadIdx.x;
    const bool valid = pos < hw;
    const int pc = valid ? pos : hw - 1;
    const T *base_box = hot_channel_ptr<T>(p.g, l, b, 0, pc), *base_cls = hot_channel_ptr<T>(p.g, l, b, 4 * kRegMax, pc);
    auto acc = [&](int c) { return to_f32(__ldg(base_box + static_cast<int64_t>(c) * hw)); };
    const int w = p.g.lvl_w[l];
    const int yy = fast_div(pc, w, p.g.lvl_w_magic[l]), xx = pc - yy * w;
    const float4 xyxy = xywh2xyxy_rn(decode_xywh(acc, xx, yy, p.g.lvl_stride[l]));
    auto score = [&](int j) { return sigmoid_rn(to_f32(__ldg(base_cls + static_cast<int64_t>(j) * hw))); };
    emit_candidates(valid, xyxy, static_cast<uint32_t>(p.g.lvl_aoff[l] + pc), static_cast<uint32_t>(p.g.lvl_aoff[l] + pos - threadIdx.x), p.g.nc,
                    p.f, score, p.st, b, r, scratch, stg);
}

// ---------------------------------------------------------------------------------------------
// K1 (decoded): candidates from an already decoded prediction (B, C, A) — the tensor
// ops.non_max_suppression receives (ops.py:167).  grid = (ceil(A/kTileA), B).
// ---------------------------------------------------------------------------------------------
struct K1DecodedParams {
    const void *pred;  // float or __half (kernel template parameter)
    int32_t channels, nc;
    int64_t anchors;
    CandFilter f;
    CandStore st;
};

template <typename T>
__global__ void __launch_bounds__(kTileA) k1_decoded(const __grid_constant__ K1DecodedParams p) {
    __shared__ int scratch[8];
    __shared__ __align__(16) unsigned char stg[kStageEmitBytes];
    const int r = blockIdx.x, b = blockIdx.y;
    const int64_t a = static_cast<int64_t>(r) * kTileA + threadIdx.x;
    const bool valid = a < p.anchors;
    const int64_t ac = valid ? a : p.anchors - 1;
    const T *base = static_cast<const T *>(p.pred) + static_cast<int64_t>(b) * p.channels * p.anchors + ac;
    const float4 xywh = make_float4(to_f32(__ldg(base)), to_f32(__ldg(base + p.anchors)), to_f32(__ldg(base + 2 * p.anchors)),
                                    to_f32(__ldg(base + 3 * p.anchors)));
    const float4 xyxy = xywh2xyxy_rn(xywh);
    auto score = [&](int j) { return to_f32(__ldg(base + static_cast<int64_t>(4 + j) * p.anchors)); };
    emit_candidates(valid, xyxy, static_cast<uint32_t>(ac), static_cast<uint32_t>(r) * kTileA, p.nc, p.f, score, p.st, b, r, scratch, stg);
}

// ---------------------------------------------------------------------------------------------
// K1 (labels): apriori labels of `save_hybrid` (ops.py:256-261): v[:, :4] = xywh2xyxy(lb[:, 1:5]),
// v[i, 4 + cls] = 1.0, extras 0, concatenated AFTER the image's own rows — here the tiles that follow the
// anchor tiles.  A label row yields one candidate (score 1.0, its class) in both best-class and multi-label
// mode, subject to `score > conf` and the `classes` filter.  grid = (label tiles, B), blockDim = kTileA.
// ---------------------------------------------------------------------------------------------
struct K1LabelParams {
    const float *labels;          // [B, max_labels, 5]
    const int32_t *label_counts;  // [B]
    int32_t max_labels, nc;
    int32_t first_tile;           // index of the first label tile inside an image
    uint32_t first_anchor;        // = A: label row i is reported as anchor A + i
    CandFilter f;
    CandStore st;
};

__global__ void __launch_bounds__(kTileA) k1_labels(const __grid_constant__ K1LabelParams p) {
    __shared__ int scratch[8];
    const int lt = blockIdx.x, b = blockIdx.y;
    const int i = lt * kTileA + threadIdx.x;
    int n = p.label_counts[b];
    n = n < 0 ? 0 : (n > p.max_labels ? p.max_labels : n);
    const bool valid = i < n;
    const float *row = p.labels + (static_cast<int64_t>(b) * p.max_labels + (valid ? i : 0)) * 5;
    const int cls = valid ? static_cast<int>(row[0]) : -1;  // lb[:, 0].long()
    const float4 xyxy = xywh2xyxy_rn(make_float4(row[1], row[2], row[3], row[4]));
    auto score = [&](int j) { return j == cls ? 1.0f : 0.0f; };
    emit_candidates(valid && cls >= 0 && cls < p.nc, xyxy, p.first_anchor + static_cast<uint32_t>(i),
                    p.first_anchor + static_cast<uint32_t>(lt) * kTileA, p.nc, p.f, score, p.st, b, p.first_tile + lt, scratch, nullptr);
}

// ---------------------------------------------------------------------------------------------
// K1 (merge): candidates for the cross-tile merge — one block per SAHI tile; rows are already
// filtered detections (x1,y1,x2,y2,conf,cls,...) which are shifted by the tile origin.
// grid = (tiles_per_frame, n_frames), blockDim = 128.
// ---------------------------------------------------------------------------------------------
struct K1MergeParams {
    const float *dets;
    const int32_t *det_counts;
    const float *origins;
    int32_t dets_per_tile, row_len;
    float *cls;     // [n_frames*cap] class id of every slot (NMS class offset)
    CandStore st;   // region = dets_per_tile, tpi = tiles_per_frame
};

__global__ void __launch_bounds__(128) k1_merge(const __grid_constant__ K1MergeParams p) {
    const int t = blockIdx.x, f = blockIdx.y;
    const int64_t tile = static_cast<int64_t>(f) * p.st.tpi + t;
    int n = p.det_counts[tile];
    n = n < 0 ? 0 : (n > p.dets_per_tile ? p.dets_per_tile : n);
    const float ox = p.origins[2 * tile], oy = p.origins[2 * tile + 1];
    const int64_t base = static_cast<int64_t>(f) * p.st.cap + static_cast<int64_t>(t) * p.st.region;
    if (threadIdx.x == 0) p.st.tile_max[tile] = n > 0 ? 0xffffffffu : 0u;  // no pruning on the merge path
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const float *row = p.dets + (tile * p.dets_per_tile + i) * p.row_len;
        p.st.box[base + i] = make_float4(__fadd_rn(row[0], ox), __fadd_rn(row[1], oy), __fadd_rn(row[2], ox),
                                         __fadd_rn(row[3], oy));
        p.st.score[base + i] = row[4];
        if ((i % kHistSample) == 0) atomicAdd(p.st.hist + static_cast<int64_t>(f) * kBuckets + score_bucket(__float_as_uint(row[4])), 1);
        p.cls[base + i] = row[5];
        p.st.key[base + i] = static_cast<uint32_t>(t) * p.dets_per_tile + i;
    }
    if (threadIdx.x == 0) p.st.tile_count[tile] = n;
}

// ---------------------------------------------------------------------------------------------
// Decode to y (API-exact Detect._inference / JDE._inference, head.py:100-131, :214-249):
// y (B, 4+nc+n_extra_raw+n_extra_sig, A).  grid = (tpi, B), blockDim = kTileA.
// ---------------------------------------------------------------------------------------------
struct DecodeYParams {
    HeadGeom g;
    void *y;  // same element type as the levels
    int64_t anchors;
};

template <typename T>
__global__ void __launch_bounds__(kTileA) k_decode_y(const __grid_constant__ DecodeYParams p) {
    const int r = blockIdx.x, b = blockIdx.y;
    const int l = tile_level(p.g, r);
    const int hw = p.g.lvl_hw[l];
    const int pos = (r - p.g.lvl_tile_begin[l]) * kTileA + threadIdx.x;
    if (pos >= hw) return;
    const T *base = static_cast<const T *>(p.g.lvl_ptr[l]) + static_cast<int64_t>(b) * p.g.no * hw + pos;
    auto acc = [&](int c) { return to_f32(__ldg(base + static_cast<int64_t>(c) * hw)); };
    const int w = p.g.lvl_w[l];
    const int yy = fast_div(pos, w, p.g.lvl_w_magic[l]), xx = pos - yy * w;
    const float4 o = decode_xywh(acc, xx, yy, p.g.lvl_stride[l]);
    const int cout = 4 + p.g.nc + p.g.n_extra_raw + p.g.n_extra_sig;
    T *yo = static_cast<T *>(p.y) + static_cast<int64_t>(b) * cout * p.anchors + p.g.lvl_aoff[l] + pos;
    from_f32(o.x, yo);
    from_f32(o.y, yo + p.anchors);
    from_f32(o.z, yo + 2 * p.anchors);
    from_f32(o.w, yo + 3 * p.anchors);
    int ci = 4 * kRegMax, co = 4;
    for (int j = 0; j < p.g.nc; ++j, ++ci, ++co) from_f32(sigmoid_rn(acc(ci)), yo + co * p.anchors);
    for (int j = 0; j < p.g.n_extra_raw; ++j, ++ci, ++co) yo[co * p.anchors] = __ldg(base + static_cast<int64_t>(ci) * hw);
    for (int j = 0; j < p.g.n_extra_sig; ++j, ++ci, ++co) from_f32(sigmoid_rn(acc(ci)), yo + co * p.anchors);
}

}  // namespace sarpost
