// K7: deferred JDE state head.
//
// The reference evaluates `state_predictor` = Linear(E, E/2) -> ReLU -> Dropout (identity in eval) -> Linear(E/2, S)
// on the embedding of EVERY anchor inside JDE.forward (nn/modules/head.py:189-190, 198-204) and applies the sigmoid in
// `_inference` (head.py:247); only the <= max_det rows that survive NMS are ever read.  The MLP is per-anchor, so the
// same numbers come out when it runs on the survivors' embeddings after the gather: B*max_det rows instead of B*A.
//
// One CTA (128 threads) = kStateRows output rows of one image.  Embeddings are staged in shared memory; W1 streams
// through a double-buffered 32-column shared tile (128-bit global loads prefetched into registers while the previous
// tile is consumed; `[j][36]` layout = conflict-free 128-bit shared loads).  Every thread owns a JT x 4 register tile
// (JT hidden units x 4 rows): per 4 columns JT + 4 shared loads feed 16*JT FMAs (fp32, k ascending).  The lanes of a
// warp share their rows, so embedding reads are broadcasts.  Layer 2 is a handful of dot products per row.
#pragma once
#include "common.cuh"

namespace sarpost {

constexpr int kStateRows = 16;     // rows per CTA (300 kept rows x 16 images = 304 CTAs: one wave at 2-3 CTAs per SM)
constexpr int kStateThreads = 128;
constexpr int kStateRT = 4;        // rows per thread (kStateRows = 4 warps x kStateRT)
constexpr int kStateKC = 32;       // W1 columns per shared tile
constexpr int kStateWS = kStateKC + 4;  // shared row stride of the W1 tile (floats)

struct StateHeadParams {
    // Row g = b*max_det + r of the batch: embedding at emb + g*emb_stride, state probabilities written to
    // state_out + g*state_stride (nullptr: not wanted), argmax over them (first maximum, as a float — the `state_id`
    // column of models/yolo/jde/predict.py:61-64) written to id_out + g*id_stride (nullptr: not wanted).
    const float *emb;
    float *state_out, *id_out;
    int64_t emb_stride, state_stride, id_stride;
    const int32_t *counts;  // [B]
    int32_t batch, max_det, embed_dim, n_state, hidden;
    int32_t w1_vec;         // W1 rows are 16-byte aligned (embed_dim % 4 == 0, aligned base): 128-bit loads
    const float *w1, *b1;   // nn.Linear(E, H): weight (H, E) row-major, bias (H)
    const float *w2, *b2;   // nn.Linear(H, S): weight (S, H) row-major, bias (S)
};

__host__ __device__ inline int state_head_smem_floats(int embed_dim, int hidden, int jt) {
    const int jl = 32 * jt, hpad = ((hidden + jl - 1) / jl) * jl, epad = ((embed_dim + kStateKC - 1) / kStateKC) * kStateKC;
    return kStateRows * epad + 2 * jl * kStateWS + kStateRows * (hpad + 1) + kStateRows * 64 /* probabilities for the argmax */;
}

// JT = hidden units per thread; one pass covers JL = 32*JT hidden units.
template <int JT>
__global__ void __launch_bounds__(kStateThreads) k7_state_head(const __grid_constant__ StateHeadParams p) {
    constexpr int JL = 32 * JT;
    constexpr int kVPer = JL * (kStateKC / 4) / kStateThreads;  // float4 of a W1 tile per thread (2 * JT)
    constexpr int kTileF = JL * kStateWS;
    extern __shared__ __align__(16) float sh_state[];
    const int b = blockIdx.y, row0 = blockIdx.x * kStateRows, tid = threadIdx.x;
    const int n_rows = min(p.counts[b], p.max_det) - row0;
    if (n_rows <= 0) return;
    const int E = p.embed_dim, H = p.hidden, S = p.n_state;
    const int hpad = ((H + JL - 1) / JL) * JL, epad = ((E + kStateKC - 1) / kStateKC) * kStateKC;
    float *EMB = sh_state;                 // [kStateRows][epad], zero padded
    float *WT = EMB + kStateRows * epad;   // 2 x [JL][kStateWS]
    float *HID = WT + 2 * kTileF;          // [kStateRows][hpad + 1]
    float *PROB = HID + kStateRows * (hpad + 1);  // [kStateRows][64]
    const int64_t g0 = static_cast<int64_t>(b) * p.max_det + row0;

    const int n_kt = epad / kStateKC, n_tiles = (hpad / JL) * n_kt;
    float4 wreg[kVPer];
    auto load_tile = [&](int j0, int k0) {  // global -> registers; 8 consecutive threads read one 128-byte run of a W1 row
#pragma unroll
        for (int i = 0; i < kVPer; ++i) {
            const int idx = tid + i * kStateThreads, j = j0 + (idx >> 3), k = k0 + (idx & 7) * 4;
            const float *src = p.w1 + static_cast<int64_t>(j) * E + k;
            if (p.w1_vec) {
                wreg[i] = (j < H && k < E) ? __ldg(reinterpret_cast<const float4 *>(src)) : make_float4(0.f, 0.f, 0.f, 0.f);
            } else {
                wreg[i].x = (j < H && k + 0 < E) ? __ldg(src + 0) : 0.0f;
                wreg[i].y = (j < H && k + 1 < E) ? __ldg(src + 1) : 0.0f;
                wreg[i].z = (j < H && k + 2 < E) ? __ldg(src + 2) : 0.0f;
                wreg[i].w = (j < H && k + 3 < E) ? __ldg(src + 3) : 0.0f;
            }
        }
    };
    auto store_tile = [&](int buf) {
#pragma unroll
        for (int i = 0; i < kVPer; ++i) {
            const int idx = tid + i * kStateThreads;
            *reinterpret_cast<float4 *>(WT + buf * kTileF + (idx >> 3) * kStateWS + (idx & 7) * 4) = wreg[i];
        }
    };
    load_tile(0, 0);
    for (int r = 0; r < kStateRows; ++r) {
        const float *src = p.emb + (g0 + r) * p.emb_stride;
        for (int k = tid; k < epad; k += kStateThreads) EMB[r * epad + k] = (r < n_rows && k < E) ? src[k] : 0.0f;
    }
    store_tile(0);
    __syncthreads();

    const int lane = tid & 31, grp = tid >> 5;
    const float *my_emb = EMB + grp * kStateRT * epad;
    float acc[JT][kStateRT];
    int j0 = 0, kt = 0;
    for (int t = 0; t < n_tiles; ++t) {
        if (kt == 0) {
#pragma unroll
            for (int i = 0; i < JT; ++i)
#pragma unroll
                for (int r = 0; r < kStateRT; ++r) acc[i][r] = 0.0f;
        }
        const int k0 = kt * kStateKC;
        const bool last_k = kt == n_kt - 1, more = t + 1 < n_tiles;
        if (more) load_tile(last_k ? j0 + JL : j0, last_k ? 0 : k0 + kStateKC);  // in flight while this tile is consumed
        const float *wbase = WT + (t & 1) * kTileF + lane * kStateWS;
        const float *ebase = my_emb + k0;
#pragma unroll
        for (int k = 0; k < kStateKC; k += 4) {
            float4 w[JT];
#pragma unroll
            for (int i = 0; i < JT; ++i) w[i] = *reinterpret_cast<const float4 *>(wbase + i * 32 * kStateWS + k);
#pragma unroll
            for (int r = 0; r < kStateRT; ++r) {
                const float4 e = *reinterpret_cast<const float4 *>(ebase + r * epad + k);
#pragma unroll
                for (int i = 0; i < JT; ++i) {
                    acc[i][r] = fmaf(w[i].x, e.x, acc[i][r]);
                    acc[i][r] = fmaf(w[i].y, e.y, acc[i][r]);
                    acc[i][r] = fmaf(w[i].z, e.z, acc[i][r]);
                    acc[i][r] = fmaf(w[i].w, e.w, acc[i][r]);
                }
            }
        }
        if (last_k) {
#pragma unroll
            for (int i = 0; i < JT; ++i) {
                const int j = j0 + lane + 32 * i;
                const float bias = j < H ? p.b1[j] : 0.0f;
#pragma unroll
                for (int r = 0; r < kStateRT; ++r) HID[(grp * kStateRT + r) * (hpad + 1) + j] = j < H ? fmaxf(acc[i][r] + bias, 0.0f) : 0.0f;
            }
        }
        if (more) store_tile((t + 1) & 1);  // the other buffer: last read before the previous barrier
        __syncthreads();
        if (last_k) {
            kt = 0;
            j0 += JL;
        } else {
            ++kt;
        }
    }
    for (int o = tid; o < kStateRows * S; o += kStateThreads) {
        const int r = o / S, s = o - r * S;
        if (r >= n_rows) continue;
        const float *h = HID + r * (hpad + 1);
        const float *w = p.w2 + static_cast<int64_t>(s) * H;
        float a = 0.0f;
        for (int j = 0; j < H; ++j) a = fmaf(__ldg(w + j), h[j], a);
        const float pr = sigmoid_rn(a + p.b2[s]);  // head.py:247
        if (p.state_out) p.state_out[(g0 + r) * p.state_stride + s] = pr;
        PROB[r * 64 + s] = pr;
    }
    if (p.id_out) {
        __syncthreads();
        if (tid < n_rows && tid < kStateRows) {
            float best = PROB[tid * 64];
            int bi = 0;
            for (int s = 1; s < S; ++s)
                if (PROB[tid * 64 + s] > best) { best = PROB[tid * 64 + s]; bi = s; }  // strict >: first maximum (torch.argmax)
            p.id_out[(g0 + tid) * p.id_stride] = static_cast<float>(bi);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Resident variant (the usual JDE sizes: E % 4 == 0, H <= 128, S <= 32, everything fits 220 KB of shared memory):
// persistent CTAs copy W1 ONCE into shared memory with cp.async (no per-tile round trips to L2), then every warp
// computes whole groups of kResRows kept rows with a JT x kResRows register tile: per 4 columns JT + kResRows 128-bit shared loads feed
// 4*JT*kResRows FMAs.  Layer 2 runs out of registers: per state class 4 FMAs per lane and a butterfly reduction.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kResWarps = 12;
constexpr int kResRows = 4;    // rows per warp pass (12 warps x 4 rows: three warps per scheduler hide the FMA and shared-memory latency)

__host__ __device__ inline int state_head_resident_smem_floats(int embed_dim, int n_state, int jt) {
    const int hp = 32 * jt;
    return hp * (embed_dim + 4) + n_state * hp + hp + 32 + kResWarps * kResRows * embed_dim;
}

__device__ __forceinline__ void cp_async_16(void *smem_dst, const void *gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}

template <int JT>
__global__ void __launch_bounds__(kResWarps * 32, 1) k7_state_head_resident(const __grid_constant__ StateHeadParams p) {
    constexpr int HP = 32 * JT;
    extern __shared__ __align__(16) float sh_state[];
    const int E = p.embed_dim, H = p.hidden, S = p.n_state, ws = E + 4, e4 = E >> 2;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float *WS = sh_state;                 // [HP][E + 4]: conflict-free 128-bit loads for lane-major rows
    float *W2S = WS + HP * ws;            // [S][HP]
    float *B1S = W2S + S * HP;            // [HP]
    float *B2S = B1S + HP;                // [32]
    float *slab = B2S + 32 + warp * (kResRows * E);  // this warp's 8 embedding rows

    for (int j = warp; j < H; j += kResWarps)  // a warp copies whole rows of W1: 512-byte runs, no index division
        for (int kq = lane; kq < e4; kq += 32) cp_async_16(WS + j * ws + kq * 4, p.w1 + static_cast<int64_t>(j) * E + kq * 4);
    asm volatile("cp.async.commit_group;" ::: "memory");
    for (int idx = H * ws + tid; idx < HP * ws; idx += kResWarps * 32) WS[idx] = 0.0f;  // padding rows (H < 32*JT)
    for (int s = warp; s < S; s += kResWarps)
        for (int j = lane; j < HP; j += 32) W2S[s * HP + j] = j < H ? __ldg(p.w2 + s * H + j) : 0.0f;
    for (int j = tid; j < HP; j += kResWarps * 32) B1S[j] = j < H ? __ldg(p.b1 + j) : 0.0f;
    if (tid < 32) B2S[tid] = tid < S ? __ldg(p.b2 + tid) : 0.0f;

    const int opi = (p.max_det + kResRows - 1) / kResRows;  // octs per image
    const int n_octs = p.batch * opi;
    const int stride = kResWarps * gridDim.x;
    bool ready = false;  // W1 landed and visible
    for (int o = warp * gridDim.x + blockIdx.x;; o += stride) {
        int n_valid = 0, b = 0, row0 = 0;
        if (o < n_octs) {
            b = o / opi;
            row0 = (o - b * opi) * kResRows;
            n_valid = min(kResRows, min(p.counts[b], p.max_det) - row0);
        }
        const int64_t g0 = static_cast<int64_t>(b) * p.max_det + row0;
        if (n_valid > 0) {
            __syncwarp();
            // the embedding columns are only read in this kernel (the state columns it writes are disjoint): read-only
            // loads, eight in flight per lane before the first store
            for (int k0 = 0; k0 < E; k0 += 32) {
                float v[kResRows];
                const int k = k0 + lane;
#pragma unroll
                for (int r = 0; r < kResRows; ++r)
                    v[r] = (r < n_valid && k < E) ? __ldg(p.emb + (g0 + r) * p.emb_stride + k) : 0.0f;
                if (k < E) {
#pragma unroll
                    for (int r = 0; r < kResRows; ++r) slab[r * E + k] = v[r];
                }
            }
            __syncwarp();
        }
        if (!ready) {  // first pass of every warp, with or without work: one block barrier
            asm volatile("cp.async.wait_all;" ::: "memory");
            __syncthreads();
            ready = true;
        }
        if (o >= n_octs) break;
        if (n_valid <= 0) continue;

        float acc[JT][kResRows];
#pragma unroll
        for (int i = 0; i < JT; ++i)
#pragma unroll
            for (int r = 0; r < kResRows; ++r) acc[i][r] = 0.0f;
        const float *wb = WS + lane * ws;
#pragma unroll 2
        for (int k = 0; k < E; k += 4) {
            float4 w[JT];
#pragma unroll
            for (int i = 0; i < JT; ++i) w[i] = *reinterpret_cast<const float4 *>(wb + i * 32 * ws + k);
#pragma unroll
            for (int r = 0; r < kResRows; ++r) {
                const float4 e = *reinterpret_cast<const float4 *>(slab + r * E + k);
#pragma unroll
                for (int i = 0; i < JT; ++i) {
                    acc[i][r] = fmaf(w[i].x, e.x, acc[i][r]);
                    acc[i][r] = fmaf(w[i].y, e.y, acc[i][r]);
                    acc[i][r] = fmaf(w[i].z, e.z, acc[i][r]);
                    acc[i][r] = fmaf(w[i].w, e.w, acc[i][r]);
                }
            }
        }
#pragma unroll
        for (int i = 0; i < JT; ++i) {
            const float bias = B1S[lane + 32 * i];
#pragma unroll
            for (int r = 0; r < kResRows; ++r) acc[i][r] = fmaxf(acc[i][r] + bias, 0.0f);  // ReLU; padding units stay 0
        }
        float res[kResRows];
#pragma unroll
        for (int r = 0; r < kResRows; ++r) res[r] = 0.0f;
        for (int s = 0; s < S; ++s) {
            float part[kResRows];
#pragma unroll
            for (int r = 0; r < kResRows; ++r) part[r] = 0.0f;
#pragma unroll
            for (int i = 0; i < JT; ++i) {
                const float w2 = W2S[s * HP + lane + 32 * i];
#pragma unroll
                for (int r = 0; r < kResRows; ++r) part[r] = fmaf(w2, acc[i][r], part[r]);
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1)
#pragma unroll
                for (int r = 0; r < kResRows; ++r) part[r] += __shfl_xor_sync(0xffffffffu, part[r], off);
#pragma unroll
            for (int r = 0; r < kResRows; ++r) res[r] = lane == s ? part[r] : res[r];
        }
        {
            const float b2 = B2S[lane];
#pragma unroll
            for (int r = 0; r < kResRows; ++r) {
                const float pr = lane < S ? sigmoid_rn(res[r] + b2) : -1.0f;  // head.py:247 (probabilities are >= 0: -1 never wins)
                if (lane < S && r < n_valid && p.state_out) p.state_out[(g0 + r) * p.state_stride + lane] = pr;
                if (p.id_out) {  // argmax over the S lanes, first maximum (torch.argmax): larger value, then lower lane
                    float bv = pr;
                    int bi = lane;
#pragma unroll
                    for (int off = 16; off > 0; off >>= 1) {
                        const float ov = __shfl_xor_sync(0xffffffffu, bv, off);
                        const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
                        if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
                    }
                    if (lane == 0 && r < n_valid) p.id_out[(g0 + r) * p.id_stride] = static_cast<float>(bi);
                }
            }
        }
    }
}

}  // namespace sarpost
