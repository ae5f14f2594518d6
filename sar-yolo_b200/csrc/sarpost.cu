// sarpost.cu — C ABI (include/sarpost.h) and host-side orchestration of the K1..K5 pipeline.
//
//   K1 candidates   k1_fused_tma | k1_fused_ldg | k1_decoded | k1_merge     (k1_candidates.cuh)
//   K2/K3/K4        k4_nms: histogram scan, lazy top-k selection + sort, NMS  (k2_select_sort.cuh, k4_nms.cuh)
//   K5 gather       k5_gather                                                (k4_nms.cuh)
//
//   K6 matching     k6_match: batched box_iou + match_predictions (validator)  (k6_match.cuh)
//
// No torch types, no exceptions across the ABI, no global mutable state (thread-local error string and
// instrumentation only).
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <vector>

#include <omp.h>

#include <nvtx3/nvToolsExt.h>  // header-only; ranges show up in Nsight tools, no-ops otherwise

#include "common.cuh"
#include "k1_candidates.cuh"
#include "k2_select_sort.cuh"
#include "k4_nms.cuh"
#include "k6_match.cuh"
#include "k7_state_head.cuh"

namespace sarpost {

// ------------------------------------------------------------------------------------------------
// thread-local error / instrumentation state
// ------------------------------------------------------------------------------------------------
thread_local char g_err[512] = "";
thread_local int g_launches = 0;
thread_local int g_timing = 0;  // 0 off, 1 time the last call, 2 accumulate every call until read
thread_local std::vector<cudaEvent_t> g_ev;  // 5 events per timed call: start, hist ready, K1, NMS, gather
thread_local int g_ev_calls = 0;             // complete event sets recorded (mode 1 keeps only the last)
thread_local int g_ev_valid = 0;
thread_local int g_last_nms_ctas = 0;            // CTAs of the last NMS kernel launched on this thread (pipeline gate)
thread_local unsigned int *g_resident_counter = nullptr;  // set by the pipeline around its call of fused_impl

static int fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define CUDA_TRY(expr)                                                                              \
    do {                                                                                            \
        cudaError_t _e = (expr);                                                                    \
        if (_e != cudaSuccess)                                                                      \
            return fail(SARPOST_ECUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)

static void stage_mark(int i, cudaStream_t s) {
    if (!g_timing) return;
    if (i == 0 && g_timing == 1) g_ev_calls = 0;  // mode 1: overwrite the single set
    const size_t at = static_cast<size_t>(g_ev_calls) * 5 + i;
    while (g_ev.size() <= at) {
        cudaEvent_t e;
        cudaEventCreate(&e);
        g_ev.push_back(e);
    }
    cudaEventRecord(g_ev[at], s);
    g_ev_valid = i;
    if (i == 4) ++g_ev_calls;
}

static inline int64_t align_up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }

// NVTX range for the host-side span of one entry point / stage (per-stage tracing)
struct NvtxRange {
    explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};

// ------------------------------------------------------------------------------------------------
// workspace layout
// ------------------------------------------------------------------------------------------------
struct Layout {
    int64_t box, score, key, cls, key_a, val_a, key_b, val_b, tile_count, tile_max, hist, kept_slot, total;
};

// bytes at the head of a workspace that must be zero when a call starts: per-image score histograms + K1's tile counter
static int64_t clean_region_bytes(int64_t batch) { return batch * kBuckets * 4 + 256; }

static Layout make_layout(int64_t batch, int64_t cap, int64_t tpi, int64_t max_det, bool with_cls) {
    Layout L;
    int64_t o = 0;
    auto take = [&](int64_t bytes) {
        const int64_t at = o;
        o = align_up(o + bytes, 256);
        return at;
    };
    const int64_t slots = batch * cap;
    L.hist = take(clean_region_bytes(batch));  // first: the only region with an entry contract (see workspace_clean);
                                               // histograms, then the decode kernel's tile counter
    L.box = take(slots * 16);
    L.score = take(slots * 4);
    L.key = take(slots * 4);
    L.cls = with_cls ? take(slots * 4) : -1;
    L.key_a = take(slots * 4);
    L.val_a = take(slots * 4);
    L.key_b = take(slots * 4);
    L.val_b = take(slots * 4);
    L.tile_count = take(batch * tpi * 4);
    L.tile_max = take(batch * tpi * 4);
    L.kept_slot = take(batch * max_det * 4);
    L.total = o;
    return L;
}

struct Pipeline {
    CandStore st;
    int32_t *tile_counter;  // behind the histograms, zero on entry like them
    uint32_t *key_a, *val_a, *key_b, *val_b;
    uint32_t *kept_slot;
    float *cls;
};

static int bind_workspace(void *ws, int64_t ws_bytes, int64_t batch, int64_t cap, int64_t tpi, int64_t region,
                          int64_t max_det, bool with_cls, Pipeline *P) {
    const Layout L = make_layout(batch, cap, tpi, max_det, with_cls);
    if (!ws) return fail(SARPOST_EINVAL, "workspace is NULL");
    if (reinterpret_cast<uintptr_t>(ws) % 256) return fail(SARPOST_EINVAL, "workspace must be 256-byte aligned");
    if (ws_bytes < L.total) return fail(SARPOST_EWORKSPACE, "workspace too small: %lld < %lld bytes", (long long)ws_bytes, (long long)L.total);
    char *base = static_cast<char *>(ws);
    P->st.box = reinterpret_cast<float4 *>(base + L.box);
    P->st.score = reinterpret_cast<float *>(base + L.score);
    P->st.key = reinterpret_cast<uint32_t *>(base + L.key);
    P->st.tile_count = reinterpret_cast<int32_t *>(base + L.tile_count);
    P->st.tile_max = reinterpret_cast<uint32_t *>(base + L.tile_max);
    P->st.cap = cap;
    P->st.tpi = static_cast<int32_t>(tpi);
    P->st.region = static_cast<int32_t>(region);
    P->key_a = reinterpret_cast<uint32_t *>(base + L.key_a);
    P->val_a = reinterpret_cast<uint32_t *>(base + L.val_a);
    P->key_b = reinterpret_cast<uint32_t *>(base + L.key_b);
    P->val_b = reinterpret_cast<uint32_t *>(base + L.val_b);
    P->st.hist = reinterpret_cast<int32_t *>(base + L.hist);
    P->tile_counter = reinterpret_cast<int32_t *>(base + L.hist + batch * kBuckets * 4);
    P->kept_slot = reinterpret_cast<uint32_t *>(base + L.kept_slot);
    P->cls = with_cls ? reinterpret_cast<float *>(base + L.cls) : nullptr;
    return SARPOST_OK;
}

// ------------------------------------------------------------------------------------------------
// parameter checking helpers
// ------------------------------------------------------------------------------------------------
static int check_params(const sarpost_nms_params_t *p, int nc) {
    if (!p) return fail(SARPOST_EINVAL, "params is NULL");
    if (!(p->conf_thres >= 0.f && p->conf_thres <= 1.f)) return fail(SARPOST_EINVAL, "Invalid Confidence threshold %g, valid values are between 0.0 and 1.0", p->conf_thres);
    if (!(p->iou_thres >= 0.0 && p->iou_thres <= 1.0)) return fail(SARPOST_EINVAL, "Invalid IoU %g, valid values are between 0.0 and 1.0", p->iou_thres);
    if (p->max_det < 1 || p->max_det > 4096) return fail(SARPOST_EUNSUPPORTED, "max_det %d outside [1, 4096]", p->max_det);
    if (p->max_nms < 1) return fail(SARPOST_EINVAL, "max_nms %d < 1", p->max_nms);
    if (nc < 1 || nc > SARPOST_MAX_CLASSES) return fail(SARPOST_EUNSUPPORTED, "nc %d outside [1, %d]", nc, SARPOST_MAX_CLASSES);
    if (p->n_classes < 0 || (p->n_classes > 0 && !p->classes)) return fail(SARPOST_EINVAL, "classes pointer/count mismatch");
    if (p->n_peers < 0 || p->n_peers > 8) return fail(SARPOST_EINVAL, "n_peers %d outside [0, 8]", p->n_peers);
    for (int q = 0; q < p->n_peers; ++q) {
        if (!p->peer_out[q] || !p->peer_counts[q]) return fail(SARPOST_EINVAL, "peer buffer %d is NULL", q);
        // the exchange of 6-column rows goes out as 16-byte stores (k5_gather)
        if (reinterpret_cast<uintptr_t>(p->peer_out[q]) % 16) return fail(SARPOST_EINVAL, "peer_out[%d] is not 16-byte aligned", q);
    }
    if (p->out_tail_cols < 0 || p->out_tail_cols > 4096) return fail(SARPOST_EINVAL, "out_tail_cols %d outside [0, 4096]", p->out_tail_cols);
    if (p->res_boxes && (p->n_peers > 0 || p->out_tail_cols > 0)) return fail(SARPOST_EINVAL, "res_boxes does not combine with peer_out / out_tail_cols");
    if (p->res_state_cols < 0) return fail(SARPOST_EINVAL, "res_state_cols %d < 0", p->res_state_cols);
    if (p->nms_cluster != 0 && p->nms_cluster != 1 && p->nms_cluster != 2 && p->nms_cluster != 4 && p->nms_cluster != 8)
        return fail(SARPOST_EINVAL, "nms_cluster %d is not one of 0, 1, 2, 4, 8", p->nms_cluster);
    return SARPOST_OK;
}

static void make_filter(const sarpost_nms_params_t *p, int nc, CandFilter *f) {
    memset(f, 0, sizeof(*f));
    f->conf = p->conf_thres;
    f->multi_label = (p->multi_label && nc > 1) ? 1 : 0;  // ops.py:239
    f->has_cls_filter = p->classes != nullptr ? 1 : 0;    // classes=[] keeps nothing, like (x[:,5:6]==classes).any(1)
    for (int i = 0; i < p->n_classes; ++i) {
        const int c = p->classes[i];
        if (c >= 0 && c < nc) f->cls_allow[c >> 5] |= 1u << (c & 31);
    }
}

// largest float <= thr : (double)iou > thr  <=>  iou > iou_thr_float(thr) for fp32 iou
static float iou_thr_float(double thr) {
    float t = static_cast<float>(thr);
    if (static_cast<double>(t) > thr) t = nextafterf(t, -INFINITY);
    return t;
}

static int fill_geom(const sarpost_head_t *h, HeadGeom *g, int64_t *anchors) {
    if (!h) return fail(SARPOST_EINVAL, "head is NULL");
    if (h->nl < 1 || h->nl > kMaxLevels) return fail(SARPOST_EINVAL, "nl %d outside [1, %d]", h->nl, kMaxLevels);
    if (h->reg_max != kRegMax) return fail(SARPOST_EUNSUPPORTED, "reg_max %d unsupported (only 16, head.py:39)", h->reg_max);
    if (h->batch < 1) return fail(SARPOST_EINVAL, "batch %d < 1", h->batch);
    if (h->dtype != SARPOST_F32 && h->dtype != SARPOST_F16) return fail(SARPOST_EUNSUPPORTED, "dtype %d unsupported (0 = f32, 1 = f16)", h->dtype);
    if (h->nc < 1 || h->nc > SARPOST_MAX_CLASSES) return fail(SARPOST_EUNSUPPORTED, "nc %d outside [1, %d]", h->nc, SARPOST_MAX_CLASSES);
    if (h->layout != SARPOST_LAYOUT_CAT && h->layout != SARPOST_LAYOUT_SPLIT) return fail(SARPOST_EINVAL, "layout %d unknown (0 = cat, 1 = split)", h->layout);
    const bool split = h->layout == SARPOST_LAYOUT_SPLIT;
    if (h->n_extra_raw < 0 || h->n_extra_sigmoid < 0) return fail(SARPOST_EINVAL, "negative extras count");
    if (!split && h->no < 4 * kRegMax + h->nc + h->n_extra_raw + h->n_extra_sigmoid)
        return fail(SARPOST_EINVAL, "no %d < 4*reg_max + nc + extras (%d)", h->no, 4 * kRegMax + h->nc + h->n_extra_raw + h->n_extra_sigmoid);
    memset(g, 0, sizeof(*g));
    g->nl = h->nl;
    g->batch = h->batch;
    g->no = h->no;
    g->nc = h->nc;
    g->n_extra_raw = h->n_extra_raw;
    g->n_extra_sig = h->n_extra_sigmoid;
    g->is_half = h->dtype == SARPOST_F16;
    g->split = split ? 1 : 0;
    g->emb_cl = (split && h->emb_channels_last) ? 1 : 0;
    int64_t a = 0, t = 0;
    for (int l = 0; l < h->nl; ++l) {
        if (h->h[l] < 1 || h->w[l] < 1) return fail(SARPOST_EINVAL, "level %d has empty shape %dx%d", l, h->h[l], h->w[l]);
        if (static_cast<int64_t>(h->h[l]) * h->w[l] >= (1 << 24)) return fail(SARPOST_EUNSUPPORTED, "level %d has more than 2^24 anchors", l);
        if (!h->data[l]) return fail(SARPOST_EINVAL, "level %d data pointer is NULL", l);
        if (split) {
            if (!h->cls[l]) return fail(SARPOST_EINVAL, "level %d class-branch pointer is NULL (split layout)", l);
            if (h->n_extra_raw > 0 && !h->emb[l]) return fail(SARPOST_EINVAL, "level %d embedding-branch pointer is NULL (split layout)", l);
            if (h->n_extra_sigmoid > 0 && !h->state[l]) return fail(SARPOST_EINVAL, "level %d state-branch pointer is NULL (split layout)", l);
            g->lvl_cls[l] = h->cls[l];
            g->lvl_emb[l] = h->emb[l];
            g->lvl_state[l] = h->state[l];
        }
        const int64_t hw = static_cast<int64_t>(h->h[l]) * h->w[l];
        g->lvl_tile_begin[l] = static_cast<int32_t>(t);
        g->lvl_hw[l] = static_cast<int32_t>(hw);
        g->lvl_w[l] = h->w[l];
        g->lvl_w_magic[l] = h->w[l] == 1 ? 0xffffffffu : static_cast<uint32_t>(((1ull << 32) + h->w[l] - 1) / h->w[l]);  // ceil(2^32 / W)
        g->lvl_aoff[l] = static_cast<int32_t>(a);
        g->lvl_stride[l] = h->stride[l];
        g->lvl_ptr[l] = h->data[l];
        a += hw;
        t += (hw + kTileA - 1) / kTileA;
    }
    for (int l = h->nl; l <= kMaxLevels; ++l) g->lvl_tile_begin[l] = static_cast<int32_t>(t);
    for (int l = h->nl; l <= kMaxLevels; ++l) g->lvl_aoff[l] = static_cast<int32_t>(a);
    g->tpi = static_cast<int32_t>(t);
    if (a * h->nc >= (1ll << 32)) return fail(SARPOST_EUNSUPPORTED, "anchors*nc = %lld does not fit 32 bits", (long long)(a * h->nc));
    *anchors = a;
    return SARPOST_OK;
}

// ------------------------------------------------------------------------------------------------
// TMA descriptors (driver entry point resolved through the runtime: no link against libcuda)
// ------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
    static PFN_encodeTiled fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return reinterpret_cast<PFN_encodeTiled>(p);
    }();
    return fn;
}

static bool tma_eligible(const HeadGeom &g) {
    if (4 * kRegMax + g.nc > 256) return false;
    for (int l = 0; l < g.nl; ++l) {
        if ((static_cast<int64_t>(g.lvl_hw[l]) * (g.is_half ? 2 : 4)) % 16) return false;
        if (reinterpret_cast<uintptr_t>(g.lvl_ptr[l]) % 16) return false;
        if (g.split && reinterpret_cast<uintptr_t>(g.lvl_cls[l]) % 16) return false;
    }
    return get_encode_fn() != nullptr;
}

// Per-device facts and one-time kernel attributes: queried / set on the first call that touches a device, so the
// launch path of every later call is free of cudaDeviceGetAttribute / cudaFuncSetAttribute round trips.
constexpr int kK1StaticSmemReserve = 1024;  // >= static shared memory of k1_fused_tma (barriers, ring, scan scratch)
struct DevInfo {
    int sms = 0, smem_optin = 0;
    bool ready = false;
};
constexpr int kMaxDevices = 64;
static DevInfo g_dev[kMaxDevices];
static std::mutex g_dev_mutex;

static int device_info(const DevInfo **out) {
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    if (dev < 0 || dev >= kMaxDevices) return fail(SARPOST_EUNSUPPORTED, "device ordinal %d outside [0, %d)", dev, kMaxDevices);
    DevInfo &d = g_dev[dev];
    if (!__atomic_load_n(&d.ready, __ATOMIC_ACQUIRE)) {
        std::lock_guard<std::mutex> lock(g_dev_mutex);
        if (!d.ready) {
            CUDA_TRY(cudaDeviceGetAttribute(&d.sms, cudaDevAttrMultiProcessorCount, dev));
            CUDA_TRY(cudaDeviceGetAttribute(&d.smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
            // dynamic shared memory ceilings, once per device: the decode kernel may take everything the SM offers, the NMS
            // kernel what max_det = 4096 needs
            // (the opt-in limit covers static + dynamic shared memory: leave room for the kernel's static arrays)
            CUDA_TRY(cudaFuncSetAttribute(k1_fused_tma<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, d.smem_optin - kK1StaticSmemReserve));
            CUDA_TRY(cudaFuncSetAttribute(k1_fused_tma<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, d.smem_optin - kK1StaticSmemReserve));
            const int nms_max = static_cast<int>(nms_smem_bytes(4096));
            CUDA_TRY(cudaFuncSetAttribute(k4_nms<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, nms_max));
            CUDA_TRY(cudaFuncSetAttribute(k4_nms<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, nms_max));
            CUDA_TRY(cudaFuncSetAttribute(k4_nms<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, nms_max));
            CUDA_TRY(cudaFuncSetAttribute(k4_nms<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, nms_max));
            CUDA_TRY((cudaFuncSetAttribute(k4_nms<1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, nms_max)));
            __atomic_store_n(&d.ready, true, __ATOMIC_RELEASE);
        }
    }
    *out = &d;
    return SARPOST_OK;
}

static int device_sm_count(int *sms, int *smem_optin) {
    const DevInfo *d = nullptr;
    if (int rc = device_info(&d)) return rc;
    *sms = d->sms;
    *smem_optin = d->smem_optin;
    return SARPOST_OK;
}

static int env_int(const char *name, int dflt) {
    const char *v = getenv(name);
    return v && *v ? atoi(v) : dflt;
}

// ------------------------------------------------------------------------------------------------
// stages.  Every stage is split into `*_prepare` (decide the launch geometry, encode tensor maps, fill the parameter
// blocks) and `*_launch` (enqueue): a one-off call does both, a plan (sarpost_plan_*) prepares once and only refreshes
// the addresses afterwards.
// ------------------------------------------------------------------------------------------------
struct EnvK1 {
    int force_ldg, ctas, stages, l2promo;
};
static EnvK1 read_env_k1() {
    EnvK1 e;
    e.force_ldg = env_int("SARPOST_K1_FORCE_LDG", 0);
    e.ctas = env_int("SARPOST_K1_CTAS", 0);
    e.stages = env_int("SARPOST_K1_STAGES", 0);
    e.l2promo = env_int("SARPOST_K1_L2PROMO", CU_TENSOR_MAP_L2_PROMOTION_L2_256B);
    return e;
}

struct K1Launch {
    bool use_tma;
    int grid, smem, l2promo;
    K1TmaParams tp;
    K1LdgParams lp;
    const void *enc_box[kMaxLevels], *enc_cls[kMaxLevels];  // addresses the tensor maps in `tp` were encoded for
};

// one 3-D map (anchors of the level, channels, images) per tensor the tile is assembled from; box = {kTileA, rows, 1}
static int k1_encode_level(K1Launch *k, int l) {
    const HeadGeom &g = k->tp.g;
    const int nch = 4 * kRegMax + g.nc;
    const int esz = g.is_half ? 2 : 4;
    PFN_encodeTiled enc = get_encode_fn();
    const CUtensorMapDataType dt = g.is_half ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
    const CUtensorMapL2promotion promo = static_cast<CUtensorMapL2promotion>(k->l2promo);
    auto encode = [&](CUtensorMap *m, const void *ptr, int channels_in_tensor, int rows) -> CUresult {
        const cuuint64_t dims[3] = {static_cast<cuuint64_t>(g.lvl_hw[l]), static_cast<cuuint64_t>(channels_in_tensor), static_cast<cuuint64_t>(g.batch)};
        const cuuint64_t strides[2] = {static_cast<cuuint64_t>(g.lvl_hw[l]) * esz, static_cast<cuuint64_t>(g.lvl_hw[l]) * channels_in_tensor * esz};
        const cuuint32_t box[3] = {kTileA, static_cast<cuuint32_t>(rows), 1};
        const cuuint32_t estr[3] = {1, 1, 1};
        return enc(m, dt, 3, const_cast<void *>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, promo,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    };
    CUresult r = g.split ? encode(&k->tp.maps[l], g.lvl_ptr[l], 4 * kRegMax, 4 * kRegMax) : encode(&k->tp.maps[l], g.lvl_ptr[l], g.no, nch);
    if (r == CUDA_SUCCESS && g.split) r = encode(&k->tp.maps_cls[l], g.lvl_cls[l], g.nc, g.nc);
    if (r != CUDA_SUCCESS) return fail(SARPOST_ECUDA, "cuTensorMapEncodeTiled failed for level %d (CUresult %d)", l, static_cast<int>(r));
    k->enc_box[l] = g.lvl_ptr[l];
    k->enc_cls[l] = g.lvl_cls[l];
    return SARPOST_OK;
}

static int k1_prepare(const HeadGeom &g, const CandFilter &f, const CandStore &st, int32_t *tile_counter, const EnvK1 &env, K1Launch *k) {
    const int nch = 4 * kRegMax + g.nc;
    const int esz = g.is_half ? 2 : 4;
    const int64_t stage_bytes = static_cast<int64_t>(nch) * kTileA * esz;
    int sms = 0, smem_optin = 0;
    if (int rc = device_sm_count(&sms, &smem_optin)) return rc;
    bool use_tma = tma_eligible(g) && !env.force_ldg;
    int stages = 0, ctas = 0;
    if (use_tma) {
        const int64_t sm_smem = 228 * 1024;
        // fp32 tiles: 3 CTAs x 2 stages of 33 KB fill the SM's shared memory and reach the HBM roofline; fp16 tiles
        // are half the size and the kernel turns issue-bound, so more resident warps (6 CTAs) pay off (measured)
        const int want_ctas = env.ctas > 0 ? env.ctas : (g.is_half ? 6 : 3);
        for (ctas = want_ctas; ctas >= 1; --ctas) {
            const int64_t per_cta = sm_smem / ctas - 1024 /*driver reserve*/ - 512 /*static + align*/;
            stages = static_cast<int>(per_cta / stage_bytes);
            if (stages > kMaxStages) stages = kMaxStages;
            if (stages >= 2) break;
        }
        if (env.stages > 0) stages = env.stages > kMaxStages ? kMaxStages : env.stages;
        if (ctas < 1 || stages < 1 || (stages < 2 && env.stages <= 0) || stages * stage_bytes + 128 > smem_optin - kK1StaticSmemReserve) use_tma = false;
    }
    memset(k, 0, sizeof(*k));
    k->use_tma = use_tma;
    k->l2promo = env.l2promo;
    if (use_tma) {
        K1TmaParams &p = k->tp;
        p.g = g;
        p.f = f;
        p.st = st;
        p.stages = stages;
        p.n_tiles = g.batch * g.tpi;
        p.tile_counter = tile_counter;
        for (int l = 0; l < g.nl; ++l)
            if (int rc = k1_encode_level(k, l)) return rc;
        k->smem = static_cast<int>(stages * stage_bytes + 128);
        k->grid = sms * ctas;
        if (k->grid > p.n_tiles) k->grid = p.n_tiles;
    } else {
        k->lp.g = g;
        k->lp.f = f;
        k->lp.st = st;
    }
    return SARPOST_OK;
}

// New level addresses for a prepared launch (same geometry): only the tensor maps whose address changed are re-encoded.
static int k1_refresh(K1Launch *k, const HeadGeom &g) {
    if (!k->use_tma) {
        k->lp.g = g;
        return SARPOST_OK;
    }
    if (!tma_eligible(g)) return fail(SARPOST_EINVAL, "plan: a level address is not 16-byte aligned (the plan was created for aligned tensors)");
    k->tp.g = g;
    for (int l = 0; l < g.nl; ++l)
        if (k->enc_box[l] != g.lvl_ptr[l] || k->enc_cls[l] != g.lvl_cls[l])
            if (int rc = k1_encode_level(k, l)) return rc;
    return SARPOST_OK;
}

static int k1_launch(const K1Launch &k, cudaStream_t s) {
    NvtxRange nvtx("sarpost:K1 decode+score+compact");
    if (k.use_tma) {
        if (k.tp.g.is_half) k1_fused_tma<__half><<<k.grid, kTileA, k.smem, s>>>(k.tp);
        else k1_fused_tma<float><<<k.grid, kTileA, k.smem, s>>>(k.tp);
    } else {
        const HeadGeom &g = k.lp.g;
        if (g.is_half) k1_fused_ldg<__half><<<dim3(g.tpi, g.batch), kTileA, 0, s>>>(k.lp);
        else k1_fused_ldg<float><<<dim3(g.tpi, g.batch), kTileA, 0, s>>>(k.lp);
    }
    ++g_launches;
    CUDA_TRY(cudaGetLastError());
    return SARPOST_OK;
}

static int launch_k1_fused(const HeadGeom &g, const CandFilter &f, const CandStore &st, int32_t *tile_counter, cudaStream_t s) {
    K1Launch k;
    if (int rc = k1_prepare(g, f, st, tile_counter, read_env_k1(), &k)) return rc;
    return k1_launch(k, s);
}

// where the gather kernel finds the extras of a kept row when the input is the raw level tensors (mode 1)
static void fill_extras_src(const HeadGeom &g, ExtrasSrc *ex) {
    memset(ex, 0, sizeof(*ex));
    ex->mode = 1;
    ex->nc = g.nc;
    ex->nm = g.n_extra_raw + g.n_extra_sig;
    ex->nl = g.nl;
    ex->no = g.no;
    ex->n_extra_raw = g.n_extra_raw;
    for (int l = 0; l <= kMaxLevels; ++l) ex->lvl_aoff[l] = g.lvl_aoff[l];
    for (int l = 0; l < kMaxLevels; ++l) {
        ex->lvl_hw[l] = g.lvl_hw[l];
        ex->lvl_ptr[l] = g.lvl_ptr[l];
        ex->lvl_emb[l] = g.lvl_emb[l];
        ex->lvl_state[l] = g.lvl_state[l];
    }
    ex->is_half = g.is_half;
    ex->split = g.split;
    ex->emb_cl = g.emb_cl;
}

// K2 + K4 + K5 on a filled candidate store.
typedef void (*NmsKernel)(const NmsParams);

struct TailLaunch {
    NmsParams np;
    GatherParams gp;
    NmsKernel kern;
    int cl, nms_grid, nms_smem;
    bool pdl;
    dim3 gather_grid;
};

static int tail_prepare(const Pipeline &P, int batch, const sarpost_nms_params_t *prm, int nc, const ExtrasSrc &ex, float *out,
                        int32_t *counts, int32_t *kept_index, int cl_hint, int forced_cl, TailLaunch *t) {
    NmsParams &np = t->np;
    np.st = P.st;
    np.tmp_key_a = P.key_a;
    np.tmp_val_a = P.val_a;
    np.tmp_key_b = P.key_b;
    np.tmp_val_b = P.val_b;
    np.kept_slot = P.kept_slot;
    np.counts = counts;
    np.max_det = prm->max_det;
    np.max_nms = prm->max_nms;
    np.nc = nc;
    np.cls_override = P.cls;
    np.max_wh = prm->agnostic ? 0.0f : prm->max_wh;
    np.thr = iou_thr_float(prm->iou_thres);
    np.stats = reinterpret_cast<long long *>(prm->stats);
    np.tile_counter = P.tile_counter;
    np.resident_counter = g_resident_counter;
    t->nms_smem = static_cast<int>(nms_smem_bytes(prm->max_det));
    // one CTA per image; when the batch leaves SMs idle, a cluster of 2 or 4 CTAs per image shares the work
    int sms = 0, smem_optin = 0;
    if (int rc = device_sm_count(&sms, &smem_optin)) return rc;
    int cl = batch * 4 <= sms ? 4 : (batch * 2 <= sms ? 2 : 1);
    if (prm->nms_cluster > 0) cl = prm->nms_cluster;
    if (cl_hint == 1 || cl_hint == 2 || cl_hint == 4) cl = cl_hint;
    if (forced_cl == 1 || forced_cl == 2 || forced_cl == 4 || forced_cl == 8) cl = forced_cl;
    t->cl = cl;
    t->nms_grid = batch * cl;
    t->kern = cl == 8 ? k4_nms<8> : cl == 4 ? k4_nms<4> : cl == 2 ? k4_nms<2> : k4_nms<1>;
    // More images than SMs (one CTA per image): the 64-register build lets two CTAs share an SM, which is worth its spills
    // only when CTAs would otherwise queue for SMs — cfg4 on one GPU (512 tiles) 1.29 -> 1.42 M tiles/s.  With SMs to spare
    // it is a loss where the NMS has real work (pipelined, clustered inputs: cfg3 136 k -> 127 k img/s, cfg5 425 k -> 327 k),
    // so smaller batches keep a whole SM per CTA.  SARPOST_NMS_TWO_PER_SM=0 / 2 turns it off / forces it for CL = 1.
    const int two = env_int("SARPOST_NMS_TWO_PER_SM", 1);
    if (cl == 1 && (two == 2 || (two == 1 && batch > sms)) && 2 * (t->nms_smem + 1024) <= smem_optin) t->kern = k4_nms<1, 2>;

    GatherParams &gp = t->gp;
    gp.st = P.st;
    gp.ex = ex;
    gp.kept_slot = P.kept_slot;
    gp.counts = counts;
    gp.out = out;
    gp.kept_index = kept_index;
    gp.rescale = ex.mode == 2 ? nullptr : prm->rescale;
    gp.max_det = prm->max_det;
    gp.tail_cols = prm->out_tail_cols;
    gp.n_peers = prm->n_peers;
    gp.peer_slot_offset = prm->peer_slot_offset;
    gp.res_boxes = ex.mode == 2 ? nullptr : prm->res_boxes;
    gp.res_embeds = prm->res_embeds;
    // which extras columns are the embedding: the raw ones of a head (mode 1); for a decoded prediction the caller says
    // how many trailing columns are state probabilities
    gp.res_n_raw = ex.mode == 1 ? ex.n_extra_raw : ex.nm - prm->res_state_cols;
    if (gp.res_boxes) {
        if (gp.res_n_raw < 0) return fail(SARPOST_EINVAL, "res_state_cols %d exceeds the %d extras columns", prm->res_state_cols, ex.nm);
        if (gp.res_n_raw > 0 && !gp.res_embeds) return fail(SARPOST_EINVAL, "res_boxes set but res_embeds is NULL (embedding has %d columns)", gp.res_n_raw);
    }
    for (int q = 0; q < 8; ++q) {
        gp.peer_out[q] = q < prm->n_peers ? prm->peer_out[q] : nullptr;
        gp.peer_counts[q] = q < prm->n_peers ? prm->peer_counts[q] : nullptr;
    }
    t->gather_grid = dim3((prm->max_det + kGatherWarps - 1) / kGatherWarps, batch);
    t->pdl = !env_int("SARPOST_NO_PDL", 0);
    return SARPOST_OK;
}

static int tail_launch(const TailLaunch &t, cudaStream_t s) {
    NvtxRange nvtx("sarpost:K2-K5 select+sort+nms+gather");
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(t.nms_grid);
    cfg.blockDim = dim3(kNmsThreads);
    cfg.dynamicSmemBytes = t.nms_smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = t.cl;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    // a programmatic dependent of the candidate kernel in front of it (see k4_nms: griddepcontrol.wait)
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = (t.pdl && !env_int("SARPOST_NO_PDL_NMS", 0)) ? 2 : 1;
    CUDA_TRY(cudaLaunchKernelEx(&cfg, t.kern, t.np));
    g_last_nms_ctas = t.nms_grid;
    ++g_launches;
    CUDA_TRY(cudaGetLastError());
    stage_mark(3, s);
    // the gather as a programmatic dependent of the NMS kernel (see k5_gather); per-stage timing puts an event record
    // between the two, which makes it an ordinary launch again
    cudaLaunchConfig_t gcfg;
    memset(&gcfg, 0, sizeof(gcfg));
    gcfg.gridDim = t.gather_grid;
    gcfg.blockDim = dim3(kGatherWarps * 32);
    gcfg.stream = s;
    cudaLaunchAttribute gattr;
    gattr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
    gattr.val.programmaticStreamSerializationAllowed = 1;
    gcfg.attrs = &gattr;
    gcfg.numAttrs = t.pdl ? 1 : 0;
    CUDA_TRY(cudaLaunchKernelEx(&gcfg, k5_gather, t.gp));
    ++g_launches;
    CUDA_TRY(cudaGetLastError());
    stage_mark(4, s);
    return SARPOST_OK;
}

static int run_tail(const Pipeline &P, int batch, const sarpost_nms_params_t *prm, int nc, const ExtrasSrc &ex, float *out,
                    int32_t *counts, int32_t *kept_index, cudaStream_t s, int cl_hint = 0) {
    TailLaunch t;
    if (int rc = tail_prepare(P, batch, prm, nc, ex, out, counts, kept_index, cl_hint, env_int("SARPOST_NMS_CLUSTER", 0), &t)) return rc;
    return tail_launch(t, s);
}

// the per-image score histogram must be zero before K1 accumulates into it
static int zero_hist(const Pipeline &P, int batch, const sarpost_nms_params_t *prm, cudaStream_t s) {
    if (!prm->workspace_clean)
        CUDA_TRY(cudaMemsetAsync(P.st.hist, 0, static_cast<size_t>(clean_region_bytes(batch)), s));
    stage_mark(1, s);
    return SARPOST_OK;
}

static int64_t tpi_upper_bound(int64_t anchors) { return (anchors + kTileA - 1) / kTileA + kMaxLevels; }

}  // namespace sarpost

using namespace sarpost;

// ================================================================================================
// C ABI
// ================================================================================================
extern "C" {

const char *sarpost_last_error(void) { return g_err; }
int32_t sarpost_version(void) { return SARPOST_VERSION; }
int32_t sarpost_last_launch_count(void) { return g_launches; }

int32_t sarpost_set_stage_timing(int32_t enabled) {
    g_timing = enabled == 2 ? 2 : (enabled ? 1 : 0);
    g_ev_valid = 0;
    g_ev_calls = 0;
    return SARPOST_OK;
}

int32_t sarpost_stage_times(float *ms4) {
    if (!ms4) return fail(SARPOST_EINVAL, "ms4 is NULL");
    if (!g_timing || g_ev_calls < 1) return fail(SARPOST_EINVAL, "no timed call recorded on this thread");
    // mode 1: the last call; mode 2: the MEAN over every call since timing was enabled / last read
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    CUDA_TRY(cudaEventSynchronize(g_ev[static_cast<size_t>(g_ev_calls - 1) * 5 + 4]));
    for (int c = 0; c < g_ev_calls; ++c) {
        const cudaEvent_t *e = &g_ev[static_cast<size_t>(c) * 5];
        float v;
        for (int i = 0; i < 3; ++i) {
            CUDA_TRY(cudaEventElapsedTime(&v, e[i + 1], e[i + 2]));
            acc[i] += v;
        }
        CUDA_TRY(cudaEventElapsedTime(&v, e[0], e[4]));
        acc[3] += v;
    }
    for (int i = 0; i < 4; ++i) ms4[i] = acc[i] / static_cast<float>(g_ev_calls);
    g_ev_calls = 0;
    return SARPOST_OK;
}

int64_t sarpost_workspace_bytes(int32_t batch, int64_t anchors, int32_t nc, int32_t multi_label, int32_t max_det) {
    if (batch < 1 || anchors < 1 || nc < 1 || max_det < 1) return fail(SARPOST_EINVAL, "bad workspace query");
    const int64_t nc_eff = (multi_label && nc > 1) ? nc : 1;
    const int64_t tpi = tpi_upper_bound(anchors);
    return make_layout(batch, tpi * kTileA * nc_eff, tpi, max_det, false).total;
}

int64_t sarpost_merge_workspace_bytes(int32_t n_frames, int32_t tiles_per_frame, int32_t dets_per_tile, int32_t max_det) {
    if (n_frames < 1 || tiles_per_frame < 1 || dets_per_tile < 1 || max_det < 1) return fail(SARPOST_EINVAL, "bad workspace query");
    return make_layout(n_frames, static_cast<int64_t>(tiles_per_frame) * dets_per_tile, tiles_per_frame, max_det, true).total;
}

int64_t sarpost_workspace_clean_bytes(int32_t batch) {
    if (batch < 1) return fail(SARPOST_EINVAL, "batch %d < 1", batch);
    return clean_region_bytes(batch);
}

int32_t sarpost_workspace_prepare(void *workspace, int64_t workspace_bytes, int32_t batch, void *stream) {
    if (!workspace || batch < 1) return fail(SARPOST_EINVAL, "bad workspace_prepare arguments");
    const int64_t n = clean_region_bytes(batch);
    if (workspace_bytes < n) return fail(SARPOST_EWORKSPACE, "workspace too small: %lld < %lld bytes", (long long)workspace_bytes, (long long)n);
    CUDA_TRY(cudaMemsetAsync(workspace, 0, static_cast<size_t>(n), static_cast<cudaStream_t>(stream)));
    return SARPOST_OK;
}

int32_t sarpost_decode(const sarpost_head_t *head, void *y, void *stream) {
    g_launches = 0;
    HeadGeom g;
    int64_t anchors = 0;
    if (int rc = fill_geom(head, &g, &anchors)) return rc;
    if (!y) return fail(SARPOST_EINVAL, "y is NULL");
    if (g.split) return fail(SARPOST_EUNSUPPORTED, "sarpost_decode reproduces JDE._inference, which takes the concatenated levels (layout 0)");
    DecodeYParams p;
    p.g = g;
    p.y = y;
    p.anchors = anchors;
    if (g.is_half) k_decode_y<__half><<<dim3(g.tpi, g.batch), kTileA, 0, static_cast<cudaStream_t>(stream)>>>(p);
    else k_decode_y<float><<<dim3(g.tpi, g.batch), kTileA, 0, static_cast<cudaStream_t>(stream)>>>(p);
    ++g_launches;
    CUDA_TRY(cudaGetLastError());
    return SARPOST_OK;
}

int32_t sarpost_nms_decoded(const void *prediction, int32_t batch, int32_t channels, int64_t anchors, int32_t nc,
                            const sarpost_nms_params_t *params, float *out, int32_t *counts, int32_t *kept_index,
                            void *workspace, int64_t workspace_bytes, void *stream) {
    g_launches = 0;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (int rc = check_params(params, nc)) return rc;
    if (!prediction || (!out && !params->res_boxes) || !counts) return fail(SARPOST_EINVAL, "NULL tensor pointer");
    if (batch < 1 || anchors < 1) return fail(SARPOST_EINVAL, "empty prediction (batch %d, anchors %lld)", batch, (long long)anchors);
    if (channels < 4 + nc) return fail(SARPOST_EINVAL, "channels %d < 4 + nc (%d)", channels, 4 + nc);
    if (anchors * nc >= (1ll << 32)) return fail(SARPOST_EUNSUPPORTED, "anchors*nc does not fit 32 bits");
    CandFilter f;
    make_filter(params, nc, &f);
    const int64_t nc_eff = f.multi_label ? nc : 1;
    const bool with_labels = params->labels != nullptr && params->max_labels > 0;
    if (with_labels && !params->label_counts) return fail(SARPOST_EINVAL, "labels given without label_counts");
    const int64_t tpi_anchor = (anchors + kTileA - 1) / kTileA;
    const int64_t tpi_label = with_labels ? (params->max_labels + kTileA - 1) / kTileA : 0;
    const int64_t tpi = tpi_anchor + tpi_label;
    if ((anchors + (with_labels ? params->max_labels : 0)) * nc >= (1ll << 32)) return fail(SARPOST_EUNSUPPORTED, "(anchors+labels)*nc does not fit 32 bits");
    const int64_t region = kTileA * nc_eff;
    Pipeline P;
    if (int rc = bind_workspace(workspace, workspace_bytes, batch, tpi * region, tpi, region, params->max_det, false, &P)) return rc;

    stage_mark(0, s);
    if (int rc = zero_hist(P, batch, params, s)) return rc;
    K1DecodedParams kp;
    kp.pred = prediction;
    kp.channels = channels;
    kp.nc = nc;
    kp.anchors = anchors;
    kp.f = f;
    kp.st = P.st;
    const bool pred_half = params->prediction_dtype == SARPOST_F16;
    if (params->prediction_dtype != SARPOST_F32 && !pred_half) return fail(SARPOST_EUNSUPPORTED, "prediction_dtype %d unsupported", params->prediction_dtype);
    if (pred_half) k1_decoded<__half><<<dim3(static_cast<unsigned>(tpi_anchor), batch), kTileA, 0, s>>>(kp);
    else k1_decoded<float><<<dim3(static_cast<unsigned>(tpi_anchor), batch), kTileA, 0, s>>>(kp);
    ++g_launches;
    CUDA_TRY(cudaGetLastError());
    if (with_labels) {
        K1LabelParams lp;
        lp.labels = params->labels;
        lp.label_counts = params->label_counts;
        lp.max_labels = params->max_labels;
        lp.nc = nc;
        lp.first_tile = static_cast<int32_t>(tpi_anchor);
        lp.first_anchor = static_cast<uint32_t>(anchors);
        lp.f = f;
        lp.st = P.st;
        k1_labels<<<dim3(static_cast<unsigned>(tpi_label), batch), kTileA, 0, s>>>(lp);
        ++g_launches;
        CUDA_TRY(cudaGetLastError());
    }
    stage_mark(2, s);

    ExtrasSrc ex;
    memset(&ex, 0, sizeof(ex));
    ex.mode = 0;
    ex.nc = nc;
    ex.nm = channels - 4 - nc;
    ex.pred = prediction;
    ex.channels = channels;
    ex.anchors = anchors;
    ex.is_half = pred_half ? 1 : 0;
    return run_tail(P, batch, params, nc, ex, out, counts, kept_index, s);
}

// decode + candidates, then select/sort/NMS/gather, all on stream s; `mid` (optional) is recorded right behind the decode
// kernel — the pipeline chains the next batch's decode kernel, on another stream, to it
// `s_tail` (optional, needs `mid`): the NMS + gather kernels go to that stream, ordered behind the decode kernel through `mid`.
static int fused_impl(const sarpost_head_t *head, const sarpost_nms_params_t *params, float *out, int32_t *counts,
                      int32_t *kept_index, void *workspace, int64_t workspace_bytes, cudaStream_t s, cudaEvent_t mid, int cl_hint,
                      cudaStream_t s_tail = nullptr) {
    HeadGeom g;
    int64_t anchors = 0;
    if (int rc = fill_geom(head, &g, &anchors)) return rc;
    if (int rc = check_params(params, g.nc)) return rc;
    if ((!out && params->n_peers == 0 && !params->res_boxes) || !counts) return fail(SARPOST_EINVAL, "NULL tensor pointer");
    CandFilter f;
    make_filter(params, g.nc, &f);
    const int64_t nc_eff = f.multi_label ? g.nc : 1;
    const int64_t region = kTileA * nc_eff;
    Pipeline P;
    if (int rc = bind_workspace(workspace, workspace_bytes, g.batch, g.tpi * region, g.tpi, region, params->max_det, false, &P)) return rc;

    stage_mark(0, s);
    if (int rc = zero_hist(P, g.batch, params, s)) return rc;
    if (int rc = launch_k1_fused(g, f, P.st, P.tile_counter, s)) return rc;
    stage_mark(2, s);
    if (mid) CUDA_TRY(cudaEventRecord(mid, s));
    if (s_tail && mid) {
        CUDA_TRY(cudaStreamWaitEvent(s_tail, mid, 0));
        s = s_tail;
    }
    ExtrasSrc ex;
    fill_extras_src(g, &ex);
    return run_tail(P, g.batch, params, g.nc, ex, out, counts, kept_index, s, cl_hint);
}

int32_t sarpost_fused(const sarpost_head_t *head, const sarpost_nms_params_t *params, float *out, int32_t *counts,
                      int32_t *kept_index, void *workspace, int64_t workspace_bytes, void *stream) {
    NvtxRange nvtx("sarpost_fused");
    g_launches = 0;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    return fused_impl(head, params, out, counts, kept_index, workspace, workspace_bytes, s, nullptr, 0);
}

// ------------------------------------------------------------------------------------------------
// Plans: sarpost_fused with everything that does not change between calls prepared once
// ------------------------------------------------------------------------------------------------
struct sarpost_plan {
    HeadGeom g;
    sarpost_nms_params_t prm;  // a copy; `classes` already folded into the candidate filter
    Pipeline P;
    K1Launch k1;
    TailLaunch tail;
};

int32_t sarpost_plan_create(const sarpost_head_t *head, const sarpost_nms_params_t *params, void *workspace,
                            int64_t workspace_bytes, sarpost_plan_t **plan) {
    if (!plan) return fail(SARPOST_EINVAL, "plan is NULL");
    *plan = nullptr;
    HeadGeom g;
    int64_t anchors = 0;
    if (int rc = fill_geom(head, &g, &anchors)) return rc;
    if (int rc = check_params(params, g.nc)) return rc;
    if (!params->workspace_clean) return fail(SARPOST_EINVAL, "a plan needs a prepared workspace (sarpost_workspace_prepare, workspace_clean = 1)");
    CandFilter f;
    make_filter(params, g.nc, &f);
    const int64_t region = kTileA * (f.multi_label ? g.nc : 1);
    sarpost_plan *pl = new (std::nothrow) sarpost_plan;
    if (!pl) return fail(SARPOST_EINVAL, "out of host memory");
    pl->g = g;
    pl->prm = *params;
    pl->prm.classes = nullptr;
    pl->prm.n_classes = 0;
    pl->prm.rescale = nullptr;  // per-run addresses travel in the io block
    pl->prm.stats = nullptr;
    pl->prm.res_boxes = pl->prm.res_embeds = nullptr;
    int rc = bind_workspace(workspace, workspace_bytes, g.batch, g.tpi * region, g.tpi, region, params->max_det, false, &pl->P);
    if (!rc) rc = k1_prepare(g, f, pl->P.st, pl->P.tile_counter, read_env_k1(), &pl->k1);
    ExtrasSrc ex;
    fill_extras_src(g, &ex);
    float dummy_out = 0.f;  // the real addresses arrive with every run; the consistency checks of the parameter block run here
    if (!rc) rc = tail_prepare(pl->P, g.batch, &pl->prm, g.nc, ex, &dummy_out, nullptr, nullptr, 0, env_int("SARPOST_NMS_CLUSTER", 0), &pl->tail);
    if (rc) {
        delete pl;
        return rc;
    }
    *plan = pl;
    return SARPOST_OK;
}

int32_t sarpost_plan_run(sarpost_plan_t *pl, const sarpost_plan_io_t *io, void *stream) {
    NvtxRange nvtx("sarpost_plan_run");
    g_launches = 0;
    if (!pl || !io) return fail(SARPOST_EINVAL, "plan / io is NULL");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    HeadGeom &g = pl->g;
    if ((!io->out && pl->prm.n_peers == 0 && !io->res_boxes) || !io->counts) return fail(SARPOST_EINVAL, "NULL tensor pointer");
    for (int l = 0; l < g.nl; ++l) {
        if (!io->data[l]) return fail(SARPOST_EINVAL, "level %d data pointer is NULL", l);
        g.lvl_ptr[l] = io->data[l];
        if (g.split) {
            if (!io->cls[l] || (g.n_extra_raw > 0 && !io->emb[l]) || (g.n_extra_sig > 0 && !io->state[l]))
                return fail(SARPOST_EINVAL, "level %d: a branch pointer is NULL (split layout)", l);
            g.lvl_cls[l] = io->cls[l];
            g.lvl_emb[l] = io->emb[l];
            g.lvl_state[l] = io->state[l];
        }
    }
    if (int rc = k1_refresh(&pl->k1, g)) return rc;
    TailLaunch &t = pl->tail;
    fill_extras_src(g, &t.gp.ex);
    t.np.counts = io->counts;
    t.np.stats = reinterpret_cast<long long *>(io->stats);
    t.np.resident_counter = nullptr;
    t.gp.counts = io->counts;
    t.gp.out = io->out;
    t.gp.kept_index = io->kept_index;
    t.gp.rescale = io->rescale;
    t.gp.res_boxes = io->res_boxes;
    t.gp.res_embeds = io->res_embeds;
    if (t.gp.res_boxes && t.gp.res_n_raw > 0 && !t.gp.res_embeds) return fail(SARPOST_EINVAL, "res_boxes set but res_embeds is NULL");
    if (t.gp.res_boxes && (pl->prm.n_peers > 0 || pl->prm.out_tail_cols > 0)) return fail(SARPOST_EINVAL, "res_boxes does not combine with peer_out / out_tail_cols");
    stage_mark(0, s);
    stage_mark(1, s);  // the workspace is clean by contract: no memset
    if (int rc = k1_launch(pl->k1, s)) return rc;
    stage_mark(2, s);
    return tail_launch(t, s);
}

void sarpost_plan_destroy(sarpost_plan_t *pl) { delete pl; }

int32_t sarpost_merge_tiles(const float *dets, const int32_t *det_counts, const float *origins, int32_t n_frames,
                            int32_t tiles_per_frame, int32_t dets_per_tile, int32_t row_len,
                            const sarpost_nms_params_t *params, float *out, int32_t *counts, int32_t *kept_index,
                            void *workspace, int64_t workspace_bytes, void *stream) {
    g_launches = 0;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (int rc = check_params(params, 1)) return rc;
    if (params->res_boxes) return fail(SARPOST_EUNSUPPORTED, "the results layout (res_boxes) is not available for sarpost_merge_tiles");
    if (!dets || !det_counts || !origins || (!out && params->n_peers == 0) || !counts) return fail(SARPOST_EINVAL, "NULL tensor pointer");
    if (n_frames < 1 || tiles_per_frame < 1 || dets_per_tile < 1 || row_len < 6) return fail(SARPOST_EINVAL, "bad merge geometry");
    Pipeline P;
    const int64_t cap = static_cast<int64_t>(tiles_per_frame) * dets_per_tile;
    if (cap >= (1ll << 31)) return fail(SARPOST_EUNSUPPORTED, "too many detections per frame");
    if (int rc = bind_workspace(workspace, workspace_bytes, n_frames, cap, tiles_per_frame, dets_per_tile, params->max_det, true, &P)) return rc;

    stage_mark(0, s);
    if (int rc = zero_hist(P, n_frames, params, s)) return rc;
    K1MergeParams kp;
    kp.dets = dets;
    kp.det_counts = det_counts;
    kp.origins = origins;
    kp.dets_per_tile = dets_per_tile;
    kp.row_len = row_len;
    kp.cls = P.cls;
    kp.st = P.st;
    k1_merge<<<dim3(tiles_per_frame, n_frames), 128, 0, s>>>(kp);
    ++g_launches;
    CUDA_TRY(cudaGetLastError());
    stage_mark(2, s);

    ExtrasSrc ex;
    memset(&ex, 0, sizeof(ex));
    ex.mode = 2;
    ex.nc = 1;
    ex.nm = row_len - 6;
    ex.dets = dets;
    ex.dets_per_tile = dets_per_tile;
    ex.row_len = row_len;
    return run_tail(P, n_frames, params, 1, ex, out, counts, kept_index, s);
}

int32_t sarpost_gather_extras(const sarpost_head_t *head, const int32_t *image_index, const int32_t *anchor_index,
                              int32_t n, float *out, void *stream) {
    g_launches = 0;
    HeadGeom g;
    int64_t anchors = 0;
    if (int rc = fill_geom(head, &g, &anchors)) return rc;
    if (n < 0 || (n > 0 && (!image_index || !anchor_index || !out))) return fail(SARPOST_EINVAL, "bad gather arguments");
    if (n == 0 || g.n_extra_raw + g.n_extra_sig == 0) return SARPOST_OK;
    GatherExtrasParams p;
    memset(&p, 0, sizeof(p));
    p.image_index = image_index;
    p.anchor_index = anchor_index;
    p.n = n;
    p.batch = g.batch;
    p.out = out;
    fill_extras_src(g, &p.ex);
    k_gather_extras<<<(n + kGatherWarps - 1) / kGatherWarps, kGatherWarps * 32, 0, static_cast<cudaStream_t>(stream)>>>(p);
    ++g_launches;
    CUDA_TRY(cudaGetLastError());
    return SARPOST_OK;
}

int32_t sarpost_match_predictions(const float *dets, const int32_t *det_counts, int32_t batch, int32_t max_det,
                                  int32_t row_len, const float *gt_boxes, const float *gt_cls, const int32_t *gt_counts,
                                  int32_t max_gt, const float *iouv, int32_t n_thr, uint8_t *correct,
                                  int32_t *matched_gt, int32_t tag_thr, void *stream) {
    g_launches = 0;
    if (!dets || !det_counts || !gt_boxes || !gt_cls || !gt_counts || !iouv || !correct) return fail(SARPOST_EINVAL, "NULL pointer");
    if (batch < 1 || max_det < 1 || max_gt < 1 || row_len < 6) return fail(SARPOST_EINVAL, "bad match geometry");
    if (n_thr < 1 || n_thr > kMaxThr) return fail(SARPOST_EINVAL, "n_thr %d outside [1, %d]", n_thr, kMaxThr);
    if (reinterpret_cast<uintptr_t>(gt_boxes) % 16) return fail(SARPOST_EINVAL, "gt_boxes must be 16-byte aligned");
    MatchParams p;
    memset(&p, 0, sizeof(p));
    p.dets = dets;
    p.det_counts = det_counts;
    p.max_det = max_det;
    p.row_len = row_len;
    p.gt_boxes = gt_boxes;
    p.gt_cls = gt_cls;
    p.gt_counts = gt_counts;
    p.max_gt = max_gt;
    for (int i = 0; i < n_thr; ++i) p.iouv[i] = iouv[i];
    p.n_thr = n_thr;
    p.correct = correct;
    p.matched_gt = matched_gt;
    p.tag_thr = tag_thr;
    const size_t smem = static_cast<size_t>(max_det) * 8 + static_cast<size_t>(max_gt) * 4;
    if (smem > 200 * 1024) return fail(SARPOST_EUNSUPPORTED, "max_det/max_gt too large for one CTA");
    CUDA_TRY(cudaFuncSetAttribute(k6_match, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    k6_match<<<batch, kMatchThreads, smem, static_cast<cudaStream_t>(stream)>>>(p);
    ++g_launches;
    CUDA_TRY(cudaGetLastError());
    return SARPOST_OK;
}

int32_t sarpost_match_from_iou(const float *iou, int32_t n_gt, int32_t n_det, int64_t iou_row_stride, const float *pred_cls,
                               const float *true_cls, const float *iouv, int32_t n_thr, uint8_t *correct, int32_t *matched_gt,
                               int32_t tag_thr, void *stream) {
    g_launches = 0;
    if (n_det < 0 || n_gt < 0) return fail(SARPOST_EINVAL, "negative matrix size");
    if (n_det == 0) return SARPOST_OK;
    if (!correct || !iouv || !pred_cls || (n_gt > 0 && (!iou || !true_cls))) return fail(SARPOST_EINVAL, "NULL pointer");
    if (n_thr < 1 || n_thr > kMaxThr) return fail(SARPOST_EINVAL, "n_thr %d outside [1, %d]", n_thr, kMaxThr);
    if (iou_row_stride < n_det) return fail(SARPOST_EINVAL, "iou_row_stride %lld < n_det %d", (long long)iou_row_stride, n_det);
    MatchIouParams p;
    memset(&p, 0, sizeof(p));
    p.iou = iou;
    p.ld = iou_row_stride;
    p.pred_cls = pred_cls;
    p.true_cls = true_cls;
    p.n_det = n_det;
    p.n_gt = n_gt;
    for (int i = 0; i < n_thr; ++i) p.iouv[i] = iouv[i];
    p.n_thr = n_thr;
    p.correct = correct;
    p.matched_gt = matched_gt;
    p.tag_thr = tag_thr;
    const size_t smem = static_cast<size_t>(n_det) * 8 + static_cast<size_t>(n_gt) * 4;
    if (smem > 200 * 1024) return fail(SARPOST_EUNSUPPORTED, "n_det/n_gt too large for one CTA");
    CUDA_TRY(cudaFuncSetAttribute(k6_match_iou, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    k6_match_iou<<<1, kMatchThreads, smem, static_cast<cudaStream_t>(stream)>>>(p);
    ++g_launches;
    CUDA_TRY(cudaGetLastError());
    return SARPOST_OK;
}

static int launch_state_head(StateHeadParams p, const float *w1, cudaStream_t s) {
    const int embed_dim = p.embed_dim, hidden = p.hidden, n_state = p.n_state, batch = p.batch, max_det = p.max_det;
    if (batch < 1 || max_det < 1) return fail(SARPOST_EINVAL, "bad state-head geometry");
    if (embed_dim < 1 || embed_dim > 1024 || hidden < 1 || hidden > 1024 || n_state < 1 || n_state > 64)
        return fail(SARPOST_EUNSUPPORTED, "state head %d -> %d -> %d outside (<=1024, <=1024, <=64)", embed_dim, hidden, n_state);
    p.w1_vec = (embed_dim % 4 == 0 && reinterpret_cast<uintptr_t>(w1) % 16 == 0) ? 1 : 0;
    const int jt = hidden > 64 ? 4 : hidden > 32 ? 2 : 1;
    const size_t res_smem = static_cast<size_t>(state_head_resident_smem_floats(embed_dim, n_state, jt)) * 4;
    if (p.w1_vec && hidden <= 128 && n_state <= 32 && res_smem <= 220 * 1024 && !getenv("SARPOST_STATE_TILED")) {
        // resident variant: W1 copied once per persistent CTA, warps own whole groups of kept rows
        int sms = 0, smem_optin = 0;
        if (int rc = device_sm_count(&sms, &smem_optin)) return rc;
        const int n_octs = batch * ((max_det + kResRows - 1) / kResRows);
        const int ctas = n_octs < sms ? n_octs : sms;  // persistent: warp w of CTA c takes octs w*ctas + c, + 12*ctas, ...
        void (*kern)(const StateHeadParams) = jt == 4 ? k7_state_head_resident<4> : jt == 2 ? k7_state_head_resident<2> : k7_state_head_resident<1>;
        CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        kern<<<ctas, kResWarps * 32, res_smem, s>>>(p);
        ++g_launches;
        CUDA_TRY(cudaGetLastError());
        return SARPOST_OK;
    }
    const size_t smem = static_cast<size_t>(state_head_smem_floats(embed_dim, hidden, jt)) * 4;
    if (smem > 220 * 1024) return fail(SARPOST_EUNSUPPORTED, "state head too large for one CTA's shared memory");
    const dim3 grid((max_det + kStateRows - 1) / kStateRows, batch);
    void (*kern)(const StateHeadParams) = jt == 4 ? k7_state_head<4> : jt == 2 ? k7_state_head<2> : k7_state_head<1>;
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    kern<<<grid, kStateThreads, smem, s>>>(p);
    ++g_launches;
    CUDA_TRY(cudaGetLastError());
    return SARPOST_OK;
}

int32_t sarpost_state_head(float *rows, const int32_t *counts, int32_t batch, int32_t max_det, int32_t row_len,
                           int32_t emb_col, int32_t embed_dim, int32_t state_col, int32_t n_state, int32_t hidden,
                           const float *w1, const float *b1, const float *w2, const float *b2, void *stream) {
    g_launches = 0;
    if (!rows || !counts || !w1 || !b1 || !w2 || !b2) return fail(SARPOST_EINVAL, "NULL pointer");
    if (emb_col < 0 || state_col < 0 || emb_col + embed_dim > row_len || state_col + n_state > row_len)
        return fail(SARPOST_EINVAL, "embedding / state columns outside the row (row_len %d)", row_len);
    if (state_col < emb_col + embed_dim && emb_col < state_col + n_state)
        return fail(SARPOST_EINVAL, "embedding and state columns overlap");
    StateHeadParams p;
    memset(&p, 0, sizeof(p));
    p.emb = rows + emb_col;
    p.emb_stride = row_len;
    p.state_out = rows + state_col;
    p.state_stride = row_len;
    p.counts = counts;
    p.batch = batch;
    p.max_det = max_det;
    p.embed_dim = embed_dim;
    p.n_state = n_state;
    p.hidden = hidden;
    p.w1 = w1;
    p.b1 = b1;
    p.w2 = w2;
    p.b2 = b2;
    return launch_state_head(p, w1, static_cast<cudaStream_t>(stream));
}

int32_t sarpost_state_ids(const float *embeds, const int32_t *counts, float *boxes7, int32_t batch, int32_t max_det,
                          int32_t embed_dim, int32_t n_state, int32_t hidden, const float *w1, const float *b1,
                          const float *w2, const float *b2, void *stream) {
    g_launches = 0;
    if (!embeds || !counts || !boxes7 || !w1 || !b1 || !w2 || !b2) return fail(SARPOST_EINVAL, "NULL pointer");
    StateHeadParams p;
    memset(&p, 0, sizeof(p));
    p.emb = embeds;
    p.emb_stride = embed_dim;
    p.id_out = boxes7 + 4;  // column 4 of x1,y1,x2,y2,state_id,conf,cls
    p.id_stride = 7;
    p.counts = counts;
    p.batch = batch;
    p.max_det = max_det;
    p.embed_dim = embed_dim;
    p.n_state = n_state;
    p.hidden = hidden;
    p.w1 = w1;
    p.b1 = b1;
    p.w2 = w2;
    p.b2 = b2;
    return launch_state_head(p, w1, static_cast<cudaStream_t>(stream));
}

int32_t sarpost_abi_sizes(int32_t *head_bytes, int32_t *params_bytes) {
    if (head_bytes) *head_bytes = static_cast<int32_t>(sizeof(sarpost_head_t));
    if (params_bytes) *params_bytes = static_cast<int32_t>(sizeof(sarpost_nms_params_t));
    return SARPOST_OK;
}

#ifdef SARPOST_PHASE_PROF
// debug builds only (not part of include/sarpost.h): cycles per K4 phase of block 0, accumulated
int32_t sarpost_debug_phase_cycles(unsigned long long *out16, int32_t reset) {
    CUDA_TRY(cudaDeviceSynchronize());
    if (out16) CUDA_TRY(cudaMemcpyFromSymbol(out16, g_phase, sizeof(unsigned long long) * 16));
    if (reset) {
        unsigned long long z[16] = {0};
        CUDA_TRY(cudaMemcpyToSymbol(g_phase, z, sizeof(z)));
    }
    return SARPOST_OK;
}
// out48 = cycles[16] | visits[16] | longest visit[16]
int32_t sarpost_debug_phase_detail(unsigned long long *out48, int32_t reset) {
    CUDA_TRY(cudaDeviceSynchronize());
    if (out48) {
        CUDA_TRY(cudaMemcpyFromSymbol(out48, g_phase, sizeof(unsigned long long) * 16));
        CUDA_TRY(cudaMemcpyFromSymbol(out48 + 16, g_phase_n, sizeof(unsigned long long) * 16));
        CUDA_TRY(cudaMemcpyFromSymbol(out48 + 32, g_phase_max, sizeof(unsigned long long) * 16));
    }
    if (reset) {
        unsigned long long z[16] = {0};
        CUDA_TRY(cudaMemcpyToSymbol(g_phase, z, sizeof(z)));
        CUDA_TRY(cudaMemcpyToSymbol(g_phase_n, z, sizeof(z)));
        CUDA_TRY(cudaMemcpyToSymbol(g_phase_max, z, sizeof(z)));
    }
    return SARPOST_OK;
}
#endif

}  // extern "C"

#include "host_ctx.inl"
