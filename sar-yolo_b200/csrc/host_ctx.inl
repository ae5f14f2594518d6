// host_ctx.inl — end-to-end entry with HOST buffers (included at the end of sarpost.cu).
//
// sarpost_fused_host: H2D of the box+cls channels only (the 262 JDE extras channels of every anchor
// never cross PCIe), fused pipeline on the context's stream, D2H of counts / rows / indices; when the
// head has extras, the raw extras of the <= max_det kept rows are packed from the host tensors (pure
// memcpy, no arithmetic), sent to the device, finished there (sigmoid of the state channels,
// head.py:247) and returned inside the output rows.

namespace sarpost {

struct ExtrasFinishParams {
    const float *rows6;   // [B, max_det, 6]
    const float *extras;  // [B, max_det, nm] raw
    const int32_t *counts;
    float *out;           // [B, max_det, 6+nm]
    int32_t max_det, nm, n_extra_raw;
};

__global__ void __launch_bounds__(256) k_extras_finish(const __grid_constant__ ExtrasFinishParams p) {
    const int b = blockIdx.y;
    const int r = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (r >= p.max_det) return;
    const int64_t row = static_cast<int64_t>(b) * p.max_det + r;
    float *o = p.out + row * (6 + p.nm);
    if (r >= p.counts[b]) {  // the whole buffer travels to the caller: rows beyond the count are defined (zero), never stale
        for (int c = lane; c < 6 + p.nm; c += 32) o[c] = 0.0f;
        return;
    }
    if (lane < 6) o[lane] = p.rows6[row * 6 + lane];
    for (int c = lane; c < p.nm; c += 32) {
        const float v = p.extras[row * p.nm + c];
        o[6 + c] = c < p.n_extra_raw ? v : sigmoid_rn(v);
    }
}

struct Buf {
    void *p = nullptr;
    int64_t bytes = 0;
    bool host = false;
    int ensure(int64_t need) {
        if (need <= bytes) return SARPOST_OK;
        if (p) {
            if (host) cudaFreeHost(p); else cudaFree(p);
            p = nullptr;
            bytes = 0;
        }
        const int64_t cap = need + need / 8;
        cudaError_t e = host ? cudaMallocHost(&p, cap) : cudaMalloc(&p, cap);
        if (e != cudaSuccess) return fail(SARPOST_ECUDA, "allocation of %lld bytes failed: %s", (long long)cap, cudaGetErrorString(e));
        bytes = cap;
        return SARPOST_OK;
    }
    void release() {
        if (p) { if (host) cudaFreeHost(p); else cudaFree(p); }
        p = nullptr;
        bytes = 0;
    }
};

}  // namespace sarpost

struct sarpost_host_ctx {
    int device = 0;
    int pack_threads = 1;  // host threads packing the extras of the kept rows (see sarpost_host_ctx_create)
    cudaStream_t s_copy = nullptr, s_main = nullptr, s_fin = nullptr;
    std::vector<cudaEvent_t> ev_copied, ev_done;
    sarpost::Buf d_levels, d_ws, d_rows6, d_counts, d_kidx, d_extras, d_out;
    sarpost::Buf h_rows6, h_counts, h_kidx, h_extras;
    int64_t last_h2d = 0, last_d2h = 0;
};

extern "C" {

int32_t sarpost_host_ctx_create(int32_t device, sarpost_host_ctx_t **ctx) {
    if (!ctx) return fail(SARPOST_EINVAL, "ctx is NULL");
    CUDA_TRY(cudaSetDevice(device));
    sarpost_host_ctx *c = new sarpost_host_ctx();
    c->device = device;
    c->h_rows6.host = c->h_counts.host = c->h_kidx.host = c->h_extras.host = true;
    {
        // Packing threads: one process per GPU shares the host's cores with its LOCAL_WORLD_SIZE - 1 siblings (torchrun
        // exports it), so the share of this process is cpus / local_world, capped at 8 (more does not help a gather of
        // <= max_det rows per image) — 8 ranks x 8 threads on a 32-vCPU host was measurable oversubscription.
        // SARPOST_HOST_THREADS overrides.
        const int cpus = omp_get_num_procs();
        const int lws = env_int("LOCAL_WORLD_SIZE", 1);
        int t = cpus / (lws > 0 ? lws : 1);
        t = t < 1 ? 1 : (t > 8 ? 8 : t);
        const int forced = env_int("SARPOST_HOST_THREADS", 0);
        c->pack_threads = forced > 0 ? forced : t;
    }
    for (cudaStream_t *st : {&c->s_copy, &c->s_main, &c->s_fin}) {
        cudaError_t e = cudaStreamCreateWithFlags(st, cudaStreamNonBlocking);
        if (e != cudaSuccess) {
            sarpost_host_ctx_destroy(c);
            return fail(SARPOST_ECUDA, "cudaStreamCreate failed: %s", cudaGetErrorString(e));
        }
    }
    *ctx = c;
    return SARPOST_OK;
}

void sarpost_host_ctx_destroy(sarpost_host_ctx_t *c) {
    if (!c) return;
    cudaSetDevice(c->device);
    for (cudaStream_t st : {c->s_copy, c->s_main, c->s_fin})
        if (st) cudaStreamSynchronize(st);
    for (sarpost::Buf *b : {&c->d_levels, &c->d_ws, &c->d_rows6, &c->d_counts, &c->d_kidx, &c->d_extras, &c->d_out,
                            &c->h_rows6, &c->h_counts, &c->h_kidx, &c->h_extras})
        b->release();
    for (cudaEvent_t e : c->ev_copied) cudaEventDestroy(e);
    for (cudaEvent_t e : c->ev_done) cudaEventDestroy(e);
    for (cudaStream_t st : {c->s_copy, c->s_main, c->s_fin})
        if (st) cudaStreamDestroy(st);
    delete c;
}

int32_t sarpost_host_ctx_last_traffic(const sarpost_host_ctx_t *c, int64_t *h2d_bytes, int64_t *d2h_bytes) {
    if (!c) return fail(SARPOST_EINVAL, "ctx is NULL");
    if (h2d_bytes) *h2d_bytes = c->last_h2d;
    if (d2h_bytes) *d2h_bytes = c->last_d2h;
    return SARPOST_OK;
}

// Pipelined over chunks of images on three streams:
//   s_copy  H2D of the chunk's box+cls channels (one contiguous copy per image and level)
//   s_main  fused pipeline of the chunk, D2H of its counts / indices (/ rows when there are no extras)
//   s_fin   H2D of the packed extras of the chunk's kept rows, finish kernel, D2H of the final rows
// While the copy engine streams chunk k+1, the GPU post-processes chunk k and the host packs the extras of
// chunk k-1 (OpenMP), so the call costs about the PCIe time of the inputs.
static int32_t fused_host_impl(sarpost_host_ctx_t *c, const sarpost_head_t *head, const sarpost_nms_params_t *params,
                               float *out, int32_t *counts, int32_t *kept_index);

int32_t sarpost_fused_host(sarpost_host_ctx_t *c, const sarpost_head_t *head, const sarpost_nms_params_t *params,
                           float *out, int32_t *counts, int32_t *kept_index) {
    if (!c) return fail(SARPOST_EINVAL, "ctx is NULL");
    const int32_t rc = fused_host_impl(c, head, params, out, counts, kept_index);
    if (rc != SARPOST_OK) {
        // a failed call may have left copies / kernels in flight that reference the context's buffers (which the next
        // call may reallocate) and the caller's host pointers: drain the three streams before handing control back
        for (cudaStream_t st : {c->s_copy, c->s_main, c->s_fin})
            if (st) cudaStreamSynchronize(st);
        cudaGetLastError();
    }
    return rc;
}

static int32_t fused_host_impl(sarpost_host_ctx_t *c, const sarpost_head_t *head, const sarpost_nms_params_t *params,
                               float *out, int32_t *counts, int32_t *kept_index) {
    HeadGeom g;
    int64_t anchors = 0;
    if (int rc = fill_geom(head, &g, &anchors)) return rc;
    if (int rc = check_params(params, g.nc)) return rc;
    if (!out || !counts) return fail(SARPOST_EINVAL, "NULL output pointer");
    if (g.is_half) return fail(SARPOST_EUNSUPPORTED, "sarpost_fused_host takes fp32 host tensors");
    if (g.split) return fail(SARPOST_EUNSUPPORTED, "sarpost_fused_host takes the concatenated level tensors (layout 0)");
    CUDA_TRY(cudaSetDevice(c->device));
    const int B = g.batch, nch = 4 * kRegMax + g.nc, nm = g.n_extra_raw + g.n_extra_sig, max_det = params->max_det;
    c->last_h2d = c->last_d2h = 0;
    int launches = 0;

    // chunking: ~128 MB of input per chunk, at least 1 image (SARPOST_HOST_CHUNK_MB overrides; smaller chunks shorten the
    // part of the call behind the last copy but measured slower: cfg3 1458 img/s at 128 MB, 1423 at 16-64 MB)
    int64_t img_bytes = 0;
    for (int l = 0; l < g.nl; ++l) img_bytes += static_cast<int64_t>(nch) * g.lvl_hw[l] * 4;
    const int64_t chunk_mb = env_int("SARPOST_HOST_CHUNK_MB", 128);
    int chunk = static_cast<int>(((chunk_mb > 0 ? chunk_mb : 128) << 20) / (img_bytes > 0 ? img_bytes : 1));
    chunk = chunk < 1 ? 1 : (chunk > B ? B : chunk);
    const int n_chunks = (B + chunk - 1) / chunk;
    while (static_cast<int>(c->ev_copied.size()) < n_chunks) {
        cudaEvent_t a, b;
        CUDA_TRY(cudaEventCreateWithFlags(&a, cudaEventDisableTiming));
        CUDA_TRY(cudaEventCreateWithFlags(&b, cudaEventDisableTiming));
        c->ev_copied.push_back(a);
        c->ev_done.push_back(b);
    }

    int64_t off[kMaxLevels + 1];
    off[0] = 0;
    for (int l = 0; l < g.nl; ++l) off[l + 1] = align_up(off[l] + static_cast<int64_t>(B) * nch * g.lvl_hw[l] * 4, 256);
    if (int rc = c->d_levels.ensure(off[g.nl])) return rc;
    const int64_t ws_bytes = sarpost_workspace_bytes(chunk, anchors, g.nc, params->multi_label, max_det);
    if (ws_bytes < 0) return static_cast<int32_t>(ws_bytes);
    if (int rc = c->d_ws.ensure(ws_bytes)) return rc;
    const int64_t rows = static_cast<int64_t>(B) * max_det;
    if (int rc = c->d_rows6.ensure(rows * 6 * 4)) return rc;
    if (int rc = c->d_counts.ensure(B * 4)) return rc;
    if (int rc = c->d_kidx.ensure(rows * 4)) return rc;
    if (int rc = c->h_rows6.ensure(rows * 6 * 4)) return rc;
    if (int rc = c->h_counts.ensure(B * 4)) return rc;
    if (int rc = c->h_kidx.ensure(rows * 4)) return rc;
    if (nm > 0) {
        if (int rc = c->h_extras.ensure(rows * nm * 4)) return rc;
        if (int rc = c->d_extras.ensure(rows * nm * 4)) return rc;
        if (int rc = c->d_out.ensure(rows * (6 + nm) * 4)) return rc;
    }
    float *d_rows6 = static_cast<float *>(c->d_rows6.p);
    int32_t *d_counts = static_cast<int32_t *>(c->d_counts.p), *d_kidx = static_cast<int32_t *>(c->d_kidx.p);
    int32_t *hc = static_cast<int32_t *>(c->h_counts.p), *hk = static_cast<int32_t *>(c->h_kidx.p);

    sarpost_nms_params_t prm = *params;
    prm.out_tail_cols = 0;  // the host entry packs its own rows (6 + nm wide)
    prm.res_boxes = prm.res_embeds = nullptr;  // results layout: device entry points only
    prm.workspace_clean = 0;  // chunks of different sizes share the context's workspace: let each call zero its histogram
    // ---- enqueue copy + compute of every chunk ----
    for (int k = 0; k < n_chunks; ++k) {
        const int b0 = k * chunk, nb = (b0 + chunk <= B) ? chunk : B - b0;
        sarpost_head_t hd = *head;
        hd.batch = nb;
        hd.no = nch;
        hd.n_extra_raw = hd.n_extra_sigmoid = 0;
        for (int l = 0; l < g.nl; ++l) {
            const int64_t width = static_cast<int64_t>(nch) * g.lvl_hw[l] * 4;
            const int64_t spitch = static_cast<int64_t>(g.no) * g.lvl_hw[l] * 4;
            char *dst = static_cast<char *>(c->d_levels.p) + off[l] + b0 * width;
            // one contiguous copy per image: a pitched 2-D copy of these 0.4-27 MB rows runs at half the PCIe rate
            for (int b = 0; b < nb; ++b)
                CUDA_TRY(cudaMemcpyAsync(dst + b * width, static_cast<const char *>(head->data[l]) + (b0 + b) * spitch, width,
                                         cudaMemcpyHostToDevice, c->s_copy));
            c->last_h2d += width * nb;
            hd.data[l] = dst;
        }
        CUDA_TRY(cudaEventRecord(c->ev_copied[k], c->s_copy));
        CUDA_TRY(cudaStreamWaitEvent(c->s_main, c->ev_copied[k], 0));
        if (nm == 0)  // these rows go to the caller as they are: rows beyond counts[b] must not carry an earlier call's data
            CUDA_TRY(cudaMemsetAsync(d_rows6 + static_cast<int64_t>(b0) * max_det * 6, 0, static_cast<size_t>(nb) * max_det * 24, c->s_main));
        if (int rc = sarpost_fused(&hd, &prm, d_rows6 + static_cast<int64_t>(b0) * max_det * 6, d_counts + b0,
                                   d_kidx + static_cast<int64_t>(b0) * max_det, c->d_ws.p, c->d_ws.bytes, c->s_main))
            return rc;
        launches += g_launches;
        CUDA_TRY(cudaMemcpyAsync(hc + b0, d_counts + b0, nb * 4, cudaMemcpyDeviceToHost, c->s_main));
        CUDA_TRY(cudaMemcpyAsync(hk + static_cast<int64_t>(b0) * max_det, d_kidx + static_cast<int64_t>(b0) * max_det,
                                 static_cast<int64_t>(nb) * max_det * 4, cudaMemcpyDeviceToHost, c->s_main));
        c->last_d2h += nb * 4 + static_cast<int64_t>(nb) * max_det * 4;
        if (nm == 0) {
            CUDA_TRY(cudaMemcpyAsync(static_cast<float *>(c->h_rows6.p) + static_cast<int64_t>(b0) * max_det * 6,
                                     d_rows6 + static_cast<int64_t>(b0) * max_det * 6, static_cast<int64_t>(nb) * max_det * 24,
                                     cudaMemcpyDeviceToHost, c->s_main));
            c->last_d2h += static_cast<int64_t>(nb) * max_det * 24;
        }
        CUDA_TRY(cudaEventRecord(c->ev_done[k], c->s_main));
    }
    // ---- per chunk: wait, pack the raw extras of the kept rows from the host tensors (memcpy only), finish ----
    float *he = static_cast<float *>(c->h_extras.p);
    for (int k = 0; k < n_chunks; ++k) {
        const int b0 = k * chunk, nb = (b0 + chunk <= B) ? chunk : B - b0;
        CUDA_TRY(cudaEventSynchronize(c->ev_done[k]));
        if (nm == 0) continue;
        const int64_t n_rows = static_cast<int64_t>(nb) * max_det;
#pragma omp parallel for schedule(static) num_threads(c->pack_threads) if (n_rows >= 64 && c->pack_threads > 1)
        for (int64_t q = 0; q < n_rows; ++q) {
            const int b = b0 + static_cast<int>(q / max_det), r = static_cast<int>(q % max_det);
            if (r >= hc[b]) continue;
            const uint32_t key = static_cast<uint32_t>(hk[static_cast<int64_t>(b) * max_det + r]);
            const uint32_t anchor = key / static_cast<uint32_t>(g.nc);
            int l = 0;
            while (l + 1 < g.nl && anchor >= static_cast<uint32_t>(g.lvl_aoff[l + 1])) ++l;
            const int64_t hw = g.lvl_hw[l];
            const float *src = static_cast<const float *>(head->data[l]) + (static_cast<int64_t>(b) * g.no + nch) * hw + (anchor - g.lvl_aoff[l]);
            float *dst = he + (static_cast<int64_t>(b) * max_det + r) * nm;
            for (int m = 0; m < nm; ++m) dst[m] = src[static_cast<int64_t>(m) * hw];
        }
        const int64_t o_e = static_cast<int64_t>(b0) * max_det * nm, o_o = static_cast<int64_t>(b0) * max_det * (6 + nm);
        CUDA_TRY(cudaMemcpyAsync(static_cast<float *>(c->d_extras.p) + o_e, he + o_e, n_rows * nm * 4, cudaMemcpyHostToDevice, c->s_fin));
        c->last_h2d += n_rows * nm * 4;
        ExtrasFinishParams ep;
        ep.rows6 = d_rows6 + static_cast<int64_t>(b0) * max_det * 6;
        ep.extras = static_cast<const float *>(c->d_extras.p) + o_e;
        ep.counts = d_counts + b0;
        ep.out = static_cast<float *>(c->d_out.p) + o_o;
        ep.max_det = max_det;
        ep.nm = nm;
        ep.n_extra_raw = g.n_extra_raw;
        k_extras_finish<<<dim3((max_det + 7) / 8, nb), 256, 0, c->s_fin>>>(ep);
        ++launches;
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaMemcpyAsync(out + o_o, static_cast<float *>(c->d_out.p) + o_o, n_rows * (6 + nm) * 4, cudaMemcpyDeviceToHost, c->s_fin));
        c->last_d2h += n_rows * (6 + nm) * 4;
    }
    CUDA_TRY(cudaStreamSynchronize(c->s_fin));
    CUDA_TRY(cudaStreamSynchronize(c->s_main));
    memcpy(counts, hc, B * 4);
    if (kept_index) memcpy(kept_index, hk, rows * 4);
    if (nm == 0) memcpy(out, c->h_rows6.p, rows * 6 * 4);
    g_launches = launches;
    return SARPOST_OK;
}

}  // extern "C"

// ================================================================================================
// Software pipeline over successive batches (device buffers): sarpost_pipeline_*
// ================================================================================================
// One call of sarpost_fused is three dependent kernels: the decode kernel K1 streams the logits at the HBM rate on every
// SM, the NMS kernel is a latency chain on a handful of SMs (its CTAs need a whole SM each), the gather is small.  Back to
// back on one stream the last two leave the memory system idle; issued from two INDEPENDENT streams they still end up
// queued behind the other batch's K1, whose persistent CTAs fill every SM the moment the previous K1 drains.
//
// Scheme A ("lagged tails", the default for depth >= 2): every decode kernel goes to ONE stream, back to back — nothing
// ever sits between K1(i) and K1(i+1).  The NMS + gather of batch i go to a high-priority stream of their own behind an
// event recorded after K1(i).  By the time that event fires K1(i+1) already holds every SM, so the NMS kernel of batch i
// is launched and stays PENDING through K1(i+1); its CTAs take the first SMs that fall free when K1(i+1) drains — before
// K1(i+2), which only becomes eligible once K1(i+1) has completed — and K1(i+2) streams on what is left (tiles are handed
// out dynamically).  Results lag one decode kernel behind; throughput is one decode kernel per batch, with no gate kernel
// and no cross-stream hop on the decode path.  depth + 1 workspaces rotate (batch i+depth+1 waits for the tail of batch i).
//
// Scheme B ("gated", SARPOST_PIPE_GATED=1, and depth 1): batch i entirely on stream i % depth, K1(i+1) chained to K1(i)
// through an event plus a one-thread gate kernel that returns when every CTA of NMS(i) has checked in (or after 30 us),
// so that NMS(i) takes its SMs before K1(i+1) floods the GPU.  Costs an event hop + the gate + a launch per batch.
struct sarpost_pipeline {
    int device = 0, depth = 2;
    std::vector<cudaStream_t> streams;
    cudaEvent_t ev_in = nullptr;
    std::vector<cudaEvent_t> ev_mid, ev_tail;
    std::vector<sarpost::Buf> ws;
    std::vector<int64_t> ws_sig;   // geometry signature of the last batch that used the slot (clean-region contract)
    std::vector<char> used;
    int64_t n_submitted = 0;
    unsigned int *d_resident = nullptr;  // device counter: NMS CTAs that have started, over the pipeline's lifetime
    unsigned int nms_ctas_total = 0;     // what it will read once every NMS kernel submitted so far is resident
    cudaStream_t s_k1 = nullptr;         // scheme A: the decode stream
    bool lagged = false;
    int slots = 1;                       // workspaces / tail streams in rotation: depth (scheme B) or depth + 1 (scheme A)
};

extern "C" {

int32_t sarpost_pipeline_create(int32_t device, int32_t depth, sarpost_pipeline_t **pl) {
    if (!pl) return fail(SARPOST_EINVAL, "pl is NULL");
    if (depth < 1 || depth > 8) return fail(SARPOST_EINVAL, "depth %d outside [1, 8]", depth);
    CUDA_TRY(cudaSetDevice(device));
    sarpost_pipeline *c = new sarpost_pipeline();
    c->device = device;
    c->depth = depth;
    c->lagged = depth > 1 && !env_int("SARPOST_PIPE_GATED", 0);
    const int slots = c->lagged ? depth + 1 : depth;
    c->slots = slots;
    c->streams.assign(slots, nullptr);
    c->ev_mid.assign(slots, nullptr);
    c->ev_tail.assign(slots, nullptr);
    c->ws.resize(slots);
    c->ws_sig.assign(slots, -1);
    c->used.assign(slots, 0);
    int prio_lo = 0, prio_hi = 0;
    bool ok = cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi) == cudaSuccess &&
              cudaEventCreateWithFlags(&c->ev_in, cudaEventDisableTiming) == cudaSuccess &&
              cudaMalloc(&c->d_resident, 256) == cudaSuccess && cudaMemset(c->d_resident, 0, 256) == cudaSuccess;
    // scheme A: the tails outrank the decode stream, so a pending NMS kernel is placed before anything else when SMs fall free
    if (ok && c->lagged) ok = cudaStreamCreateWithPriority(&c->s_k1, cudaStreamNonBlocking, prio_lo) == cudaSuccess;
    for (int i = 0; ok && i < slots; ++i)
        ok = (c->lagged ? cudaStreamCreateWithPriority(&c->streams[i], cudaStreamNonBlocking, prio_hi)
                        : cudaStreamCreateWithFlags(&c->streams[i], cudaStreamNonBlocking)) == cudaSuccess &&
             cudaEventCreateWithFlags(&c->ev_mid[i], cudaEventDisableTiming) == cudaSuccess &&
             cudaEventCreateWithFlags(&c->ev_tail[i], cudaEventDisableTiming) == cudaSuccess;
    if (!ok) {
        const char *msg = cudaGetErrorString(cudaGetLastError());
        sarpost_pipeline_destroy(c);
        return fail(SARPOST_ECUDA, "pipeline stream/event creation failed: %s", msg);
    }
    *pl = c;
    return SARPOST_OK;
}

void sarpost_pipeline_destroy(sarpost_pipeline_t *c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->s_k1) cudaStreamSynchronize(c->s_k1);
    for (cudaStream_t st : c->streams)
        if (st) cudaStreamSynchronize(st);
    for (sarpost::Buf &b : c->ws) b.release();
    for (cudaEvent_t e : c->ev_mid) if (e) cudaEventDestroy(e);
    for (cudaEvent_t e : c->ev_tail) if (e) cudaEventDestroy(e);
    if (c->ev_in) cudaEventDestroy(c->ev_in);
    if (c->d_resident) cudaFree(c->d_resident);
    if (c->s_k1) cudaStreamDestroy(c->s_k1);
    for (cudaStream_t st : c->streams)
        if (st) cudaStreamDestroy(st);
    delete c;
}

int32_t sarpost_pipeline_submit(sarpost_pipeline_t *c, const sarpost_head_t *head, const sarpost_nms_params_t *params, float *out,
                                int32_t *counts, int32_t *kept_index, void *in_stream) {
    NvtxRange nvtx("sarpost_pipeline_submit");
    g_launches = 0;
    if (!c || !head || !params) return fail(SARPOST_EINVAL, "NULL argument");
    if (params->n_peers > 0) return fail(SARPOST_EUNSUPPORTED, "the pipeline does not combine with the peer_out exchange");
    CUDA_TRY(cudaSetDevice(c->device));
    HeadGeom g;
    int64_t anchors = 0;
    if (int rc = fill_geom(head, &g, &anchors)) return rc;
    const int slot = static_cast<int>(c->n_submitted % c->slots);
    cudaStream_t st = c->streams[slot];
    const int64_t need = sarpost_workspace_bytes(g.batch, anchors, g.nc, params->multi_label, params->max_det);
    if (need < 0) return static_cast<int32_t>(need);
    sarpost::Buf &ws = c->ws[slot];
    if (need > ws.bytes) {
        CUDA_TRY(cudaStreamSynchronize(st));  // about to free memory the slot's previous batch may still be using
        if (c->s_k1) CUDA_TRY(cudaStreamSynchronize(c->s_k1));
        if (int rc = ws.ensure(need)) return rc;
        c->ws_sig[slot] = -1;
    }
    sarpost_nms_params_t prm = *params;
    const int64_t sig = (static_cast<int64_t>(g.batch) << 40) ^ (anchors << 8) ^ (g.nc & 0xff) ^ (static_cast<int64_t>(params->multi_label != 0) << 62) ^
                        (static_cast<int64_t>(params->max_det) << 24);
    prm.workspace_clean = c->ws_sig[slot] == sig ? 1 : 0;  // same geometry as last time: the NMS kernel left the head of the slot zeroed
    c->ws_sig[slot] = sig;
    // the NMS kernel shares the GPU with a later batch's decode: one CTA per image instead of a cluster when the batch is
    // large enough to keep the latency chain off the critical path
    const int cl_hint = (c->depth > 1 && g.batch >= 8) ? 1 : 0;
    // the inputs are produced on the caller's stream: this batch waits for what is enqueued there so far
    CUDA_TRY(cudaEventRecord(c->ev_in, static_cast<cudaStream_t>(in_stream)));
    int rc;
    if (c->lagged) {
        // scheme A: decode on the decode stream (behind the previous batch's decode kernel by stream order), tail on the slot's
        // own stream behind ev_mid[slot]
        CUDA_TRY(cudaStreamWaitEvent(c->s_k1, c->ev_in, 0));
        if (c->used[slot]) CUDA_TRY(cudaStreamWaitEvent(c->s_k1, c->ev_tail[slot], 0));  // the slot's workspace is free again
        g_resident_counter = nullptr;
        rc = fused_impl(head, &prm, out, counts, kept_index, ws.p, ws.bytes, c->s_k1, c->ev_mid[slot], cl_hint, st);
    } else {
        CUDA_TRY(cudaStreamWaitEvent(st, c->ev_in, 0));
        // ... and its decode kernel for the decode kernel of the batch before it (on the previous stream of the rotation)
        if (c->depth > 1 && c->n_submitted > 0) {
            const int prev = static_cast<int>((c->n_submitted - 1) % c->slots);
            CUDA_TRY(cudaStreamWaitEvent(st, c->ev_mid[prev], 0));
            // ... once the previous batch's NMS kernel (launched right behind that decode kernel) has taken its SMs
            k_gate<<<1, 32, 0, st>>>(c->d_resident, c->nms_ctas_total, 30000u);
            CUDA_TRY(cudaGetLastError());
        }
        g_resident_counter = c->depth > 1 ? c->d_resident : nullptr;
        g_last_nms_ctas = 0;
        rc = fused_impl(head, &prm, out, counts, kept_index, ws.p, ws.bytes, st, c->ev_mid[slot], cl_hint);
        g_resident_counter = nullptr;
        if (c->depth > 1 && c->n_submitted > 0) ++g_launches;  // the gate kernel
        c->nms_ctas_total += static_cast<unsigned int>(g_last_nms_ctas);  // 0 when the launch itself failed
    }
    if (rc != SARPOST_OK) {
        c->ws_sig[slot] = -1;
        return rc;
    }
    CUDA_TRY(cudaEventRecord(c->ev_tail[slot], st));
    c->used[slot] = 1;
    ++c->n_submitted;
    return SARPOST_OK;
}

int32_t sarpost_pipeline_wait(sarpost_pipeline_t *c, void *stream) {
    if (!c) return fail(SARPOST_EINVAL, "pl is NULL");
    CUDA_TRY(cudaSetDevice(c->device));
    for (int i = 0; i < c->slots; ++i)  // the newest batch of every stream of the rotation
        if (c->used[i]) CUDA_TRY(cudaStreamWaitEvent(static_cast<cudaStream_t>(stream), c->ev_tail[i], 0));
    return SARPOST_OK;
}

}  // extern "C"
