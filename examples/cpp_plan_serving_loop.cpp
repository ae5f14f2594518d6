// examples/cpp_plan_serving_loop.cpp — a per-frame serving loop on DEVICE buffers through the C ABI (no Python, no torch).
//
// What a tracker built on the reference does once per frame — JDE head outputs -> decode -> non_max_suppression
// (ultralytics/models/yolo/jde/predict.py:29-78) — with everything that does not change between frames prepared once:
// sarpost_plan_create freezes geometry, thresholds, workspace binding, launch configuration and the TMA tensor maps;
// sarpost_plan_run then only takes the addresses of this frame's level tensors (here two input sets alternate, as a
// double-buffered producer would hand them over) and enqueues three kernels on the caller's stream.
// Build (see tests/test_abi.py::test_cpp_plan_example_builds_and_links):
//   g++ -std=c++17 -O2 -Iinclude -I/usr/local/cuda/include examples/cpp_plan_serving_loop.cpp -Lsar-yolo_b200 -lsarpost \
//       -L/usr/local/cuda/lib64 -lcudart -Wl,-rpath,$PWD/sar-yolo_b200 -o /tmp/cpp_plan_serving_loop
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include "sarpost.h"

#define CK(expr)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (expr);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            std::fprintf(stderr, "%s failed: %s\n", #expr, cudaGetErrorString(e_));                \
            return 1;                                                                              \
        }                                                                                          \
    } while (0)

#define SP(expr)                                                                                   \
    do {                                                                                           \
        int rc_ = (expr);                                                                          \
        if (rc_ != SARPOST_OK) {                                                                   \
            std::fprintf(stderr, "%s failed (%d): %s\n", #expr, rc_, sarpost_last_error());        \
            return 1;                                                                              \
        }                                                                                          \
    } while (0)

int main(int argc, char **argv) {
    const int frames = argc > 1 ? std::atoi(argv[1]) : 8;
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev < 1) {
        std::fprintf(stderr, "no usable CUDA device\n");
        return 2;
    }
    const int strides[3] = {8, 16, 32};
    const int img = 640, nc = 1, embed = 256, state = 6, batch = 1, max_det = 300;
    const int no = 4 * 16 + nc + embed + state, nm = embed + state;

    sarpost_head_t head = {};
    head.nl = 3;
    head.batch = batch;
    head.nc = nc;
    head.reg_max = 16;
    head.n_extra_raw = embed;
    head.n_extra_sigmoid = state;
    head.no = no;
    head.dtype = SARPOST_F32;
    head.layout = SARPOST_LAYOUT_CAT;

    // two sets of device level tensors holding different random logits
    std::mt19937 rng(11);
    std::normal_distribution<float> box(0.f, 2.f), cls(-4.f, 2.f), extra(0.f, 1.f);
    float *d_level[2][3];
    int64_t anchors = 0;
    for (int l = 0; l < 3; ++l) {
        const int h = img / strides[l], w = img / strides[l];
        head.h[l] = h;
        head.w[l] = w;
        head.stride[l] = static_cast<float>(strides[l]);
        anchors += static_cast<int64_t>(h) * w;
        std::vector<float> host(static_cast<size_t>(batch) * no * h * w);
        for (int set = 0; set < 2; ++set) {
            for (int b = 0; b < batch; ++b)
                for (int c = 0; c < no; ++c)
                    for (int i = 0; i < h * w; ++i)
                        host[(static_cast<size_t>(b) * no + c) * h * w + i] = c < 64 ? box(rng) : (c < 64 + nc ? cls(rng) : extra(rng));
            CK(cudaMalloc(&d_level[set][l], host.size() * sizeof(float)));
            CK(cudaMemcpy(d_level[set][l], host.data(), host.size() * sizeof(float), cudaMemcpyHostToDevice));
        }
        head.data[l] = d_level[0][l];
    }

    sarpost_nms_params_t prm = {};
    prm.conf_thres = 0.25f;
    prm.iou_thres = 0.7;
    prm.max_det = max_det;
    prm.max_nms = 30000;
    prm.max_wh = 7680.f;
    prm.workspace_clean = 1;  // a plan's workspace is prepared once and stays clean from call to call

    cudaStream_t stream;
    CK(cudaStreamCreate(&stream));
    const int64_t ws_bytes = sarpost_workspace_bytes(batch, anchors, nc, 0, max_det);
    if (ws_bytes < 0) {
        std::fprintf(stderr, "workspace query failed: %s\n", sarpost_last_error());
        return 1;
    }
    void *d_ws = nullptr;
    CK(cudaMalloc(&d_ws, static_cast<size_t>(ws_bytes)));
    SP(sarpost_workspace_prepare(d_ws, ws_bytes, batch, stream));

    sarpost_plan_t *plan = nullptr;
    SP(sarpost_plan_create(&head, &prm, d_ws, ws_bytes, &plan));

    float *d_out = nullptr;
    int32_t *d_counts = nullptr;
    CK(cudaMalloc(&d_out, static_cast<size_t>(batch) * max_det * (6 + nm) * sizeof(float)));
    CK(cudaMalloc(&d_counts, batch * sizeof(int32_t)));

    cudaEvent_t t0, t1;
    CK(cudaEventCreate(&t0));
    CK(cudaEventCreate(&t1));
    sarpost_plan_io_t io = {};
    io.out = d_out;
    io.counts = d_counts;
    int32_t kept[2] = {0, 0};
    for (int f = 0; f < frames; ++f) {
        for (int l = 0; l < 3; ++l) io.data[l] = d_level[f & 1][l];  // this frame's tensors: only the addresses change
        if (f == 2) CK(cudaEventRecord(t0, stream));              // the first frames include one-time driver work
        SP(sarpost_plan_run(plan, &io, stream));
        if (f < 2) {
            CK(cudaMemcpyAsync(&kept[f], d_counts, sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
            CK(cudaStreamSynchronize(stream));
        }
    }
    CK(cudaEventRecord(t1, stream));
    CK(cudaStreamSynchronize(stream));
    float ms = 0.f;
    if (frames > 2) CK(cudaEventElapsedTime(&ms, t0, t1));
    std::printf("frames %d: %d / %d detections in the two input sets, %.1f us per frame (device, frames 3..)\n", frames, kept[0], kept[1],
                frames > 2 ? 1e3f * ms / (frames - 2) : 0.f);

    sarpost_plan_destroy(plan);
    for (int set = 0; set < 2; ++set)
        for (int l = 0; l < 3; ++l) cudaFree(d_level[set][l]);
    cudaFree(d_out);
    cudaFree(d_counts);
    cudaFree(d_ws);
    cudaEventDestroy(t0);
    cudaEventDestroy(t1);
    cudaStreamDestroy(stream);
    return 0;
}
