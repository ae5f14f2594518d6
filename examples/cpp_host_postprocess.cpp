// examples/cpp_host_postprocess.cpp — calling libsarpost through its C ABI from C++ (no Python, no torch).
//
// The role the reference's examples/YOLOv8-LibTorch-CPP-Inference/main.cc plays for its own post-processing:
// a C++ program holding raw head outputs in HOST memory hands them to sarpost_fused_host and gets the final
// rows (x1,y1,x2,y2,conf,cls,extras) back.  Build (see tests/test_abi.py::test_cpp_example_builds):
//   g++ -std=c++17 -O2 -Iinclude examples/cpp_host_postprocess.cpp -Lsar-yolo_b200 -lsarpost \
//       -Wl,-rpath,$PWD/sar-yolo_b200 -o /tmp/cpp_host_postprocess
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include "sarpost.h"

int main(int argc, char **argv) {
    const int batch = argc > 1 ? std::atoi(argv[1]) : 2;
    const int strides[3] = {8, 16, 32};
    const int img = 640, nc = 1, embed = 256, state = 6;

    sarpost_head_t head = {};
    head.nl = 3;
    head.batch = batch;
    head.nc = nc;
    head.reg_max = 16;
    head.n_extra_raw = embed;
    head.n_extra_sigmoid = state;
    head.no = 4 * 16 + nc + embed + state;
    head.dtype = SARPOST_F32;

    std::mt19937 rng(7);
    std::normal_distribution<float> box(0.f, 2.f), cls(-4.f, 2.f), extra(0.f, 1.f);
    std::vector<std::vector<float>> levels(3);
    for (int l = 0; l < 3; ++l) {
        const int h = img / strides[l], w = img / strides[l];
        head.h[l] = h;
        head.w[l] = w;
        head.stride[l] = static_cast<float>(strides[l]);
        levels[l].resize(static_cast<size_t>(batch) * head.no * h * w);
        for (int b = 0; b < batch; ++b)
            for (int c = 0; c < head.no; ++c)
                for (int i = 0; i < h * w; ++i)
                    levels[l][(static_cast<size_t>(b) * head.no + c) * h * w + i] = c < 64 ? box(rng) : (c < 64 + nc ? cls(rng) : extra(rng));
        head.data[l] = levels[l].data();
    }

    sarpost_nms_params_t prm = {};
    prm.conf_thres = 0.25f;
    prm.iou_thres = 0.7;
    prm.max_det = 300;
    prm.max_nms = 30000;
    prm.max_wh = 7680.f;

    sarpost_host_ctx_t *ctx = nullptr;
    if (sarpost_host_ctx_create(0, &ctx) != SARPOST_OK) {
        std::fprintf(stderr, "no usable CUDA device: %s\n", sarpost_last_error());
        return 2;
    }
    const int row_len = 6 + embed + state;
    std::vector<float> out(static_cast<size_t>(batch) * prm.max_det * row_len);
    std::vector<int32_t> counts(batch);
    const int rc = sarpost_fused_host(ctx, &head, &prm, out.data(), counts.data(), nullptr);
    if (rc != SARPOST_OK) {
        std::fprintf(stderr, "sarpost_fused_host failed (%d): %s\n", rc, sarpost_last_error());
        sarpost_host_ctx_destroy(ctx);
        return 1;
    }
    int64_t h2d = 0, d2h = 0;
    sarpost_host_ctx_last_traffic(ctx, &h2d, &d2h);
    for (int b = 0; b < batch; ++b) {
        const float *r = out.data() + static_cast<size_t>(b) * prm.max_det * row_len;
        std::printf("image %d: %d detections; best = [%.1f %.1f %.1f %.1f] conf %.3f cls %.0f\n", b, counts[b], r[0], r[1], r[2], r[3], r[4], r[5]);
    }
    std::printf("H2D %lld bytes, D2H %lld bytes, %d kernel launches\n", static_cast<long long>(h2d), static_cast<long long>(d2h), sarpost_last_launch_count());
    sarpost_host_ctx_destroy(ctx);
    return 0;
}
