// C++ drop-in check: drives include/dvfe/feature_tracker.hpp (the reference-shaped C++ API) exactly like
// FeatureTrack() in dynamic_vins/src/system/main.cpp:178-330 does, on frames dumped by the python test, and writes the
// features in the reference's SerializePointFeature text format (utils/io/feature_serialization.cpp:26-38).
//   usage: test_feature_tracker <config.yaml> <frames.bin> <n_frames> <stereo 0|1> <out_prefix>
// frames.bin: n_frames x (time0 f64, gray0 H*W, [gray1 H*W])
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <string>
#include <vector>

#include "dvfe/feature_tracker.hpp"

static void write_points(const char* prefix, int stream, int k, const dynamic_vins::FeatureBackground& fb) {
    char path[512];
    if (stream < 0) std::snprintf(path, sizeof(path), "%s_%d_point.txt", prefix, k);
    else std::snprintf(path, sizeof(path), "%s_s%d_%d_point.txt", prefix, stream, k);
    std::FILE* fo = std::fopen(path, "w");
    for (const auto& kv : fb.points) {
        std::fprintf(fo, "%d %u", kv.second.size() == 1 ? 0 : 1, kv.first);
        for (const auto& obs : kv.second)
            for (double v : obs.second) std::fprintf(fo, " %.17g", v);
        std::fprintf(fo, "\n");
    }
    std::fclose(fo);
}

// batch mode: test_feature_tracker batch <config.yaml> <frames.bin> <n_frames> <n_streams> <out_prefix> <n_groups>
// frames.bin: n_frames x (n_streams x time0 f64, n_streams x gray0, n_streams x gray1); pipelined (two frames in flight)
static int run_batch(int argc, char** argv) {
    if (argc < 8) { std::fprintf(stderr, "usage (batch)\n"); return 2; }
    const int n_frames = std::atoi(argv[4]), B = std::atoi(argv[5]), G = std::atoi(argv[7]);
    dynamic_vins::BatchFeatureTracker tracker(argv[2], B, G);
    const int W = tracker.config().width, H = tracker.config().height;
    const size_t P = (size_t)W * H;
    std::ifstream fin(argv[3], std::ios::binary);
    // the frames of both in-flight steps must stay valid: keep them all
    std::vector<std::vector<uint8_t>> L(n_frames, std::vector<uint8_t>(B * P)), R(n_frames, std::vector<uint8_t>(B * P));
    std::vector<std::vector<double>> T(n_frames, std::vector<double>(B));
    for (int k = 0; k < n_frames; k++) {
        fin.read(reinterpret_cast<char*>(T[k].data()), (std::streamsize)(B * sizeof(double)));
        fin.read(reinterpret_cast<char*>(L[k].data()), (std::streamsize)(B * P));
        fin.read(reinterpret_cast<char*>(R[k].data()), (std::streamsize)(B * P));
    }
    for (int k = 0; k < n_frames; k++) {
        tracker.TrackImageAsync(L[k].data(), R[k].data(), P, W, T[k]);
        if (k > 0) {
            tracker.Wait();
            for (int s = 0; s < B; s++) write_points(argv[6], s, k - 1, tracker.Features(s));
        }
    }
    tracker.Wait();
    for (int s = 0; s < B; s++) write_points(argv[6], s, n_frames - 1, tracker.Features(s));
    return 0;
}

int main(int argc, char** argv) {
    if (argc < 6) { std::fprintf(stderr, "usage\n"); return 2; }
    if (std::string(argv[1]) == "batch") {
        try { return run_batch(argc, argv); } catch (const std::exception& e) { std::fprintf(stderr, "error: %s\n", e.what()); return 1; }
    }
    try {
        dynamic_vins::FeatureTracker tracker(argv[1]);
        const int n_frames = std::atoi(argv[3]);
        const bool stereo = std::atoi(argv[4]) != 0;
        const int W = tracker.config().width, H = tracker.config().height;
        std::ifstream fin(argv[2], std::ios::binary);
        std::vector<uint8_t> g0((size_t)W * H), g1((size_t)W * H);
        for (int k = 0; k < n_frames; k++) {
            dynamic_vins::SemanticImage img;
            fin.read(reinterpret_cast<char*>(&img.time0), sizeof(double));
            fin.read(reinterpret_cast<char*>(g0.data()), (std::streamsize)g0.size());
            img.gray0 = {g0.data(), H, W, W};
            if (stereo) {
                fin.read(reinterpret_cast<char*>(g1.data()), (std::streamsize)g1.size());
                img.gray1 = {g1.data(), H, W, W};
            }
            img.seq = (unsigned)k;
            dynamic_vins::FeatureBackground fb = tracker.TrackImage(img);
            write_points(argv[5], -1, k, fb);
        }
        // the reference throws on a wrong settings path (front_end_parameters.cpp:20-22)
        bool threw = false;
        try { dynamic_vins::FeatureTracker bad("/nonexistent/config.yaml"); } catch (const std::runtime_error&) { threw = true; }
        if (!threw) { std::fprintf(stderr, "bad config path did not throw\n"); return 3; }
    } catch (const std::exception& e) {
        std::fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
    return 0;
}
