// include/dvfe/frontend_io.hpp: the hand-off between the front-end and the estimator.
//   test_frontend_io host <tmp_dir>
//       no device needed: FeatureQueue (basic/feature_queue.h:19-71) semantics, the text round trip of
//       SerializePointFeature / DeserializePointFeature (utils/io/feature_serialization.cpp:26-70) and ToFeatureFrame with
//       Eigen::Matrix<double,7,1> (here the stand-in Eigen of oracle/shim, the reference's own Eigen in an integration).
//   test_frontend_io dynamic <config.yaml> <frames.bin> <n_frames> <out_prefix> <max_instances>
//       FeatureTrackFrame (system/main.cpp:178-330) as the producer thread, a consumer thread taking FrontendFeature frames
//       from the FeatureQueue like Estimator::ProcessImage's caller does; writes the same files as test_feature_tracker dynamic.
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <thread>

#include "dvshim_eigen.hpp"      // namespace Eigen (stand-in): Matrix<double,7,1>

#include "dvfe/frontend_io.hpp"

using namespace dynamic_vins;

#define CHECK(cond)                                                                  \
    do {                                                                             \
        if (!(cond)) { std::fprintf(stderr, "CHECK failed line %d: %s\n", __LINE__, #cond); return 1; } \
    } while (0)

static FrontendFeature make_frame(unsigned seq, int n_pts) {
    FrontendFeature f;
    f.time = 100.0 + 0.05 * seq;
    f.seq_id = seq;
    for (int i = 0; i < n_pts; i++) {
        Vec7d a, b;
        for (int k = 0; k < 7; k++) { a[k] = std::sin(0.37 * (i + 1) * (k + 1) + seq) / 3.0; b[k] = std::cos(1.1 * i + k) * 1e-3 + seq; }
        auto& obs = f.features.points[(unsigned)(3 * i + 1)];
        obs.emplace_back(0, a);
        if (i % 3 != 0) obs.emplace_back(1, b);
    }
    return f;
}

static int run_host(const char* tmp_dir) {
    // ---- FeatureQueue --------------------------------------------------------------------------------------------------
    FeatureQueue q;
    CHECK(q.empty() && q.size() == 0 && !q.front_time().has_value());
    const auto t0 = std::chrono::steady_clock::now();
    CHECK(!q.request().has_value());                               // 30 ms timed wait on an empty queue
    const double waited = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    CHECK(waited >= 25.0 && waited < 2000.0);
    for (unsigned k = 0; k < 130; k++) { FrontendFeature f = make_frame(k, 4); q.push_back(f); }
    CHECK(q.size() == kImageQueueSize);                            // frames beyond the bound are dropped, not queued
    CHECK(q.front_time().has_value() && *q.front_time() == 100.0);
    for (unsigned k = 0; k < 100; k++) {                           // FIFO
        auto f = q.request();
        CHECK(f.has_value() && f->seq_id == k && f->features.points.size() == 4);
    }
    CHECK(q.empty());
    { FrontendFeature f = make_frame(7, 1); q.push_back(f); }
    q.clear();
    CHECK(q.empty());
    // producer / consumer threads
    std::atomic<int> got{0};
    std::atomic<bool> in_order{true};
    std::thread consumer([&] {
        unsigned next = 0;
        int idle = 0;
        while (next < 50 && idle < 100) {
            auto f = q.request();
            if (!f) { idle++; continue; }
            if (f->seq_id != next) in_order = false;
            next++; got++;
        }
    });
    for (unsigned k = 0; k < 50; k++) { FrontendFeature f = make_frame(k, 2); q.push_back(f); std::this_thread::sleep_for(std::chrono::milliseconds(1)); }
    consumer.join();
    CHECK(got == 50 && in_order);

    // ---- text format -----------------------------------------------------------------------------------------------------
    const FrontendFeature fr = make_frame(3, 40);
    const std::string path = std::string(tmp_dir) + "/3_point.txt";
    SerializePointFeature(path, fr.features.points);
    const auto back = DeserializePointFeature(path);
    CHECK(back.size() == fr.features.points.size());
    for (const auto& kv : fr.features.points) {
        auto it = back.find(kv.first);
        CHECK(it != back.end() && it->second.size() == kv.second.size());
        for (size_t o = 0; o < kv.second.size(); o++) {
            CHECK(it->second[o].first == kv.second[o].first);
            for (int k = 0; k < 7; k++) CHECK(it->second[o].second[k] == kv.second[o].second[k]);     // bit-exact round trip
        }
    }
    {   // the line layout of the reference: "<0|1> <id> <7 or 14 numbers>"
        std::ifstream fin(path);
        std::string line;
        int n_lines = 0;
        while (std::getline(fin, line)) {
            std::istringstream ss(line);
            std::vector<std::string> tok;
            for (std::string t; ss >> t;) tok.push_back(t);
            CHECK(tok.size() == (tok[0] == "1" ? 16u : 9u));
            n_lines++;
        }
        CHECK(n_lines == 40);
    }

    // ---- the estimator's map type ------------------------------------------------------------------------------------------
    using EigenVec7 = Eigen::Matrix<double, 7, 1>;
    const std::map<unsigned int, std::vector<std::pair<int, EigenVec7>>> image = ToFeatureFrame<EigenVec7>(fr.features);
    CHECK(image.size() == fr.features.points.size());
    for (const auto& kv : image) {          // what FeatureManager::AddFeatureCheckParallax asserts (feature_manager.cpp:73-77)
        CHECK(kv.second[0].first == 0);
        CHECK(kv.second.size() == 1 || (kv.second.size() == 2 && kv.second[1].first == 1));
        const auto& src = fr.features.points.at(kv.first);
        for (size_t o = 0; o < kv.second.size(); o++)
            for (int k = 0; k < 7; k++) CHECK(kv.second[o].second(k) == src[o].second[k]);
    }
    const FeatureBackground again = FromFeatureFrame<EigenVec7>(image);
    CHECK(again.points == fr.features.points);
    static_assert(sizeof(Vec7d) == 7 * sizeof(double), "Vec7d is 7 contiguous doubles, like Eigen::Matrix<double,7,1>");
    std::printf("frontend_io host checks ok\n");
    return 0;
}

// ---- dynamic mode through FeatureTrackFrame + FeatureQueue ---------------------------------------------------------------
struct DynFrame {
    double time0 = 0;
    std::vector<uint8_t> g0, g1, inv;
    int exist = 0;
    std::vector<std::vector<uint8_t>> masks;
    std::vector<Box2D::Ptr> boxes;
};

static void read_dyn_frame(std::ifstream& fin, size_t P, DynFrame& f) {
    f.g0.resize(P); f.g1.resize(P); f.inv.resize(P);
    fin.read(reinterpret_cast<char*>(&f.time0), sizeof(double));
    fin.read(reinterpret_cast<char*>(f.g0.data()), (std::streamsize)P);
    fin.read(reinterpret_cast<char*>(f.g1.data()), (std::streamsize)P);
    fin.read(reinterpret_cast<char*>(f.inv.data()), (std::streamsize)P);
    int n = 0;
    fin.read(reinterpret_cast<char*>(&f.exist), 4);
    fin.read(reinterpret_cast<char*>(&n), 4);
    f.masks.resize(n);
    for (int b = 0; b < n; b++) {
        int v[5];
        fin.read(reinterpret_cast<char*>(v), 20);
        f.masks[b].resize((size_t)v[3] * v[4]);
        fin.read(reinterpret_cast<char*>(f.masks[b].data()), (std::streamsize)f.masks[b].size());
        auto box = std::make_shared<Box2D>();
        box->track_id = v[0]; box->id = b;
        box->rect = {(float)v[1], (float)v[2], (float)v[3], (float)v[4]};
        box->min_pt = {(float)v[1], (float)v[2]};
        box->max_pt = {(float)(v[1] + v[3]), (float)(v[2] + v[4])};
        box->roi = std::make_shared<InstRoi>();
        box->roi->mask_cv = {f.masks[b].data(), v[4], v[3], v[3]};
        f.boxes.push_back(box);
    }
}

static void write_frame(const char* prefix, const FrontendFeature& frame) {
    char path[512];
    std::snprintf(path, sizeof(path), "%s_%u_point.txt", prefix, frame.seq_id);
    SerializePointFeature(path, frame.features.points);
    std::snprintf(path, sizeof(path), "%s_%u_inst.txt", prefix, frame.seq_id);
    std::FILE* fo = std::fopen(path, "w");
    for (const auto& kv : frame.instances)
        for (const auto& fp : kv.second.features) {
            const FeaturePoint& f = *fp.second;
            std::fprintf(fo, "%u %u %d %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.9g %d\n", kv.first, fp.first,
                         f.is_stereo ? 1 : 0, f.point[0], f.point[1], f.point[2], f.vel[0], f.vel[1], f.point_right[0],
                         f.point_right[1], f.point_right[2], f.vel_right[0], f.vel_right[1], (double)f.disp,
                         kv.second.box2d ? kv.second.box2d->track_id : -1);
        }
    std::fclose(fo);
}

static int run_dynamic(int argc, char** argv) {
    if (argc < 7) { std::fprintf(stderr, "usage (dynamic)\n"); return 2; }
    const int n_frames = std::atoi(argv[4]);
    FeatureTracker::Ptr feature_tracker(new FeatureTracker(argv[2], 0, std::atoi(argv[6])));
    InstsFeatManager::Ptr insts_tracker = std::make_shared<InstsFeatManager>(std::string(argv[2]));
    const int W = feature_tracker->config().width, H = feature_tracker->config().height;
    std::ifstream fin(argv[3], std::ios::binary);
    std::vector<float> disp((size_t)W * H);
    for (size_t i = 0; i < disp.size(); i++) disp[i] = -(float)(i % 4099) - 1.f;
    FeatureQueue feature_queue;
    std::atomic<int> consumed{0};
    std::atomic<bool> ordered{true};
    std::thread estimator([&] {                                      // the consumer side: frames arrive whole and in order
        int idle = 0;
        while (consumed < n_frames && idle < 2000) {
            if (auto frame = feature_queue.request()) {
                if ((int)frame->seq_id != consumed) ordered = false;
                write_frame(argv[5], *frame);
                consumed++;
                idle = 0;
            } else idle++;
        }
    });
    std::vector<DynFrame> frames(n_frames);
    for (int k = 0; k < n_frames; k++) {
        DynFrame& f = frames[k];
        read_dyn_frame(fin, (size_t)W * H, f);
        SemanticImage img;
        img.time0 = f.time0; img.seq = (unsigned)k;
        img.gray0 = {f.g0.data(), H, W, W};
        img.gray1 = {f.g1.data(), H, W, W};
        img.inv_merge_mask = {f.inv.data(), H, W, W};
        img.disp = {disp.data(), H, W, W * (int)sizeof(float)};
        img.exist_inst = f.exist != 0;
        img.boxes2d = f.boxes;
        FeatureTrackFrame(*feature_tracker, insts_tracker.get(), img, SlamMode::kDynamic, &feature_queue);
    }
    estimator.join();
    if (consumed != n_frames || !ordered) { std::fprintf(stderr, "consumer saw %d of %d frames\n", (int)consumed, n_frames); return 3; }
    return 0;
}

int main(int argc, char** argv) {
    try {
        if (argc >= 3 && std::string(argv[1]) == "host") return run_host(argv[2]);
        if (argc >= 2 && std::string(argv[1]) == "dynamic") return run_dynamic(argc, argv);
    } catch (const std::exception& e) { std::fprintf(stderr, "error: %s\n", e.what()); return 1; }
    std::fprintf(stderr, "usage: test_frontend_io host <tmp_dir> | dynamic <config.yaml> <frames.bin> <n_frames> <out_prefix> <max_instances>\n");
    return 2;
}
