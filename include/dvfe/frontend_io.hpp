// Hand-off between the front-end and the estimator, on top of include/dvfe/feature_tracker.hpp:
//   FeatureQueue               dynamic_vins/src/basic/feature_queue.h:19-71   (bounded list, 30 ms timed request)
//   FeatureTrackFrame          dynamic_vins/src/system/main.cpp:178-330       (one iteration of the FeatureTrack() thread)
//   ToFeatureFrame<Vec7>       the map type Estimator::ProcessImage / FeatureManager::AddFeatureCheckParallax consume
//                              (dynamic_vins/src/estimator/feature_manager.cpp:61-121):
//                              std::map<unsigned, std::vector<std::pair<int, Eigen::Matrix<double,7,1>>>>
//   SerializePointFeature / DeserializePointFeature    dynamic_vins/src/utils/io/feature_serialization.cpp:26-70
// Header only.  Eigen is not a dependency: ToFeatureFrame is a template over the 7-vector type, an integrator instantiates it
// with Eigen::Matrix<double,7,1> (tests/cpp/test_frontend_io.cpp does, against a stand-in Eigen).
#pragma once
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <fstream>
#include <list>
#include <mutex>
#include <optional>
#include <sstream>

#include "feature_tracker.hpp"

namespace dynamic_vins {

constexpr int kImageQueueSize = 100;     // utils/parameters.h:48

// basic/feature_queue.h:19-71 — same members, same behaviour: push_back drops the frame when the list is full, request waits
// up to 30 ms for a frame.
class FeatureQueue {
public:
    using Ptr = std::shared_ptr<FeatureQueue>;

    void push_back(FrontendFeature& frame) {
        std::unique_lock<std::mutex> lock(queue_mutex);
        if (frame_list.size() < (size_t)kImageQueueSize) frame_list.push_back(frame);
        queue_cond.notify_one();
    }
    std::optional<FrontendFeature> request() {
        std::unique_lock<std::mutex> lock(queue_mutex);
        if (!queue_cond.wait_for(lock, std::chrono::milliseconds(30), [&] { return !frame_list.empty(); })) return std::nullopt;
        FrontendFeature frame = std::move(frame_list.front());
        frame_list.pop_front();
        return frame;
    }
    int size() {
        std::unique_lock<std::mutex> lock(queue_mutex);
        return (int)frame_list.size();
    }
    bool empty() {
        std::unique_lock<std::mutex> lock(queue_mutex);
        return frame_list.empty();
    }
    void clear() {
        std::unique_lock<std::mutex> lock(queue_mutex);
        frame_list.clear();
    }
    std::optional<double> front_time() {
        std::unique_lock<std::mutex> lock(queue_mutex);
        if (frame_list.empty()) return std::nullopt;
        return frame_list.front().time;
    }

private:
    std::mutex queue_mutex;
    std::condition_variable queue_cond;
    std::list<FrontendFeature> frame_list;
};

// FeatureBackground::points with the estimator's 7-vector type (Vec7 = Eigen::Matrix<double,7,1>): x y z u v vx vy
template <class Vec7>
std::map<unsigned int, std::vector<std::pair<int, Vec7>>> ToFeatureFrame(const FeatureBackground& features) {
    std::map<unsigned int, std::vector<std::pair<int, Vec7>>> out;
    for (const auto& kv : features.points) {
        auto& dst = out[kv.first];
        dst.reserve(kv.second.size());
        for (const auto& obs : kv.second) {
            Vec7 v;
            for (int i = 0; i < 7; i++) v[i] = obs.second[i];
            dst.emplace_back(obs.first, v);
        }
    }
    return out;
}
template <class Vec7>
FeatureBackground FromFeatureFrame(const std::map<unsigned int, std::vector<std::pair<int, Vec7>>>& points) {
    FeatureBackground out;
    for (const auto& kv : points) {
        auto& dst = out.points[kv.first];
        for (const auto& obs : kv.second) {
            Vec7d v;
            for (int i = 0; i < 7; i++) v[i] = obs.second[i];
            dst.emplace_back(obs.first, v);
        }
    }
    return out;
}

// utils/io/feature_serialization.cpp:26-38: one line per feature, "0 id v0..v6" (mono) or "1 id v0..v6 w0..w6" (stereo).
// Numbers are written with 17 significant digits (the reference uses fmt's shortest round-trip form; both read back to the same
// doubles, which is what DeserializePointFeature's std::stod sees).
inline void SerializePointFeature(const std::string& path, const std::map<unsigned int, std::vector<std::pair<int, Vec7d>>>& points) {
    std::ofstream fout(path.data(), std::ios::out);
    char buf[64];
    for (const auto& kv : points) {
        fout << (kv.second.size() == 1 ? "0 " : "1 ") << kv.first;
        for (size_t o = 0; o < kv.second.size() && o < 2; o++)
            for (double v : kv.second[o].second) {
                std::snprintf(buf, sizeof(buf), " %.17g", v);
                fout << buf;
            }
        fout << std::endl;
    }
    fout.close();
}

// utils/io/feature_serialization.cpp:45-70
inline std::map<unsigned int, std::vector<std::pair<int, Vec7d>>> DeserializePointFeature(const std::string& path) {
    std::map<unsigned int, std::vector<std::pair<int, Vec7d>>> points;
    std::ifstream fin(path.data(), std::ios::in);
    std::string line;
    while (std::getline(fin, line)) {
        std::istringstream ss(line);
        std::vector<std::string> tokens;
        for (std::string t; ss >> t;) tokens.push_back(t);
        if (tokens.size() < 9) continue;
        const int id = std::stoi(tokens[1]);
        Vec7d v;
        for (int i = 0; i < 7; i++) v[i] = std::stod(tokens[2 + i]);
        points[id].push_back({0, v});
        if (tokens[0] == "1" && tokens.size() >= 16) {
            for (int i = 0; i < 7; i++) v[i] = std::stod(tokens[9 + i]);
            points[id].push_back({1, v});
        }
    }
    fin.close();
    return points;
}

enum class SlamMode { kRaw, kNaive, kDynamic };      // utils/parameters.h SLAM

// Static-instance removal of FeatureTrack() (system/main.cpp:219-242): the ROI mask of every instance the estimator marked
// static is cleared from merge_mask and set in inv_merge_mask.  The masks are the caller's writable buffers (the views in
// SemanticImage are const).  Returns the number of static instances.
inline int PunchOutStaticInstances(InstsFeatManager& insts, uint8_t* merge_mask, uint8_t* inv_merge_mask, int width, int height) {
    int static_inst_cnt = 0;
    insts.ExecInst([&](unsigned int key, InstFeat& inst) {
        auto it = insts.estimated_info.find(key);
        if (!inst.roi || !inst.box2d || it == insts.estimated_info.end() || !it->second.is_static) return;
        static_inst_cnt++;
        const GrayImage& m = inst.roi->mask_cv;
        if (dvfe_op_punch_out(merge_mask, inv_merge_mask, width, height, m.data, m.step, (int)inst.box2d->rect.tl().x,
                              (int)inst.box2d->rect.tl().y, m.cols, m.rows) != DVFE_OK)
            throw std::runtime_error(dvfe_last_error(nullptr));
    });
    return static_inst_cnt;
}

// One iteration of the FeatureTrack() thread (system/main.cpp:178-330) for a VIODE-style input (instances arrive with track
// ids in img.boxes2d): fills a FrontendFeature and, when `queue` is given, pushes it (main.cpp:300-318).
// merge_mask / inv_merge_mask: writable full-size masks behind img.merge_mask / img.inv_merge_mask, needed only when the
// estimator reported static instances (nullptr skips the punch-out).
inline FrontendFeature FeatureTrackFrame(FeatureTracker& feature_tracker, InstsFeatManager* insts_tracker, SemanticImage& img,
                                         SlamMode slam, FeatureQueue* queue = nullptr, uint8_t* merge_mask = nullptr,
                                         uint8_t* inv_merge_mask = nullptr) {
    FrontendFeature frame;
    frame.time = img.time0;
    frame.seq_id = img.seq;
    if (slam == SlamMode::kDynamic) {
        if (insts_tracker == nullptr) throw std::runtime_error("FeatureTrackFrame: dynamic mode needs an InstsFeatManager");
        for (auto& kv : insts_tracker->instances) {          // main.cpp:198-202
            kv.second.is_curr_visible = false;
            kv.second.box2d.reset();
            kv.second.box3d.reset();
        }
        insts_tracker->AddViodeInstances(img);                // :209
        if (merge_mask != nullptr && inv_merge_mask != nullptr)
            PunchOutStaticInstances(*insts_tracker, merge_mask, inv_merge_mask, feature_tracker.config().width,
                                    feature_tracker.config().height);       // :219-242
        frame.features = feature_tracker.TrackSemanticImage(img);           // :250
        insts_tracker->InstsTrack(img);                                      // :247 (the reference's second thread)
        frame.instances = insts_tracker->Output();                           // :254
    } else if (slam == SlamMode::kNaive) {
        frame.features = feature_tracker.TrackImageNaive(img);              // :276
    } else {
        frame.features = feature_tracker.TrackImage(img);                   // :283
    }
    if (queue != nullptr) queue->push_back(frame);
    return frame;
}

}  // namespace dynamic_vins
