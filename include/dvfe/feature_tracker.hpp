// C++ host mirror of the reference's front-end classes over the C ABI (include/dvfe.h).
//
// Same names, argument meaning and error behaviour as
//   dynamic_vins::FeatureTracker    dynamic_vins/src/front_end/background_tracker.h:44-49
//   dynamic_vins::InstsFeatManager  dynamic_vins/src/front_end/dynamic_tracker.h:44-83
// with the OpenCV / Eigen / ROS types replaced by plain views so the header has no dependency:
//   cv::Mat (CV_8UC1)                      -> GrayImage {data, rows, cols, step}
//   Eigen::Matrix<double,7,1> (Vec7d)      -> std::array<double,7>  (layout compatible: 7 contiguous doubles)
//   FeatureBackground / FeatureInstance    -> same std::map shapes (basic/frontend_feature.h:34-73)
// The reference throws std::runtime_error for a bad settings path or empty input; so does this shim (the C ABI
// underneath returns codes).  Header only; link with -ldvfe.
#pragma once
#include <array>
#include <map>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../dvfe.h"

namespace dynamic_vins {

using Vec7d = std::array<double, 7>;

struct GrayImage {                       // a CV_8UC1 cv::Mat header
    const uint8_t* data = nullptr;
    int rows = 0, cols = 0, step = 0;
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
};

struct Box2D {                           // basic/box2d.h:24-56 (fields the path reads) + InstRoi::mask_cv
    unsigned int track_id = 0;
    int x = 0, y = 0, w = 0, h = 0;      // rect
    const uint8_t* mask = nullptr;       // roi->mask_cv, h x w, 255 = object
    int mask_step = 0;
};

struct SemanticImage {                   // basic/semantic_image.h:30-65 (fields the path reads)
    GrayImage gray0, gray1;
    double time0 = 0.0;
    unsigned int seq = 0;
    GrayImage inv_merge_mask;            // 0 = object, 255 = background
    bool exist_inst = false;
    std::vector<Box2D> boxes2d;
};

struct FeatureBackground {               // basic/frontend_feature.h:34-47
    std::map<unsigned int, std::vector<std::pair<int, Vec7d>>> points;
};

struct FeaturePoint {                    // basic/point_feature.h:22-100 (fields Output() fills)
    std::array<double, 3> point{}, point_right{};
    std::array<double, 2> vel{}, vel_right{};
    bool is_stereo = false;
    double disp = 0.0;
};

struct FeatureInstance {                 // basic/frontend_feature.h:49-58
    std::map<unsigned int, FeaturePoint> features;
};

class FeatureTracker {
public:
    // FeatureTracker(const string& config_path): fe_para::SetParameters + camera yaml files
    explicit FeatureTracker(const std::string& config_path, int device = 0, int max_instances = -1) {
        if (dvfe_config_from_yaml(config_path.c_str(), &cfg_) != DVFE_OK) throw std::runtime_error(dvfe_last_error(nullptr));
        cfg_.device = device;
        if (max_instances >= 0) cfg_.max_instances = max_instances;
        if (dvfe_create(&cfg_, &h_) != DVFE_OK) throw std::runtime_error(dvfe_last_error(nullptr));
    }
    explicit FeatureTracker(const dvfe_config& cfg) : cfg_(cfg) {
        if (dvfe_create(&cfg_, &h_) != DVFE_OK) throw std::runtime_error(dvfe_last_error(nullptr));
    }
    ~FeatureTracker() { dvfe_destroy(h_); }
    FeatureTracker(const FeatureTracker&) = delete;
    FeatureTracker& operator=(const FeatureTracker&) = delete;

    // front_end/background_tracker.cpp:52-158
    FeatureBackground TrackImage(SemanticImage& img) {
        check_input(img);
        const bool right = cfg_.stereo && !img.gray1.empty();
        if (dvfe_track_image(h_, img.gray0.data, right ? img.gray1.data : nullptr, (size_t)img.gray0.step * img.gray0.rows,
                             img.gray0.step, &img.time0) != DVFE_OK)
            throw std::runtime_error(dvfe_last_error(h_));
        return SetOutputFeats();
    }
    // front_end/background_tracker.cpp:757-837
    FeatureBackground TrackSemanticImage(SemanticImage& img) {
        check_input(img);
        const bool right = cfg_.stereo && !img.gray1.empty();
        const int exist = img.exist_inst ? 1 : 0;
        if (img.exist_inst && (img.inv_merge_mask.empty() || img.inv_merge_mask.step != img.gray0.step / channels_))
            throw std::runtime_error("TrackSemanticImage: inv_merge_mask must have the layout of gray0");
        if (dvfe_track_semantic_image(h_, img.gray0.data, right ? img.gray1.data : nullptr,
                                      img.exist_inst ? img.inv_merge_mask.data : nullptr,
                                      (size_t)img.gray0.step * img.gray0.rows, img.gray0.step, &exist,
                                      &img.time0) != DVFE_OK)
            throw std::runtime_error(dvfe_last_error(h_));
        return SetOutputFeats();
    }
    // ImageProcessor::Run moved onto the device (image_process/image_process.cpp:105-126): with SetColorInput(true) the
    // gray0/gray1 views of SemanticImage carry color0/color1 (BGR, step >= 3 * cols; inv_merge_mask stays 1 byte/px
    // with step = color step / 3); SetUndistortMaps hands over cam_s.left/right_undist_map1/2 (CV_16SC2 + CV_16UC1,
    // utils/camera_model.cpp:483-497) and every image of that camera is remapped before the gray conversion.
    void SetColorInput(bool bgr) {
        if (dvfe_set_input(h_, bgr ? 3 : 1) != DVFE_OK) throw std::runtime_error(dvfe_last_error(h_));
        channels_ = bgr ? 3 : 1;
    }
    void SetUndistortMaps(int cam, const int16_t* map1, const uint16_t* map2) {
        if (dvfe_set_undistort_maps(h_, cam, map1, map2) != DVFE_OK) throw std::runtime_error(dvfe_last_error(h_));
    }
    dvfe_tracker* handle() { return h_; }
    const dvfe_config& config() const { return cfg_; }

private:
    void check_input(const SemanticImage& img) const {
        if (img.gray0.empty() || img.gray0.cols != cfg_.width || img.gray0.rows != cfg_.height)
            throw std::runtime_error("FeatureTracker: gray0 is empty or does not match image_width/image_height");
    }
    // front_end/background_tracker.cpp:340-392
    FeatureBackground SetOutputFeats() {
        std::vector<dvfe_obs> rec(2 * (size_t)cfg_.max_cnt);
        int n = 0;
        if (dvfe_get_features(h_, 0, rec.data(), (int)rec.size(), &n) != DVFE_OK) throw std::runtime_error(dvfe_last_error(h_));
        FeatureBackground fm;
        for (int i = 0; i < n; i++) {
            Vec7d v;
            for (int k = 0; k < 7; k++) v[k] = rec[i].v[k];
            fm.points[rec[i].id].emplace_back(rec[i].cam, v);
        }
        return fm;
    }
    dvfe_config cfg_{};
    dvfe_tracker* h_ = nullptr;
    int channels_ = 1;
};

// The reference builds InstsFeatManager from the config path and shares the process-global feature-id counter with
// FeatureTracker; here both live in one dvfe_tracker, so the manager is attached to its FeatureTracker.
class InstsFeatManager {
public:
    explicit InstsFeatManager(FeatureTracker& tracker) : t_(tracker) {}

    // system/main.cpp:198-210 (reset + AddViodeInstances) + front_end/dynamic_tracker.cpp:348-493.
    // Call after FeatureTracker::TrackSemanticImage of the same frame.
    void InstsTrack(SemanticImage img) {
        std::vector<dvfe_inst_in> in(img.boxes2d.size());
        for (size_t i = 0; i < in.size(); i++) {
            const Box2D& b = img.boxes2d[i];
            in[i].track_id = b.track_id; in[i].x = b.x; in[i].y = b.y; in[i].w = b.w; in[i].h = b.h;
            in[i].mask = b.mask; in[i].mask_pitch = b.mask_step;
        }
        if (dvfe_insts_track(t_.handle(), 0, in.data(), (int)in.size(), img.time0) != DVFE_OK)
            throw std::runtime_error(dvfe_last_error(t_.handle()));
    }
    // front_end/dynamic_tracker.cpp:521-577
    std::map<unsigned int, FeatureInstance> Output() {
        const dvfe_config& c = t_.config();
        std::vector<dvfe_inst_obs> rec((size_t)(c.max_instances > 0 ? c.max_instances : 1) * (size_t)(c.max_dynamic_cnt > 0 ? c.max_dynamic_cnt : 1));
        int n = 0;
        if (dvfe_insts_output(t_.handle(), 0, rec.data(), (int)rec.size(), &n) != DVFE_OK)
            throw std::runtime_error(dvfe_last_error(t_.handle()));
        std::map<unsigned int, FeatureInstance> out;
        for (int i = 0; i < n; i++) {
            FeaturePoint f;
            for (int k = 0; k < 3; k++) { f.point[k] = rec[i].point[k]; f.point_right[k] = rec[i].point_right[k]; }
            for (int k = 0; k < 2; k++) { f.vel[k] = rec[i].vel[k]; f.vel_right[k] = rec[i].vel_right[k]; }
            f.is_stereo = rec[i].is_stereo != 0;
            f.disp = rec[i].disp;
            out[rec[i].inst_id].features[rec[i].id] = f;
        }
        return out;
    }

private:
    FeatureTracker& t_;
};

// Many cameras on one GPU: B independent FeatureTracker states advanced together, pipelined.  Not in the reference (one
// tracker configuration per process there); this is the host-side shape of the B200 deployment: one object per
// GPU, `TrackImageAsync(frame k+1)` then `Wait()` -> results of frame k, so uploads overlap the kernels.  The images of
// stream s start at `base + s * stream_stride`.
class BatchFeatureTracker {
public:
    BatchFeatureTracker(const std::string& config_path, int n_streams, int n_groups = 4, int max_instances = 0) {
        if (dvfe_config_from_yaml(config_path.c_str(), &cfg_) != DVFE_OK) throw std::runtime_error(dvfe_last_error(nullptr));
        cfg_.n_streams = n_streams; cfg_.n_groups = n_groups; cfg_.max_instances = max_instances;
        if (dvfe_create(&cfg_, &h_) != DVFE_OK) throw std::runtime_error(dvfe_last_error(nullptr));
    }
    explicit BatchFeatureTracker(const dvfe_config& cfg) : cfg_(cfg) {
        if (dvfe_create(&cfg_, &h_) != DVFE_OK) throw std::runtime_error(dvfe_last_error(nullptr));
    }
    ~BatchFeatureTracker() { dvfe_destroy(h_); }
    BatchFeatureTracker(const BatchFeatureTracker&) = delete;
    BatchFeatureTracker& operator=(const BatchFeatureTracker&) = delete;

    int streams() const { return cfg_.n_streams; }
    const dvfe_config& config() const { return cfg_; }
    dvfe_tracker* handle() { return h_; }

    // TrackImage of all cameras, enqueued; `time0` has one entry per stream; the images stay valid until the matching Wait()
    void TrackImageAsync(const uint8_t* left, const uint8_t* right, size_t stream_stride, int step, const std::vector<double>& time0) {
        check(time0);
        if (dvfe_track_image_async(h_, left, right, stream_stride, step, time0.data()) != DVFE_OK)
            throw std::runtime_error(dvfe_last_error(h_));
    }
    // TrackSemanticImage + InstsTrack of all cameras, enqueued (dynamic mode; FeatureTrack() in system/main.cpp:247-254)
    void TrackDynamicAsync(const uint8_t* left, const uint8_t* right, const uint8_t* inv_merge_mask, size_t stream_stride, int step,
                           const std::vector<int>& exist_inst, const std::vector<std::vector<Box2D>>& boxes2d,
                           const std::vector<double>& time0) {
        check(time0);
        if ((int)exist_inst.size() != cfg_.n_streams || (int)boxes2d.size() != cfg_.n_streams)
            throw std::runtime_error("BatchFeatureTracker: one exist_inst / box list per stream expected");
        std::vector<dvfe_inst_in> in;
        std::vector<int> n(boxes2d.size());
        for (size_t s = 0; s < boxes2d.size(); s++) {
            n[s] = (int)boxes2d[s].size();
            for (const Box2D& b : boxes2d[s]) in.push_back({b.track_id, b.x, b.y, b.w, b.h, b.mask, b.mask_step});
        }
        if (dvfe_track_dynamic_async(h_, left, right, inv_merge_mask, stream_stride, step, exist_inst.data(),
                                     in.empty() ? nullptr : in.data(), n.data(), time0.data()) != DVFE_OK)
            throw std::runtime_error(dvfe_last_error(h_));
    }
    // the oldest enqueued frame is finished: Features() / InstsOutput() return its results
    void Wait() {
        if (dvfe_wait(h_) != DVFE_OK) throw std::runtime_error(dvfe_last_error(h_));
    }
    // SetOutputFeats() of one camera (front_end/background_tracker.cpp:340-392)
    FeatureBackground Features(int stream) {
        std::vector<dvfe_obs> rec(2 * (size_t)cfg_.max_cnt);
        int n = 0;
        if (dvfe_get_features(h_, stream, rec.data(), (int)rec.size(), &n) != DVFE_OK) throw std::runtime_error(dvfe_last_error(h_));
        FeatureBackground fb;
        for (int i = 0; i < n; i++) {
            Vec7d v;
            for (int k = 0; k < 7; k++) v[k] = rec[i].v[k];
            fb.points[rec[i].id].emplace_back(rec[i].cam, v);
        }
        return fb;
    }
    // InstsFeatManager::Output() of one camera (front_end/dynamic_tracker.cpp:521-577)
    std::map<unsigned int, FeatureInstance> InstsOutput(int stream) {
        std::vector<dvfe_inst_obs> rec((size_t)(cfg_.max_instances > 0 ? cfg_.max_instances : 1) *
                                       (size_t)(cfg_.max_dynamic_cnt > 0 ? cfg_.max_dynamic_cnt : 1));
        int n = 0;
        if (dvfe_insts_output(h_, stream, rec.data(), (int)rec.size(), &n) != DVFE_OK) throw std::runtime_error(dvfe_last_error(h_));
        std::map<unsigned int, FeatureInstance> out;
        for (int i = 0; i < n; i++) {
            FeaturePoint f;
            for (int k = 0; k < 3; k++) { f.point[k] = rec[i].point[k]; f.point_right[k] = rec[i].point_right[k]; }
            for (int k = 0; k < 2; k++) { f.vel[k] = rec[i].vel[k]; f.vel_right[k] = rec[i].vel_right[k]; }
            f.is_stereo = rec[i].is_stereo != 0;
            f.disp = rec[i].disp;
            out[rec[i].inst_id].features[rec[i].id] = f;
        }
        return out;
    }

private:
    void check(const std::vector<double>& time0) const {
        if ((int)time0.size() != cfg_.n_streams) throw std::runtime_error("BatchFeatureTracker: one time0 per stream expected");
    }
    dvfe_config cfg_{};
    dvfe_tracker* h_ = nullptr;
};

}  // namespace dynamic_vins
