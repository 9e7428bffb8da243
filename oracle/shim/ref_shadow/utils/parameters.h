// ORACLE shim (test infrastructure).  Stands in for dynamic_vins/src/utils/parameters.h when reference front-end
// sources are compiled into oracle/_ref/: the real header drags in ROS, PCL and the estimator.  Only the names the
// front-end reads are declared (the `Config` statics, /root/reference/dynamic_vins/src/utils/parameters.h:62-116);
// the harness sets them (oracle/ref/ref_glue.cpp).
#pragma once
#include <atomic>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include <eigen3/Eigen/Dense>
#include <opencv2/opencv.hpp>
#include <spdlog/spdlog.h>

#include "utils/camera_model.h"
#include "utils/log_utils.h"

namespace dynamic_vins {

constexpr double kFocalLength = 460.0;
constexpr int kWinSize = 10;
constexpr int kNumFeat = 1000;
constexpr int kQueueSize = 200;
constexpr double kDelay = 0.005;
constexpr int kImageQueueSize = 100;

enum class SLAM { kRaw, kNaive, kDynamic };
enum class DatasetType { kViode, kKitti, kEuRoc, kCustom };

class Config {
public:
    using Ptr = std::shared_ptr<Config>;
    inline static std::string kExCalibResultPath;
    inline static int kCamNum;
    inline static bool is_stereo;
    inline static int use_imu;
    inline static std::map<int, Eigen::Vector3d> pts_gt;
    inline static std::string FISHEYE_MASK;
    inline static int kInputHeight, kInputWidth, kInputChannel = 3;
    inline static SLAM slam;
    inline static DatasetType dataset;
    inline static std::string dataset_name;
    inline static bool is_input_seg;
    inline static bool is_only_frontend;
    inline static bool is_only_imgprocess;
    inline static bool use_line;
    inline static bool is_undistort_input{false};
    inline static int is_estimate_ex;
    inline static int is_estimate_td;
    inline static std::string kBasicDir;
    inline static std::string kDatasetSequence;
    inline static bool use_dense_flow{false};
    inline static bool use_background_flow{false};
    inline static bool use_plane_constraint{false};
    inline static bool use_det3d{false};
    inline static bool dst_mode{false};
    inline static bool is_vertical_draw{false};
    inline static std::atomic_bool ok{true};
};
using cfg = Config;

}  // namespace dynamic_vins
