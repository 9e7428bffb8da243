// ORACLE / TEST INFRASTRUCTURE ONLY — never linked into, loaded by or shipped with the product.
//
// C entry points over the REFERENCE's own front-end code, compiled unmodified from /root/reference by oracle/ref/Makefile
// into oracle/_ref/libdvref.so:
//     camera_models/src/camera_models/{PinholeCamera,Camera}.cc
//     dynamic_vins/src/front_end/{feature_utils,instance_feature,background_tracker,dynamic_tracker}.cpp
//     dynamic_vins/src/basic/semantic_image.cpp, dynamic_vins/src/utils/io/feature_serialization.cpp
// against the stand-in third-party headers in oracle/shim/ (Eigen / OpenCV containers; OpenCV image algorithms are forwarded
// through hooks to cv2 by the Python harness, oracle/ref_lib.py).  This file only (a) defines the few out-of-scope symbols
// those translation units reference (line detector, yaml parameter loading, camera globals), (b) converts flat C arrays to
// the reference's types and back.  It contains no front-end arithmetic of its own.
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <stdexcept>

// InstsFeatManager::RejectWithF is a private member (never called by the reference itself): the checker reaches it by
// parsing the reference's class declarations with `private` spelled `public` in THIS translation unit only (standard headers
// are included before).  The reference's own translation units are compiled untouched.
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <set>
#include <sstream>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>
#include "dvshim_cv.hpp"
#include "dvshim_eigen.hpp"
#define private public
#include "front_end/background_tracker.h"
#include "front_end/dynamic_tracker.h"
#undef private
#include "front_end/feature_utils.h"
#include "front_end/front_end_parameters.h"
#include "utils/io/feature_serialization.h"
#include "basic/feature_queue.h"
#include "camodocal/camera_models/PinholeCamera.h"

// ---- shim state ----------------------------------------------------------------------------------------------------
namespace dvshim {
Hooks& hooks() { static Hooks h; return h; }
}

// ---- out-of-scope symbols the reference translation units reference -------------------------------------------------
namespace dynamic_vins {

CameraInfo cam_s, cam_t, cam_v;                                   // utils/camera_model.cpp:21-23 (set by dvref_configure)
std::vector<Eigen::Matrix3d> R_IC;
std::vector<Eigen::Vector3d> T_IC;

// front_end/front_end_parameters.cpp reads a yaml file through cv::FileStorage; the harness sets the statics directly
void FrontendParemater::SetParameters(const std::string&) {}

// line features are off on the parity path (cfg::use_line = false); the constructor runs, nothing else may
LineDetector::LineDetector(const std::string&) {}
FrameLines::Ptr LineDetector::Detect(cv::Mat&, const cv::Mat&) { dvshim_unreachable("LineDetector::Detect"); }
void LineDetector::TrackLeftLine(const FrameLines::Ptr&, const FrameLines::Ptr&) { dvshim_unreachable("LineDetector"); }
void LineDetector::TrackRightLine(const FrameLines::Ptr&, const FrameLines::Ptr&) { dvshim_unreachable("LineDetector"); }
void LineDetector::VisualizeLine(cv::Mat&, const FrameLines::Ptr&) { dvshim_unreachable("LineDetector"); }
void LineDetector::VisualizeRightLine(cv::Mat&, const FrameLines::Ptr&, bool) { dvshim_unreachable("LineDetector"); }
void FrameLines::SetLines() { dvshim_unreachable("FrameLines"); }
void FrameLines::UndistortedLineEndPoints(camodocal::CameraPtr&) { dvshim_unreachable("FrameLines"); }
float Box2D::IoU(const cv::Rect2f&, const cv::Rect2f&) { dvshim_unreachable("Box2D::IoU"); }
void Box3D::VisCorners2d(cv::Mat&, const cv::Scalar&, camodocal::CameraPtr&) { dvshim_unreachable("Box3D::VisCorners2d"); }

}  // namespace dynamic_vins

using namespace dynamic_vins;

static thread_local char g_err[512] = "";
static int fail(const std::exception& e) {
    std::snprintf(g_err, sizeof(g_err), "%s", e.what());
    return -1;
}

extern "C" {

const char* dvref_last_error(void) { return g_err; }
const char* dvref_sources(void) {
    return "camera_models/src/camera_models/PinholeCamera.cc camera_models/src/camera_models/Camera.cc "
           "dynamic_vins/src/front_end/feature_utils.cpp dynamic_vins/src/front_end/instance_feature.cpp "
           "dynamic_vins/src/front_end/background_tracker.cpp dynamic_vins/src/front_end/dynamic_tracker.cpp "
           "dynamic_vins/src/basic/semantic_image.cpp dynamic_vins/src/utils/io/feature_serialization.cpp";
}

void dvref_set_hooks(void* lk, void* gftt, void* erode, void* circle, void* bgr2gray) {
    auto& h = dvshim::hooks();
    h.calc_optical_flow_pyr_lk = reinterpret_cast<decltype(h.calc_optical_flow_pyr_lk)>(lk);
    h.good_features_to_track = reinterpret_cast<decltype(h.good_features_to_track)>(gftt);
    h.erode_rect = reinterpret_cast<decltype(h.erode_rect)>(erode);
    h.circle_filled = reinterpret_cast<decltype(h.circle_filled)>(circle);
    h.bgr2gray = reinterpret_cast<decltype(h.bgr2gray)>(bgr2gray);
}
void dvref_set_hook_fundamental(void* fm) {
    auto& h = dvshim::hooks();
    h.find_fundamental_mat = reinterpret_cast<decltype(h.find_fundamental_mat)>(fm);
}
void dvref_set_hook_gftt_cuda(void* f) {
    auto& h = dvshim::hooks();
    h.good_features_cuda = reinterpret_cast<decltype(h.good_features_cuda)>(f);
}

// ---- camodocal::PinholeCamera ------------------------------------------------------------------------------------
// cam = {k1, k2, p1, p2, fx, fy, cx, cy}
void* dvref_camera_new(int w, int h, const double* cam) {
    auto* p = new camodocal::CameraPtr(new camodocal::PinholeCamera("cam", w, h, cam[0], cam[1], cam[2], cam[3], cam[4], cam[5],
                                                                    cam[6], cam[7]));
    return p;
}
void dvref_camera_free(void* c) { delete static_cast<camodocal::CameraPtr*>(c); }

// PinholeCamera::liftProjective: uv [n][2] -> P [n][3]
void dvref_lift_projective(void* c, const double* uv, int n, double* P) {
    camodocal::CameraPtr& cam = *static_cast<camodocal::CameraPtr*>(c);
    for (int i = 0; i < n; i++) {
        Eigen::Vector3d b;
        cam->liftProjective(Eigen::Vector2d(uv[2 * i], uv[2 * i + 1]), b);
        P[3 * i] = b.x(); P[3 * i + 1] = b.y(); P[3 * i + 2] = b.z();
    }
}
// PinholeCamera::distortion: p_u [n][2] -> d_u [n][2]
void dvref_distortion(void* c, const double* pu, int n, double* du) {
    auto* cam = static_cast<camodocal::PinholeCamera*>(static_cast<camodocal::CameraPtr*>(c)->get());
    for (int i = 0; i < n; i++) {
        Eigen::Vector2d d;
        cam->distortion(Eigen::Vector2d(pu[2 * i], pu[2 * i + 1]), d);
        du[2 * i] = d(0); du[2 * i + 1] = d(1);
    }
}
// PinholeCamera::spaceToPlane: P [n][3] -> p [n][2]
void dvref_space_to_plane(void* c, const double* P, int n, double* p) {
    camodocal::CameraPtr& cam = *static_cast<camodocal::CameraPtr*>(c);
    for (int i = 0; i < n; i++) {
        Eigen::Vector2d q;
        cam->spaceToPlane(Eigen::Vector3d(P[3 * i], P[3 * i + 1], P[3 * i + 2]), q);
        p[2 * i] = q(0); p[2 * i + 1] = q(1);
    }
}
// UndistortedPts (front_end/feature_utils.cpp:193-203): pts [n][2] float -> un [n][2] float
void dvref_undistorted_pts(void* c, const float* pts, int n, float* un) {
    camodocal::CameraPtr& cam = *static_cast<camodocal::CameraPtr*>(c);
    std::vector<cv::Point2f> v(n);
    for (int i = 0; i < n; i++) v[i] = cv::Point2f(pts[2 * i], pts[2 * i + 1]);
    std::vector<cv::Point2f> out = UndistortedPts(v, cam);
    for (int i = 0; i < n; i++) { un[2 * i] = out[i].x; un[2 * i + 1] = out[i].y; }
}

// ---- front_end/feature_utils.h inline helpers ----------------------------------------------------------------------
int dvref_in_border(float x, float y, int row, int col) { return InBorder(cv::Point2f(x, y), row, col) ? 1 : 0; }
float dvref_point_distance(float x1, float y1, float x2, float y2) { return PointDistance(cv::Point2f(x1, y1), cv::Point2f(x2, y2)); }
int dvref_cv_round_f(float v) { return cvRound(v); }
// ReduceVector<cv::Point2f> / <int> / <unsigned>: compacts in place, returns the new size
int dvref_reduce_points(float* pts, const unsigned char* status, int n) {
    std::vector<cv::Point2f> v(n);
    for (int i = 0; i < n; i++) v[i] = cv::Point2f(pts[2 * i], pts[2 * i + 1]);
    ReduceVector(v, std::vector<uchar>(status, status + n));
    for (size_t i = 0; i < v.size(); i++) { pts[2 * i] = v[i].x; pts[2 * i + 1] = v[i].y; }
    return (int)v.size();
}
int dvref_reduce_ints(int* a, const unsigned char* status, int n) {
    std::vector<int> v(a, a + n);
    ReduceVector(v, std::vector<uchar>(status, status + n));
    std::copy(v.begin(), v.end(), a);
    return (int)v.size();
}
// SetStatusByMask (feature_utils.h:149-151)
void dvref_set_status_by_mask(unsigned char* status, const float* pts, int n, const unsigned char* mask, int rows, int cols) {
    std::vector<uchar> st(status, status + n);
    std::vector<cv::Point2f> v(n);
    for (int i = 0; i < n; i++) v[i] = cv::Point2f(pts[2 * i], pts[2 * i + 1]);
    cv::Mat m(rows, cols, CV_8UC1, (void*)mask);
    SetStatusByMask(st, v, m);
    std::copy(st.begin(), st.end(), status);
}
// PtsVelocity (feature_utils.cpp:274-296)
void dvref_pts_velocity(double dt, const unsigned* ids, const float* cur_un, int n, const unsigned* prev_ids, const float* prev_un,
                        int m, float* vel) {
    std::vector<unsigned> idv(ids, ids + n);
    std::vector<cv::Point2f> cur(n), out;
    for (int i = 0; i < n; i++) cur[i] = cv::Point2f(cur_un[2 * i], cur_un[2 * i + 1]);
    std::map<unsigned, cv::Point2f> prev;
    for (int i = 0; i < m; i++) prev[prev_ids[i]] = cv::Point2f(prev_un[2 * i], prev_un[2 * i + 1]);
    PtsVelocity(dt, idv, cur, prev, out);
    for (int i = 0; i < n; i++) { vel[2 * i] = out[i].x; vel[2 * i + 1] = out[i].y; }
}
// FeatureTrackByLK (feature_utils.cpp:35-69); fe_para::is_flow_back is read, not the argument (reference quirk Q1)
int dvref_feature_track_by_lk(const unsigned char* img1, const unsigned char* img2, int rows, int cols, const float* pts1, int n,
                              float* pts2, unsigned char* status, int flow_back) {
    try {
        cv::Mat a(rows, cols, CV_8UC1, (void*)img1), b(rows, cols, CV_8UC1, (void*)img2);
        std::vector<cv::Point2f> p1(n), p2;
        for (int i = 0; i < n; i++) p1[i] = cv::Point2f(pts1[2 * i], pts1[2 * i + 1]);
        fe_para::is_flow_back = flow_back;
        std::vector<uchar> st = FeatureTrackByLK(a, b, p1, p2, flow_back != 0);
        for (int i = 0; i < n; i++) { pts2[2 * i] = p2[i].x; pts2[2 * i + 1] = p2[i].y; status[i] = st[i]; }
        return n;
    } catch (const std::exception& e) { return fail(e); }
}
// InstanceImagePadding (feature_utils.cpp:406-413): both outputs are max(rows) x max(cols)
void dvref_instance_image_padding(const unsigned char* img1, int r1, int c1, const unsigned char* img2, int r2, int c2,
                                  unsigned char* out1, unsigned char* out2) {
    cv::Mat a(r1, c1, CV_8UC1, (void*)img1), b(r2, c2, CV_8UC1, (void*)img2);
    auto [pa, pb] = InstanceImagePadding(a, b);
    for (int r = 0; r < pa.rows; r++) {
        std::memcpy(out1 + (size_t)r * pa.cols, pa.ptr(r), pa.cols);
        std::memcpy(out2 + (size_t)r * pb.cols, pb.ptr(r), pb.cols);
    }
}
// ErodeMask (feature_utils.h:141-146)
void dvref_erode_mask(const unsigned char* in, int rows, int cols, int k, unsigned char* out) {
    cv::Mat a = cv::Mat(rows, cols, CV_8UC1, (void*)in).clone(), b;
    ErodeMask(a, b, k);
    for (int r = 0; r < rows; r++) std::memcpy(out + (size_t)r * cols, b.ptr(r), cols);
}

// ---- the trackers ---------------------------------------------------------------------------------------------------
typedef struct dvref_config {
    int max_cnt, max_dynamic_cnt, min_dist, min_dynamic_dist, flow_back, use_mask_morphology, mask_morphology_size;
    int width, height, stereo, dynamic;
    double cam0[8], cam1[8];          // k1 k2 p1 p2 fx fy cx cy
} dvref_config;

typedef struct dvref_obs {            // one (feature id, camera) record of FeatureBackground::points
    unsigned id;
    int cam;
    double v[7];                      // x y z u v vx vy
} dvref_obs;

typedef struct dvref_inst_obs {       // one feature of FeatureInstance::features
    unsigned inst_id, id;
    int is_stereo, reserved;
    double point[3], vel[2], point_right[3], vel_right[2];
    double disp;
} dvref_inst_obs;

typedef struct dvref_box {            // Box2D + InstRoi as SemanticImage::SetMaskAndRoi leaves them
    unsigned track_id;
    int x, y, w, h;
    const unsigned char* mask;        // h x w, 255 = object
    int mask_pitch;
} dvref_box;

struct RefFrontEnd {
    std::unique_ptr<FeatureTracker> tracker;
    std::unique_ptr<InstsFeatManager> insts;
    int rows = 0, cols = 0;
};

void* dvref_front_end_new(const dvref_config* c) {
    fe_para::kMaxCnt = c->max_cnt;
    fe_para::kMaxDynamicCnt = c->max_dynamic_cnt;
    fe_para::kMinDist = c->min_dist;
    fe_para::kMinDynamicDist = c->min_dynamic_dist;
    fe_para::kFThreshold = 1.0;
    fe_para::is_show_track = 0;
    fe_para::is_flow_back = c->flow_back;
    fe_para::use_mask_morphology = c->use_mask_morphology != 0;
    fe_para::kMaskMorphologySize = c->mask_morphology_size;
    fe_para::kInputHeight = c->height;
    fe_para::kInputWidth = c->width;
    cfg::is_stereo = c->stereo != 0;
    cfg::slam = c->dynamic ? SLAM::kDynamic : SLAM::kRaw;
    cfg::dataset = DatasetType::kCustom;
    cfg::dataset_name = "custom";
    cfg::use_line = false;
    cfg::use_det3d = false;
    cfg::kInputHeight = c->height;
    cfg::kInputWidth = c->width;
    cam_t.cam0 = camodocal::CameraPtr(new camodocal::PinholeCamera("cam0", c->width, c->height, c->cam0[0], c->cam0[1], c->cam0[2],
                                                                   c->cam0[3], c->cam0[4], c->cam0[5], c->cam0[6], c->cam0[7]));
    cam_t.cam1 = camodocal::CameraPtr(new camodocal::PinholeCamera("cam1", c->width, c->height, c->cam1[0], c->cam1[1], c->cam1[2],
                                                                   c->cam1[3], c->cam1[4], c->cam1[5], c->cam1[6], c->cam1[7]));
    cam_s = cam_t;
    InstFeat::global_id_count = 1;                      // one front-end per process at a time (the reference's static)
    auto* fe = new RefFrontEnd();
    fe->tracker.reset(new FeatureTracker(""));
    if (c->dynamic) fe->insts.reset(new InstsFeatManager(""));
    fe->rows = c->height; fe->cols = c->width;
    return fe;
}
void dvref_front_end_free(void* p) { delete static_cast<RefFrontEnd*>(p); }

static int flatten(const FeatureBackground& fb, dvref_obs* out, int cap) {
    int n = 0;
    for (auto& [id, obs] : fb.points)
        for (auto& [cam, v] : obs) {
            if (n < cap) {
                out[n].id = id; out[n].cam = cam;
                for (int k = 0; k < 7; k++) out[n].v[k] = v(k);
            }
            n++;
        }
    return n;
}

static SemanticImage make_image(RefFrontEnd* fe, const unsigned char* gray0, const unsigned char* gray1, double time0, unsigned seq) {
    SemanticImage img;
    img.gray0 = cv::Mat(fe->rows, fe->cols, CV_8UC1, (void*)gray0).clone();
    if (gray1) img.gray1 = cv::Mat(fe->rows, fe->cols, CV_8UC1, (void*)gray1).clone();
    img.color0 = cv::Mat(fe->rows, fe->cols, CV_8UC3);      // only its size is read on this path
    img.time0 = time0; img.time1 = time0; img.seg0_time = time0; img.seg1_time = time0;
    img.seq = seq;
    return img;
}

// FeatureTracker::TrackImage (front_end/background_tracker.cpp:52-158) -> number of records
int dvref_track_image(void* p, const unsigned char* gray0, const unsigned char* gray1, double time0, dvref_obs* out, int cap) {
    try {
        auto* fe = static_cast<RefFrontEnd*>(p);
        SemanticImage img = make_image(fe, gray0, gray1, time0, 0);
        return flatten(fe->tracker->TrackImage(img), out, cap);
    } catch (const std::exception& e) { return fail(e); }
}

// One frame of dynamic mode as system/main.cpp:193-252 drives it: reset visibility, AddViodeInstances (instances arrive with
// track ids), TrackSemanticImage, InstsTrack, Output.  The reference runs InstsTrack on a second thread that races the
// background thread for InstFeat::global_id_count; here the two run one after the other, background first.
int dvref_track_dynamic(void* p, const unsigned char* gray0, const unsigned char* gray1, const unsigned char* merge_mask,
                        const unsigned char* inv_merge_mask, const float* disp, const dvref_box* boxes, int n_boxes, double time0,
                        unsigned seq, dvref_obs* out, int cap, int* n_out, dvref_inst_obs* iout, int icap, int* n_iout) {
    try {
        auto* fe = static_cast<RefFrontEnd*>(p);
        SemanticImage img = make_image(fe, gray0, gray1, time0, seq);
        img.gray0_gpu.upload(img.gray0);
        if (gray1) img.gray1_gpu.upload(img.gray1);
        img.exist_inst = n_boxes > 0;
        if (merge_mask) { img.merge_mask = cv::Mat(fe->rows, fe->cols, CV_8UC1, (void*)merge_mask).clone(); img.merge_mask_gpu.upload(img.merge_mask); }
        if (inv_merge_mask) {
            img.inv_merge_mask = cv::Mat(fe->rows, fe->cols, CV_8UC1, (void*)inv_merge_mask).clone();
            img.inv_merge_mask_gpu.upload(img.inv_merge_mask);
        }
        img.disp = disp ? cv::Mat(fe->rows, fe->cols, CV_32FC1, (void*)disp).clone() : cv::Mat(fe->rows, fe->cols, CV_32FC1, cv::Scalar(0));
        for (int i = 0; i < n_boxes; i++) {                                  // basic/semantic_image.cpp:48-59
            auto b = std::make_shared<Box2D>();
            b->id = i; b->track_id = (int)boxes[i].track_id; b->class_id = 0; b->score = 1.f;
            b->min_pt = cv::Point2f((float)boxes[i].x, (float)boxes[i].y);
            b->max_pt = cv::Point2f((float)(boxes[i].x + boxes[i].w), (float)(boxes[i].y + boxes[i].h));
            b->rect = cv::Rect2f((float)boxes[i].x, (float)boxes[i].y, (float)boxes[i].w, (float)boxes[i].h);
            b->roi = std::make_shared<InstRoi>();
            b->roi->mask_cv = cv::Mat(boxes[i].h, boxes[i].w, CV_8UC1, (void*)boxes[i].mask, (size_t)boxes[i].mask_pitch).clone();
            b->roi->mask_gpu.upload(b->roi->mask_cv);
            b->roi->roi_gpu = img.gray0_gpu(b->rect);
            b->roi->roi_gpu.download(b->roi->roi_gray);
            img.boxes2d.push_back(b);
        }
        for (auto& [inst_id, inst] : fe->insts->instances) {                 // system/main.cpp:198-202
            inst.is_curr_visible = false;
            inst.box2d.reset();
            inst.box3d.reset();
        }
        fe->insts->AddViodeInstances(img);                                   // :209
        *n_out = flatten(fe->tracker->TrackSemanticImage(img), out, cap);    // :250
        fe->insts->InstsTrack(img);                                          // :247
        std::map<unsigned int, FeatureInstance> res = fe->insts->Output();   // :254
        int n = 0;
        for (auto& [key, fi] : res)
            for (auto& [fid, f] : fi.features) {
                if (n < icap) {
                    dvref_inst_obs& o = iout[n];
                    o.inst_id = key; o.id = fid; o.is_stereo = f->is_stereo ? 1 : 0; o.reserved = 0;
                    for (int k = 0; k < 3; k++) { o.point[k] = f->point(k); o.point_right[k] = f->point_right(k); }
                    for (int k = 0; k < 2; k++) { o.vel[k] = f->vel(k); o.vel_right[k] = f->vel_right(k); }
                    o.disp = f->disp;
                }
                n++;
            }
        *n_iout = n;
        return 0;
    } catch (const std::exception& e) { return fail(e); }
}

// basic/feature_queue.h:19-71, the reference's FeatureQueue itself (header-only), one object per handle
void* dvref_queue_new(void) { return new FeatureQueue(); }
void dvref_queue_free(void* q) { delete static_cast<FeatureQueue*>(q); }
void dvref_queue_push(void* q, unsigned seq, double time) {
    FrontendFeature f;
    f.seq_id = seq; f.time = time;
    static_cast<FeatureQueue*>(q)->push_back(f);
}
int dvref_queue_request(void* q, unsigned* seq, double* time) {
    auto f = static_cast<FeatureQueue*>(q)->request();
    if (!f) return 0;
    *seq = f->seq_id; *time = f->time;
    return 1;
}
int dvref_queue_size(void* q) { return static_cast<FeatureQueue*>(q)->size(); }
int dvref_queue_empty(void* q) { return static_cast<FeatureQueue*>(q)->empty() ? 1 : 0; }
void dvref_queue_clear(void* q) { static_cast<FeatureQueue*>(q)->clear(); }
int dvref_queue_front_time(void* q, double* time) {
    auto t = static_cast<FeatureQueue*>(q)->front_time();
    if (!t) return 0;
    *time = *t;
    return 1;
}

// SerializePointFeature / DeserializePointFeature (utils/io/feature_serialization.cpp:26-70) on flat records (id, cam, 7 doubles),
// sorted by (id, cam).  fmt::format is the stand-in of oracle/shim/dvshim_fmt.hpp (shortest round-trip numbers).
int dvref_serialize_points(const char* path, const dvref_obs* obs, int n) {
    try {
        std::map<unsigned int, std::vector<std::pair<int, Eigen::Matrix<double, 7, 1>>>> points;
        for (int i = 0; i < n; i++) {
            Eigen::Matrix<double, 7, 1> v;
            for (int k = 0; k < 7; k++) v(k, 0) = obs[i].v[k];
            points[obs[i].id].push_back({obs[i].cam, v});
        }
        SerializePointFeature(path, points);
        return 0;
    } catch (const std::exception& e) { return fail(e); }
}
int dvref_deserialize_points(const char* path, dvref_obs* out, int cap) {
    try {
        FeatureBackground fb;
        fb.points = DeserializePointFeature(path);
        return flatten(fb, out, cap);
    } catch (const std::exception& e) { return fail(e); }
}

// SemanticImage::SetMaskAndRoi (basic/semantic_image.cpp:20-63) on an N x rows x cols instance-mask tensor (int8 values as the
// segmentation network leaves them) and N boxes: merge_mask, inv_merge_mask (rows x cols each), and per box the ROI mask
// full_mask(rect) and the gray crop gray0(rect), written back to back into roi_masks / roi_grays (w*h bytes per box, in order).
int dvref_set_mask_and_roi(const signed char* masks, int n, int rows, int cols, const unsigned char* gray0, const int* rects /* x y w h */,
                           unsigned char* merge_mask, unsigned char* inv_merge_mask, unsigned char* roi_masks, unsigned char* roi_grays) {
    try {
        SemanticImage img;
        img.gray0 = cv::Mat(rows, cols, CV_8UC1, (void*)gray0).clone();
        img.gray0_gpu.upload(img.gray0);
        std::vector<int64_t> vals((size_t)n * rows * cols);
        for (size_t i = 0; i < vals.size(); i++) vals[i] = masks[i];
        img.mask_tensor = torch::Tensor({(int64_t)n, (int64_t)rows, (int64_t)cols}, vals.data(), torch::kInt8);
        for (int i = 0; i < n; i++) {
            auto b = std::make_shared<Box2D>();
            b->id = i; b->track_id = i + 1;
            b->rect = cv::Rect2f((float)rects[4 * i], (float)rects[4 * i + 1], (float)rects[4 * i + 2], (float)rects[4 * i + 3]);
            img.boxes2d.push_back(b);
        }
        img.SetMaskAndRoi();
        if (!img.exist_inst) return 0;
        for (int r = 0; r < rows; r++) {
            std::memcpy(merge_mask + (size_t)r * cols, img.merge_mask.data + (size_t)r * img.merge_mask.step, (size_t)cols);
            std::memcpy(inv_merge_mask + (size_t)r * cols, img.inv_merge_mask.data + (size_t)r * img.inv_merge_mask.step, (size_t)cols);
        }
        size_t off = 0;
        for (auto& b : img.boxes2d) {
            const cv::Mat& m = b->roi->mask_cv;
            const cv::Mat& g = b->roi->roi_gray;
            for (int r = 0; r < m.rows; r++) {
                std::memcpy(roi_masks + off + (size_t)r * m.cols, m.data + (size_t)r * m.step, (size_t)m.cols);
                std::memcpy(roi_grays + off + (size_t)r * g.cols, g.data + (size_t)r * g.step, (size_t)g.cols);
            }
            off += (size_t)m.rows * m.cols;
        }
        return n;
    } catch (const std::exception& e) { return fail(e); }
}

// FeatureTracker::TrackImageNaive (front_end/background_tracker.cpp:400-516): the all-cv::cuda flow.  The cv::cuda objects are
// hosted (oracle/shim/dvshim_cv.hpp): GpuMat = Mat, SparsePyrLKOpticalFlow and the morphology filter forward to the CPU hooks, the
// corner detector to the harness's restatement of cv::cuda::GoodFeaturesToTrackDetector.
int dvref_track_image_naive(void* p, const unsigned char* gray0, const unsigned char* gray1, const unsigned char* inv_merge_mask,
                            int exist_inst, double time0, unsigned seq, dvref_obs* out, int cap) {
    try {
        auto* fe = static_cast<RefFrontEnd*>(p);
        SemanticImage img = make_image(fe, gray0, gray1, time0, seq);
        img.gray0_gpu.upload(img.gray0);
        if (gray1) img.gray1_gpu.upload(img.gray1);
        img.exist_inst = exist_inst != 0;
        if (inv_merge_mask) {
            img.inv_merge_mask = cv::Mat(fe->rows, fe->cols, CV_8UC1, (void*)inv_merge_mask).clone();
            img.inv_merge_mask_gpu.upload(img.inv_merge_mask);
        }
        return flatten(fe->tracker->TrackImageNaive(img), out, cap);
    } catch (const std::exception& e) { return fail(e); }
}

// InstsFeatManager::RejectWithF on a free-standing instance holding (curr_points, last_points); needs a dynamic-mode front end
// (cam_t.cam0 and fe_para::kFThreshold are the globals dvref_front_end_new set).  Returns the size of the status vector.
int dvref_reject_with_f(void* p, const float* cur, const float* prev, int n, int col, int row, unsigned char* status) {
    try {
        auto* fe = static_cast<RefFrontEnd*>(p);
        if (!fe->insts) throw std::runtime_error("dvref_reject_with_f needs a dynamic-mode front end");
        InstFeat inst;
        for (int i = 0; i < n; i++) {
            inst.curr_points.emplace_back(cur[2 * i], cur[2 * i + 1]);
            inst.last_points.emplace_back(prev[2 * i], prev[2 * i + 1]);
        }
        const std::vector<uchar> st = fe->insts->RejectWithF(inst, col, row);
        for (size_t i = 0; i < st.size(); i++) status[i] = st[i];
        return (int)st.size();
    } catch (const std::exception& e) { return fail(e); }
}

// InstFeat::DetectExtraPoints on a free-standing instance: ROI mask + box, full-size CV_32F disparity map; cam_s = {fx, fy, cx,
// cy, baseline} (CameraInfo's float members).  out = 3 doubles per point; returns the count.
int dvref_detect_extra_points(const unsigned char* mask, int rows, int cols, const float* disp, int disp_rows, int disp_cols, int box_x,
                              int box_y, const float* cam, double* out, int cap) {
    try {
        cam_s.fx0 = cam[0]; cam_s.fy0 = cam[1]; cam_s.cx0 = cam[2]; cam_s.cy0 = cam[3]; cam_s.baseline = cam[4];
        InstFeat inst;
        inst.roi = std::make_shared<InstRoi>();
        inst.roi->mask_cv = cv::Mat(rows, cols, CV_8UC1, const_cast<unsigned char*>(mask)).clone();
        inst.roi->roi_gray = cv::Mat(rows, cols, CV_8UC1, cv::Scalar(0));
        inst.box2d = std::make_shared<Box2D>();
        inst.box2d->rect = cv::Rect2f((float)box_x, (float)box_y, (float)cols, (float)rows);
        cv::Mat d(disp_rows, disp_cols, CV_32FC1, const_cast<float*>(disp));
        inst.DetectExtraPoints(d);
        const int n = (int)inst.extra_points3d.size();
        for (int i = 0; i < n && i < cap; i++) {
            out[3 * i] = inst.extra_points3d[i].x(); out[3 * i + 1] = inst.extra_points3d[i].y(); out[3 * i + 2] = inst.extra_points3d[i].z();
        }
        return n;
    } catch (const std::exception& e) { return fail(e); }
}

// background InstFeat state after a frame (for teacher-forced comparisons): ids, track_cnt, last_points
int dvref_instance_count(void* p) {
    auto* fe = static_cast<RefFrontEnd*>(p);
    return fe->insts ? (int)fe->insts->instances.size() : 0;
}
// per-instance bookkeeping: writes up to cap rows of {key, lost_num, is_curr_visible, n_points}
int dvref_instance_table(void* p, int* rows, int cap) {
    auto* fe = static_cast<RefFrontEnd*>(p);
    if (!fe->insts) return 0;
    std::map<unsigned, const InstFeat*> sorted;
    for (auto& [k, inst] : fe->insts->instances) sorted[k] = &inst;
    int n = 0;
    for (auto& [k, inst] : sorted) {
        if (n < cap) {
            rows[4 * n] = (int)k; rows[4 * n + 1] = inst->lost_num; rows[4 * n + 2] = inst->is_curr_visible ? 1 : 0;
            rows[4 * n + 3] = (int)inst->last_points.size();
        }
        n++;
    }
    return n;
}

}  // extern "C"
