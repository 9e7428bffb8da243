// ORACLE / TEST INFRASTRUCTURE ONLY.  Stand-in for the OpenCV 3.4 C++ API surface that the reference sources compiled
// into oracle/_ref/ (oracle/ref/Makefile) name.  Containers (Mat, Point_, Size_, Rect_, Scalar_, Vec) are real; the
// image ALGORITHMS the front-end calls (calcOpticalFlowPyrLK, goodFeaturesToTrack, erode, circle, ...) are not
// re-implemented here: they are forwarded through `dvshim::hooks` to whatever the test harness registers (cv2 itself,
// through ctypes callbacks), so the reference's own glue code runs on top of the real OpenCV arithmetic.  Entry points
// that the parity path never reaches (calibration, yaml, drawing, cv::cuda) are declared so the reference sources parse
// and abort when called.  Not a product file; never shipped.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <list>
#include <map>
#include <memory>
#include <set>
#include <sstream>
#include <string>
#include <vector>

typedef unsigned char uchar;
typedef unsigned short ushort;

#define CV_8U 0
#define CV_8S 1
#define CV_16U 2
#define CV_16S 3
#define CV_32S 4
#define CV_32F 5
#define CV_64F 6
#define CV_CN_SHIFT 3
#define CV_MAKETYPE(depth, cn) (((depth) & 7) + (((cn) - 1) << CV_CN_SHIFT))
#define CV_INTER_LINEAR 1
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_16SC2 CV_MAKETYPE(CV_16S, 2)
#define CV_16UC1 CV_MAKETYPE(CV_16U, 1)
#define CV_32SC1 CV_MAKETYPE(CV_32S, 1)
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)
#define CV_32FC2 CV_MAKETYPE(CV_32F, 2)
#define CV_32FC3 CV_MAKETYPE(CV_32F, 3)
#define CV_64FC1 CV_MAKETYPE(CV_64F, 1)
#define CV_GRAY2BGR 8
#define CV_GRAY2RGB 8
#define CV_BGR2GRAY 6
#define CV_RGB2GRAY 7

[[noreturn]] inline void dvshim_unreachable(const char* what) {
    std::fprintf(stderr, "oracle/shim: %s is outside the parity path and has no implementation\n", what);
    std::abort();
}

// cvRound: OpenCV rounds to nearest, ties to even (SSE2 cvtsd2si / lrint under the default rounding mode)
inline int cvRound(double v) { return (int)std::lrint(v); }
inline int cvRound(float v) { return (int)std::lrintf(v); }
inline int cvRound(int v) { return v; }
inline int cvFloor(double v) { int i = (int)v; return i - (i > v); }
inline int cvCeil(double v) { int i = (int)v; return i + (i < v); }

namespace cv {

template <class T>
struct saturate_helper { static T cast(double v) { return (T)v; } };
template <class T, class U>
inline T saturate_cast(U v) { return saturate_helper<T>::cast((double)v); }
template <>
struct saturate_helper<uchar> { static uchar cast(double v) { int i = cvRound(v); return (uchar)(i < 0 ? 0 : i > 255 ? 255 : i); } };
template <>
struct saturate_helper<int> { static int cast(double v) { return cvRound(v); } };

template <class T>
struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T x_, T y_) : x(x_), y(y_) {}
    template <class U>
    Point_(const Point_<U>& o) : x(saturate_cast<T>(o.x)), y(saturate_cast<T>(o.y)) {}
    Point_ operator+(const Point_& o) const { return Point_(x + o.x, y + o.y); }
    Point_ operator-(const Point_& o) const { return Point_(x - o.x, y - o.y); }
    Point_& operator+=(const Point_& o) { x += o.x; y += o.y; return *this; }
    Point_& operator-=(const Point_& o) { x -= o.x; y -= o.y; return *this; }
    bool operator==(const Point_& o) const { return x == o.x && y == o.y; }
};
template <class T, class S>
inline Point_<T> operator*(const Point_<T>& p, S s) { return Point_<T>(saturate_cast<T>(p.x * s), saturate_cast<T>(p.y * s)); }
template <class T>
inline Point_<T> operator/(const Point_<T>& p, double s) { return Point_<T>(saturate_cast<T>(p.x / s), saturate_cast<T>(p.y / s)); }
typedef Point_<int> Point2i;
typedef Point_<int> Point;
typedef Point_<float> Point2f;
typedef Point_<double> Point2d;
template <class T>
inline double norm(const Point_<T>& p) { return std::sqrt((double)p.x * p.x + (double)p.y * p.y); }

template <class T>
struct Point3_ {
    T x, y, z;
    Point3_() : x(0), y(0), z(0) {}
    Point3_(T x_, T y_, T z_) : x(x_), y(y_), z(z_) {}
};
typedef Point3_<float> Point3f;
typedef Point3_<double> Point3d;

template <class T>
struct Size_ {
    T width, height;
    Size_() : width(0), height(0) {}
    Size_(T w, T h) : width(w), height(h) {}
    bool operator==(const Size_& o) const { return width == o.width && height == o.height; }
    bool operator!=(const Size_& o) const { return !(*this == o); }
    T area() const { return width * height; }
    bool empty() const { return width <= 0 || height <= 0; }
};
typedef Size_<int> Size;
typedef Size_<int> Size2i;
typedef Size_<float> Size2f;

template <class T>
struct Rect_ {
    T x, y, width, height;
    Rect_() : x(0), y(0), width(0), height(0) {}
    Rect_(T x_, T y_, T w, T h) : x(x_), y(y_), width(w), height(h) {}
    template <class U>
    Rect_(const Rect_<U>& o) : x(saturate_cast<T>(o.x)), y(saturate_cast<T>(o.y)), width(saturate_cast<T>(o.width)), height(saturate_cast<T>(o.height)) {}
    Rect_(const Point_<T>& a, const Point_<T>& b)
        : x(std::min(a.x, b.x)), y(std::min(a.y, b.y)), width(std::max(a.x, b.x) - std::min(a.x, b.x)),
          height(std::max(a.y, b.y) - std::min(a.y, b.y)) {}
    Rect_(const Point_<T>& o, const Size_<T>& s) : x(o.x), y(o.y), width(s.width), height(s.height) {}
    Point_<T> tl() const { return Point_<T>(x, y); }
    Point_<T> br() const { return Point_<T>(x + width, y + height); }
    Size_<T> size() const { return Size_<T>(width, height); }
    T area() const { return width * height; }
    bool empty() const { return width <= 0 || height <= 0; }
    bool contains(const Point_<T>& p) const { return x <= p.x && p.x < x + width && y <= p.y && p.y < y + height; }
};
template <class T>
inline Rect_<T> operator&(const Rect_<T>& a, const Rect_<T>& b) {
    T x1 = std::max(a.x, b.x), y1 = std::max(a.y, b.y);
    T x2 = std::min(a.x + a.width, b.x + b.width), y2 = std::min(a.y + a.height, b.y + b.height);
    if (x2 <= x1 || y2 <= y1) return Rect_<T>();
    return Rect_<T>(x1, y1, x2 - x1, y2 - y1);
}
typedef Rect_<int> Rect;
typedef Rect_<int> Rect2i;
typedef Rect_<float> Rect2f;

template <class T, int N>
struct Vec {
    T val[N];
    Vec() { for (int i = 0; i < N; i++) val[i] = T(0); }
    Vec(T a, T b) { val[0] = a; val[1] = b; }
    Vec(T a, T b, T c) { val[0] = a; val[1] = b; val[2] = c; }
    Vec(T a, T b, T c, T d) { val[0] = a; val[1] = b; val[2] = c; val[3] = d; }
    T& operator[](int i) { return val[i]; }
    const T& operator[](int i) const { return val[i]; }
};
typedef Vec<uchar, 3> Vec3b;
typedef Vec<float, 2> Vec2f;
typedef Vec<float, 3> Vec3f;
typedef Vec<float, 4> Vec4f;
typedef Vec<double, 3> Vec3d;
typedef Vec<int, 4> Vec4i;

template <class T>
struct Scalar_ : Vec<T, 4> {
    Scalar_() {}
    Scalar_(T a, T b = 0, T c = 0, T d = 0) : Vec<T, 4>(a, b, c, d) {}
    static Scalar_ all(T v) { return Scalar_(v, v, v, v); }
};
typedef Scalar_<double> Scalar;

struct TermCriteria {
    enum Type { COUNT = 1, MAX_ITER = COUNT, EPS = 2 };
    int type, maxCount;
    double epsilon;
    TermCriteria() : type(0), maxCount(0), epsilon(0) {}
    TermCriteria(int t, int c, double e) : type(t), maxCount(c), epsilon(e) {}
};

enum { OPTFLOW_USE_INITIAL_FLOW = 4, OPTFLOW_LK_GET_MIN_EIGENVALS = 8 };
enum BorderTypes { BORDER_CONSTANT = 0, BORDER_REPLICATE = 1, BORDER_REFLECT = 2, BORDER_WRAP = 3, BORDER_REFLECT_101 = 4,
                   BORDER_DEFAULT = 4 };
enum DecompTypes { DECOMP_LU = 0, DECOMP_SVD = 1, DECOMP_EIG = 2, DECOMP_CHOLESKY = 3, DECOMP_QR = 4, DECOMP_NORMAL = 16 };
enum MorphShapes { MORPH_RECT = 0, MORPH_CROSS = 1, MORPH_ELLIPSE = 2 };
enum MorphTypes { MORPH_ERODE = 0, MORPH_DILATE = 1 };
enum HersheyFonts { FONT_HERSHEY_SIMPLEX = 0 };
enum ColorConversionCodes { COLOR_BGR2GRAY = 6, COLOR_GRAY2BGR = 8 };
enum InterpolationFlags { INTER_NEAREST = 0, INTER_LINEAR = 1 };
enum { FM_RANSAC = 8 };
enum LineTypes { FILLED = -1, LINE_4 = 4, LINE_8 = 8, LINE_AA = 16 };

inline int dvshim_elem_size(int type) {
    static const int depth_bytes[8] = {1, 1, 2, 2, 4, 4, 8, 2};
    return depth_bytes[type & 7] * ((type >> CV_CN_SHIFT) + 1);
}

class Mat;
struct MatExpr;

// Reference-counted dense 2-D array; `data`/`step` address either owned storage or the caller's memory.
class Mat {
public:
    int rows = 0, cols = 0, flags = 0;
    size_t step = 0;
    uchar* data = nullptr;

    Mat() {}
    Mat(int r, int c, int type) { create(r, c, type); }
    Mat(Size s, int type) { create(s.height, s.width, type); }
    Mat(int r, int c, int type, const Scalar& v) { create(r, c, type); setTo(v); }
    Mat(Size s, int type, const Scalar& v) { create(s.height, s.width, type); setTo(v); }
    Mat(int r, int c, int type, void* ext, size_t ext_step = 0) : rows(r), cols(c), flags(type), data((uchar*)ext) {
        step = ext_step ? ext_step : (size_t)c * dvshim_elem_size(type);
    }
    Mat(const Mat& m, const Rect& roi) : rows(roi.height), cols(roi.width), flags(m.flags), step(m.step), own_(m.own_) {
        data = m.data + (size_t)roi.y * m.step + (size_t)roi.x * m.elemSize();
    }
    void create(int r, int c, int type) {
        if (r == rows && c == cols && type == flags && data) return;
        rows = r; cols = c; flags = type;
        step = (size_t)c * dvshim_elem_size(type);
        own_ = std::make_shared<std::vector<uchar>>((size_t)r * step, (uchar)0);
        data = own_->data();
    }
    void create(Size s, int type) { create(s.height, s.width, type); }
    int type() const { return flags; }
    int depth() const { return flags & 7; }
    int channels() const { return (flags >> CV_CN_SHIFT) + 1; }
    size_t elemSize() const { return (size_t)dvshim_elem_size(flags); }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    Size size() const { return Size(cols, rows); }
    size_t total() const { return (size_t)rows * cols; }
    bool isContinuous() const { return step == (size_t)cols * elemSize(); }
    void release() { rows = cols = 0; data = nullptr; own_.reset(); }
    Mat clone() const {
        Mat m;
        if (empty()) return m;
        m.create(rows, cols, flags);
        for (int r = 0; r < rows; r++) std::memcpy(m.data + r * m.step, data + r * step, (size_t)cols * elemSize());
        return m;
    }
    void copyTo(Mat& dst) const { dst = clone(); }
    Mat operator()(const Rect& roi) const { return Mat(*this, roi); }
    Mat& setTo(const Scalar& v) {
        for (int r = 0; r < rows; r++)
            for (int c = 0; c < cols; c++)
                for (int k = 0; k < channels(); k++) put(r, c, k, v[k]);
        return *this;
    }
    Mat& operator=(const Scalar& v) { return setTo(v); }
    template <class T> T* ptr(int r = 0) { return (T*)(data + (size_t)r * step); }
    template <class T> const T* ptr(int r = 0) const { return (const T*)(data + (size_t)r * step); }
    uchar* ptr(int r = 0) { return data + (size_t)r * step; }
    const uchar* ptr(int r = 0) const { return data + (size_t)r * step; }
    template <class T> T& at(int r, int c) { return ((T*)(data + (size_t)r * step))[c]; }
    template <class T> const T& at(int r, int c) const { return ((const T*)(data + (size_t)r * step))[c]; }
    template <class T> T& at(int i) { return rows == 1 ? at<T>(0, i) : at<T>(i, 0); }
    template <class T> const T& at(int i) const { return rows == 1 ? at<T>(0, i) : at<T>(i, 0); }
    // Mat::at<T>(Point): a Point2f argument converts to Point by saturate_cast<int> = cvRound (ties to even), as in OpenCV
    template <class T> T& at(Point p) { return at<T>(p.y, p.x); }
    template <class T> const T& at(Point p) const { return at<T>(p.y, p.x); }
    static Mat zeros(int r, int c, int type) { return Mat(r, c, type); }
    static Mat zeros(Size s, int type) { return Mat(s.height, s.width, type); }
    static Mat ones(int r, int c, int type) { Mat m(r, c, type); m.setTo(Scalar(1)); return m; }
    static Mat eye(int r, int c, int type) { Mat m(r, c, type); for (int i = 0; i < r && i < c; i++) m.put(i, i, 0, 1.0); return m; }
    void convertTo(Mat&, int, double = 1, double = 0) const { dvshim_unreachable("cv::Mat::convertTo"); }
    std::shared_ptr<std::vector<uchar>> own_;

private:
    void put(int r, int c, int k, double v) {
        uchar* p = data + (size_t)r * step + (size_t)c * elemSize();
        switch (depth()) {
            case CV_8U: ((uchar*)p)[k] = saturate_cast<uchar>(v); break;
            case CV_16S: ((short*)p)[k] = (short)cvRound(v); break;
            case CV_16U: ((ushort*)p)[k] = (ushort)cvRound(v); break;
            case CV_32S: ((int*)p)[k] = cvRound(v); break;
            case CV_32F: ((float*)p)[k] = (float)v; break;
            case CV_64F: ((double*)p)[k] = v; break;
            default: dvshim_unreachable("cv::Mat::setTo for this depth");
        }
    }
};
template <class T>
class Mat_ : public Mat {
public:
    Mat_() {}
    Mat_(int r, int c) { dvshim_unreachable("cv::Mat_"); (void)r; (void)c; }
    T& operator()(int r, int c) { return at<T>(r, c); }
};

// Array argument wrappers: the reference passes cv::Mat and std::vector<Point2f>/<uchar>/<float>
class _InputArray {
public:
    _InputArray() {}
    _InputArray(const Mat& m) : mat_(&m) {}
    bool needed() const { return mat_ != nullptr; }
    Mat getMat() const { return mat_ ? *mat_ : Mat(); }
    void create(int, int, int) const { dvshim_unreachable("cv::OutputArray::create"); }
protected:
    const Mat* mat_ = nullptr;
};
typedef const _InputArray& InputArray;
typedef const _InputArray& OutputArray;
typedef const _InputArray& InputOutputArray;
inline const _InputArray& noArray() { static _InputArray none; return none; }

class FileNode {
public:
    bool isNone() const { return true; }
    bool empty() const { return true; }
    FileNode operator[](const char*) const { return FileNode(); }
    FileNode operator[](const std::string&) const { return FileNode(); }
    operator int() const { dvshim_unreachable("cv::FileNode"); }
    operator float() const { dvshim_unreachable("cv::FileNode"); }
    operator double() const { dvshim_unreachable("cv::FileNode"); }
    operator std::string() const { dvshim_unreachable("cv::FileNode"); }
};
template <class T>
inline void operator>>(const FileNode&, T&) { dvshim_unreachable("cv::FileNode >>"); }
class FileStorage {
public:
    enum Mode { READ = 0, WRITE = 1 };
    FileStorage() {}
    FileStorage(const std::string&, int) {}
    bool isOpened() const { return false; }      // yaml files are parsed by the harness, never by the shim
    void release() {}
    FileNode operator[](const char*) const { return FileNode(); }
    FileNode operator[](const std::string&) const { return FileNode(); }
};
template <class T>
inline FileStorage& operator<<(FileStorage& fs, const T&) { return fs; }

template <class T>
class Ptr : public std::shared_ptr<T> {
public:
    Ptr() {}
    Ptr(T* p) : std::shared_ptr<T>(p) {}
    Ptr(const std::shared_ptr<T>& p) : std::shared_ptr<T>(p) {}
    bool empty() const { return !*this; }
};
template <class T, class... A>
inline Ptr<T> makePtr(A&&... a) { return Ptr<T>(std::make_shared<T>(std::forward<A>(a)...)); }

}  // namespace cv

// ---- algorithm hooks: set by the harness (oracle/ref/ref_glue.cpp: dvref_set_hooks) ------------------------------
namespace dvshim {
struct Hooks {
    // cv::calcOpticalFlowPyrLK(prev, next, prevPts, nextPts (in/out), status, err, winSize, maxLevel, criteria, flags)
    void (*calc_optical_flow_pyr_lk)(const uchar* prev, const uchar* next, int rows, int cols, int step_prev, int step_next,
                                     const float* prev_pts, float* next_pts, int n, uchar* status, int win, int max_level,
                                     int crit_type, int crit_count, double crit_eps, int flags) = nullptr;
    // cv::goodFeaturesToTrack(image, corners, maxCorners, qualityLevel, minDistance, mask): returns the corner count
    int (*good_features_to_track)(const uchar* img, int rows, int cols, int step, float* corners, int max_corners, double quality,
                                  double min_dist, const uchar* mask, int mask_step) = nullptr;
    // cv::erode(src, dst, rectangular kernel k x k, anchor centre)
    void (*erode_rect)(const uchar* src, uchar* dst, int rows, int cols, int src_step, int dst_step, int k) = nullptr;
    // cv::circle(img, center, radius, color, FILLED)
    void (*circle_filled)(uchar* img, int rows, int cols, int step, int cx, int cy, int radius, int color) = nullptr;
    // cv::cvtColor(src BGR, dst gray, COLOR_BGR2GRAY)
    void (*bgr2gray)(const uchar* src, uchar* dst, int rows, int cols, int src_step, int dst_step) = nullptr;
    // cv::findFundamentalMat(points1, points2, method, param1, param2, mask): fills mask[n], returns 1 when a matrix was found
    int (*find_fundamental_mat)(const float* pts1, const float* pts2, int n, int method, double param1, double param2,
                                uchar* mask) = nullptr;
    // cv::cuda::GoodFeaturesToTrackDetector::detect(image, corners, mask) of a detector created with (maxCorners, qualityLevel,
    // minDistance, blockSize 3): returns the corner count (the harness restates the published cv::cuda algorithm over cv2)
    int (*good_features_cuda)(const uchar* img, int rows, int cols, int step, float* corners, int max_corners, double quality,
                              double min_dist, const uchar* mask, int mask_step) = nullptr;
};
Hooks& hooks();
}  // namespace dvshim

namespace cv {

struct dvshim_pts {       // a vector<Point2f> or a Mat of CV_32FC2 as a flat float array
    static float* of(std::vector<Point2f>& v) { return v.empty() ? nullptr : &v[0].x; }
    static const float* of(const std::vector<Point2f>& v) { return v.empty() ? nullptr : &v[0].x; }
};

inline void calcOpticalFlowPyrLK(const Mat& prev, const Mat& next, const std::vector<Point2f>& prevPts, std::vector<Point2f>& nextPts,
                                 std::vector<uchar>& status, std::vector<float>& err, Size winSize = Size(21, 21), int maxLevel = 3,
                                 TermCriteria criteria = TermCriteria(TermCriteria::COUNT + TermCriteria::EPS, 30, 0.01), int flags = 0,
                                 double minEigThreshold = 1e-4) {
    (void)minEigThreshold;
    if (!dvshim::hooks().calc_optical_flow_pyr_lk) dvshim_unreachable("cv::calcOpticalFlowPyrLK (no hook registered)");
    const int n = (int)prevPts.size();
    if (!(flags & OPTFLOW_USE_INITIAL_FLOW)) nextPts.assign(n, Point2f());
    status.assign(n, 0);
    err.assign(n, 0.f);
    if (n == 0) return;
    dvshim::hooks().calc_optical_flow_pyr_lk(prev.data, next.data, prev.rows, prev.cols, (int)prev.step, (int)next.step,
                                             dvshim_pts::of(prevPts), dvshim_pts::of(nextPts), n, status.data(), winSize.width,
                                             maxLevel, criteria.type, criteria.maxCount, criteria.epsilon, flags);
}

inline void goodFeaturesToTrack(const Mat& image, std::vector<Point2f>& corners, int maxCorners, double qualityLevel,
                                double minDistance, const Mat& mask = Mat(), int blockSize = 3, bool useHarris = false,
                                double k = 0.04) {
    (void)blockSize; (void)useHarris; (void)k;
    if (!dvshim::hooks().good_features_to_track) dvshim_unreachable("cv::goodFeaturesToTrack (no hook registered)");
    corners.clear();
    if (maxCorners <= 0) dvshim_unreachable("cv::goodFeaturesToTrack with maxCorners <= 0");
    std::vector<float> buf((size_t)2 * maxCorners);
    const int n = dvshim::hooks().good_features_to_track(image.data, image.rows, image.cols, (int)image.step, buf.data(), maxCorners,
                                                         qualityLevel, minDistance, mask.empty() ? nullptr : mask.data, (int)mask.step);
    for (int i = 0; i < n; i++) corners.emplace_back(buf[2 * i], buf[2 * i + 1]);
}

inline Mat getStructuringElement(int shape, Size ksize, Point anchor = Point(-1, -1)) {
    (void)anchor;
    if (shape != MORPH_RECT) dvshim_unreachable("cv::getStructuringElement (non-rectangular)");
    return Mat(ksize.height, ksize.width, CV_8UC1, Scalar(1));
}
inline void erode(const Mat& src, Mat& dst, const Mat& kernel, Point anchor = Point(-1, -1), int iterations = 1) {
    (void)anchor; (void)iterations;
    if (!dvshim::hooks().erode_rect) dvshim_unreachable("cv::erode (no hook registered)");
    Mat out(src.rows, src.cols, src.type());        // in-place calls (src is dst) are allowed by OpenCV
    dvshim::hooks().erode_rect(src.data, out.data, src.rows, src.cols, (int)src.step, (int)out.step, kernel.rows);
    dst = out;
}
inline void circle(Mat& img, Point center, int radius, const Scalar& color, int thickness = 1, int lineType = 8, int shift = 0) {
    (void)lineType; (void)shift;
    if (thickness >= 0) dvshim_unreachable("cv::circle (outline)");
    if (!dvshim::hooks().circle_filled) dvshim_unreachable("cv::circle (no hook registered)");
    dvshim::hooks().circle_filled(img.data, img.rows, img.cols, (int)img.step, center.x, center.y, radius, (int)color[0]);
}
// cv::circle(img, Point2f, ...): the Point2f -> Point conversion rounds (saturate_cast<int> = cvRound)
inline void cvtColor(const Mat& src, Mat& dst, int code, int dstCn = 0) {
    (void)dstCn;
    if (code != COLOR_BGR2GRAY) dvshim_unreachable("cv::cvtColor (other than BGR2GRAY)");
    if (!dvshim::hooks().bgr2gray) dvshim_unreachable("cv::cvtColor (no hook registered)");
    Mat out(src.rows, src.cols, CV_8UC1);
    dvshim::hooks().bgr2gray(src.data, out.data, src.rows, src.cols, (int)src.step, (int)out.step);
    dst = out;
}
// zero / constant padding on the bottom and right is the only form the front-end uses (InstanceImagePadding)
inline void copyMakeBorder(const Mat& src, Mat& dst, int top, int bottom, int left, int right, int borderType,
                           const Scalar& value = Scalar()) {
    if (borderType != BORDER_CONSTANT) dvshim_unreachable("cv::copyMakeBorder (non-constant border)");
    Mat out(src.rows + top + bottom, src.cols + left + right, src.type(), value);
    for (int r = 0; r < src.rows; r++)
        std::memcpy(out.data + (size_t)(r + top) * out.step + (size_t)left * out.elemSize(), src.data + (size_t)r * src.step,
                    (size_t)src.cols * src.elemSize());
    dst = out;
}

// ---- declared for parsing only ------------------------------------------------------------------------------------
inline Mat findHomography(const std::vector<Point2f>&, const std::vector<Point2f>&, int = 0, double = 3) { dvshim_unreachable("cv::findHomography"); }
inline Mat findFundamentalMat(const std::vector<Point2f>& points1, const std::vector<Point2f>& points2, int method, double param1,
                              double param2, std::vector<uchar>& mask) {
    if (!dvshim::hooks().find_fundamental_mat) dvshim_unreachable("cv::findFundamentalMat (no hook registered)");
    if (points1.size() < 7) return Mat();          // cv::findFundamentalMat: "if( npoints < 7 ) return Mat();" — the mask is not created
    mask.assign(points1.size(), 0);
    dvshim::hooks().find_fundamental_mat(points1.empty() ? nullptr : &points1[0].x, points2.empty() ? nullptr : &points2[0].x,
                                         (int)points1.size(), method, param1, param2, mask.empty() ? nullptr : mask.data());
    return Mat();      // the callers on the parity path (RejectWithF) use the mask only
}
inline bool solve(const Mat&, const Mat&, Mat&, int = DECOMP_LU) { dvshim_unreachable("cv::solve"); }
inline void convertMaps(const Mat&, const Mat&, Mat&, Mat&, int, bool = false) { dvshim_unreachable("cv::convertMaps"); }
template <class A, class B>
inline bool solvePnP(const A&, const B&, const Mat&, const _InputArray&, Mat&, Mat&, bool = false, int = 0) { dvshim_unreachable("cv::solvePnP"); }
inline void Rodrigues(const Mat&, Mat&) { dvshim_unreachable("cv::Rodrigues"); }
inline Size getTextSize(const std::string&, int, double, int, int*) { dvshim_unreachable("cv::getTextSize"); }
inline void rectangle(Mat&, Point, Point, const Scalar&, int = 1, int = 8, int = 0) { dvshim_unreachable("cv::rectangle"); }
inline void rectangle(Mat&, Rect, const Scalar&, int = 1, int = 8, int = 0) { dvshim_unreachable("cv::rectangle"); }
inline void putText(Mat&, const std::string&, Point, int, double, Scalar, int = 1, int = 8, bool = false) { dvshim_unreachable("cv::putText"); }
inline void line(Mat&, Point, Point, const Scalar&, int = 1, int = 8, int = 0) { dvshim_unreachable("cv::line"); }
inline void arrowedLine(Mat&, Point, Point, const Scalar&, int = 1, int = 8, int = 0, double = 0.1) { dvshim_unreachable("cv::arrowedLine"); }
inline void remap(const Mat&, Mat&, const Mat&, const Mat&, int, int = BORDER_CONSTANT, const Scalar& = Scalar()) { dvshim_unreachable("cv::remap"); }
inline void hconcat(const Mat&, const Mat&, Mat&) { dvshim_unreachable("cv::hconcat"); }
inline void vconcat(const Mat&, const Mat&, Mat&) { dvshim_unreachable("cv::vconcat"); }
inline int countNonZero(const Mat&) { dvshim_unreachable("cv::countNonZero"); }
inline void imshow(const std::string&, const Mat&) { dvshim_unreachable("cv::imshow"); }
inline int waitKey(int = 0) { dvshim_unreachable("cv::waitKey"); }
inline Mat imread(const std::string&, int = 1) { dvshim_unreachable("cv::imread"); }
inline bool imwrite(const std::string&, const Mat&) { dvshim_unreachable("cv::imwrite"); }

template <class T, int R, int C, class E>
inline void cv2eigen(const Mat&, E&) { dvshim_unreachable("cv::cv2eigen"); }
template <class E>
inline void cv2eigen(const Mat&, E&) { dvshim_unreachable("cv::cv2eigen"); }
template <class E>
inline void eigen2cv(const E&, Mat&) { dvshim_unreachable("cv::eigen2cv"); }

namespace cuda {
// The reference's cv::cuda objects are host-backed here: a GpuMat is a Mat, upload/download are copies, the morphology
// filter and the sparse LK object forward to the same hooks as their CPU counterparts (so the cv::cuda CALL PATTERN of the
// reference runs -- backward pass over all levels, 1 px round-trip test -- on the CPU arithmetic; cv::cuda's own fp32
// texture arithmetic is not reproduced).
class Stream {
public:
    static Stream& Null() { static Stream s; return s; }
};
class GpuMat {
public:
    int rows = 0, cols = 0;
    Mat m;
    GpuMat() {}
    GpuMat(const Mat& host) { upload(host); }
    GpuMat(int r, int c, int type) { create(r, c, type); }
    GpuMat(Size s, int type, void* data) : rows(s.height), cols(s.width), m(s.height, s.width, type, data) {}   // a header over the caller's bytes
    void create(int r, int c, int type) { m.create(r, c, type); rows = r; cols = c; }
    bool empty() const { return m.empty(); }
    void upload(const Mat& host) { m = host.clone(); rows = m.rows; cols = m.cols; }
    void download(Mat& host) const {
        host.create(m.rows, m.cols, m.type());
        for (int r = 0; r < m.rows; r++) std::memcpy(host.data + (size_t)r * host.step, m.data + (size_t)r * m.step, (size_t)m.cols * m.elemSize());
    }
    GpuMat clone() const { GpuMat g; g.m = m.clone(); g.rows = rows; g.cols = cols; return g; }
    Size size() const { return Size(cols, rows); }
    int type() const { return m.type(); }
    int channels() const { return m.channels(); }
    GpuMat operator()(const Rect& r) const { GpuMat g; g.m = m(r); g.rows = r.height; g.cols = r.width; return g; }
    void copyTo(GpuMat& dst) const { dst = clone(); }
    void setTo(const Scalar& v) { m.setTo(v); }
};
class SparsePyrLKOpticalFlow {
public:
    virtual ~SparsePyrLKOpticalFlow() {}
    virtual void calc(const GpuMat& prev, const GpuMat& next, const GpuMat& prevPts, GpuMat& nextPts, GpuMat& status) = 0;
    static Ptr<SparsePyrLKOpticalFlow> create(Size winSize = Size(21, 21), int maxLevel = 3, int iters = 30, bool useInitialFlow = false);
};
class dvshim_HostedLK : public SparsePyrLKOpticalFlow {
public:
    dvshim_HostedLK(Size w, int l, int it, bool init) : win_(w), max_level_(l), iters_(it), use_initial_(init) {}
    void calc(const GpuMat& prev, const GpuMat& next, const GpuMat& prevPts, GpuMat& nextPts, GpuMat& status) override {
        if (!dvshim::hooks().calc_optical_flow_pyr_lk) dvshim_unreachable("cv::cuda::SparsePyrLKOpticalFlow::calc (no hook registered)");
        const int n = prevPts.cols;
        if (!use_initial_ || nextPts.cols != n) nextPts.create(1, n, CV_32FC2);
        status.create(1, n, CV_8UC1);
        if (n == 0) return;
        dvshim::hooks().calc_optical_flow_pyr_lk(prev.m.data, next.m.data, prev.rows, prev.cols, (int)prev.m.step, (int)next.m.step,
                                                 (const float*)prevPts.m.data, (float*)nextPts.m.data, n, status.m.data, win_.width,
                                                 max_level_, TermCriteria::COUNT + TermCriteria::EPS, iters_, 0.01,
                                                 use_initial_ ? OPTFLOW_USE_INITIAL_FLOW : 0);
    }
private:
    Size win_;
    int max_level_, iters_;
    bool use_initial_;
};
inline Ptr<SparsePyrLKOpticalFlow> SparsePyrLKOpticalFlow::create(Size winSize, int maxLevel, int iters, bool useInitialFlow) {
    return Ptr<SparsePyrLKOpticalFlow>(std::shared_ptr<SparsePyrLKOpticalFlow>(new dvshim_HostedLK(winSize, maxLevel, iters, useInitialFlow)));
}
class CornersDetector {
public:
    virtual ~CornersDetector() {}
    virtual void detect(const GpuMat&, GpuMat&, const GpuMat&) = 0;
};
class dvshim_HostedGftt : public CornersDetector {
public:
    dvshim_HostedGftt(int n, double q, double d) : max_corners_(n), quality_(q), min_dist_(d) {}
    void detect(const GpuMat& image, GpuMat& corners, const GpuMat& mask) override {
        if (!dvshim::hooks().good_features_cuda) dvshim_unreachable("cv::cuda::CornersDetector::detect (no hook registered)");
        std::vector<float> buf((size_t)2 * (max_corners_ > 0 ? max_corners_ : 1));
        const int n = dvshim::hooks().good_features_cuda(image.m.data, image.rows, image.cols, (int)image.m.step, buf.data(), max_corners_,
                                                         quality_, min_dist_, mask.empty() ? nullptr : mask.m.data,
                                                         mask.empty() ? 0 : (int)mask.m.step);
        if (n <= 0) { corners = GpuMat(); return; }              // _corners.release()
        corners.create(1, n, CV_32FC2);
        std::memcpy(corners.m.data, buf.data(), sizeof(float) * 2 * (size_t)n);
    }
private:
    int max_corners_;
    double quality_, min_dist_;
};
inline Ptr<CornersDetector> createGoodFeaturesToTrackDetector(int, int maxCorners = 1000, double qualityLevel = 0.01, double minDistance = 0.0,
                                                              int blockSize = 3, bool useHarris = false, double = 0.04) {
    if (blockSize != 3 || useHarris) dvshim_unreachable("cv::cuda::createGoodFeaturesToTrackDetector (blockSize != 3 or Harris)");
    return Ptr<CornersDetector>(std::shared_ptr<CornersDetector>(new dvshim_HostedGftt(maxCorners, qualityLevel, minDistance)));
}
class Filter {
public:
    virtual ~Filter() {}
    virtual void apply(const GpuMat& src, GpuMat& dst) = 0;
};
class dvshim_HostedErode : public Filter {
public:
    explicit dvshim_HostedErode(const Mat& k) : kernel_(k) {}
    void apply(const GpuMat& src, GpuMat& dst) override {
        Mat out;
        cv::erode(src.m, out, kernel_);
        dst.m = out; dst.rows = out.rows; dst.cols = out.cols;
    }
private:
    Mat kernel_;
};
inline Ptr<Filter> createMorphologyFilter(int op, int, const Mat& kernel, Point = Point(-1, -1), int = 1) {
    if (op != MORPH_ERODE) dvshim_unreachable("cv::cuda::createMorphologyFilter (other than erode)");
    return Ptr<Filter>(std::shared_ptr<Filter>(new dvshim_HostedErode(kernel)));
}
inline void cvtColor(const GpuMat&, GpuMat&, int, int = 0) { dvshim_unreachable("cv::cuda::cvtColor"); }
inline void bitwise_not(const GpuMat& src, GpuMat& dst) {
    if (src.m.type() != CV_8UC1) dvshim_unreachable("cv::cuda::bitwise_not (other than CV_8UC1)");
    GpuMat out(src.rows, src.cols, src.m.type());
    for (int r = 0; r < src.rows; r++)
        for (int c = 0; c < src.cols; c++) out.m.data[(size_t)r * out.m.step + c] = (uchar)~src.m.data[(size_t)r * src.m.step + c];
    dst = out;
}
inline void scaleAdd(const GpuMat&, double, const GpuMat&, GpuMat&) { dvshim_unreachable("cv::cuda::scaleAdd"); }
}  // namespace cuda

}  // namespace cv
