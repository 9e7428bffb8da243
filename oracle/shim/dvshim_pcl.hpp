// ORACLE shim (test infrastructure): the PCL names the reference's dynamic_tracker.cpp mentions (DetectExtraPoints, off the
// parity path).  Everything aborts when reached.
#pragma once
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <string>
#include <vector>
namespace pcl {
[[noreturn]] inline void dvshim_no_pcl() { std::fprintf(stderr, "oracle/shim: PCL is outside the parity path\n"); std::abort(); }
struct PointXYZ { float x = 0, y = 0, z = 0; PointXYZ() {} PointXYZ(float a, float b, float c) : x(a), y(b), z(c) {} };
struct PointXYZRGB { float x = 0, y = 0, z = 0; unsigned char r = 0, g = 0, b = 0; PointXYZRGB() {} PointXYZRGB(unsigned char r_, unsigned char g_, unsigned char b_) : r(r_), g(g_), b(b_) {} };
struct PointIndices { std::vector<int> indices; };
template <class P>
class PointCloud {
public:
    typedef std::shared_ptr<PointCloud<P>> Ptr;
    typedef std::shared_ptr<const PointCloud<P>> ConstPtr;
    std::vector<P> points;
    unsigned width = 0, height = 0;
    bool is_dense = true;
    size_t size() const { return points.size(); }
    bool empty() const { return points.empty(); }
    void push_back(const P& p) { points.push_back(p); }
    P& operator[](size_t i) { return points[i]; }
    const P& operator[](size_t i) const { return points[i]; }
    typename std::vector<P>::iterator begin() { return points.begin(); }
    typename std::vector<P>::iterator end() { return points.end(); }
    Ptr makeShared() const { return Ptr(new PointCloud<P>(*this)); }
};
namespace search {
template <class P>
class KdTree {
public:
    typedef std::shared_ptr<KdTree<P>> Ptr;
    void setInputCloud(const typename PointCloud<P>::ConstPtr&) { dvshim_no_pcl(); }
};
}
template <class P>
class EuclideanClusterExtraction {
public:
    void setClusterTolerance(double) {}
    void setMinClusterSize(int) {}
    void setMaxClusterSize(int) {}
    void setSearchMethod(const typename search::KdTree<P>::Ptr&) {}
    void setInputCloud(const typename PointCloud<P>::ConstPtr&) {}
    void extract(std::vector<PointIndices>&) { dvshim_no_pcl(); }
};
template <class P>
class RadiusOutlierRemoval {
public:
    void setInputCloud(const typename PointCloud<P>::ConstPtr& c) { in_ = c; }
    void setRadiusSearch(double) {}
    void setMinNeighborsInRadius(int) {}
    void filter(PointCloud<P>& out) { if (in_ && !in_->empty()) dvshim_no_pcl(); out.points.clear(); }
    typename PointCloud<P>::ConstPtr in_;
};
namespace io {
template <class P> inline int savePCDFileASCII(const std::string&, const PointCloud<P>&) { dvshim_no_pcl(); }
template <class P> inline int savePCDFile(const std::string&, const PointCloud<P>&) { dvshim_no_pcl(); }
}
template <class P> inline void getMinMax3D(const PointCloud<P>&, P&, P&) { dvshim_no_pcl(); }
}
