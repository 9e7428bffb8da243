// ORACLE shim (test infrastructure).  Stands in for dynamic_vins/src/utils/convert_utils.h (ROS / tf message
// conversions): only the two point-cloud helpers dynamic_tracker.cpp names are declared; extra-point clustering is outside
// the parity path and aborts when reached.
#pragma once
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>

#include "basic/def.h"

namespace dynamic_vins {
inline pcl::PointCloud<pcl::PointXYZ>::Ptr EigenToPclXYZ(const std::vector<Vec3d>& pts) {
    pcl::PointCloud<pcl::PointXYZ>::Ptr pc(new pcl::PointCloud<pcl::PointXYZ>);
    for (auto& p : pts) pc->points.emplace_back((float)p.x(), (float)p.y(), (float)p.z());
    return pc;
}
inline pcl::PointCloud<pcl::PointXYZRGB>::Ptr EigenToPclXYZRGB(const std::vector<Vec3d>&) { pcl::dvshim_no_pcl(); }
template <class P>
inline std::vector<Vec3d> PclToEigen(const typename pcl::PointCloud<P>::Ptr&) { pcl::dvshim_no_pcl(); }
}
