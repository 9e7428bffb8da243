// ORACLE shim (test infrastructure).  Stands in for dynamic_vins/src/mot/deep_sort.h (DeepSORT re-identification network,
// out of scope): the type InstsFeatManager holds a pointer to.  Its methods abort when reached.
#pragma once
#include <array>
#include <memory>
#include <string>
#include <vector>

#include <opencv2/opencv.hpp>
#include <torch/torch.h>

#include "basic/box2d.h"
#include "basic/def.h"

namespace dynamic_vins {
class DeepSORT {
public:
    using Ptr = std::unique_ptr<DeepSORT>;
    explicit DeepSORT(const std::string&, const std::array<int64_t, 2>&) {}
    std::vector<Box2D::Ptr> update(const std::vector<Box2D::Ptr>&, cv::Mat) { torch::dvshim_no_tensor(); }
};
}
