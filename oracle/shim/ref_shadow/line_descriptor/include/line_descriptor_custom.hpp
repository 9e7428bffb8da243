// ORACLE shim (test infrastructure).  The reference's line_descriptor sources are not vendored in /root/reference and line
// features are off on the parity path (`use_line: 0`); only the KeyLine type has to exist for the headers to parse.
#pragma once
#include <opencv2/opencv.hpp>
namespace cv { namespace line_descriptor {
struct KeyLine {
    float angle = 0, response = 0, lineLength = 0;
    int class_id = 0, octave = 0, numOfPixels = 0;
    Point2f pt;
    float startPointX = 0, startPointY = 0, endPointX = 0, endPointY = 0;
    float sPointInOctaveX = 0, sPointInOctaveY = 0, ePointInOctaveX = 0, ePointInOctaveY = 0;
    Point2f getStartPoint() const { return Point2f(startPointX, startPointY); }
    Point2f getEndPoint() const { return Point2f(endPointX, endPointY); }
};
class LSDDetectorC { public: struct LSDOptions { int refine = 0; double scale = 0, sigma_scale = 0, quant = 0, ang_th = 0, log_eps = 0, density_th = 0; int n_bins = 0; double min_length = 0; }; };
class BinaryDescriptor {};
class BinaryDescriptorMatcher {};
} }
