// ORACLE shim (test infrastructure).  Stands in for dynamic_vins/src/det2d/detector2d.h (SOLOv2 / TensorRT instance
// segmentation, out of scope).  background_tracker.h includes it but uses nothing from it.
#pragma once
#include "basic/def.h"
#include "basic/semantic_image.h"
