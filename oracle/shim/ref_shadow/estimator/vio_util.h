// ORACLE shim (test infrastructure).  Stands in for dynamic_vins/src/estimator/vio_util.h (estimator helpers, out of scope).
// background_tracker.h includes it but the front-end's point path uses nothing from it.
#pragma once
#include "basic/def.h"
#include "basic/box3d.h"
#include "basic/point_landmark.h"
