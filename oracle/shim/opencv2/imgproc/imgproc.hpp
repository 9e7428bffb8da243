// ORACLE shim (test infrastructure): forwards to the OpenCV API stand-in
#pragma once
#include "dvshim_eigen.hpp"
#include "dvshim_cv.hpp"
