// ORACLE shim (test infrastructure): PCL is named by reference headers; extra-point clustering is outside the parity path
#pragma once
#include "dvshim_pcl.hpp"
