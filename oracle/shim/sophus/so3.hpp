// ORACLE shim (test infrastructure): Sophus is named by a reference header but unused on the parity path
#pragma once
