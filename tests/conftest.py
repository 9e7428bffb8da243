import os
import sys
import zlib

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _has_gpu() -> bool:
    try:
        import ctypes
        cudart = ctypes.CDLL("libcudart.so")
        n = ctypes.c_int(0)
        return cudart.cudaGetDeviceCount(ctypes.byref(n)) == 0 and n.value > 0
    except OSError:
        try:
            import torch
            return torch.cuda.is_available()
        except Exception:
            return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build libdvfe.so and the oracle's C restatement once (cross-compiles without a GPU)."""
    import __graft_entry__ as g
    g.build()


def crc(a: np.ndarray) -> int:
    return zlib.crc32(np.ascontiguousarray(a).tobytes())


def load_golden(name: str):
    return np.load(os.path.join(GOLDEN_DIR, name), allow_pickle=False)


def feature_map_arrays(points: dict):
    """{id: [(cam, v7), ...]} -> (ids, cams, v) flat arrays in (id, cam) order."""
    ids, cams, vs = [], [], []
    for fid in sorted(points):
        for cam, v in points[fid]:
            ids.append(fid); cams.append(cam); vs.append(np.asarray(v, np.float64))
    return (np.asarray(ids, np.uint32), np.asarray(cams, np.int32),
            np.asarray(vs, np.float64).reshape(-1, 7))
