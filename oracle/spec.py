"""ORACLE (test infrastructure): ctypes access to oracle/spec.c (plain-C arithmetic restatement).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libdvfe_spec.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "spec.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.spec_lk.restype = C.c_int
        _lib.spec_good_features.restype = C.c_int
        _lib.spec_gftt_select.restype = C.c_int
        _lib.spec_pyr_levels.restype = C.c_int
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def _u8(a):
    a = np.ascontiguousarray(a, dtype=np.uint8)
    return a


def pyr_down(img):
    img = _u8(img)
    h, w = img.shape
    out = np.zeros(((h + 1) // 2, (w + 1) // 2), np.uint8)
    lib().spec_pyr_down(_p(img, C.c_uint8), w, h, w, _p(out, C.c_uint8), out.shape[1])
    return out


def pyr_levels(w, h, max_level):
    return lib().spec_pyr_levels(w, h, max_level)


def scharr(img):
    img = _u8(img)
    h, w = img.shape
    out = np.zeros((h, w, 2), np.int16)
    lib().spec_scharr(_p(img, C.c_uint8), w, h, w, _p(out, C.c_int16))
    return out


def lk(img1, img2, pts1, pts2_init=None, max_level=3, exact_int=False):
    img1, img2 = _u8(img1), _u8(img2)
    h, w = img1.shape
    p1 = np.ascontiguousarray(pts1, np.float32).reshape(-1, 2)
    n = len(p1)
    use_init = pts2_init is not None
    p2 = np.ascontiguousarray(pts2_init, np.float32).reshape(-1, 2).copy() if use_init else np.zeros((n, 2), np.float32)
    st = np.zeros(n, np.uint8)
    lib().spec_lk(_p(img1, C.c_uint8), _p(img2, C.c_uint8), w, h, w, w, _p(p1, C.c_float), _p(p2, C.c_float),
                  _p(st, C.c_uint8), n, max_level, int(use_init), int(exact_int))
    return p2, st


def feature_track_by_lk(img1, img2, pts1, flow_back=True, max_level=3, exact_int=False, return_rev=False):
    img1, img2 = _u8(img1), _u8(img2)
    h, w = img1.shape
    p1 = np.ascontiguousarray(pts1, np.float32).reshape(-1, 2)
    n = len(p1)
    p2 = np.zeros((n, 2), np.float32)
    rev = np.zeros((n, 2), np.float32)
    st = np.zeros(n, np.uint8)
    lib().spec_feature_track_by_lk(_p(img1, C.c_uint8), _p(img2, C.c_uint8), w, h, w, w, _p(p1, C.c_float),
                                   _p(p2, C.c_float), _p(st, C.c_uint8), n, int(flow_back), max_level,
                                   int(exact_int), _p(rev, C.c_float))
    return (p2, st, rev) if return_rev else (p2, st)


def min_eigen_val(img):
    img = _u8(img)
    h, w = img.shape
    out = np.zeros((h, w), np.float32)
    lib().spec_min_eigen_val(_p(img, C.c_uint8), w, h, w, _p(out, C.c_float))
    return out


def gftt_select(eig, mask, max_corners, quality, min_dist, unmasked_max=False):
    """stages 2..6 of goodFeaturesToTrack on a response map; unmasked_max: cv::cuda detector's threshold (whole-image maximum)"""
    eig = np.ascontiguousarray(eig, np.float32)
    h, w = eig.shape
    out = np.zeros((max(max_corners, 1) if max_corners > 0 else h * w, 2), np.float32)
    m = _u8(mask) if mask is not None else None
    ncand = C.c_int(0)
    lib().spec_gftt_select_ex.restype = C.c_int
    n = lib().spec_gftt_select_ex(_p(eig, C.c_float), w, h, _p(m, C.c_uint8) if m is not None else None, w,
                                  int(max_corners), C.c_double(quality), C.c_double(min_dist), _p(out, C.c_float),
                                  C.byref(ncand), int(bool(unmasked_max)))
    return out[:n].copy(), ncand.value


def good_features(img, mask, max_corners, quality, min_dist):
    return gftt_select(min_eigen_val(img), mask, max_corners, quality, min_dist)[0]


def disc_mask(mask, pts, r):
    mask = _u8(mask).copy()
    h, w = mask.shape
    p = np.ascontiguousarray(pts, np.float32).reshape(-1, 2)
    lib().spec_disc_mask(_p(mask, C.c_uint8), w, h, w, _p(p, C.c_float), len(p), int(r))
    return mask


def erode_rect(mask, k):
    mask = _u8(mask)
    h, w = mask.shape
    out = np.zeros_like(mask)
    lib().spec_erode_rect(_p(mask, C.c_uint8), w, h, w, _p(out, C.c_uint8), w, int(k))
    return out


def lift(cam, pts, off=(0.0, 0.0)):
    c = np.array([cam[k] for k in ("fx", "fy", "cx", "cy", "k1", "k2", "p1", "p2")], np.float64)
    p = np.ascontiguousarray(pts, np.float32).reshape(-1, 2)
    out = np.zeros_like(p)
    lib().spec_lift(_p(c, C.c_double), _p(p, C.c_float), len(p), C.c_float(off[0]), C.c_float(off[1]),
                    _p(out, C.c_float))
    return out


def remap(src, map1, map2):
    src = _u8(src)
    h, w = src.shape[:2]
    ch = 1 if src.ndim == 2 else src.shape[2]
    m1 = np.ascontiguousarray(map1, np.int16)
    m2 = np.ascontiguousarray(map2, np.uint16)
    out = np.zeros_like(src)
    lib().spec_remap(_p(src, C.c_uint8), w, h, ch, w * ch, _p(m1, C.c_int16), _p(m2, C.c_uint16), _p(out, C.c_uint8), w * ch)
    return out


def bgr_to_gray(bgr):
    bgr = _u8(bgr)
    h, w, _ = bgr.shape
    out = np.zeros((h, w), np.uint8)
    lib().spec_bgr_to_gray(_p(bgr, C.c_uint8), w, h, 3 * w, _p(out, C.c_uint8), w)
    return out


def find_fundamental_mat(pts1, pts2, thresh=1.0, conf=0.99):
    """cv::findFundamentalMat(pts1, pts2, FM_RANSAC, thresh, conf, mask): (F 3x3 | None, mask u8[n] | None)."""
    a = np.ascontiguousarray(pts1, np.float32).reshape(-1, 2)
    b = np.ascontiguousarray(pts2, np.float32).reshape(-1, 2)
    n = len(a)
    F = np.zeros(9, np.float64)
    mask = np.zeros(max(n, 1), np.uint8)
    lib().spec_find_fundamental_mat.restype = C.c_int
    ok = lib().spec_find_fundamental_mat(_p(a, C.c_float), _p(b, C.c_float), n, C.c_double(thresh), C.c_double(conf),
                                         _p(F, C.c_double), _p(mask, C.c_uint8))
    return (F.reshape(3, 3), mask[:n].copy()) if ok else (None, None)


def reject_with_f(cam, cur_pts, prev_pts, col, row, f_threshold=1.0, return_un=False):
    """InstsFeatManager::RejectWithF: status u8[n] (empty when n < 7)."""
    c = np.array([cam[k] for k in ("fx", "fy", "cx", "cy", "k1", "k2", "p1", "p2")], np.float64)
    a = np.ascontiguousarray(cur_pts, np.float32).reshape(-1, 2)
    b = np.ascontiguousarray(prev_pts, np.float32).reshape(-1, 2)
    n = len(a)
    st = np.zeros(max(n, 1), np.uint8)
    un = np.zeros((2, max(n, 1), 2), np.float32)
    lib().spec_reject_with_f.restype = C.c_int
    k = lib().spec_reject_with_f(_p(c, C.c_double), _p(a, C.c_float), _p(b, C.c_float), n, int(col), int(row), C.c_double(f_threshold),
                                 _p(st, C.c_uint8), _p(un, C.c_float))
    return (st[:k].copy(), un[:, :n].copy()) if return_un else st[:k].copy()


def detect_extra_points(mask, disp, box_xy, fx, fy, cx, cy, baseline):
    """InstFeat::DetectExtraPoints: (n, 3) float64 points (x, y, depth) of one instance ROI."""
    mask = _u8(mask)
    rows, cols = mask.shape
    d = np.ascontiguousarray(disp, np.float32)
    out = np.zeros((rows * cols + 1, 3), np.float64)
    lib().spec_detect_extra_points.restype = C.c_int
    n = lib().spec_detect_extra_points(_p(mask, C.c_uint8), rows, cols, cols, _p(d, C.c_float), d.shape[1], int(box_xy[0]), int(box_xy[1]),
                                       C.c_float(fx), C.c_float(fy), C.c_float(cx), C.c_float(cy), C.c_float(baseline),
                                       _p(out, C.c_double))
    return out[:n].copy()
