"""Seam-level operators (include/dvfe.h `dvfe_op_*`) on numpy arrays.  Each call goes through the C ABI into
the CUDA kernels; nothing here computes on the CPU."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

from . import _lib as L


def _u8(a):
    return np.ascontiguousarray(a, dtype=np.uint8)


def build_pyramid(img, max_level: int = 3):
    """cv::buildOpticalFlowPyramid(img, Size(21,21), max_level) -> list of level images."""
    img = _u8(img)
    h, w = img.shape
    outs, lw, lh = [], w, h
    for _ in range(max_level + 1):
        outs.append(np.zeros((lh, lw), np.uint8))
        lw, lh = (lw + 1) // 2, (lh + 1) // 2
    arr = (C.c_void_p * len(outs))(*[o.ctypes.data for o in outs])
    ws = np.zeros(8, np.int32)
    hs = np.zeros(8, np.int32)
    n = C.c_int(0)
    L.check(L.lib().dvfe_op_build_pyramid(L.ptr(img), w, h, img.strides[0], max_level, arr, L.ptr(ws), L.ptr(hs),
                                          C.byref(n)))
    return [outs[i][:hs[i], :ws[i]] for i in range(n.value)]


def build_pyramid_bordered(img, level: int, border: int = 21, max_level: int = 3):
    """One pyramid level with `border` pixels of its REFLECT_101 border (what LK reads), dense."""
    img = _u8(img)
    h, w = img.shape
    lw, lh = w, h
    for _ in range(level):
        lw, lh = (lw + 1) // 2, (lh + 1) // 2
    out = np.zeros((lh + 2 * border, lw + 2 * border), np.uint8)
    L.check(L.lib().dvfe_op_build_pyramid_bordered(L.ptr(img), w, h, img.strides[0], max_level, level, border, L.ptr(out)))
    return out


def feature_track_by_lk(img1, img2, pts1, flow_back: bool = True, max_level: int = 3, mask=None,
                        return_rev: bool = False):
    """FeatureTrackByLK (dynamic_vins/src/front_end/feature_utils.cpp:35-69) -> (pts2, status[, rev])."""
    img1, img2 = _u8(img1), _u8(img2)
    h, w = img1.shape
    p1 = np.ascontiguousarray(pts1, np.float32).reshape(-1, 2)
    n = len(p1)
    p2 = np.zeros((n, 2), np.float32)
    rev = np.zeros((n, 2), np.float32)
    st = np.zeros(n, np.uint8)
    m = _u8(mask) if mask is not None else None
    L.check(L.lib().dvfe_op_lk(L.ptr(img1), L.ptr(img2), w, h, w, L.ptr(p1), n, int(flow_back), max_level,
                               L.ptr(m), w, L.ptr(p2), L.ptr(st), L.ptr(rev)))
    return (p2, st, rev) if return_rev else (p2, st)


def min_eigen_val(img):
    img = _u8(img)
    h, w = img.shape
    out = np.zeros((h, w), np.float32)
    L.check(L.lib().dvfe_op_min_eigen_val(L.ptr(img), w, h, w, L.ptr(out)))
    return out


def good_features(img, max_corners: int, quality: float, min_dist: float, mask=None, eig=None,
                  return_n_candidates: bool = False, cuda_semantics: bool = False):
    """cv::goodFeaturesToTrack(img, K, quality, min_dist, mask); `eig` overrides the response map.  cuda_semantics: the
    cv::cuda detector's threshold (0.01 x the maximum of the whole response map), DetectShiTomasiCornersGpu."""
    img = _u8(img) if img is not None else None
    h, w = (img.shape if img is not None else eig.shape)
    e = np.ascontiguousarray(eig, np.float32) if eig is not None else None
    m = _u8(mask) if mask is not None else None
    out = np.zeros((max_corners, 2), np.float32)
    n, nc = C.c_int(0), C.c_int(0)
    fn = L.lib().dvfe_op_good_features_cuda if cuda_semantics else L.lib().dvfe_op_good_features
    L.check(fn(L.ptr(img), w, h, w, L.ptr(e), L.ptr(m), w, int(max_corners), float(quality), float(min_dist), L.ptr(out), C.byref(n),
               C.byref(nc)))
    res = out[:n.value].copy()
    return (res, nc.value) if return_n_candidates else res


def disc_mask(mask, pts, radius: int):
    m = _u8(mask).copy()
    h, w = m.shape
    p = np.ascontiguousarray(pts, np.float32).reshape(-1, 2)
    L.check(L.lib().dvfe_op_disc_mask(L.ptr(m), w, h, w, L.ptr(p), len(p), int(radius)))
    return m


def erode_rect(mask, k: int):
    m = _u8(mask)
    h, w = m.shape
    out = np.zeros_like(m)
    L.check(L.lib().dvfe_op_erode_rect(L.ptr(m), w, h, w, int(k), L.ptr(out)))
    return out


def lift_projective(cam: dict, pts, off=(0.0, 0.0)):
    c = L.Camera(**{k: float(cam[k]) for k in ("fx", "fy", "cx", "cy", "k1", "k2", "p1", "p2")})
    p = np.ascontiguousarray(pts, np.float32).reshape(-1, 2)
    out = np.zeros_like(p)
    L.check(L.lib().dvfe_op_lift_projective(C.byref(c), L.ptr(p), len(p), float(off[0]), float(off[1]), L.ptr(out)))
    return out


def bgr_to_gray(bgr):
    """cv::cvtColor(color, gray, CV_BGR2GRAY)"""
    b = np.ascontiguousarray(bgr, dtype=np.uint8)
    h, w, _ = b.shape
    out = np.zeros((h, w), np.uint8)
    L.check(L.lib().dvfe_op_bgr_to_gray(L.ptr(b), w, h, b.strides[0], L.ptr(out)))
    return out


def remap(src, map1=None, map2=None, to_gray=False):
    """cv::remap(src, map1, map2, INTER_LINEAR) with CV_16SC2 / CV_16UC1 maps (+ cvtColor(BGR2GRAY) if to_gray)"""
    a = np.ascontiguousarray(src, dtype=np.uint8)
    h, w = a.shape[:2]
    ch = 1 if a.ndim == 2 else a.shape[2]
    m1 = None if map1 is None else np.ascontiguousarray(map1, np.int16)
    m2 = None if map2 is None else np.ascontiguousarray(map2, np.uint16)
    if m1 is not None:
        assert m1.shape == (h, w, 2) and m2.shape == (h, w)
    out = np.zeros((h, w) if (ch == 1 or to_gray) else (h, w, ch), np.uint8)
    L.check(L.lib().dvfe_op_remap(L.ptr(a), w, h, ch, a.strides[0], L.ptr(m1), L.ptr(m2), int(bool(to_gray)), L.ptr(out)))
    return out


def merge_masks(masks):
    """(n, h, w) instance masks -> (merge_mask 255 = object, inv_merge_mask)"""
    m = np.ascontiguousarray(masks, dtype=np.uint8)
    n, h, w = m.shape
    merge, inv = np.zeros((h, w), np.uint8), np.zeros((h, w), np.uint8)
    L.check(L.lib().dvfe_op_merge_masks(L.ptr(m) if n else None, n, w, h, L.ptr(merge), L.ptr(inv)))
    return merge, inv


def punch_out(merge_mask, inv_merge_mask, roi_mask, rect):
    """remove a static instance from the merged masks (system/main.cpp:219-242); returns new (merge_mask, inv_merge_mask)"""
    m = np.ascontiguousarray(merge_mask, np.uint8).copy()
    iv = np.ascontiguousarray(inv_merge_mask, np.uint8).copy()
    r = np.ascontiguousarray(roi_mask, np.uint8)
    h, w = m.shape
    x, y, rw, rh = rect
    assert r.shape == (rh, rw)
    L.check(L.lib().dvfe_op_punch_out(L.ptr(m), L.ptr(iv), w, h, L.ptr(r), r.strides[0], int(x), int(y), int(rw), int(rh)))
    return m, iv


def reject_with_f(cam: dict, cur_pts, prev_pts, col: int, row: int, f_threshold: float = 1.0):
    """InstsFeatManager::RejectWithF (front_end/dynamic_tracker.cpp:831-849) -> status u8 (empty when fewer than 7 points)."""
    c = L.Camera(**{k: float(cam[k]) for k in ("fx", "fy", "cx", "cy", "k1", "k2", "p1", "p2")})
    a = np.ascontiguousarray(cur_pts, np.float32).reshape(-1, 2)
    b = np.ascontiguousarray(prev_pts, np.float32).reshape(-1, 2)
    if len(a) != len(b):
        raise ValueError("reject_with_f: cur_pts and prev_pts must have the same length")
    st = np.zeros(max(len(a), 1), np.uint8)
    n = C.c_int(0)
    L.check(L.lib().dvfe_op_reject_with_f(C.byref(c), L.ptr(a), L.ptr(b), len(a), int(col), int(row), float(f_threshold), L.ptr(st),
                                          C.byref(n)))
    return st[:n.value].copy()


def detect_extra_points(roi_mask, disp, box_xy, fx: float, fy: float, cx: float, cy: float, baseline: float):
    """InstFeat::DetectExtraPoints (front_end/instance_feature.cpp:413-461) -> (n, 3) float64 (x, y, depth)."""
    m = _u8(roi_mask)
    rows, cols = m.shape
    d = np.ascontiguousarray(disp, np.float32)
    step = int(max(np.sqrt(0.8 * rows * cols / 1000.), 2.))
    cap = ((cols + step - 1) // step) * ((rows + step - 1) // step)
    out = np.zeros((cap, 3), np.float64)
    n = C.c_int(0)
    L.check(L.lib().dvfe_op_detect_extra_points(L.ptr(m), rows, cols, m.strides[0], L.ptr(d), d.shape[1], d.shape[0], d.strides[0] // 4,
                                                int(box_xy[0]), int(box_xy[1]), float(fx), float(fy), float(cx), float(cy), float(baseline),
                                                L.ptr(out), cap, C.byref(n)))
    return out[:n.value].copy()
