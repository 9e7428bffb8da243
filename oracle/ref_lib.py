"""ORACLE / TEST INFRASTRUCTURE ONLY.  Python harness over oracle/_ref/libdvref.so: the reference's own front-end
sources (camera_models/src/camera_models/{PinholeCamera,Camera}.cc, dynamic_vins/src/front_end/{feature_utils,
instance_feature,background_tracker,dynamic_tracker}.cpp) compiled unmodified by oracle/ref/Makefile against the stand-in
third-party headers in oracle/shim/.  The OpenCV image algorithms those sources call are served by cv2 through the ctypes
callbacks registered here, so what runs is: reference glue + reference camera model + real OpenCV arithmetic.

Used by tests/ to pin oracle/cv_front_end.py (the restatement) and, on the GPU box, to check the CUDA path against
reference-compiled code directly.  Never imported by the product.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, List, Optional, Tuple

import cv2
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libdvref.so")

_u8p = C.POINTER(C.c_ubyte)
_f32p = C.POINTER(C.c_float)


def available() -> bool:
    return os.path.exists(LIB_PATH)


class Obs(C.Structure):
    _fields_ = [("id", C.c_uint), ("cam", C.c_int), ("v", C.c_double * 7)]


class InstObs(C.Structure):
    _fields_ = [("inst_id", C.c_uint), ("id", C.c_uint), ("is_stereo", C.c_int), ("reserved", C.c_int),
                ("point", C.c_double * 3), ("vel", C.c_double * 2), ("point_right", C.c_double * 3),
                ("vel_right", C.c_double * 2), ("disp", C.c_double)]


class Box(C.Structure):
    _fields_ = [("track_id", C.c_uint), ("x", C.c_int), ("y", C.c_int), ("w", C.c_int), ("h", C.c_int),
                ("mask", C.c_void_p), ("mask_pitch", C.c_int)]


class Config(C.Structure):
    _fields_ = [("max_cnt", C.c_int), ("max_dynamic_cnt", C.c_int), ("min_dist", C.c_int), ("min_dynamic_dist", C.c_int),
                ("flow_back", C.c_int), ("use_mask_morphology", C.c_int), ("mask_morphology_size", C.c_int),
                ("width", C.c_int), ("height", C.c_int), ("stereo", C.c_int), ("dynamic", C.c_int),
                ("cam0", C.c_double * 8), ("cam1", C.c_double * 8)]


def _view(ptr, rows: int, cols: int, step: int, ch: int = 1) -> np.ndarray:
    """numpy view (no copy) of a pitched u8 image owned by the C++ side"""
    buf = (C.c_ubyte * (rows * step)).from_address(C.cast(ptr, C.c_void_p).value)
    a = np.frombuffer(buf, np.uint8).reshape(rows, step)
    a = a[:, :cols * ch]
    return a.reshape(rows, cols, ch) if ch > 1 else a


_LK_T = C.CFUNCTYPE(None, _u8p, _u8p, C.c_int, C.c_int, C.c_int, C.c_int, _f32p, _f32p, C.c_int, _u8p, C.c_int, C.c_int,
                    C.c_int, C.c_int, C.c_double, C.c_int)
_GFTT_T = C.CFUNCTYPE(C.c_int, _u8p, C.c_int, C.c_int, C.c_int, _f32p, C.c_int, C.c_double, C.c_double, _u8p, C.c_int)
_ERODE_T = C.CFUNCTYPE(None, _u8p, _u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int)
_CIRCLE_T = C.CFUNCTYPE(None, _u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int)
_GRAY_T = C.CFUNCTYPE(None, _u8p, _u8p, C.c_int, C.c_int, C.c_int, C.c_int)


def _lk(prev, nxt, rows, cols, sp, sn, p_prev, p_next, n, status, win, max_level, ct, cc, ce, flags):
    a, b = _view(prev, rows, cols, sp), _view(nxt, rows, cols, sn)
    p0 = np.ctypeslib.as_array(p_prev, (n, 2)).astype(np.float32).reshape(-1, 1, 2)
    out = np.ctypeslib.as_array(p_next, (n, 2))
    init = out.astype(np.float32).reshape(-1, 1, 2).copy() if flags & cv2.OPTFLOW_USE_INITIAL_FLOW else None
    p1, st, _ = cv2.calcOpticalFlowPyrLK(np.ascontiguousarray(a), np.ascontiguousarray(b), p0, init, winSize=(win, win),
                                         maxLevel=max_level, criteria=(ct, cc, ce), flags=flags)
    out[:] = p1.reshape(-1, 2)
    np.ctypeslib.as_array(status, (n,))[:] = st.reshape(-1)


def _gftt(img, rows, cols, step, corners, max_corners, quality, min_dist, mask, mask_step):
    a = np.ascontiguousarray(_view(img, rows, cols, step))
    m = np.ascontiguousarray(_view(mask, rows, cols, mask_step)) if mask else None
    pts = cv2.goodFeaturesToTrack(a, max_corners, quality, min_dist, mask=m)
    if pts is None:
        return 0
    pts = pts.reshape(-1, 2)
    np.ctypeslib.as_array(corners, (max_corners, 2))[:len(pts)] = pts
    return len(pts)


def _erode(src, dst, rows, cols, ss, ds, k):
    a = np.ascontiguousarray(_view(src, rows, cols, ss))
    _view(dst, rows, cols, ds)[:] = cv2.erode(a, cv2.getStructuringElement(cv2.MORPH_RECT, (k, k), (-1, -1)))


def _circle(img, rows, cols, step, cx, cy, r, color):
    a = np.ascontiguousarray(_view(img, rows, cols, step))
    cv2.circle(a, (cx, cy), r, int(color), -1)
    _view(img, rows, cols, step)[:] = a


def _gray(src, dst, rows, cols, ss, ds):
    a = np.ascontiguousarray(_view(src, rows, cols, ss, 3))
    _view(dst, rows, cols, ds)[:] = cv2.cvtColor(a, cv2.COLOR_BGR2GRAY)


_FM_T = C.CFUNCTYPE(C.c_int, _f32p, _f32p, C.c_int, C.c_int, C.c_double, C.c_double, _u8p)


def _fundamental(p1, p2, n, method, param1, param2, mask):
    a = np.ctypeslib.as_array(p1, (n, 2)).copy()
    b = np.ctypeslib.as_array(p2, (n, 2)).copy()
    F, m = cv2.findFundamentalMat(a, b, method, param1, param2)
    if F is None or m is None:          # OpenCV leaves the mask unwritten: reported as zeros (already zero-filled by the shim)
        return 0
    np.ctypeslib.as_array(mask, (n,))[:] = m.ravel()
    return 1


def _gftt_cuda(img, rows, cols, step, corners, max_corners, quality, min_dist, mask, mask_step):
    """cv::cuda::GoodFeaturesToTrackDetector::detect, restated from its published source over cv2 (oracle/cv_front_end.py
    good_features_cuda_semantics: whole-image maximum for the quality threshold, CPU response arithmetic)"""
    from oracle import cv_front_end as cvfe
    a = np.ascontiguousarray(_view(img, rows, cols, step))
    m = np.ascontiguousarray(_view(mask, rows, cols, mask_step)) if mask else None
    assert abs(quality - 0.01) < 1e-12
    pts = cvfe.good_features_cuda_semantics(a, max_corners, min_dist, m)
    if len(pts) == 0:
        return 0
    np.ctypeslib.as_array(corners, (max_corners, 2))[:len(pts)] = pts
    return len(pts)


_CALLBACKS = (_LK_T(_lk), _GFTT_T(_gftt), _ERODE_T(_erode), _CIRCLE_T(_circle), _GRAY_T(_gray))   # keep alive
_FM_CALLBACK = _FM_T(_fundamental)
_GFTT_CUDA_CALLBACK = _GFTT_T(_gftt_cuda)
_lib = None


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError("oracle/_ref/libdvref.so is missing: run `make -C oracle/ref` where /root/reference exists")
        L = C.CDLL(LIB_PATH)
        L.dvref_last_error.restype = C.c_char_p
        L.dvref_sources.restype = C.c_char_p
        L.dvref_camera_new.restype = C.c_void_p
        L.dvref_camera_new.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_double)]
        L.dvref_camera_free.argtypes = [C.c_void_p]
        for f in ("dvref_lift_projective", "dvref_distortion", "dvref_space_to_plane"):
            getattr(L, f).argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.dvref_undistorted_pts.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.dvref_in_border.argtypes = [C.c_float, C.c_float, C.c_int, C.c_int]
        L.dvref_point_distance.restype = C.c_float
        L.dvref_point_distance.argtypes = [C.c_float] * 4
        L.dvref_cv_round_f.argtypes = [C.c_float]
        L.dvref_reduce_points.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.dvref_reduce_ints.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.dvref_set_status_by_mask.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int]
        L.dvref_pts_velocity.argtypes = [C.c_double, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.dvref_feature_track_by_lk.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p,
                                                C.c_void_p, C.c_int]
        L.dvref_instance_image_padding.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.dvref_erode_mask.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.dvref_front_end_new.restype = C.c_void_p
        L.dvref_front_end_new.argtypes = [C.POINTER(Config)]
        L.dvref_front_end_free.argtypes = [C.c_void_p]
        L.dvref_track_image.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.POINTER(Obs), C.c_int]
        L.dvref_track_dynamic.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(Box),
                                          C.c_int, C.c_double, C.c_uint, C.POINTER(Obs), C.c_int, C.POINTER(C.c_int),
                                          C.POINTER(InstObs), C.c_int, C.POINTER(C.c_int)]
        L.dvref_instance_table.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.dvref_set_hooks(*[C.cast(cb, C.c_void_p) for cb in _CALLBACKS])
        L.dvref_set_hook_fundamental(C.cast(_FM_CALLBACK, C.c_void_p))
        L.dvref_set_hook_gftt_cuda(C.cast(_GFTT_CUDA_CALLBACK, C.c_void_p))
        L.dvref_serialize_points.argtypes = [C.c_char_p, C.POINTER(Obs), C.c_int]
        L.dvref_deserialize_points.argtypes = [C.c_char_p, C.POINTER(Obs), C.c_int]
        L.dvref_set_mask_and_roi.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                             C.c_void_p, C.c_void_p]
        L.dvref_track_image_naive.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_uint, C.POINTER(Obs), C.c_int]
        L.dvref_reject_with_f.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.dvref_detect_extra_points.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                                C.c_void_p, C.c_int]
        _lib = L
    return _lib


def _cam8(cam: dict):
    return (C.c_double * 8)(cam.get("k1", 0.0), cam.get("k2", 0.0), cam.get("p1", 0.0), cam.get("p2", 0.0),
                            cam["fx"], cam["fy"], cam["cx"], cam["cy"])


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class RefCamera:
    """camodocal::PinholeCamera (camera_models/src/camera_models/PinholeCamera.cc), reference-compiled"""

    def __init__(self, cam: dict, w: int = 1280, h: int = 720):
        self.h = lib().dvref_camera_new(w, h, _cam8(cam))

    def __del__(self):
        if getattr(self, "h", None) and _lib is not None:
            _lib.dvref_camera_free(self.h)
            self.h = None

    def _call(self, fn, a: np.ndarray, k_out: int) -> np.ndarray:
        a = np.ascontiguousarray(a, np.float64)
        out = np.zeros((len(a), k_out), np.float64)
        getattr(lib(), fn)(self.h, _ptr(a), len(a), _ptr(out))
        return out

    def lift_projective(self, uv):
        return self._call("dvref_lift_projective", uv, 3)

    def distortion(self, pu):
        return self._call("dvref_distortion", pu, 2)

    def space_to_plane(self, P):
        return self._call("dvref_space_to_plane", P, 2)

    def undistorted_pts(self, pts):
        pts = np.ascontiguousarray(pts, np.float32)
        out = np.zeros_like(pts)
        lib().dvref_undistorted_pts(self.h, _ptr(pts), len(pts), _ptr(out))
        return out


def in_border(x, y, rows, cols) -> bool:
    return bool(lib().dvref_in_border(float(np.float32(x)), float(np.float32(y)), rows, cols))


def point_distance(p, q) -> np.float32:
    return np.float32(lib().dvref_point_distance(*[float(np.float32(v)) for v in (p[0], p[1], q[0], q[1])]))


def cv_round(v) -> int:
    return int(lib().dvref_cv_round_f(float(np.float32(v))))


def reduce_points(pts, status) -> np.ndarray:
    pts = np.ascontiguousarray(pts, np.float32).copy()
    status = np.ascontiguousarray(status, np.uint8)
    n = lib().dvref_reduce_points(_ptr(pts), _ptr(status), len(pts))
    return pts[:n]


def set_status_by_mask(status, pts, mask) -> np.ndarray:
    status = np.ascontiguousarray(status, np.uint8).copy()
    pts = np.ascontiguousarray(pts, np.float32)
    mask = np.ascontiguousarray(mask, np.uint8)
    lib().dvref_set_status_by_mask(_ptr(status), _ptr(pts), len(pts), _ptr(mask), mask.shape[0], mask.shape[1])
    return status


def pts_velocity(dt, ids, cur_un, prev_ids, prev_un) -> np.ndarray:
    ids = np.ascontiguousarray(ids, np.uint32)
    cur_un = np.ascontiguousarray(cur_un, np.float32)
    prev_ids = np.ascontiguousarray(prev_ids, np.uint32)
    prev_un = np.ascontiguousarray(prev_un, np.float32)
    out = np.zeros_like(cur_un)
    lib().dvref_pts_velocity(float(dt), _ptr(ids), _ptr(cur_un), len(ids), _ptr(prev_ids), _ptr(prev_un), len(prev_ids), _ptr(out))
    return out


def feature_track_by_lk(img1, img2, pts1, flow_back=True) -> Tuple[np.ndarray, np.ndarray]:
    img1, img2 = np.ascontiguousarray(img1), np.ascontiguousarray(img2)
    pts1 = np.ascontiguousarray(pts1, np.float32)
    pts2 = np.zeros_like(pts1)
    st = np.zeros(len(pts1), np.uint8)
    rc = lib().dvref_feature_track_by_lk(_ptr(img1), _ptr(img2), img1.shape[0], img1.shape[1], _ptr(pts1), len(pts1), _ptr(pts2),
                                         _ptr(st), int(flow_back))
    if rc < 0:
        raise RuntimeError(lib().dvref_last_error().decode())
    return pts2, st


def instance_image_padding(a, b):
    a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
    r, c = max(a.shape[0], b.shape[0]), max(a.shape[1], b.shape[1])
    oa, ob = np.zeros((r, c), np.uint8), np.zeros((r, c), np.uint8)
    lib().dvref_instance_image_padding(_ptr(a), a.shape[0], a.shape[1], _ptr(b), b.shape[0], b.shape[1], _ptr(oa), _ptr(ob))
    return oa, ob


def erode_mask(mask, k) -> np.ndarray:
    mask = np.ascontiguousarray(mask, np.uint8)
    out = np.zeros_like(mask)
    lib().dvref_erode_mask(_ptr(mask), mask.shape[0], mask.shape[1], int(k), _ptr(out))
    return out


class RefFrontEnd:
    """FeatureTracker (+ InstsFeatManager in dynamic mode) of the reference, reference-compiled; `step(frame)` returns the
    same structure as oracle.cv_front_end.FrontEnd.step.  One instance per process at a time (InstFeat::global_id_count
    is a static of the reference)."""

    def __init__(self, params, cam0: dict, cam1: Optional[dict], mode: str, width: int, height: int):
        c = Config()
        c.max_cnt, c.max_dynamic_cnt = params.max_cnt, params.max_dynamic_cnt
        c.min_dist, c.min_dynamic_dist = params.min_dist, params.min_dynamic_dist
        c.flow_back = int(params.flow_back)
        c.use_mask_morphology, c.mask_morphology_size = int(params.use_mask_morphology), params.mask_morphology_size
        c.width, c.height, c.stereo, c.dynamic = width, height, int(params.is_stereo), int(mode in ("dynamic", "naive"))
        c.cam0 = _cam8(cam0)
        c.cam1 = _cam8(cam1 if cam1 is not None else cam0)
        self.mode, self.w, self.h_img = mode, width, height
        self.cap = 4 * max(params.max_cnt, 1)
        self.h = lib().dvref_front_end_new(C.byref(c))
        self.seq = 0

    def close(self):
        if getattr(self, "h", None):
            lib().dvref_front_end_free(self.h)
            self.h = None

    def __del__(self):
        if _lib is not None:
            self.close()

    @staticmethod
    def _points(obs, n) -> Dict[int, List[Tuple[int, np.ndarray]]]:
        out: Dict[int, List[Tuple[int, np.ndarray]]] = {}
        for i in range(n):
            out.setdefault(int(obs[i].id), []).append((int(obs[i].cam), np.array(obs[i].v[:], np.float64)))
        return out

    def step(self, frame, disp: Optional[np.ndarray] = None) -> dict:
        g0 = np.ascontiguousarray(frame.gray0)
        g1 = None if frame.gray1 is None else np.ascontiguousarray(frame.gray1)
        obs = (Obs * self.cap)()
        if self.mode == "raw":
            n = lib().dvref_track_image(self.h, _ptr(g0), _ptr(g1), float(frame.time0), obs, self.cap)
            if n < 0:
                raise RuntimeError(lib().dvref_last_error().decode())
            assert n <= self.cap
            return {"features": self._points(obs, n), "instances": {}}
        if self.mode == "naive":
            im = None if frame.inv_merge_mask is None else np.ascontiguousarray(frame.inv_merge_mask)
            n = lib().dvref_track_image_naive(self.h, _ptr(g0), _ptr(g1), _ptr(im), int(bool(frame.exist_inst)), float(frame.time0),
                                              self.seq, obs, self.cap)
            self.seq += 1
            if n < 0:
                raise RuntimeError(lib().dvref_last_error().decode())
            assert n <= self.cap
            return {"features": self._points(obs, n), "instances": {}}
        boxes = (Box * max(len(frame.boxes), 1))()
        keep = []
        for i, b in enumerate(frame.boxes):
            m = np.ascontiguousarray(b["mask"], np.uint8)
            keep.append(m)
            x, y, w, h = b["rect"]
            boxes[i] = Box(int(b["track_id"]), x, y, w, h, m.ctypes.data, m.strides[0])
        mm = None if frame.merge_mask is None else np.ascontiguousarray(frame.merge_mask)
        im = None if frame.inv_merge_mask is None else np.ascontiguousarray(frame.inv_merge_mask)
        dd = None if disp is None else np.ascontiguousarray(disp, np.float32)
        icap = 64 * 1024
        iobs = (InstObs * icap)()
        n, ni = C.c_int(0), C.c_int(0)
        rc = lib().dvref_track_dynamic(self.h, _ptr(g0), _ptr(g1), _ptr(mm), _ptr(im), _ptr(dd), boxes, len(frame.boxes),
                                       float(frame.time0), self.seq, obs, self.cap, C.byref(n), iobs, icap, C.byref(ni))
        self.seq += 1
        if rc < 0:
            raise RuntimeError(lib().dvref_last_error().decode())
        assert n.value <= self.cap and ni.value <= icap
        insts: Dict[int, dict] = {}
        for i in range(ni.value):
            o = iobs[i]
            insts.setdefault(int(o.inst_id), {"features": {}})["features"][int(o.id)] = dict(
                point=np.array(o.point[:]), vel=np.array(o.vel[:]), point_right=np.array(o.point_right[:]),
                vel_right=np.array(o.vel_right[:]), is_stereo=bool(o.is_stereo), disp=float(o.disp))
        return {"features": self._points(obs, n.value), "instances": dict(sorted(insts.items()))}

    def reject_with_f(self, cur_pts, prev_pts, col: int, row: int) -> np.ndarray:
        """InstsFeatManager::RejectWithF on (curr_points, last_points) with this front end's cam0; F_threshold = 1.0"""
        a = np.ascontiguousarray(cur_pts, np.float32).reshape(-1, 2)
        b = np.ascontiguousarray(prev_pts, np.float32).reshape(-1, 2)
        st = np.zeros(max(len(a), 1), np.uint8)
        n = lib().dvref_reject_with_f(self.h, _ptr(a), _ptr(b), len(a), int(col), int(row), _ptr(st))
        if n < 0:
            raise RuntimeError(lib().dvref_last_error().decode())
        return st[:n].copy()

    def instance_table(self) -> np.ndarray:
        rows = np.zeros((256, 4), np.int32)
        n = lib().dvref_instance_table(self.h, _ptr(rows), 256)
        return rows[:n]


def detect_extra_points(mask, disp, box_xy, fx, fy, cx, cy, baseline) -> np.ndarray:
    """InstFeat::DetectExtraPoints (front_end/instance_feature.cpp:413-461), reference-compiled -> (n, 3) float64"""
    mask = np.ascontiguousarray(mask, np.uint8)
    disp = np.ascontiguousarray(disp, np.float32)
    cam = np.array([fx, fy, cx, cy, baseline], np.float32)
    cap = mask.size
    out = np.zeros((cap, 3), np.float64)
    n = lib().dvref_detect_extra_points(_ptr(mask), mask.shape[0], mask.shape[1], _ptr(disp), disp.shape[0], disp.shape[1],
                                        int(box_xy[0]), int(box_xy[1]), _ptr(cam), _ptr(out), cap)
    if n < 0:
        raise RuntimeError(lib().dvref_last_error().decode())
    return out[:n].copy()


def set_mask_and_roi(mask_stack, rects, gray0):
    """SemanticImage::SetMaskAndRoi (basic/semantic_image.cpp:20-63), reference-compiled over a stand-in integer tensor:
    mask_stack int8 [N, H, W] (the segmentation output, any non-zero value = object), rects [(x, y, w, h)] ->
    (merge_mask, inv_merge_mask, [roi_mask], [roi_gray])"""
    ms = np.ascontiguousarray(mask_stack, np.int8)
    n, h, w = ms.shape
    g = np.ascontiguousarray(gray0, np.uint8)
    rc = np.ascontiguousarray(np.array(rects, np.int32).reshape(-1, 4))
    merge, inv = np.zeros((h, w), np.uint8), np.zeros((h, w), np.uint8)
    total = int(sum(int(r[2]) * int(r[3]) for r in rc))
    rm, rg = np.zeros(max(total, 1), np.uint8), np.zeros(max(total, 1), np.uint8)
    k = lib().dvref_set_mask_and_roi(_ptr(ms), n, h, w, _ptr(g), _ptr(rc), _ptr(merge), _ptr(inv), _ptr(rm), _ptr(rg))
    if k < 0:
        raise RuntimeError(lib().dvref_last_error().decode())
    masks, grays, off = [], [], 0
    for x, y, bw, bh in rc:
        masks.append(rm[off:off + bw * bh].reshape(bh, bw).copy())
        grays.append(rg[off:off + bw * bh].reshape(bh, bw).copy())
        off += bw * bh
    return merge, inv, masks, grays


def serialize_point_features(path: str, points: Dict[int, List[Tuple[int, np.ndarray]]]) -> None:
    """SerializePointFeature (utils/io/feature_serialization.cpp:26-38), reference-compiled (fmt::format is a stand-in that prints
    the shortest round-trip form)"""
    flat = [(fid, cam, v) for fid in sorted(points) for cam, v in points[fid]]
    obs = (Obs * max(len(flat), 1))()
    for i, (fid, cam, v) in enumerate(flat):
        obs[i].id, obs[i].cam = fid, cam
        for k in range(7):
            obs[i].v[k] = float(v[k])
    if lib().dvref_serialize_points(path.encode(), obs, len(flat)) < 0:
        raise RuntimeError(lib().dvref_last_error().decode())


def deserialize_point_features(path: str, cap: int = 65536) -> Dict[int, List[Tuple[int, np.ndarray]]]:
    """DeserializePointFeature (utils/io/feature_serialization.cpp:45-70), reference-compiled"""
    obs = (Obs * cap)()
    n = lib().dvref_deserialize_points(path.encode(), obs, cap)
    if n < 0:
        raise RuntimeError(lib().dvref_last_error().decode())
    assert n <= cap
    return RefFrontEnd._points(obs, n)
