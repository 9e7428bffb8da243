"""ORACLE (test infrastructure) — the frame preparation upstream of the feature tracker, restated over cv2.

Follows (reference file:line)
  * utils/camera_model.cpp:479-501  InitOneCamera(), cfg::is_undistort_input: new_K = getOptimalNewCameraMatrix(K, D, size,
    alpha = 0, size); initUndistortRectifyMap(K, D, Mat(), new_K, size, CV_16SC2, map1, map2); afterwards the camera used
    by liftProjective is the distortion-free pinhole (new_K, D = 0) (:504-510).
  * image_process/image_process.cpp:105-126  ImageProcessor::Run(): color = remap(color, map1, map2, INTER_LINEAR) for both
    cameras when is_undistort_input, then SemanticImage::SetGrayImageGpu() = cvtColor(BGR2GRAY) (basic/semantic_image.cpp:76-93).
  * basic/semantic_image.cpp:103-117  SetBackgroundMask(): merge_mask is remapped with the left maps before bitwise_not.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.  PARITY UNPINNED for THIS module: it is a
sequence of OpenCV calls (pinned against cv2 4.13 in tests/test_oracle_pinned.py; the reference pins OpenCV 3.4.16) and its
reference sources (image_process.cpp, utils/camera_model.cpp: yaml / ROS parameter plumbing) are not among the translation units
compiled into oracle/_ref; the front end proper (oracle/cv_front_end.py) is pinned against reference-compiled code."""
from __future__ import annotations

from typing import Optional, Tuple

import cv2
import numpy as np


def undistort_maps(cam: dict, width: int, height: int) -> Tuple[np.ndarray, np.ndarray, dict]:
    """(map1 CV_16SC2, map2 CV_16UC1, camera after undistortion) as InitOneCamera() builds them (utils/camera_model.cpp:441-501).
    Followed to the letter, including two properties of the reference a clean-room version would not have:
      * K0 and D0 are cv::Mat_<float>: the intrinsics are narrowed to float before OpenCV sees them (:446-450);
      * D0 = (k1, k2, 0, p1, p2) (:450, commented "k1 k2 k3 p1 p2"), which OpenCV reads as (k1, k2, p1, p2, k3): the maps are
        built with p1 = 0, p2 = the camera's p1 and k3 = the camera's p2.
    The camera used afterwards is the distortion-free pinhole with new_K read back as float (SetCameraIntrinsicByK, :347-358)."""
    K = np.array([[cam["fx"], 0, cam["cx"]], [0, cam["fy"], cam["cy"]], [0, 0, 1]], np.float32)
    D = np.array([cam.get("k1", 0), cam.get("k2", 0), 0, cam.get("p1", 0), cam.get("p2", 0)], np.float32).reshape(5, 1)
    new_k, _ = cv2.getOptimalNewCameraMatrix(K, D, (width, height), 0, (width, height))
    map1, map2 = cv2.initUndistortRectifyMap(K, D, None, new_k, (width, height), cv2.CV_16SC2)
    nk = np.asarray(new_k, np.float32)              # cam.K0 = new_K0; fx0 = K0.at<float>(0,0) ...
    new_cam = dict(fx=float(nk[0, 0]), fy=float(nk[1, 1]), cx=float(nk[0, 2]), cy=float(nk[1, 2]),
                   k1=0.0, k2=0.0, p1=0.0, p2=0.0)
    return map1, map2, new_cam


def run(color0: np.ndarray, color1: Optional[np.ndarray], maps0=None, maps1=None):
    """ImageProcessor::Run up to the gray images: (gray0, gray1)"""
    if maps0 is not None:
        color0 = cv2.remap(color0, maps0[0], maps0[1], cv2.INTER_LINEAR)
    if color1 is not None and maps1 is not None:
        color1 = cv2.remap(color1, maps1[0], maps1[1], cv2.INTER_LINEAR)
    gray0 = cv2.cvtColor(color0, cv2.COLOR_BGR2GRAY) if color0.ndim == 3 else color0
    gray1 = None
    if color1 is not None:
        gray1 = cv2.cvtColor(color1, cv2.COLOR_BGR2GRAY) if color1.ndim == 3 else color1
    return gray0, gray1


def background_mask(merge_mask: np.ndarray, maps0=None) -> np.ndarray:
    """SetBackgroundMask(): inv_merge_mask = bitwise_not(remap(merge_mask))"""
    if maps0 is not None:
        merge_mask = cv2.remap(merge_mask, maps0[0], maps0[1], cv2.INTER_LINEAR)
    return cv2.bitwise_not(merge_mask)


def punch_out_static(merge_mask: np.ndarray, roi_mask: np.ndarray, rect) -> Tuple[np.ndarray, np.ndarray]:
    """FeatureTrack(): remove the mask of a static instance (system/main.cpp:219-242), then inv = bitwise_not(merge) (:238-240)"""
    m = merge_mask.copy()
    x, y, w, h = rect
    for row in range(roi_mask.shape[0]):
        for col in range(roi_mask.shape[1]):
            if roi_mask[row, col] >= 0.5:
                m[row + y, col + x] = 0
    return m, cv2.bitwise_not(m)
