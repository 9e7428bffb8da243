"""Generates the golden fixtures under tests/golden/ from the ORACLE (oracle/cv_front_end.py over cv2 4.13.0)
on seeded synthetic inputs.  Run here (CPU container):  python tests/golden/make_golden.py
The reference ships no golden vectors for this path; these are outputs of the third-party
library the reference calls (OpenCV) at the reference's call sites, plus the restated glue.
Inputs are NOT stored: they are regenerated from the seed; their CRC32 is stored so that a generator drift is
reported as such and not as a parity failure."""
import os
import sys
import zlib

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from dynamic_vins_b200 import synth  # noqa: E402
from oracle import cv_front_end as cvfe  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def crc(a):
    return zlib.crc32(np.ascontiguousarray(a).tobytes())


def flat(points):
    ids, cams, vs = [], [], []
    for fid in sorted(points):
        for cam, v in points[fid]:
            ids.append(fid); cams.append(cam); vs.append(v)
    return np.asarray(ids, np.uint32), np.asarray(cams, np.int32), np.asarray(vs, np.float64).reshape(-1, 7)


def params(c):
    return cvfe.FrontEndParams(max_cnt=c["max_cnt"], min_dist=c["min_dist"],
                               max_dynamic_cnt=c.get("max_dynamic_cnt", 50), min_dynamic_dist=c.get("min_dynamic_dist", 5),
                               use_mask_morphology=c.get("use_mask_morphology", 0),
                               mask_morphology_size=c.get("mask_morphology_size", 5), is_stereo=c["stereo"])


def tracker_golden(name, n_frames, mode="raw"):
    c = synth.CONFIGS[name]
    st = synth.make_stream(name, 0)
    fe = cvfe.FrontEnd(params(c), c["cam0"], c["cam1"], mode)
    out = {"n_frames": np.int32(n_frames)}
    for k in range(n_frames):
        fr = st.frame(k)
        res = fe.step(fr)
        ids, cams, v = flat(res["features"])
        out[f"f{k}_ids"], out[f"f{k}_cams"], out[f"f{k}_v"] = ids, cams, v
        out[f"f{k}_crc0"] = np.uint32(crc(fr.gray0))
        out[f"f{k}_crc1"] = np.uint32(crc(fr.gray1) if fr.gray1 is not None else 0)
        if mode == "dynamic":
            rows = []
            for inst_id, inst in res["instances"].items():
                for fid, f in inst["features"].items():
                    rows.append([inst_id, fid, int(f["is_stereo"]), *f["point"], *f["vel"], *f["point_right"],
                                 *f["vel_right"], *f["uv"]])
            out[f"f{k}_inst"] = np.asarray(rows, np.float64).reshape(-1, 15)
    np.savez_compressed(os.path.join(OUT, f"tracker_{name}_{mode}.npz"), **out)
    print(name, mode, "frames", n_frames, "last n_obs", len(out[f"f{n_frames-1}_ids"]))


def stage_golden():
    """Stage-level known answers straight from cv2 (KITTI-shaped frame pair, seed 2000)."""
    st = synth.make_stream("c2_kitti_stereo", 0)
    f0, f1 = st.frame(0), st.frame(1)
    g = f0.gray0
    out = {"crc_f0": np.uint32(crc(f0.gray0)), "crc_f1": np.uint32(crc(f1.gray0)), "crc_r0": np.uint32(crc(f0.gray1))}
    pyr = cv2.buildOpticalFlowPyramid(g, (21, 21), 3, withDerivatives=False)[1]
    for l in range(4):
        out[f"pyr{l}_crc"] = np.uint32(crc(np.ascontiguousarray(pyr[l])))
    p = cv2.goodFeaturesToTrack(g, 200, 0.01, 30).reshape(-1, 2)
    out["gftt_200_30"] = p
    mask = np.full(g.shape, 255, np.uint8)
    cvfe.draw_discs(mask, p[:120], 30)
    out["mask_crc"] = np.uint32(crc(mask))
    out["gftt_masked_80_30"] = cv2.goodFeaturesToTrack(g, 80, 0.01, 30, mask=mask).reshape(-1, 2)
    out["gftt_1000_10"] = cv2.goodFeaturesToTrack(g, 1000, 0.01, 10).reshape(-1, 2)
    p2, st_ = cvfe.feature_track_by_lk(f0.gray0, f1.gray0, p, True, 3)
    out["lk_pts1"], out["lk_pts2"], out["lk_status"] = p, p2, st_
    r2, rst = cvfe.feature_track_by_lk(f0.gray0, f0.gray1, p, True, 3)
    out["lkr_pts2"], out["lkr_status"] = r2, rst
    m = np.zeros((120, 160), np.uint8); m[20:90, 30:130] = 255; m[0:14, 0:22] = 255; m[100:, 140:] = 255
    for k in (5, 10, 20):
        out[f"erode{k}_crc"] = np.uint32(crc(cvfe.erode_mask(m, k)))
    out["erode_in"] = m
    cam = cvfe.PinholeCamera(**synth.EUROC_CAM0)
    q = np.stack([np.linspace(3, 740, 40), np.linspace(470, 5, 40)], 1).astype(np.float32)
    out["lift_in"], out["lift_out"] = q, cam.undistort_points(q)
    np.savez_compressed(os.path.join(OUT, "stages_kitti.npz"), **out)
    print("stages: lk ok", int(st_.sum()), "stereo ok", int(rst.sum()))


def prep_golden():
    """cv2.remap (fixed-point maps, INTER_LINEAR) and cvtColor(BGR2GRAY) known answers on a small random image"""
    rng = np.random.default_rng(77)
    h, w = 64, 96
    bgr = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    map1, map2 = synth.random_maps(w, h, 78)
    out = {"bgr": bgr, "map1": map1, "map2": map2,
           "remap_bgr": cv2.remap(bgr, map1, map2, cv2.INTER_LINEAR),
           "remap_gray": cv2.remap(bgr[..., 1].copy(), map1, map2, cv2.INTER_LINEAR),
           "gray": cv2.cvtColor(bgr, cv2.COLOR_BGR2GRAY)}
    out["remap_then_gray"] = cv2.cvtColor(out["remap_bgr"], cv2.COLOR_BGR2GRAY)
    # the EuRoC undistortion maps as utils/camera_model.cpp:479-501 builds them: keep their checksums and the new intrinsics
    from oracle import image_process as ip
    c = synth.CONFIGS["c1_euroc_mono"]
    m1, m2, cam = ip.undistort_maps(c["cam0"], c["width"], c["height"])
    out["euroc_map1_crc"], out["euroc_map2_crc"] = np.uint32(crc(m1)), np.uint32(crc(m2))
    out["euroc_new_k"] = np.array([cam["fx"], cam["fy"], cam["cx"], cam["cy"]])
    g = synth.make_stream("c1_euroc_mono", 0).frame(0).gray0
    out["euroc_undist_gray_crc"] = np.uint32(crc(ip.run(synth.colorize(g), None, (m1, m2))[0]))
    np.savez_compressed(os.path.join(OUT, "prep.npz"), **out)
    print("prep golden written")


if __name__ == "__main__":
    stage_golden()
    prep_golden()
    tracker_golden("c1_euroc_mono", 6)
    tracker_golden("c2_kitti_stereo", 6)
    if "--dynamic" in sys.argv or True:
        tracker_golden("c3_zed_dynamic", 4, "dynamic")
