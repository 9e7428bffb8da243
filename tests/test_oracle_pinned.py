"""CPU tests: the oracle's plain-C restatement (oracle/spec.c) is pinned against cv2 (the third-party library the
reference calls) and against the committed golden fixtures; the cv2-backed front-end restatement reproduces
the golden tracker outputs.  These run without a GPU."""
import cv2
import numpy as np
import pytest

from conftest import crc, feature_map_arrays, load_golden
from dynamic_vins_b200 import synth
from oracle import cv_front_end as cvfe
from oracle import spec


@pytest.fixture(scope="module")
def kitti_pair():
    st = synth.make_stream("c2_kitti_stereo", 0)
    return st.frame(0), st.frame(1)


def test_synth_matches_golden_crc(kitti_pair):
    g = load_golden("stages_kitti.npz")
    f0, f1 = kitti_pair
    assert crc(f0.gray0) == int(g["crc_f0"]) and crc(f1.gray0) == int(g["crc_f1"]) and crc(f0.gray1) == int(g["crc_r0"])


@pytest.mark.parametrize("shape", [(375, 1242), (480, 752), (61, 77), (50, 60)])
def test_pyr_down_bit_exact(shape):
    rng = np.random.default_rng(shape[0])
    a = rng.integers(0, 256, shape, dtype=np.uint8)
    for _ in range(3):
        if min(a.shape) < 3:
            break
        b = cv2.pyrDown(a)
        assert np.array_equal(spec.pyr_down(a), b)
        a = b


@pytest.mark.parametrize("wh,expect", [((50, 60), 1), ((43, 44), 1), ((42, 100), 0), ((90, 90), 2), ((176, 200), 3),
                                        ((1280, 720), 3)])
def test_pyramid_truncation_rule(wh, expect):
    # cv::buildOpticalFlowPyramid keeps level l only while the next size stays > winSize
    assert spec.pyr_levels(wh[0], wh[1], 3) == expect
    img = np.zeros((wh[1], wh[0]), np.uint8)
    assert cv2.buildOpticalFlowPyramid(img, (21, 21), 3, withDerivatives=False)[0] == expect


def test_pyramid_golden(kitti_pair):
    g = load_golden("stages_kitti.npz")
    a = kitti_pair[0].gray0
    for l in range(4):
        assert crc(a) == int(g[f"pyr{l}_crc"])
        a = spec.pyr_down(a)


def test_scharr_bit_exact(kitti_pair):
    pyr = cv2.buildOpticalFlowPyramid(kitti_pair[0].gray0, (21, 21), 3, withDerivatives=True)[1]
    for l in range(4):
        img, der = np.ascontiguousarray(pyr[2 * l]), np.ascontiguousarray(pyr[2 * l + 1])
        assert np.array_equal(spec.scharr(img), der)


def test_lk_spec_vs_cv2(kitti_pair):
    """status bits identical; positions within 1e-3 px of cv2 (north_star tolerance: 0.02 px); the exact-integer
    accumulation mode (what the CUDA kernel does) is at least as close to cv2 as float accumulation."""
    g = load_golden("stages_kitti.npz")
    f0, f1 = kitti_pair
    p = g["lk_pts1"]
    for exact in (False, True):
        p2, st = spec.feature_track_by_lk(f0.gray0, f1.gray0, p, True, 3, exact_int=exact)
        assert np.array_equal(st, g["lk_status"])
        ok = st == 1
        assert np.abs(p2[ok] - g["lk_pts2"][ok]).max() < 1e-3
        r2, rst = spec.feature_track_by_lk(f0.gray0, f0.gray1, p, True, 3, exact_int=exact)
        assert np.array_equal(rst, g["lkr_status"])
        assert np.abs(r2[rst == 1] - g["lkr_pts2"][rst == 1]).max() < 1e-3


def test_lk_small_roi_and_border_points():
    """ragged inputs: pyramid truncation on small crops, points on / outside the border, flat regions"""
    rng = np.random.default_rng(5)
    st = synth.SynthStream(200, 120, seed=11, stereo=False)
    a, b = st.frame(0).gray0[:58, :70].copy(), st.frame(1).gray0[:58, :70].copy()
    pts = np.array([[0, 0], [69, 57], [35.5, 29.25], [-3.0, 10.0], [80.0, 20.0], [10.2, 50.9], [60.1, 3.3]], np.float32)
    p_cv, s_cv = cvfe.feature_track_by_lk(a, b, pts, True, 3)
    p_sp, s_sp = spec.feature_track_by_lk(a, b, pts, True, 3, exact_int=True)
    assert np.array_equal(s_cv, s_sp)
    if s_cv.any():
        assert np.abs(p_cv[s_cv == 1] - p_sp[s_sp == 1]).max() < 1e-3
    flat = np.full((100, 100), 77, np.uint8)
    _, s = spec.feature_track_by_lk(flat, flat, np.array([[50, 50]], np.float32), True, 3, exact_int=True)
    assert s[0] == 0     # minEig below threshold


@pytest.mark.parametrize("cfg", ["c1_euroc_mono", "c2_kitti_stereo"])
def test_min_eigen_val_close_to_cv2(cfg):
    g = synth.make_stream(cfg, 0).frame(0).gray0
    e_cv = cv2.cornerMinEigenVal(g, 3, ksize=3)
    e_sp = spec.min_eigen_val(g)
    # bit-identical in the SIMD body; OpenCV's scalar tail columns differ at the 1-ulp level
    assert (e_cv != e_sp).mean() < 0.005
    assert np.abs(e_cv - e_sp).max() <= 1e-7 * max(1.0, float(e_cv.max()))


def test_gftt_spec_vs_cv2_and_golden(kitti_pair):
    g = load_golden("stages_kitti.npz")
    img = kitti_pair[0].gray0
    assert np.array_equal(spec.good_features(img, None, 200, 0.01, 30), g["gftt_200_30"])
    assert np.array_equal(spec.good_features(img, None, 1000, 0.01, 10), g["gftt_1000_10"])
    mask = spec.disc_mask(np.full(img.shape, 255, np.uint8), g["gftt_200_30"][:120], 30)
    assert crc(mask) == int(g["mask_crc"])
    assert np.array_equal(spec.good_features(img, mask, 80, 0.01, 30), g["gftt_masked_80_30"])
    # NMS given the cv2 response map: identical survivors and order
    eig = cv2.cornerMinEigenVal(img, 3, ksize=3)
    sel, _ = spec.gftt_select(eig, mask, 80, 0.01, 30)
    assert np.array_equal(sel, g["gftt_masked_80_30"])


@pytest.mark.parametrize("r", [4, 5, 10, 20, 25, 30])
def test_disc_mask_vs_cv2(r):
    rng = np.random.default_rng(r)
    pts = np.stack([rng.uniform(-10, 210, 30), rng.uniform(-10, 160, 30)], 1).astype(np.float32)
    m = np.full((150, 200), 255, np.uint8)
    want = m.copy()
    cvfe.draw_discs(want, pts, r)
    assert np.array_equal(spec.disc_mask(m, pts, r), want)


def test_erode_vs_cv2_and_golden():
    g = load_golden("stages_kitti.npz")
    m = g["erode_in"]
    for k in (5, 10, 20):
        out = spec.erode_rect(m, k)
        assert np.array_equal(out, cvfe.erode_mask(m, k))
        assert crc(out) == int(g[f"erode{k}_crc"])


def test_lift_projective_vs_python_restatement():
    g = load_golden("stages_kitti.npz")
    assert np.array_equal(spec.lift(synth.EUROC_CAM0, g["lift_in"]), g["lift_out"])
    cam = cvfe.PinholeCamera(**synth.HD_CAM1)
    q = g["lift_in"] * 2.0
    assert np.array_equal(spec.lift(synth.HD_CAM1, q, off=(13.0, 7.0)), cam.undistort_points(q, off=(13.0, 7.0)))
    assert np.array_equal(spec.lift(synth.KITTI_CAM, q), cvfe.PinholeCamera(**synth.KITTI_CAM).undistort_points(q))


@pytest.mark.parametrize("name,mode", [("c1_euroc_mono", "raw"), ("c2_kitti_stereo", "raw")])
def test_cv_front_end_reproduces_golden(name, mode):
    g = load_golden(f"tracker_{name}_{mode}.npz")
    c = synth.CONFIGS[name]
    st = synth.make_stream(name, 0)
    fe = cvfe.FrontEnd(cvfe.FrontEndParams(max_cnt=c["max_cnt"], min_dist=c["min_dist"], is_stereo=c["stereo"]),
                       c["cam0"], c["cam1"], mode)
    for k in range(int(g["n_frames"])):
        fr = st.frame(k)
        assert crc(fr.gray0) == int(g[f"f{k}_crc0"])
        ids, cams, v = feature_map_arrays(fe.step(fr)["features"])
        assert np.array_equal(ids, g[f"f{k}_ids"]) and np.array_equal(cams, g[f"f{k}_cams"])
        assert np.array_equal(v, g[f"f{k}_v"])


def test_feature_track_by_lk_throws_on_empty():
    # front_end/feature_utils.cpp:39-41
    img = np.zeros((64, 64), np.uint8)
    with pytest.raises(RuntimeError):
        cvfe.feature_track_by_lk(img, img, np.zeros((0, 2), np.float32))


def test_remap_and_gray_spec_vs_cv2_and_golden():
    """§8f N1/N2 frame preparation: the C restatement of cv::remap (CV_16SC2 maps, INTER_LINEAR, constant border) and of
    cvtColor(BGR2GRAY) equals cv2 bit for bit, incl. taps outside the image, and reproduces the committed golden"""
    g = load_golden("prep.npz")
    assert np.array_equal(spec.remap(g["bgr"], g["map1"], g["map2"]), g["remap_bgr"])
    assert np.array_equal(spec.remap(g["bgr"][..., 1].copy(), g["map1"], g["map2"]), g["remap_gray"])
    assert np.array_equal(spec.bgr_to_gray(g["bgr"]), g["gray"])
    assert np.array_equal(spec.bgr_to_gray(spec.remap(g["bgr"], g["map1"], g["map2"])), g["remap_then_gray"])
    rng = np.random.default_rng(5)
    for (h, w) in [(33, 47), (1, 9), (120, 64)]:
        src = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        m1, m2 = synth.random_maps(w, h, 100 + h, outside=4.0)
        assert np.array_equal(spec.remap(src, m1, m2), cv2.remap(src, m1, m2, cv2.INTER_LINEAR))
        # float maps through cv2.convertMaps give the same fixed-point maps as the generator's rule
        fx = (m1[..., 0].astype(np.float32) + (m2 & 31).astype(np.float32) / 32)
        fy = (m1[..., 1].astype(np.float32) + (m2 >> 5).astype(np.float32) / 32)
        c1, c2 = cv2.convertMaps(fx, fy, cv2.CV_16SC2)
        assert np.array_equal(c1, m1) and np.array_equal(c2, m2)


def test_undistort_maps_golden():
    from oracle import image_process as ip
    g = load_golden("prep.npz")
    c = synth.CONFIGS["c1_euroc_mono"]
    m1, m2, cam = ip.undistort_maps(c["cam0"], c["width"], c["height"])
    assert crc(m1) == int(g["euroc_map1_crc"]) and crc(m2) == int(g["euroc_map2_crc"])
    assert np.allclose([cam["fx"], cam["fy"], cam["cx"], cam["cy"]], g["euroc_new_k"], rtol=0, atol=1e-9)
    gray = synth.make_stream("c1_euroc_mono", 0).frame(0).gray0
    und = ip.run(synth.colorize(gray), None, (m1, m2))[0]
    assert crc(und) == int(g["euroc_undist_gray_crc"])
    assert np.array_equal(und, spec.bgr_to_gray(spec.remap(synth.colorize(gray), m1, m2)))


def test_cuda_detector_semantics_restatement(kitti_pair):
    """oracle.cv_front_end.good_features_cuda_semantics (cv::cuda::GoodFeaturesToTrackDetector restated from its published source,
    UNPINNED: no CUDA build of OpenCV here) against an independent numpy evaluation of the same rules, and against
    cv2.goodFeaturesToTrack where the two detectors must agree (no mask: masked maximum == whole-image maximum)"""
    g = kitti_pair[0].gray0
    h, w = g.shape
    no_mask = cvfe.good_features_cuda_semantics(g, 120, 20, None)
    assert np.array_equal(no_mask, cv2.goodFeaturesToTrack(g, 120, 0.01, 20).reshape(-1, 2))
    mask = np.zeros((h, w), np.uint8)
    mask[h // 2:, : w // 2] = 255
    g = g.copy()                                                           # weak texture under the mask: the strongest corners,
    g[mask != 0] = (128 + (g[mask != 0].astype(np.int32) - 128) // 6).astype(np.uint8)   # hence the threshold, lie outside it
    got = cvfe.good_features_cuda_semantics(g, 120, 20, mask)
    eig = cv2.cornerMinEigenVal(g, 3, ksize=3)
    thr = np.float32(float(eig.max()) * 0.01)                              # whole-image maximum
    dil = cv2.dilate(eig, np.ones((3, 3), np.uint8))
    ys, xs = np.nonzero((eig > thr) & (eig == dil) & (mask != 0))
    keep = (ys >= 1) & (ys < h - 1) & (xs >= 1) & (xs < w - 1)
    ys, xs = ys[keep], xs[keep]
    order = np.lexsort((-(ys * w + xs), -eig[ys, xs].astype(np.float64)))   # value descending, then address descending
    acc = []
    for i in order:
        x, y = int(xs[i]), int(ys[i])
        if all((x - ax) ** 2 + (y - ay) ** 2 >= 400 for ax, ay in acc):
            acc.append((x, y))
            if len(acc) == 120:
                break
    assert np.array_equal(got, np.array(acc, np.float32))
    # with no cap on the count the higher threshold of the cv::cuda rule keeps fewer corners than the CPU detector
    n_cuda = len(cvfe.good_features_cuda_semantics(g, 100000, 3, mask))
    n_cpu = len(cv2.goodFeaturesToTrack(g, 100000, 0.01, 3, mask=mask))
    assert n_cuda < n_cpu, "the masked case must exercise the different threshold"
