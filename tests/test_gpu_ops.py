"""GPU parity tests of the seam-level operators, through the C ABI (include/dvfe.h dvfe_op_*).
Bar: bit-exact for integer/byte/index work and against the oracle's exact-integer arithmetic; positions within
0.02 px of cv2 (north_star tolerance); NMS survivors identical given equal response maps."""
import cv2
import numpy as np
import pytest

from conftest import crc, load_golden
import dynamic_vins_b200 as dv
from dynamic_vins_b200 import ops, synth
from oracle import cv_front_end as cvfe
from oracle import spec

pytestmark = pytest.mark.gpu
POS_TOL = 0.02     # px, BASELINE.json north_star


@pytest.fixture(scope="module")
def kitti_pair():
    st = synth.make_stream("c2_kitti_stereo", 0)
    return st.frame(0), st.frame(1)


@pytest.mark.parametrize("shape", [(375, 1242), (480, 752), (61, 77), (50, 60), (33, 45), (720, 1280)])
def test_pyramid_bit_exact(shape):
    rng = np.random.default_rng(shape[1])
    a = rng.integers(0, 256, shape, dtype=np.uint8)
    levels = ops.build_pyramid(a, 3)
    assert len(levels) == spec.pyr_levels(shape[1], shape[0], 3) + 1
    ref = a
    for lv in levels:
        assert np.array_equal(lv, ref)
        ref = spec.pyr_down(ref)


@pytest.mark.parametrize("shape", [(375, 1242), (480, 752), (720, 1280), (61, 77), (50, 60), (97, 203), (1080, 1920), (188, 621), (200, 528), (51, 96), (54, 176), (412, 2064)])
def test_pyramid_border_is_reflect_101(shape):
    """every level WITH the border LK reads (cv::buildOpticalFlowPyramid stores a winSize = 21 px REFLECT_101 border): levels >= 1
    write it from the down-sampling kernel itself when a single reflection reaches every border pixel, small levels through the
    border kernel; both must equal copyMakeBorder of the exact interior"""
    import cv2
    rng = np.random.default_rng(shape[0] * 7 + shape[1])
    a = rng.integers(0, 256, shape, dtype=np.uint8)
    ref = a
    for l in range(spec.pyr_levels(shape[1], shape[0], 3) + 1):
        got = ops.build_pyramid_bordered(a, l, 21)
        want = cv2.copyMakeBorder(ref, 21, 21, 21, 21, cv2.BORDER_REFLECT_101)
        assert got.shape == want.shape
        assert np.array_equal(got, want), f"level {l} of {shape}: {int((got != want).sum())} border/interior bytes differ"
        ref = spec.pyr_down(ref)


def test_pyramid_golden(kitti_pair):
    g = load_golden("stages_kitti.npz")
    for l, lv in enumerate(ops.build_pyramid(kitti_pair[0].gray0, 3)):
        assert crc(lv) == int(g[f"pyr{l}_crc"])


def test_lk_bit_exact_vs_exact_integer_oracle(kitti_pair):
    g = load_golden("stages_kitti.npz")
    f0, f1 = kitti_pair
    p = g["lk_pts1"]
    for img2 in (f1.gray0, f0.gray1):
        p2, st, rev = ops.feature_track_by_lk(f0.gray0, img2, p, True, 3, return_rev=True)
        q2, qst, qrev = spec.feature_track_by_lk(f0.gray0, img2, p, True, 3, exact_int=True, return_rev=True)
        assert np.array_equal(st, qst)
        assert np.array_equal(p2, q2)                 # every point, including failed ones
        assert np.array_equal(rev[st == 1], qrev[st == 1])


def test_lk_vs_cv2_golden(kitti_pair):
    g = load_golden("stages_kitti.npz")
    f0, f1 = kitti_pair
    p2, st = ops.feature_track_by_lk(f0.gray0, f1.gray0, g["lk_pts1"], True, 3)
    assert np.array_equal(st, g["lk_status"])                      # status bits: bit-exact
    assert np.abs(p2 - g["lk_pts2"])[st == 1].max() <= POS_TOL      # positions: within 0.02 px
    r2, rst = ops.feature_track_by_lk(f0.gray0, f0.gray1, g["lk_pts1"], True, 3)
    assert np.array_equal(rst, g["lkr_status"])
    assert np.abs(r2 - g["lkr_pts2"])[rst == 1].max() <= POS_TOL


def test_lk_without_flow_back_and_other_levels(kitti_pair):
    f0, f1 = kitti_pair
    p = load_golden("stages_kitti.npz")["lk_pts1"][:64]
    for fb, lvl in [(False, 3), (True, 1), (True, 0), (False, 2)]:
        p2, st = ops.feature_track_by_lk(f0.gray0, f1.gray0, p, fb, lvl)
        q2, qst = spec.feature_track_by_lk(f0.gray0, f1.gray0, p, fb, lvl, exact_int=True)
        assert np.array_equal(st, qst) and np.array_equal(p2[st == 1], q2[st == 1])


def test_lk_ragged_inputs():
    """small crops (pyramid truncation), points on/outside the border, a flat image, a region mask"""
    st = synth.SynthStream(200, 120, seed=11, stereo=False)
    A, B = st.frame(0).gray0, st.frame(1).gray0
    pts = np.array([[0, 0], [69, 57], [35.5, 29.25], [-3.0, 10.0], [80.0, 20.0], [10.2, 50.9], [60.1, 3.3],
                    [1.0, 1.0], [0.49, 30.0], [68.51, 30.0]], np.float32)
    for (h, w) in [(58, 70), (44, 100), (42, 100), (90, 90), (120, 200)]:
        a, b = np.ascontiguousarray(A[:h, :w]), np.ascontiguousarray(B[:h, :w])
        p2, s = ops.feature_track_by_lk(a, b, pts, True, 3)
        q2, qs = spec.feature_track_by_lk(a, b, pts, True, 3, exact_int=True)
        c2, cs = cvfe.feature_track_by_lk(a, b, pts, True, 3)
        assert np.array_equal(s, qs) and np.array_equal(p2[s == 1], q2[s == 1])
        assert np.array_equal(s, cs)
        if cs.any():
            assert np.abs(p2 - c2)[cs == 1].max() <= POS_TOL
    flat = np.full((100, 100), 77, np.uint8)
    _, s = ops.feature_track_by_lk(flat, flat, np.array([[50, 50], [20, 70]], np.float32), True, 3)
    assert not s.any()
    # mask test of InstFeat::TrackLeft: status cleared where mask[cvRound(pt2)] == 0
    grid = np.stack(np.meshgrid(np.arange(20, 190, 17), np.arange(15, 110, 13)), -1).reshape(-1, 2).astype(np.float32)
    p2, s_nomask = ops.feature_track_by_lk(A, B, grid, True, 3)
    mask = np.full(A.shape, 255, np.uint8)
    mask[:, 100:] = 0
    _, s_mask = ops.feature_track_by_lk(A, B, grid, True, 3, mask=mask)
    want = s_nomask.copy()
    for i, q in enumerate(p2):
        if want[i] and cvfe.mask_at(mask, q) == 0:
            want[i] = 0
    assert np.array_equal(s_mask, want) and s_mask.sum() < s_nomask.sum()


def test_lk_empty_input_is_an_error():
    img = np.zeros((64, 64), np.uint8)
    with pytest.raises(dv.DvfeError) as e:
        ops.feature_track_by_lk(img, img, np.zeros((0, 2), np.float32))
    assert e.value.code == -1      # the reference throws std::runtime_error (feature_utils.cpp:39-41)


@pytest.mark.parametrize("cfg", ["c1_euroc_mono", "c2_kitti_stereo", "c5_zed_streams"])
def test_response_map(cfg):
    g = synth.make_stream(cfg, 0).frame(0).gray0
    e = ops.min_eigen_val(g)
    assert np.array_equal(e, spec.min_eigen_val(g))          # same arithmetic as the C restatement: bit-exact
    e_cv = cv2.cornerMinEigenVal(g, 3, ksize=3)
    assert (e != e_cv).mean() < 0.005                         # cv2's scalar tail columns differ by ~1 ulp
    assert np.abs(e - e_cv).max() <= 1e-7 * max(1.0, float(e_cv.max()))


@pytest.mark.parametrize("shape", [(37, 41), (16, 130), (100, 33)])
def test_response_map_odd_shapes(shape):
    rng = np.random.default_rng(shape[0])
    a = rng.integers(0, 256, shape, dtype=np.uint8)
    assert np.array_equal(ops.min_eigen_val(a), spec.min_eigen_val(a))


def test_good_features_vs_golden(kitti_pair):
    g = load_golden("stages_kitti.npz")
    img = kitti_pair[0].gray0
    assert np.array_equal(ops.good_features(img, 200, 0.01, 30), g["gftt_200_30"])
    assert np.array_equal(ops.good_features(img, 1000, 0.01, 10), g["gftt_1000_10"])
    mask = ops.disc_mask(np.full(img.shape, 255, np.uint8), g["gftt_200_30"][:120], 30)
    assert crc(mask) == int(g["mask_crc"])
    assert np.array_equal(ops.good_features(img, 80, 0.01, 30, mask=mask), g["gftt_masked_80_30"])


@pytest.mark.parametrize("K,md", [(150, 30), (400, 25), (1000, 10), (50, 4), (2000, 3), (7, 1)])
def test_nms_bit_exact_given_equal_response(kitti_pair, K, md):
    """NMS survivors, their order and count given the oracle's (cv2) response map"""
    img = kitti_pair[0].gray0
    eig = cv2.cornerMinEigenVal(img, 3, ksize=3)
    rng = np.random.default_rng(K)
    pts = np.stack([rng.uniform(0, img.shape[1], 60), rng.uniform(0, img.shape[0], 60)], 1).astype(np.float32)
    mask = spec.disc_mask(np.full(img.shape, 255, np.uint8), pts, max(md, 3))
    for m in (None, mask):
        got, ncand = ops.good_features(img, K, 0.01, md, mask=m, eig=eig, return_n_candidates=True)
        want, ncand_ref = spec.gftt_select(eig, m, K, 0.01, md)
        assert ncand == ncand_ref
        assert np.array_equal(got, want)
        cvp = cv2.goodFeaturesToTrack(img, K, 0.01, md, mask=m)
        assert np.array_equal(got, cvp.reshape(-1, 2))


def test_good_features_cuda_detector_semantics(kitti_pair):
    """DetectShiTomasiCornersGpu's rule (quality threshold from the whole-image maximum): the op against the restatement, on a
    given response map (bit-exact) and through the fused response kernel"""
    import cv2
    from oracle import cv_front_end as cvfe
    g = kitti_pair[0].gray0
    h, w = g.shape
    mask = np.zeros((h, w), np.uint8)
    mask[h // 2:, : w // 2] = 255
    mask[: h // 3, w // 2:] = 255
    g = g.copy()                                          # weak texture under the mask: the whole-image maximum lies outside it
    g[mask != 0] = (128 + (g[mask != 0].astype(np.int32) - 128) // 6).astype(np.uint8)
    eig = cv2.cornerMinEigenVal(g, 3, ksize=3)
    for K, md in ((150, 30), (400, 10), (7, 3)):
        want, _ = spec.gftt_select(eig, mask, K, 0.01, md, unmasked_max=True)
        got = ops.good_features(None, K, 0.01, md, mask=mask, eig=eig, cuda_semantics=True)
        assert np.array_equal(got, want)
        fused = ops.good_features(g, K, 0.01, md, mask=mask, cuda_semantics=True)
        assert np.array_equal(fused, cvfe.good_features_cuda_semantics(g, K, md, mask))
        cpu_rule, _ = spec.gftt_select(eig, mask, K, 0.01, md)
        assert np.array_equal(ops.good_features(None, K, 0.01, md, mask=mask, eig=eig), cpu_rule)
    assert len(ops.good_features(g, 2000, 0.01, 3, mask=mask, cuda_semantics=True)) < len(ops.good_features(g, 2000, 0.01, 3, mask=mask))


def test_good_features_degenerate():
    flat = np.full((64, 80), 9, np.uint8)
    assert len(ops.good_features(flat, 10, 0.01, 5)) == 0
    img = synth.make_stream("c1_euroc_mono", 0).frame(0).gray0
    none = np.zeros(img.shape, np.uint8)
    assert len(ops.good_features(img, 10, 0.01, 5, mask=none)) == 0


@pytest.mark.parametrize("r", [0, 1, 4, 5, 10, 25, 30])
def test_disc_mask_bit_exact(r):
    rng = np.random.default_rng(r)
    pts = np.stack([rng.uniform(-10, 210, 30), rng.uniform(-10, 160, 30)], 1).astype(np.float32)
    pts[0] = (0.5, 1.5)      # cvRound is round-half-to-even
    pts[1] = (2.5, 3.5)
    m = np.full((150, 200), 255, np.uint8)
    want = m.copy()
    cvfe.draw_discs(want, pts, r)
    assert np.array_equal(ops.disc_mask(m, pts, r), want)


def test_erode_bit_exact():
    g = load_golden("stages_kitti.npz")
    m = g["erode_in"]
    for k in (1, 2, 3, 5, 10, 20):
        out = ops.erode_rect(m, k)
        assert np.array_equal(out, cvfe.erode_mask(m, k))
    for k in (5, 10, 20):
        assert crc(ops.erode_rect(m, k)) == int(g[f"erode{k}_crc"])
    fr = synth.make_stream("c3_zed_dynamic", 0).frame(0)
    assert np.array_equal(ops.erode_rect(fr.inv_merge_mask, 20), cvfe.erode_mask(fr.inv_merge_mask, 20))
    # the kernels take a word-wise AND path on binary masks and fall back to the byte-wise minimum otherwise: arbitrary u8
    # data, widths that are not a multiple of 4, binary masks with a few grey pixels, all-255 and all-0 masks
    rng = np.random.default_rng(12)
    for (h, w) in ((37, 53), (64, 96), (5, 7), (90, 121)):
        grey = rng.integers(0, 256, (h, w), dtype=np.uint8)
        binary = (rng.random((h, w)) > 0.2).astype(np.uint8) * 255
        mixed = binary.copy()
        mixed[rng.integers(0, h, 6), rng.integers(0, w, 6)] = 97
        for m2 in (grey, binary, mixed, np.full((h, w), 255, np.uint8), np.zeros((h, w), np.uint8)):
            for k in (1, 2, 3, 4, 5, 7, 20):
                assert np.array_equal(ops.erode_rect(m2, k), cvfe.erode_mask(m2, k)), (h, w, k)


def test_lift_projective_bit_exact():
    g = load_golden("stages_kitti.npz")
    assert np.array_equal(ops.lift_projective(synth.EUROC_CAM0, g["lift_in"]), g["lift_out"])
    q = g["lift_in"] * 2.0
    for cam in (synth.HD_CAM0, synth.HD_CAM1, synth.KITTI_CAM, synth.ZED_UN_CAM1):
        assert np.array_equal(ops.lift_projective(cam, q, off=(13.0, 7.0)),
                              cvfe.PinholeCamera(**cam).undistort_points(q, off=(13.0, 7.0)))


# ---- full-size properties (BASELINE.json config 4: 1920x1080, 1000 points) ---------------------------
def test_full_size_properties():
    st = synth.make_stream("c4_hd_stereo", 0)
    f0, f1 = st.frame(0), st.frame(1)
    pts = ops.good_features(f0.gray0, 1000, 0.01, 20)
    assert len(pts) == 1000
    # min-distance property and integer coordinates inside the border
    d = pts[:, None, :] - pts[None, :, :]
    d2 = (d ** 2).sum(-1) + np.eye(len(pts)) * 1e9
    assert d2.min() >= 20 * 20
    assert np.array_equal(pts, np.rint(pts)) and pts.min() >= 1 and (pts[:, 0] < 1919).all() and (pts[:, 1] < 1079).all()
    # acceptance order = descending response
    e = ops.min_eigen_val(f0.gray0)
    lam = e[pts[:, 1].astype(int), pts[:, 0].astype(int)]
    assert (np.diff(lam) <= 0).all()
    # identity tracking: the same image gives zero flow and status 1
    p2, s = ops.feature_track_by_lk(f0.gray0, f0.gray0, pts, True, 3)
    assert s.all() and np.abs(p2 - pts).max() < 1e-3
    # frame-to-frame and stereo tracking agree with cv2 within tolerance on the full-size frame
    p2, s = ops.feature_track_by_lk(f0.gray0, f1.gray0, pts, True, 3)
    c2, cs = cvfe.feature_track_by_lk(f0.gray0, f1.gray0, pts, True, 3)
    assert np.array_equal(s, cs) and np.abs(p2 - c2)[cs == 1].max() <= POS_TOL and cs.mean() > 0.9


def test_frame_prep_ops():
    """BGR->gray and instance-mask merge, bit-exact vs cv2 / numpy"""
    rng = np.random.default_rng(3)
    for shape in [(123, 457), (720, 1280), (5, 7)]:
        bgr = rng.integers(0, 256, shape + (3,), dtype=np.uint8)
        assert np.array_equal(ops.bgr_to_gray(bgr), cv2.cvtColor(bgr, cv2.COLOR_BGR2GRAY))
    fr = synth.make_stream("c3_zed_dynamic", 0).frame(2)
    full = np.zeros((len(fr.boxes), 720, 1280), np.uint8)
    for i, b in enumerate(fr.boxes):
        x, y, w, h = b["rect"]
        full[i, y:y + h, x:x + w] = b["mask"] // 255         # 0/1 masks like the SOLOv2 tensor
    merge, inv = ops.merge_masks(full)
    assert np.array_equal(merge, fr.merge_mask) and np.array_equal(inv, fr.inv_merge_mask)
    merge, inv = ops.merge_masks(np.zeros((0, 8, 9), np.uint8))
    assert not merge.any() and (inv == 255).all()


@pytest.mark.parametrize("shape", [(64, 96), (37, 131), (1, 5), (480, 752)])
def test_remap_bit_exact(shape):
    """dvfe_op_remap == cv2.remap(INTER_LINEAR) with fixed-point maps, taps outside the image included; 1 and 3 channels,
    with and without the fused gray conversion; NULL maps = identity"""
    h, w = shape
    rng = np.random.default_rng(h * 1000 + w)
    bgr = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    m1, m2 = synth.random_maps(w, h, 11 + h, outside=5.0)
    ref = cv2.remap(bgr, m1, m2, cv2.INTER_LINEAR)
    assert np.array_equal(ops.remap(bgr, m1, m2), ref)
    assert np.array_equal(ops.remap(bgr, m1, m2, to_gray=True), cv2.cvtColor(ref, cv2.COLOR_BGR2GRAY))
    g = bgr[..., 2].copy()
    assert np.array_equal(ops.remap(g, m1, m2), cv2.remap(g, m1, m2, cv2.INTER_LINEAR))
    assert np.array_equal(ops.remap(bgr, None, None, to_gray=True), cv2.cvtColor(bgr, cv2.COLOR_BGR2GRAY))
    assert np.array_equal(ops.remap(g, None, None), g)


def test_remap_golden_and_euroc_undistortion():
    from oracle import image_process as ip
    g = load_golden("prep.npz")
    assert np.array_equal(ops.remap(g["bgr"], g["map1"], g["map2"]), g["remap_bgr"])
    assert np.array_equal(ops.remap(g["bgr"], g["map1"], g["map2"], to_gray=True), g["remap_then_gray"])
    c = synth.CONFIGS["c1_euroc_mono"]
    m1, m2, _ = ip.undistort_maps(c["cam0"], c["width"], c["height"])
    gray = synth.make_stream("c1_euroc_mono", 0).frame(0).gray0
    assert crc(ops.remap(synth.colorize(gray), m1, m2, to_gray=True)) == int(g["euroc_undist_gray_crc"])


def test_static_instance_punch_out():
    """system/main.cpp:219-242: the ROI mask of a static instance is removed from merge_mask, inv_merge_mask follows"""
    from oracle import image_process as ip
    fr = synth.make_stream("c3_zed_dynamic", 1).frame(3)
    merge, inv = fr.merge_mask.copy(), fr.inv_merge_mask.copy()
    for b in fr.boxes[:3]:
        want_m, want_i = ip.punch_out_static(merge, b["mask"], b["rect"])
        merge, inv = ops.punch_out(merge, inv, b["mask"], b["rect"])
        assert np.array_equal(merge, want_m) and np.array_equal(inv, want_i)
    assert merge.sum() < fr.merge_mask.sum()
