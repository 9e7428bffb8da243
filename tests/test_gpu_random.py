"""Randomised GPU parity tests (seeded): many image sizes, point sets that sit on and outside the borders, random
masks and selection parameters — every result must equal the oracle's plain-C restatement bit for bit."""
import numpy as np
import pytest

from dynamic_vins_b200 import ops, synth
from oracle import spec

pytestmark = pytest.mark.gpu


def textured(rng, h, w):
    img = synth.make_canvas(rng, h, w, n_rect=max(2, (h * w) // 4000), sigma=1.5)
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)


@pytest.mark.parametrize("seed", range(6))
def test_random_pyramids(seed):
    rng = np.random.default_rng(100 + seed)
    for _ in range(6):
        h, w = int(rng.integers(23, 400)), int(rng.integers(23, 500))
        img = rng.integers(0, 256, (h, w), dtype=np.uint8)
        lv = int(rng.integers(0, 5))
        levels = ops.build_pyramid(img, lv)
        assert len(levels) == spec.pyr_levels(w, h, lv) + 1
        ref = img
        for l in levels:
            assert np.array_equal(l, ref), (h, w, lv)
            ref = spec.pyr_down(ref)


@pytest.mark.parametrize("seed", range(8))
def test_random_lk(seed):
    """forward(+backward) LK on random sizes; a third of the points lie within a window of the border or outside
    the image (template taps outside the image use the ZERO derivative border and the REFLECT image border)"""
    rng = np.random.default_rng(200 + seed)
    for _ in range(4):
        h, w = int(rng.integers(30, 260)), int(rng.integers(30, 340))
        a = textured(rng, h + 8, w + 8)
        sx, sy = rng.uniform(-3, 3, 2)
        # second image: a sub-pixel shifted view of the same texture
        yy, xx = np.mgrid[0:h, 0:w].astype(np.float64)
        b = np.clip(np.rint(synth._bilinear(a.astype(np.float64), xx + 4 + sx, yy + 4 + sy)), 0, 255).astype(np.uint8)
        a0 = np.ascontiguousarray(a[4:4 + h, 4:4 + w])
        n = 60
        pts = np.stack([rng.uniform(0, w, n), rng.uniform(0, h, n)], 1)
        edge = rng.integers(0, n, n // 3)
        pts[edge, 0] = rng.choice([rng.uniform(-12, 12), rng.uniform(w - 12, w + 12)], len(edge))
        pts[edge[::2], 1] = rng.choice([rng.uniform(-12, 12), rng.uniform(h - 12, h + 12)], len(edge[::2]))
        pts = pts.astype(np.float32)
        fb = bool(rng.integers(0, 2))
        lv = int(rng.integers(0, 5))
        p2, st, rev = ops.feature_track_by_lk(a0, b, pts, fb, lv, return_rev=True)
        q2, qst, qrev = spec.feature_track_by_lk(a0, b, pts, fb, lv, exact_int=True, return_rev=True)
        assert np.array_equal(st, qst), (h, w, fb, lv)
        # forward results are defined for every point (failed ones included)
        assert np.array_equal(p2, q2), (h, w, fb, lv)
        if fb:
            assert np.array_equal(rev[st == 1], qrev[st == 1])


@pytest.mark.parametrize("seed", range(6))
def test_random_good_features(seed):
    rng = np.random.default_rng(300 + seed)
    for _ in range(4):
        h, w = int(rng.integers(12, 300)), int(rng.integers(12, 400))
        img = textured(rng, h, w) if rng.integers(0, 2) else rng.integers(0, 256, (h, w), dtype=np.uint8)
        assert np.array_equal(ops.min_eigen_val(img), spec.min_eigen_val(img)), (h, w)
        mask = None
        if rng.integers(0, 2):
            mask = np.full((h, w), 255, np.uint8)
            m = int(rng.integers(1, 30))
            p = np.stack([rng.uniform(-5, w + 5, m), rng.uniform(-5, h + 5, m)], 1).astype(np.float32)
            r = int(rng.integers(1, 20))
            mask = ops.disc_mask(mask, p, r)
            assert np.array_equal(mask, spec.disc_mask(np.full((h, w), 255, np.uint8), p, r))
            if rng.integers(0, 2):
                mask[:, : w // 3] = 0
        K = int(rng.integers(1, 600))
        md = float(rng.integers(1, 30))
        got = ops.good_features(img, K, 0.01, md, mask=mask)
        want = spec.good_features(img, mask, K, 0.01, md)
        assert np.array_equal(got, want), (h, w, K, md, mask is not None)


@pytest.mark.parametrize("seed", range(3))
def test_random_erode(seed):
    rng = np.random.default_rng(400 + seed)
    for _ in range(6):
        h, w = int(rng.integers(3, 200)), int(rng.integers(3, 300))
        m = (rng.random((h, w)) > 0.15).astype(np.uint8) * 255
        k = int(rng.integers(1, 24))
        assert np.array_equal(ops.erode_rect(m, k), spec.erode_rect(m, k)), (h, w, k)
