"""Deterministic synthetic camera streams for parity tests and the benchmark.

The reference ships no data and no fixtures for the front-end, so
every input on the test/bench path is generated here from a seed:

* a textured canvas (Gaussian-blurred uniform noise + flat random rectangles, so
  there are real corners and real flat regions),
* frame k = the canvas seen through a smooth per-frame similarity (sub-pixel
  translation, small rotation and scale), bilinear-sampled,
* right image = the same view shifted by a row-dependent disparity (8..40 px at
  720p, scaled with the image width) so left->right LK converges inside a
  4-image pyramid,
* dynamic mode: K moving textured objects with per-object binary masks, integer
  boxes and track ids 1..K (the shape `SemanticImage::SetMaskAndRoi`
  produces, /root/reference/dynamic_vins/src/basic/semantic_image.cpp:20-63).

Only numpy is used: the generator must give identical bytes on the CPU box and on
the GPU box, and must not depend on anything under `oracle/`.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np


def _gauss_blur(img: np.ndarray, sigma: float) -> np.ndarray:
    r = int(3 * sigma + 0.5)
    x = np.arange(-r, r + 1, dtype=np.float64)
    k = np.exp(-0.5 * (x / sigma) ** 2)
    k /= k.sum()
    pad = np.pad(img, ((r, r), (0, 0)), mode="reflect")
    out = np.zeros_like(img)
    for i, w in enumerate(k):
        out += w * pad[i:i + img.shape[0], :]
    pad = np.pad(out, ((0, 0), (r, r)), mode="reflect")
    out2 = np.zeros_like(img)
    for i, w in enumerate(k):
        out2 += w * pad[:, i:i + img.shape[1]]
    return out2


def make_canvas(rng: np.random.Generator, h: int, w: int, n_rect: int = 60,
                sigma: float = 2.0) -> np.ndarray:
    """Float64 canvas in [0,255]."""
    noise = rng.random((h, w))
    tex = _gauss_blur(noise, sigma)
    tex -= tex.min()
    tex *= 255.0 / max(tex.max(), 1e-12)
    for _ in range(n_rect):
        rw = int(rng.integers(w // 40 + 4, w // 8 + 8))
        rh = int(rng.integers(h // 40 + 4, h // 8 + 8))
        x0 = int(rng.integers(0, max(1, w - rw)))
        y0 = int(rng.integers(0, max(1, h - rh)))
        tex[y0:y0 + rh, x0:x0 + rw] = float(rng.integers(10, 246))
    return tex


def _bilinear(canvas: np.ndarray, xs: np.ndarray, ys: np.ndarray) -> np.ndarray:
    h, w = canvas.shape
    xs = np.clip(xs, 0.0, w - 1.001)
    ys = np.clip(ys, 0.0, h - 1.001)
    x0 = np.floor(xs).astype(np.int64)
    y0 = np.floor(ys).astype(np.int64)
    ax = xs - x0
    ay = ys - y0
    v = ((1 - ax) * (1 - ay) * canvas[y0, x0] + ax * (1 - ay) * canvas[y0, x0 + 1]
         + (1 - ax) * ay * canvas[y0 + 1, x0] + ax * ay * canvas[y0 + 1, x0 + 1])
    return v


@dataclass
class SynthObject:
    track_id: int
    w: int
    h: int
    texture: np.ndarray          # (h, w) float64
    shape_mask: np.ndarray       # (h, w) bool (rectangle or ellipse)
    x0: float
    y0: float
    vx: float
    vy: float
    disparity: float


@dataclass
class SynthFrame:
    """One time step of one stream; mirrors the fields of the reference's
    `SemanticImage` that the front-end reads
    (/root/reference/dynamic_vins/src/basic/semantic_image.h:30-65)."""
    seq: int
    time0: float
    gray0: np.ndarray
    gray1: Optional[np.ndarray]
    # dynamic mode only
    exist_inst: bool = False
    merge_mask: Optional[np.ndarray] = None       # 255 = object
    inv_merge_mask: Optional[np.ndarray] = None   # 0 = object
    boxes: List[dict] = field(default_factory=list)  # {track_id, rect=(x,y,w,h), mask (h,w) u8 0/255}


class SynthStream:
    """A seeded camera stream. `frame(k)` is pure (depends on seed and k only)."""

    def __init__(self, width: int, height: int, seed: int, stereo: bool = True,
                 n_objects: int = 0, margin: int = 96):
        self.w, self.h, self.seed, self.stereo = width, height, seed, stereo
        rng = np.random.default_rng(seed)
        self.margin = margin
        self.canvas = make_canvas(rng, height + 2 * margin, width + 2 * margin + 64)
        ang = rng.uniform(0, 2 * np.pi)
        speed = rng.uniform(1.0, 4.0)
        self.vx, self.vy = speed * np.cos(ang), speed * np.sin(ang)
        self.omega = np.deg2rad(rng.uniform(-0.2, 0.2))
        self.dscale = rng.uniform(-0.002, 0.002)
        # disparity (px) as a function of the row: near the bottom = closer = larger
        s = width / 1280.0
        self.d_top, self.d_bot = 8.0 * s + 2.0, 40.0 * s
        self.objects: List[SynthObject] = []
        for i in range(n_objects):
            ow = int(rng.integers(width // 12, width // 6))
            oh = int(rng.integers(height // 8, height // 4))
            tex = make_canvas(rng, oh, ow, n_rect=6, sigma=1.5)
            if i % 2 == 0:
                sm = np.ones((oh, ow), dtype=bool)
            else:
                yy, xx = np.mgrid[0:oh, 0:ow]
                sm = (((xx - (ow - 1) / 2) / (ow / 2)) ** 2 + ((yy - (oh - 1) / 2) / (oh / 2)) ** 2) <= 1.0
            # lay the objects out on a jittered grid so they do not start overlapped
            gx, gy = i % 4, i // 4
            x0 = (gx + 0.15 + 0.5 * rng.random()) * width / 4.4
            y0 = (gy + 0.15 + 0.5 * rng.random()) * height / 2.6
            a = rng.uniform(0, 2 * np.pi)
            sp = rng.uniform(1.0, 5.0)
            self.objects.append(SynthObject(i + 1, ow, oh, tex, sm, x0, y0,
                                            sp * np.cos(a), sp * np.sin(a),
                                            float(rng.uniform(20.0, 48.0) * s)))
        yy, xx = np.mgrid[0:height, 0:width]
        self._xx = xx.astype(np.float64)
        self._yy = yy.astype(np.float64)

    # -- geometry ---------------------------------------------------------
    def _view(self, k: float, right: bool) -> np.ndarray:
        cx, cy = self.w / 2.0, self.h / 2.0
        th = self.omega * k
        sc = 1.0 + self.dscale * k
        c, s = np.cos(th) * sc, np.sin(th) * sc
        dx, dy = self._xx - cx, self._yy - cy
        xs = c * dx - s * dy + cx + self.margin + 32 + self.vx * k
        ys = s * dx + c * dy + cy + self.margin + self.vy * k
        if right:
            disp = self.d_top + (self.d_bot - self.d_top) * (self._yy / max(self.h - 1, 1))
            xs = xs + disp
        return _bilinear(self.canvas, xs, ys)

    def _paint_objects(self, img: np.ndarray, k: float, right: bool, boxes: Optional[list],
                       merge: Optional[np.ndarray]) -> None:
        for ob in self.objects:
            # sub-pixel object motion; objects bounce off the borders (triangle wave)
            def tri(p, lo, hi):
                span = hi - lo
                if span <= 0:
                    return lo
                q = (p - lo) % (2 * span)
                return lo + (q if q <= span else 2 * span - q)
            px = tri(ob.x0 + ob.vx * k, 2.0, self.w - ob.w - 3.0)
            py = tri(ob.y0 + ob.vy * k, 2.0, self.h - ob.h - 3.0)
            if right:
                px = px - ob.disparity
            ix, iy = int(np.floor(px)), int(np.floor(py))
            fx, fy = px - ix, py - iy
            # integer box that contains the sub-pixel-shifted object
            bx0, by0 = ix, iy
            bw, bh = ob.w + 1, ob.h + 1
            x_lo, y_lo = max(bx0, 0), max(by0, 0)
            x_hi, y_hi = min(bx0 + bw, self.w), min(by0 + bh, self.h)
            if x_hi - x_lo < 4 or y_hi - y_lo < 4:
                continue
            yy, xx = np.mgrid[y_lo:y_hi, x_lo:x_hi]
            u = xx - px
            v = yy - py
            inside = (u >= 0) & (u <= ob.w - 1) & (v >= 0) & (v <= ob.h - 1)
            uu = np.clip(u, 0, ob.w - 1.001)
            vv = np.clip(v, 0, ob.h - 1.001)
            tex = _bilinear(ob.texture, uu, vv)
            sm = ob.shape_mask[np.clip(np.rint(vv).astype(int), 0, ob.h - 1),
                               np.clip(np.rint(uu).astype(int), 0, ob.w - 1)] & inside
            region = img[y_lo:y_hi, x_lo:x_hi]
            region[sm] = tex[sm]
            if boxes is not None and not right:
                m = np.zeros((y_hi - y_lo, x_hi - x_lo), dtype=np.uint8)
                m[sm] = 255
                ys_any = np.flatnonzero(m.any(axis=1))
                xs_any = np.flatnonzero(m.any(axis=0))
                if len(ys_any) == 0:
                    continue
                ty0, ty1 = ys_any[0], ys_any[-1] + 1
                tx0, tx1 = xs_any[0], xs_any[-1] + 1
                m = np.ascontiguousarray(m[ty0:ty1, tx0:tx1])
                rect = (int(x_lo + tx0), int(y_lo + ty0), int(tx1 - tx0), int(ty1 - ty0))
                boxes.append({"track_id": ob.track_id, "rect": rect, "mask": m})
                merge[rect[1]:rect[1] + rect[3], rect[0]:rect[0] + rect[2]] |= m

    def frame(self, k: int, time0: Optional[float] = None, pos: Optional[float] = None) -> SynthFrame:
        """Frame at sequence index k. `pos` overrides the motion parameter (used by the
        ping-pong playback of the benchmark), `time0` the time stamp (default 0.05*k,
        the dataloader convention, /root/reference/dynamic_vins/src/utils/io/dataloader.cpp:81)."""
        p = float(k if pos is None else pos)
        left = self._view(p, False)
        boxes, merge = ([], np.zeros((self.h, self.w), np.uint8)) if self.objects else (None, None)
        if self.objects:
            self._paint_objects(left, p, False, boxes, merge)
        g0 = np.clip(np.rint(left), 0, 255).astype(np.uint8)
        g1 = None
        if self.stereo:
            right = self._view(p, True)
            if self.objects:
                self._paint_objects(right, p, True, None, None)
            g1 = np.clip(np.rint(right), 0, 255).astype(np.uint8)
        fr = SynthFrame(seq=k, time0=0.05 * k if time0 is None else time0, gray0=g0, gray1=g1)
        if self.objects:
            fr.exist_inst = len(boxes) > 0
            fr.merge_mask = merge
            fr.inv_merge_mask = (255 - merge).astype(np.uint8)
            fr.boxes = boxes
        return fr


def label_image(frame: SynthFrame):
    """The frame's instance masks as ONE u8 label image: bit b set = the pixel belongs to boxes[b] (the form
    `dvfe_track_dynamic_ex(..., DVFE_DYN_LABELS)` takes; what an instance-segmentation network's mask stack collapses to).
    Returns (labels, boxes with `label_bit` set).  At most 8 boxes."""
    assert len(frame.boxes) <= 8
    lab = np.zeros(frame.gray0.shape, np.uint8)
    boxes = []
    for b, box in enumerate(frame.boxes):
        x, y, w, h = box["rect"]
        lab[y:y + h, x:x + w] |= ((box["mask"] != 0).astype(np.uint8) << b)
        boxes.append(dict(box, label_bit=b))
    return lab, boxes


def pingpong_positions(n_unique: int, n_steps: int) -> List[int]:
    """0,1,..,n-1,n-2,..,1,0,1,.. : temporally coherent motion from a finite set of frames."""
    if n_unique <= 1:
        return [0] * n_steps
    period = 2 * (n_unique - 1)
    out = []
    for i in range(n_steps):
        q = i % period
        out.append(q if q < n_unique else period - q)
    return out


# ---- named configurations (BASELINE.json `configs`) ----------

EUROC_CAM0 = dict(fx=458.654, fy=457.296, cx=367.215, cy=248.375,
                  k1=-2.8340811e-01, k2=7.395907e-02, p1=1.9359e-04, p2=1.76187114e-05)
EUROC_CAM1 = dict(fx=457.587, fy=456.134, cx=379.999, cy=255.238,
                  k1=-2.8368365e-01, k2=7.451284e-02, p1=-1.0473e-04, p2=-3.55590700e-05)
KITTI_CAM = dict(fx=721.5377, fy=721.5377, cx=609.5593, cy=172.854, k1=0.0, k2=0.0, p1=0.0, p2=0.0)
# cfgd/custom/zed_1280x720_vision_only/un_cam{0,1}_pinhole.yaml
ZED_UN_CAM0 = dict(fx=5.7817315673828125e+02, fy=6.6596881103515625e+02, cx=6.7666424560546875e+02,
                   cy=3.6173339843750000e+02, k1=0.0, k2=0.0, p1=0.0, p2=0.0)
ZED_UN_CAM1 = dict(fx=5.7601123046875000e+02, fy=6.6383435058593750e+02, cx=6.9023547363281250e+02,
                   cy=3.7148672485351562e+02, k1=0.0, k2=0.0, p1=0.0, p2=0.0)
# cfgd/custom/stereo_1920x1080/cam{0,1}_pinhole.yaml
HD_CAM0 = dict(fx=1493.115757464805, fy=1487.400629263316, cx=961.5215876339698, cy=571.4387116745389,
               k1=-0.1125156715251281, k2=0.4670225696460769, p1=0.003901433044648148, p2=0.002572620348408325)
HD_CAM1 = dict(fx=1485.889400339916, fy=1466.199860203776, cx=1018.364007519278, cy=550.7153227893042,
               k1=-0.001486495030972591, k2=-0.134507205122525, p1=0.001032068050248756, p2=0.01681776826086429)

CONFIGS = {
    # C1  EuRoC-shaped mono (cfgd/euroc/euroc.yaml:23-24,76-77; cam0_pinhole.yaml)
    "c1_euroc_mono": dict(width=752, height=480, stereo=False, max_cnt=150, min_dist=30,
                          cam0=EUROC_CAM0, cam1=EUROC_CAM1, n_objects=0, config_id=1),
    # C2  KITTI-shaped stereo (cfgd/kitti/calibration.yaml:18-19,50-51)
    "c2_kitti_stereo": dict(width=1242, height=375, stereo=True, max_cnt=200, min_dist=30,
                            cam0=KITTI_CAM, cam1=KITTI_CAM, n_objects=0, config_id=2),
    # C3  ZED 1280x720 stereo dynamic (cfgd/custom/zed_1280x720_vision_only/dynamic.yaml:29-30,70-81)
    "c3_zed_dynamic": dict(width=1280, height=720, stereo=True, max_cnt=400, min_dist=25,
                           max_dynamic_cnt=50, min_dynamic_dist=4, use_mask_morphology=1,
                           mask_morphology_size=20, cam0=ZED_UN_CAM0, cam1=ZED_UN_CAM1,
                           n_objects=8, config_id=3),
    # C4  1920x1080 stereo, 1000 pts (cfgd/custom/stereo_1920x1080/custom.yaml:28-29 + BASELINE override)
    "c4_hd_stereo": dict(width=1920, height=1080, stereo=True, max_cnt=1000, min_dist=20,
                         cam0=HD_CAM0, cam1=HD_CAM1, n_objects=0, config_id=4),
    # C5  64 independent 1280x720 raw-stereo streams (per-stream shape)
    "c5_zed_streams": dict(width=1280, height=720, stereo=True, max_cnt=400, min_dist=25,
                           cam0=ZED_UN_CAM0, cam1=ZED_UN_CAM1, n_objects=0, config_id=5),
}


def make_stream(config: str, stream_id: int = 0) -> SynthStream:
    c = CONFIGS[config]
    return SynthStream(c["width"], c["height"], seed=1000 * c["config_id"] + stream_id,
                       stereo=c["stereo"], n_objects=c["n_objects"])


def colorize(gray: np.ndarray) -> np.ndarray:
    """deterministic BGR rendition of a gray frame (distinct per-channel gains/offsets, integer arithmetic) for the
    colour-input tests: SemanticImage::color0/color1 are 8UC3 BGR"""
    g = gray.astype(np.int32)
    b = g
    gr = (g * 235 + 3072) >> 8
    r = np.minimum(255, (g * 271) >> 8)
    return np.stack([b, gr, r], axis=-1).astype(np.uint8)


def random_maps(width: int, height: int, seed: int, outside: float = 3.0):
    """random float sampling maps (some of them outside the image) in cv::remap's fixed-point form:
    map1 int16 (h, w, 2) = floor(x), floor(y) ; map2 uint16 (h, w) = (fy << 5) | fx in 1/32 px"""
    rng = np.random.default_rng(seed)
    ix = np.rint(rng.uniform(-outside, width - 1 + outside, (height, width)) * 32).astype(np.int64)
    iy = np.rint(rng.uniform(-outside, height - 1 + outside, (height, width)) * 32).astype(np.int64)
    map1 = np.stack([ix >> 5, iy >> 5], axis=-1).astype(np.int16)
    map2 = (((iy & 31) << 5) | (ix & 31)).astype(np.uint16)
    return map1, map2
