"""time the ingest kernel (remap + BGR2GRAY into level 0) inside the tracker: device-resident BGR frames, 64 streams of 720p"""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from dynamic_vins_b200 import BatchTracker, make_config, synth
import cv2

c = synth.CONFIGS["c5_zed_streams"]
W, H, S = c["width"], c["height"], int(sys.argv[1]) if len(sys.argv) > 1 else 64
cam = dict(synth.EUROC_CAM0, cx=W / 2, cy=H / 2, fx=700.0, fy=700.0)
K = np.array([[cam["fx"], 0, cam["cx"]], [0, cam["fy"], cam["cy"]], [0, 0, 1]], np.float64)
D = np.array([cam["k1"], cam["k2"], cam["p1"], cam["p2"]], np.float64)
newK, _ = cv2.getOptimalNewCameraMatrix(K, D, (W, H), 0, (W, H))          # utils/camera_model.cpp:479-501
m1, m2 = cv2.initUndistortRectifyMap(K, D, None, newK, (W, H), cv2.CV_16SC2)
st = synth.SynthStream(W, H, seed=5000, stereo=True)
fr = [st.frame(k) for k in range(3)]
col = [(torch.from_numpy(synth.colorize(f.gray0)).cuda(), torch.from_numpy(synth.colorize(f.gray1)).cuda()) for f in fr]
frames = [(a[None].expand(S, -1, -1, -1).contiguous(), b[None].expand(S, -1, -1, -1).contiguous()) for a, b in col]
gray = [(a[..., 0].contiguous(), b[..., 0].contiguous()) for a, b in frames]
for mode in ("gray", "bgr", "bgr+undistort", "gray+undistort"):
    trk = BatchTracker(make_config(W, H, c["max_cnt"], c["min_dist"], c["cam0"], c["cam1"], stereo=True, n_streams=S, n_groups=4))
    ch = 3 if "bgr" in mode else 1
    trk.set_input(ch)
    if "undistort" in mode:
        trk.set_undistort_maps(0, m1, m2); trk.set_undistort_maps(1, m1, m2)
    src = frames if ch == 3 else gray
    n = 60
    for i in range(n + 10):
        if i == 10:
            torch.cuda.synchronize(); t0 = time.perf_counter()
        a, b = src[i % 3 if (i // 3) % 2 == 0 else 2 - i % 3]
        trk.track_image_device_async(a.data_ptr(), b.data_ptr(), H * W * ch, W * ch, np.full(S, 0.05 * (i + 1)))
        if i > 0:
            trk.wait()
    trk.wait()
    dt = time.perf_counter() - t0
    print(f"{mode:16s} {dt / n * 1e3:.3f} ms/step  {S * n / dt:.0f} frames/s  obs {len(trk.features(0))}")
    trk.close()
