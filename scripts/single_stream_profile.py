"""Development aid: per-stage device time of one frame step for B = 1 and a few batch sizes (c5 shape)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dynamic_vins_b200 import BatchTracker, make_config, synth

c = synth.CONFIGS["c5_zed_streams"]
st = synth.SynthStream(c["width"], c["height"], seed=77, stereo=True)
frames = [st.frame(k) for k in range(4)]
for B in (1, 4, 16):
    trk = BatchTracker(make_config(c["width"], c["height"], c["max_cnt"], c["min_dist"], c["cam0"], c["cam1"], stereo=True, n_streams=B))
    L = [np.stack([f.gray0] * B) for f in frames]
    R = [np.stack([f.gray1] * B) for f in frames]
    order = synth.pingpong_positions(4, 45)
    ts = []
    for i, k in enumerate(order):
        if i == 5:
            trk.profile(True)
        t0 = time.perf_counter()
        trk.track_image(L[k], R[k], 0.05 * (i + 1))
        ts.append((time.perf_counter() - t0) * 1e3)
    prof, n = trk.profile_read()
    print("B=%d wall %.3f ms/step; device stages (us):" % (B, np.median(ts[5:])), {k: round(v / n * 1e3, 1) for k, v in prof.items()},
          "sum %.3f ms" % (sum(prof.values()) / n))
    trk.close()
