"""host time of the pipelined calls per step (enqueue vs wait) for a small workload (C1: 64 x 752x480 mono), pinned host images"""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from dynamic_vins_b200 import BatchTracker, make_config, synth, _lib as L
import ctypes as C

name = sys.argv[1] if len(sys.argv) > 1 else "c1_euroc_mono"
c = synth.CONFIGS[name]
W, H, S = c["width"], c["height"], 64
stereo = bool(c["stereo"])
st = synth.SynthStream(W, H, seed=3, stereo=stereo)
fr = [st.frame(k) for k in range(4)]
Ls = [torch.from_numpy(np.stack([f.gray0] * S)).pin_memory().numpy() for f in fr]
Rs = [torch.from_numpy(np.stack([f.gray1] * S)).pin_memory().numpy() for f in fr] if stereo else None
trk = BatchTracker(make_config(W, H, c["max_cnt"], c["min_dist"], c["cam0"], c["cam1"], stereo=stereo, n_streams=S))
order = synth.pingpong_positions(4, 400)
lib = L.lib()
tm = np.full(S, 0.0)
ta = tw = 0.0
N = 300
for i in range(N + 20):
    k = order[i]
    tm[:] = 0.05 * (i + 1)
    t0 = time.perf_counter()
    lib.dvfe_track_image_async(trk._h, Ls[k].ctypes.data, Rs[k].ctypes.data if stereo else None, H * W, W, tm.ctypes.data)
    t1 = time.perf_counter()
    if i > 0:
        lib.dvfe_wait(trk._h)
    t2 = time.perf_counter()
    if i == 20:
        T0 = t0
    if i >= 20:
        ta += t1 - t0; tw += t2 - t1
lib.dvfe_wait(trk._h)
tot = time.perf_counter() - T0
print(f"{name}: {tot / N * 1e3:.3f} ms/step ({S * N / tot:.0f} frames/s): enqueue {ta / N * 1e3:.3f} ms, wait {tw / N * 1e3:.3f} ms")
