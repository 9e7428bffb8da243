"""Experiment: G trackers of 64/G streams each on separate CUDA streams vs one tracker of 64 streams."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dynamic_vins_b200 import BatchTracker, make_config, synth
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import gpu_frames

c = synth.CONFIGS["c5_zed_streams"]
W, H, S, T = c["width"], c["height"], 64, 6
dev = torch.device("cuda", 0)
frames = torch.empty((T, 2, S, H, W), dtype=torch.uint8, device=dev)
for s in range(S):
    frames[:, :, s] = gpu_frames(synth.SynthStream(W, H, seed=5000 + s, stereo=True), T, dev)
torch.cuda.synchronize()
order = synth.pingpong_positions(T, 400)
P = W * H
for G in (1, 2, 4):
    n = S // G
    trks = []
    for g in range(G):
        t = BatchTracker(make_config(W, H, c["max_cnt"], c["min_dist"], c["cam0"], c["cam1"], stereo=True, n_streams=n))
        trks.append(t)
    def submit(i):
        f = frames[order[i]]
        for g, t in enumerate(trks):
            t.track_image_device_async(f[0, g * n].data_ptr(), f[1, g * n].data_ptr(), P, W, 0.05 * (i + 1))
    def wait():
        for t in trks:
            t.wait()
    for i in range(10):
        submit(i); wait()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    K = 150
    submit(10)
    for i in range(11, 10 + K):
        submit(i); wait()
    wait()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print("G=%d trackers x %d streams: %.0f frames/s (%.3f ms per 64-stream step)" % (G, n, S * K / dt, dt / K * 1e3))
    for t in trks:
        t.close()
