"""where does a dynamic-mode step (64 streams x 8 instances) spend its time: semantic background call, python marshalling of
the box list, the C call of dvfe_insts_track_batch"""
import sys, time, ctypes as C
import numpy as np, torch
sys.path.insert(0, ".")
from dynamic_vins_b200 import BatchTracker, make_config, synth, _lib as L

name = "c3_zed_dynamic"
c = dict(synth.CONFIGS[name])
S = int(sys.argv[1]) if len(sys.argv) > 1 else 64
G = int(sys.argv[2]) if len(sys.argv) > 2 else 1
T = 6
base = [synth.make_stream(name, s) for s in range(8)]
frames = [[st.frame(k) for st in base] for k in range(T)]
def stack(k, attr):
    return torch.from_numpy(np.stack([getattr(frames[k][s % 8], attr) for s in range(S)])).pin_memory().numpy()
Ls = [stack(k, "gray0") for k in range(T)]; Rs = [stack(k, "gray1") for k in range(T)]; Ms = [stack(k, "inv_merge_mask") for k in range(T)]
boxes = [[frames[k][s % 8].boxes for s in range(S)] for k in range(T)]
cfg = make_config(c["width"], c["height"], c["max_cnt"], c["min_dist"], c["cam0"], c["cam1"], stereo=True, n_streams=S,
                  max_dynamic_cnt=c["max_dynamic_cnt"], min_dynamic_dist=c["min_dynamic_dist"],
                  use_mask_morphology=c["use_mask_morphology"], mask_morphology_size=c["mask_morphology_size"],
                  max_instances=8, n_groups=G)
trk = BatchTracker(cfg)
# pre-marshalled box lists (what a C++ caller holds anyway)
pre = []
for k in range(T):
    flat = [b for bs in boxes[k] for b in bs]
    arr = (L.InstIn * len(flat))()
    keep = []
    for i, b in enumerate(flat):
        m = np.ascontiguousarray(b["mask"], np.uint8); keep.append(m)
        x, y, w, h = b["rect"]
        arr[i].track_id, arr[i].x, arr[i].y, arr[i].w, arr[i].h = int(b["track_id"]), x, y, w, h
        arr[i].mask, arr[i].mask_pitch = m.ctypes.data, m.strides[0]
    counts = np.asarray([len(bs) for bs in boxes[k]], np.int32)
    pre.append((arr, counts, keep))
order = synth.pingpong_positions(T, 200)
ts, tm, tc = 0.0, 0.0, 0.0
N = 40
for i in range(N + 5):
    k = order[i]; t = 0.05 * (i + 1)
    a = time.perf_counter()
    trk.track_semantic_image(Ls[k], Rs[k], Ms[k], [1] * S, t)
    b = time.perf_counter()
    arr, counts, _ = pre[k]
    tt = np.full(S, t)
    L.check(L.lib().dvfe_insts_track_batch(trk._h, arr, L.ptr(counts), L.ptr(tt)))
    d = time.perf_counter()
    if i >= 5:
        ts += b - a; tc += d - b
print(f"S={S} G={G}: semantic {ts/N*1e3:.2f} ms  insts_track_batch(C) {tc/N*1e3:.2f} ms  -> {S*N/(ts+tc):.0f} frames/s")
t0 = time.perf_counter()
for i in range(10):
    trk.insts_track_batch(boxes[order[i]], 10 + 0.05 * i)
print(f"python-marshalled insts_track_batch {(time.perf_counter()-t0)/10*1e3:.2f} ms")

# pipelined
N = 60
for i in range(N + 5):
    k = order[i]; t = 20 + 0.05 * (i + 1)
    if i == 5:
        trk.wait(); t0 = time.perf_counter()
    trk.track_dynamic_async(Ls[k], Rs[k], Ms[k], [1] * S, pre[k], t)
    if i > 0:
        trk.wait()
trk.wait()
dt = time.perf_counter() - t0
print(f"pipelined dvfe_track_dynamic_async: {dt/N*1e3:.2f} ms/step -> {S*N/dt:.0f} frames/s")
