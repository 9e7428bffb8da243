"""Debug aid: pipelined (track_image_async + wait, stream groups) vs synchronous records on the same frames, repeated; prints
per (repeat, frame, stream) whether the bytes differ and how (ids / cams).  Env DVFE_GRAPHS / DVFE_TMA select code paths."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from dynamic_vins_b200 import BatchTracker, make_config, synth

name, B, T = "c2_kitti_stereo", 3, 5
c = dict(synth.CONFIGS[name]); c.pop("n_objects", None); c.pop("config_id", None)
streams = [synth.make_stream(name, 20 + s) for s in range(B)]
frames = [[s.frame(k) for s in streams] for k in range(T)]
L = [np.stack([f.gray0 for f in fr]) for fr in frames]; R = [np.stack([f.gray1 for f in fr]) for fr in frames]
tm = [np.array([fr.time0 + 0.002 * i for i, fr in enumerate(frs)]) for frs in frames]
ref = BatchTracker(make_config(n_streams=B, **c))
want = []
for k in range(T):
    ref.track_image(L[k], R[k], tm[k]); want.append([ref.features(s).copy() for s in range(B)])
ref.close()
groups = int(os.environ.get("GROUPS", "2"))
bad = 0
for rep in range(int(os.environ.get("REPS", "12"))):
    t = BatchTracker(make_config(n_streams=B, n_groups=groups, **c))
    got = []
    t.track_image_async(L[0], R[0], tm[0])
    for k in range(1, T):
        t.track_image_async(L[k], R[k], tm[k]); t.wait(); got.append([t.features(s).copy() for s in range(B)])
    t.wait(); got.append([t.features(s).copy() for s in range(B)])
    t.close()
    for k in range(T):
        for s in range(B):
            a, b = got[k][s], want[k][s]
            if a.tobytes() != b.tobytes():
                bad += 1
                print(f"rep {rep} frame {k} stream {s}: n {len(a)} vs {len(b)}; left {int((a['cam']==0).sum())} vs {int((b['cam']==0).sum())}; "
                      f"right {int((a['cam']==1).sum())} vs {int((b['cam']==1).sum())}")
print("graphs", os.environ.get("DVFE_GRAPHS", "1"), "tma", os.environ.get("DVFE_TMA", "1"), "groups", groups, "-> mismatching (rep, frame, stream):", bad)
