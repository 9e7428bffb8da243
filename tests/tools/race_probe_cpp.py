"""Debug aid: the C++ BatchFeatureTracker driver (fast host, two frames truly in flight) run repeatedly on the same frames;
reports which (frame, stream) outputs differ from the synchronous python-driven tracker.  Env DVFE_GRAPHS / DVFE_TMA bisect."""
import os, subprocess, sys, tempfile, pathlib
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from dynamic_vins_b200 import BatchTracker, make_config, obs_to_map, synth
from test_gpu_cpp_dynamic import write_config, build_driver

name, B, T = "c2_kitti_stereo", 3, 5
tmp = pathlib.Path(tempfile.mkdtemp())
cfg = write_config(tmp, name, "raw")
exe = build_driver(tmp)
c = dict(synth.CONFIGS[name]); c.pop("n_objects", None); c.pop("config_id", None)
streams = [synth.make_stream(name, 20 + s) for s in range(B)]
ref = BatchTracker(make_config(n_streams=B, **c))
want = []
with open(tmp / "frames.bin", "wb") as f:
    for k in range(T):
        frs = [s.frame(k) for s in streams]
        tm = np.array([fr.time0 + 0.002 * i for i, fr in enumerate(frs)], np.float64)
        L = np.stack([fr.gray0 for fr in frs]); R = np.stack([fr.gray1 for fr in frs])
        f.write(tm.tobytes()); f.write(L.tobytes()); f.write(R.tobytes())
        ref.track_image(L, R, tm)
        want.append([obs_to_map(ref.features(s)) for s in range(B)])
ref.close()
for env in ({}, {"DVFE_REUSE": "0"}, {"DVFE_GRAPHS": "0"}, {"CUDA_LAUNCH_BLOCKING": "1"}):
    bad = []
    for rep in range(10):
        subprocess.check_call([exe, "batch", cfg, str(tmp / "frames.bin"), str(T), str(B), str(tmp / "out"), "2"], env=dict(os.environ, **env))
        for k in range(T):
            for s in range(B):
                lines = open(tmp / f"out_s{s}_{k}_point.txt").read().strip().split("\n")
                w = want[k][s]
                nl = sum(1 for ln in lines if ln.split()[0] == "1"); nw = sum(1 for o in w.values() if len(o) == 2)
                if len(lines) != len(w) or nl != nw:
                    bad.append((rep, k, s, len(lines), len(w), nl, nw))
    print(env, "mismatches:", bad[:8], len(bad))
