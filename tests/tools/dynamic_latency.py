"""Development aid: per-frame latency of the dynamic mode (TrackSemanticImage + InstsTrack + Output), B = 1,
BASELINE.json config 3 (1280x720 stereo, 8 instances), against the cv2 oracle on the same frames."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from dynamic_vins_b200 import BatchTracker, make_config, synth
from oracle import cv_front_end as cvfe

name = "c3_zed_dynamic"
c = dict(synth.CONFIGS[name]); c.pop("n_objects"); c.pop("config_id")
st = synth.make_stream(name, 0)
frames = [st.frame(k) for k in range(12)]
trk = BatchTracker(make_config(max_instances=8, **c))
P = cvfe.FrontEndParams(max_cnt=c["max_cnt"], min_dist=c["min_dist"], max_dynamic_cnt=50, min_dynamic_dist=4,
                        use_mask_morphology=1, mask_morphology_size=20, is_stereo=True)
fe = cvfe.FrontEnd(P, c["cam0"], c["cam1"], "dynamic")
tg, tc = [], []
for fr in frames:
    t0 = time.perf_counter()
    trk.track_semantic_image(fr.gray0, fr.gray1, fr.inv_merge_mask, fr.exist_inst, fr.time0)
    t1 = time.perf_counter()
    trk.insts_track(0, fr.boxes, fr.time0)
    out = trk.insts_output(0)
    t2 = time.perf_counter()
    fe.step(fr)
    t3 = time.perf_counter()
    tg.append(((t1 - t0) * 1e3, (t2 - t1) * 1e3)); tc.append((t3 - t2) * 1e3)
tg = np.array(tg[2:]); tc = np.array(tc[2:])
print("gpu: semantic %.3f ms, instances %.3f ms per frame; cv2 oracle %.1f ms per frame; speed-up %.0fx" %
      (np.median(tg[:, 0]), np.median(tg[:, 1]), np.median(tc), np.median(tc) / np.median(tg.sum(1))))
