"""Ad-hoc GPU diagnostics (development aid, not part of the test suite)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import cv2
from dynamic_vins_b200 import ops, synth, make_config, BatchTracker, obs_to_map
from oracle import spec
from oracle import cv_front_end as cvfe

def main():
    s = synth.make_stream("c2_kitti_stereo", 0)
    f0, f1 = s.frame(0), s.frame(1)
    g = f0.gray0
    # pyramid
    pyr = ops.build_pyramid(g, 3)
    a = g
    for l, p in enumerate(pyr):
        print("pyr", l, p.shape, np.array_equal(p, a))
        a = spec.pyr_down(a)
    # erode / disc / lift
    m = np.zeros((100, 120), np.uint8); m[20:70, 30:100] = 255; m[0:10, 0:15] = 255
    for k in (5, 10, 20):
        print("erode", k, np.array_equal(ops.erode_rect(m, k), spec.erode_rect(m, k)))
    rng = np.random.default_rng(0)
    pts = np.stack([rng.uniform(-5, 125, 40), rng.uniform(-5, 105, 40)], 1).astype(np.float32)
    mm = np.full((100, 120), 255, np.uint8)
    print("disc", np.array_equal(ops.disc_mask(mm, pts, 7), spec.disc_mask(mm, pts, 7)))
    cam = synth.EUROC_CAM0
    print("lift", np.array_equal(ops.lift_projective(cam, pts * 5), spec.lift(cam, pts * 5)))
    # response
    e_gpu = ops.min_eigen_val(g); e_sp = spec.min_eigen_val(g); e_cv = cv2.cornerMinEigenVal(g, 3, ksize=3)
    print("eig vs spec: mismatch", (e_gpu != e_sp).mean(), np.abs(e_gpu - e_sp).max(), " vs cv2:", (e_gpu != e_cv).mean(), np.abs(e_gpu - e_cv).max())
    # gftt
    for (K, md) in [(200, 30), (1000, 10), (50, 4)]:
        c_gpu, nc = ops.good_features(g, K, 0.01, md, return_n_candidates=True)
        c_cv = cv2.goodFeaturesToTrack(g, K, 0.01, md).reshape(-1, 2)
        c_gpu2 = ops.good_features(g, K, 0.01, md, eig=e_cv)
        print("gftt", K, md, "ncand", nc, "n", len(c_gpu), len(c_cv), "identical", np.array_equal(c_gpu, c_cv), "given cv eig identical", np.array_equal(c_gpu2, c_cv))
    # LK
    p = cv2.goodFeaturesToTrack(g, 200, 0.01, 30).reshape(-1, 2)
    t = time.time(); p2, st, rev = ops.feature_track_by_lk(f0.gray0, f1.gray0, p, True, 3, return_rev=True); print("lk time", time.time() - t)
    q2, qst, qrev = spec.feature_track_by_lk(f0.gray0, f1.gray0, p, True, 3, exact_int=True, return_rev=True)
    c2, cst = cvfe.feature_track_by_lk(f0.gray0, f1.gray0, p, True, 3)
    print("lk vs spec(exact): status eq", np.array_equal(st, qst), "pts bitexact", (p2 == q2).all(1).mean(), "max", np.abs(p2 - q2).max())
    print("lk vs cv2: status eq", np.array_equal(st, cst), "max", np.abs(p2 - c2)[cst == 1].max(), "n ok", cst.sum())
    p2, st = ops.feature_track_by_lk(f0.gray0, f0.gray1, p, True, 3)
    q2, qst = spec.feature_track_by_lk(f0.gray0, f0.gray1, p, True, 3, exact_int=True)
    print("stereo lk vs spec: status eq", np.array_equal(st, qst), (p2 == q2).all(1).mean(), np.abs(p2 - q2).max(), st.sum())

    # full tracker, free running
    for name in ("c2_kitti_stereo", "c1_euroc_mono"):
        c = synth.CONFIGS[name]
        cfg = make_config(c["width"], c["height"], c["max_cnt"], c["min_dist"], c["cam0"], c["cam1"], stereo=c["stereo"])
        trk = BatchTracker(cfg)
        P = cvfe.FrontEndParams(max_cnt=c["max_cnt"], min_dist=c["min_dist"], is_stereo=c["stereo"])
        ref = cvfe.FrontEnd(P, c["cam0"], c["cam1"], "raw")
        st_ = synth.make_stream(name, 0)
        for k in range(6):
            fr = st_.frame(k)
            t = time.time(); trk.track_image(fr.gray0, fr.gray1, fr.time0); dt = time.time() - t
            got = obs_to_map(trk.features(0))
            want = ref.step(fr)["features"]
            same_ids = sorted(got) == sorted(want)
            md = 0.0; nobs_eq = True
            for fid in set(got) & set(want):
                if len(got[fid]) != len(want[fid]): nobs_eq = False; continue
                for (c0, v0), (c1, v1) in zip(got[fid], want[fid]):
                    md = max(md, np.abs(v0 - v1).max())
            print(name, "frame", k, "n", len(got), len(want), "ids eq", same_ids, "nobs eq", nobs_eq, "max diff", md, "ms", dt * 1e3)
        trk.close()

if __name__ == "__main__":
    main()
