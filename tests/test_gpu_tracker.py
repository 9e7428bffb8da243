"""GPU parity tests of the frame step (TrackImage / TrackSemanticImage) through the C ABI, against the committed
golden fixtures and the cv2 oracle run live on the same seeded frames (SURVEY.md Appendix D protocol)."""
import os

import cv2
import numpy as np
import pytest

from conftest import crc, feature_map_arrays, load_golden
import dynamic_vins_b200 as dv
from dynamic_vins_b200 import BatchTracker, FeatureTracker, make_config, obs_to_map, synth
from oracle import cv_front_end as cvfe

pytestmark = pytest.mark.gpu
POS_TOL = 0.02        # px  (north_star)


def cfg_of(name, n_streams=1, **kw):
    c = dict(synth.CONFIGS[name])
    c.update(kw)
    c.pop("n_objects", None); c.pop("config_id", None)
    return make_config(n_streams=n_streams, **c)


def params_of(name):
    c = synth.CONFIGS[name]
    return cvfe.FrontEndParams(max_cnt=c["max_cnt"], min_dist=c["min_dist"],
                               max_dynamic_cnt=c.get("max_dynamic_cnt", 50), min_dynamic_dist=c.get("min_dynamic_dist", 5),
                               use_mask_morphology=c.get("use_mask_morphology", 0),
                               mask_morphology_size=c.get("mask_morphology_size", 5), is_stereo=c["stereo"])


def compare_records(rec, ids, cams, v, cam0, dt_min=0.05):
    """integer outputs bit-exact; pixel positions within 0.02 px; normalised coordinates and velocities within the
    bound 0.02 px implies through the camera model and dt"""
    assert np.array_equal(rec["id"], ids), "feature-id assignment differs"
    assert np.array_equal(rec["cam"], cams), "camera lists differ (stereo status bits)"
    if len(ids) == 0:
        return 0.0
    gv = rec["v"]
    err_px = np.abs(gv[:, 3:5] - v[:, 3:5]).max()
    assert err_px <= POS_TOL
    un_tol = 1.5 * POS_TOL / min(cam0["fx"], cam0["fy"])
    assert np.abs(gv[:, 0:2] - v[:, 0:2]).max() <= un_tol
    assert np.array_equal(gv[:, 2], v[:, 2])
    assert np.abs(gv[:, 5:7] - v[:, 5:7]).max() <= 2 * un_tol / dt_min
    return err_px


@pytest.mark.parametrize("name", ["c1_euroc_mono", "c2_kitti_stereo"])
def test_free_running_vs_golden(name):
    g = load_golden(f"tracker_{name}_raw.npz")
    st = synth.make_stream(name, 0)
    trk = FeatureTracker(cfg_of(name))
    worst = 0.0
    for k in range(int(g["n_frames"])):
        fr = st.frame(k)
        assert crc(fr.gray0) == int(g[f"f{k}_crc0"]), "synthetic generator drifted; regenerate tests/golden"
        trk.batch.track_image(fr.gray0, fr.gray1, fr.time0)
        worst = max(worst, compare_records(trk.batch.features(0), g[f"f{k}_ids"], g[f"f{k}_cams"], g[f"f{k}_v"],
                                           synth.CONFIGS[name]["cam0"]))
    print(f"{name}: worst pixel error vs golden {worst:.2e} px")


def oracle_state(fe: cvfe.FrontEnd):
    bg = fe.tracker.bg
    n = len(bg.ids)
    rp = np.zeros((n, 2), np.float32)
    rv = np.zeros(n, np.uint8)
    pu = np.zeros((n, 2), np.float32)
    for i, fid in enumerate(bg.ids):
        pu[i] = bg.prev_id_pts[fid]
        if fid in bg.right_prev_id_pts:
            rp[i] = bg.right_prev_id_pts[fid]
            rv[i] = 1
    return dict(n=n, next_id=fe.idc.next, prev_time=fe.tracker.prev_time, ids=np.asarray(bg.ids, np.uint32),
                track_cnt=np.asarray(bg.track_cnt, np.int32), last_points=bg.last_points.copy(), prev_un=pu,
                right_prev_un=rp, right_prev_valid=rv)


@pytest.mark.parametrize("name,n_frames", [("c2_kitti_stereo", 10), ("c4_hd_stereo", 4)])
def test_teacher_forced_vs_oracle(name, n_frames):
    """every frame starts from the ORACLE's state (ids, track_cnt, points, velocity maps, id counter) and must
    reproduce that frame's outputs: integer outputs bit-exact, positions within 0.02 px"""
    c = synth.CONFIGS[name]
    st = synth.make_stream(name, 1)
    fe = cvfe.FrontEnd(params_of(name), c["cam0"], c["cam1"], "raw")
    trk = BatchTracker(cfg_of(name))
    for k in range(n_frames):
        fr = st.frame(k)
        if k > 0:
            trk.set_state(0, oracle_state(fe))
        want = fe.step(fr)["features"]
        trk.track_image(fr.gray0, fr.gray1, fr.time0)
        ids, cams, v = feature_map_arrays(want)
        compare_records(trk.features(0), ids, cams, v, c["cam0"])
        # state after the frame agrees as well (bit-exact integers)
        s_gpu, s_ref = trk.get_state(0), oracle_state(fe)
        assert s_gpu["n"] == s_ref["n"] and s_gpu["next_id"] == s_ref["next_id"]
        assert np.array_equal(s_gpu["ids"], s_ref["ids"]) and np.array_equal(s_gpu["track_cnt"], s_ref["track_cnt"])
        assert np.array_equal(s_gpu["right_prev_valid"], s_ref["right_prev_valid"])
    trk.close()


def test_batch_equals_single_streams():
    """B streams in one tracker == B single-stream trackers, bit for bit (streams are independent)"""
    name, B, T = "c2_kitti_stereo", 3, 4
    streams = [synth.make_stream(name, s) for s in range(B)]
    batch = BatchTracker(cfg_of(name, n_streams=B))
    singles = [BatchTracker(cfg_of(name)) for _ in range(B)]
    for k in range(T):
        frs = [s.frame(k) for s in streams]
        L = np.stack([f.gray0 for f in frs])
        R = np.stack([f.gray1 for f in frs])
        batch.track_image(L, R, [f.time0 for f in frs])
        for s in range(B):
            singles[s].track_image(frs[s].gray0, frs[s].gray1, frs[s].time0)
            a, b = batch.features(s), singles[s].features(0)
            assert a.tobytes() == b.tobytes()
    batch.close()
    [s.close() for s in singles]


def test_state_roundtrip_and_determinism():
    name = "c1_euroc_mono"
    st = synth.make_stream(name, 2)
    a, b = BatchTracker(cfg_of(name)), BatchTracker(cfg_of(name))
    for k in range(3):
        fr = st.frame(k)
        a.track_image(fr.gray0, None, fr.time0)
        b.track_image(fr.gray0, None, fr.time0)
    assert a.features(0).tobytes() == b.features(0).tobytes()
    s = a.get_state(0)
    assert s["n"] == len(a.features(0)) and s["next_id"] == s["ids"].max() + 1 and (s["track_cnt"] >= 1).all()
    b.set_state(0, s)
    fr = st.frame(3)
    a.track_image(fr.gray0, None, fr.time0)
    b.track_image(fr.gray0, None, fr.time0)
    assert a.features(0).tobytes() == b.features(0).tobytes()
    a.close(); b.close()


def test_mono_frame_in_stereo_config_and_first_frame_velocity():
    name = "c2_kitti_stereo"
    c = synth.CONFIGS[name]
    st = synth.make_stream(name, 3)
    fe = cvfe.FrontEnd(params_of(name), c["cam0"], c["cam1"], "raw")
    trk = BatchTracker(cfg_of(name))
    for k in range(4):
        fr = st.frame(k)
        right = None if k == 2 else fr.gray1      # a dropped right image: the stereo block is skipped (:107)
        fr.gray1 = right
        want = fe.step(fr)["features"]
        trk.track_image(fr.gray0, right, fr.time0)
        ids, cams, v = feature_map_arrays(want)
        rec = trk.features(0)
        compare_records(rec, ids, cams, v, c["cam0"])
        if k == 0:
            assert (rec["v"][:, 5:7] == 0).all()      # first frame: all velocities are zero
        if k == 2:
            assert (rec["cam"] == 0).all()
    trk.close()


def test_semantic_background_vs_oracle():
    """TrackSemanticImage: eroded inverse instance mask gates tracking and detection (C3, background part)"""
    name = "c3_zed_dynamic"
    c = synth.CONFIGS[name]
    st = synth.make_stream(name, 0)
    P = params_of(name)
    ref = cvfe.FeatureTracker(P, cvfe.PinholeCamera(**c["cam0"]), cvfe.PinholeCamera(**c["cam1"]))
    trk = FeatureTracker(cfg_of(name))
    for k in range(5):
        fr = st.frame(k)
        if k == 3:                      # a frame without detections: the mask is all 255
            fr.exist_inst, fr.boxes = False, []
        want = ref.track_semantic_image(fr.gray0, fr.gray1, fr.time0, fr.inv_merge_mask, fr.exist_inst)
        trk.batch.track_semantic_image(fr.gray0, fr.gray1, fr.inv_merge_mask, fr.exist_inst, fr.time0)
        ids, cams, v = feature_map_arrays(want)
        rec = trk.batch.features(0)
        compare_records(rec, ids, cams, v, c["cam0"])
        if fr.exist_inst:
            # no background feature sits on an (un-eroded) object pixel of the eroded region mask
            region = cvfe.erode_mask(fr.inv_merge_mask, P.mask_morphology_size)
            left = rec[rec["cam"] == 0]
            new = left  # tracked points were filtered, new points were detected inside the region
            px = np.rint(new["v"][:, 3:5]).astype(int)
            assert (region[px[:, 1], px[:, 0]] != 0).all()


def test_reference_api_names():
    name = "c1_euroc_mono"
    trk = FeatureTracker(cfg_of(name))
    fr = synth.make_stream(name, 0).frame(0)
    out = trk.TrackImage(fr)
    assert len(out) == 150 and all(len(v) == 1 and v[0][0] == 0 and v[0][1].shape == (7,) for v in out.values())
    assert sorted(out) == list(range(1, 151))       # global_id_count starts at 1
    txt = dv.tracker.serialize_point_features(out)
    assert txt.count("\n") == 150 and txt.startswith("0 1 ")


# ---- dynamic mode: TrackSemanticImage + InstsTrack + Output (BASELINE.json config 3) --------------------------
def inst_rows(records):
    """dvfe_inst_obs records -> rows comparable with the oracle's Output()"""
    return [(int(r["inst_id"]), int(r["id"]), int(r["is_stereo"]), r["point"].copy(), r["vel"].copy(),
             r["point_right"].copy(), r["vel_right"].copy(), r["uv"].copy()) for r in records]


def oracle_inst_rows(instances):
    rows = []
    for inst_id, inst in instances.items():
        for fid, f in inst["features"].items():
            rows.append((inst_id, fid, int(f["is_stereo"]), f["point"], f["vel"], f["point_right"], f["vel_right"], f["uv"]))
    return rows


def compare_instances(got, want, cam0, dt_min=0.05):
    assert [(a[0], a[1], a[2]) for a in got] == [(b[0], b[1], b[2]) for b in want], \
        "instance ids / feature ids / stereo bits differ"
    un_tol = 1.5 * POS_TOL / min(cam0["fx"], cam0["fy"])
    for a, b in zip(got, want):
        assert np.abs(a[7] - b[7]).max() <= POS_TOL                  # ROI-local pixel position
        assert np.abs(a[3] - b[3]).max() <= un_tol and np.abs(a[5] - b[5]).max() <= un_tol
        assert np.abs(a[4] - b[4]).max() <= 2 * un_tol / dt_min and np.abs(a[6] - b[6]).max() <= 2 * un_tol / dt_min


def test_dynamic_mode_vs_oracle():
    """8 moving objects with per-instance masks; objects drop out for 1-2 frames (kept, lost_num <= 3), for good
    (erased after > 3 lost frames) and a frame has no detections at all"""
    name = "c3_zed_dynamic"
    c = synth.CONFIGS[name]
    st = synth.make_stream(name, 0)
    fe = cvfe.FrontEnd(params_of(name), c["cam0"], c["cam1"], "dynamic")
    trk = BatchTracker(cfg_of(name, max_instances=12))
    n_inst_feats = 0
    for k in range(12):
        fr = st.frame(k)
        drop = set()
        if k in (3, 4):
            drop = {2}                  # instance 2 is missed for two frames and comes back
        if k >= 5:
            drop = {7}                  # instance 7 disappears for good
        if k == 8:
            drop = set(range(1, 9))     # no detections at all in this frame
        fr.boxes = [b for b in fr.boxes if b["track_id"] not in drop]
        if drop:
            merge = np.zeros_like(fr.merge_mask)
            for b in fr.boxes:
                x, y, w, h = b["rect"]
                merge[y:y + h, x:x + w] |= b["mask"]
            fr.merge_mask, fr.inv_merge_mask, fr.exist_inst = merge, (255 - merge).astype(np.uint8), len(fr.boxes) > 0
        want = fe.step(fr)
        trk.track_semantic_image(fr.gray0, fr.gray1, fr.inv_merge_mask, fr.exist_inst, fr.time0)
        trk.insts_track(0, fr.boxes, fr.time0)
        ids, cams, v = feature_map_arrays(want["features"])
        compare_records(trk.features(0), ids, cams, v, c["cam0"])
        got_i, want_i = inst_rows(trk.insts_output(0)), oracle_inst_rows(want["instances"])
        compare_instances(got_i, want_i, c["cam0"])
        n_inst_feats += len(got_i)
        if k == 8:
            assert len(got_i) == 0
    assert n_inst_feats > 2000
    trk.close()


def test_dynamic_mode_vs_golden():
    name = "c3_zed_dynamic"
    g = load_golden(f"tracker_{name}_dynamic.npz")
    c = synth.CONFIGS[name]
    st = synth.make_stream(name, 0)
    trk = BatchTracker(cfg_of(name, max_instances=8))
    for k in range(int(g["n_frames"])):
        fr = st.frame(k)
        assert crc(fr.gray0) == int(g[f"f{k}_crc0"])
        trk.track_semantic_image(fr.gray0, fr.gray1, fr.inv_merge_mask, fr.exist_inst, fr.time0)
        trk.insts_track(0, fr.boxes, fr.time0)
        compare_records(trk.features(0), g[f"f{k}_ids"], g[f"f{k}_cams"], g[f"f{k}_v"], c["cam0"])
        rows = g[f"f{k}_inst"]
        rec = trk.insts_output(0)
        assert np.array_equal(rec["inst_id"], rows[:, 0].astype(np.uint32))
        assert np.array_equal(rec["id"], rows[:, 1].astype(np.uint32))
        assert np.array_equal(rec["is_stereo"], rows[:, 2].astype(np.int32))
        assert np.abs(rec["uv"] - rows[:, 13:15]).max() <= POS_TOL


def test_instance_capacity_and_argument_errors():
    name = "c3_zed_dynamic"
    fr = synth.make_stream(name, 0).frame(0)
    trk = BatchTracker(cfg_of(name, max_instances=4))
    trk.track_semantic_image(fr.gray0, fr.gray1, fr.inv_merge_mask, fr.exist_inst, fr.time0)
    with pytest.raises(dv.DvfeError) as e:
        trk.insts_track(0, fr.boxes, fr.time0)          # 8 instances > 4 slots
    assert e.value.code == -4
    raw = BatchTracker(cfg_of("c2_kitti_stereo"))       # max_instances = 0
    with pytest.raises(dv.DvfeError):
        raw.insts_track(0, [], 0.0)


def test_cpp_reference_shaped_api(tmp_path):
    """the C++ mirror of FeatureTracker (include/dvfe/feature_tracker.hpp) driven like FeatureTrack() in
    system/main.cpp, from a yaml config in the reference's format; output compared with the oracle through the
    reference's SerializePointFeature text format"""
    import subprocess
    from conftest import ROOT
    name = "c2_kitti_stereo"
    c = synth.CONFIGS[name]
    for i, cam in enumerate((c["cam0"], c["cam1"])):
        (tmp_path / f"cam{i}.yaml").write_text(
            "%YAML:1.0\n---\nmodel_type: PINHOLE\ncamera_name: camera\n"
            f"image_width: {c['width']}\nimage_height: {c['height']}\ndistortion_parameters:\n"
            f"   k1: {cam['k1']!r}\n   k2: {cam['k2']!r}\n   p1: {cam['p1']!r}\n   p2: {cam['p2']!r}\n"
            f"projection_parameters:\n   fx: {cam['fx']!r}\n   fy: {cam['fy']!r}\n   cx: {cam['cx']!r}\n   cy: {cam['cy']!r}\n")
    (tmp_path / "cfg.yaml").write_text(
        "%YAML:1.0\n\nnum_of_cam: 2\nslam_type: \"raw\"\n"
        f"image_width: {c['width']}\nimage_height: {c['height']}\ncam0_calib: \"cam0.yaml\"\ncam1_calib: \"cam1.yaml\"\n"
        f"max_cnt: {c['max_cnt']}\nmin_dist: {c['min_dist']}\nF_threshold: 1.0\nshow_track: 0\nflow_back: 1\n"
        "min_dynamic_dist: 5\nmax_dynamic_cnt: 50\nuse_mask_morphology: 0\nmask_morphology_size: 5\n")
    exe = str(tmp_path / "test_feature_tracker")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "test_feature_tracker.cpp"),
                           "-L" + os.path.join(ROOT, "dynamic_vins_b200"), "-ldvfe",
                           "-Wl,-rpath," + os.path.join(ROOT, "dynamic_vins_b200"), "-o", exe])
    st = synth.make_stream(name, 5)
    n_frames = 4
    fe = cvfe.FrontEnd(params_of(name), c["cam0"], c["cam1"], "raw")
    want = []
    with open(tmp_path / "frames.bin", "wb") as f:
        for k in range(n_frames):
            fr = st.frame(k)
            f.write(np.float64(fr.time0).tobytes()); f.write(fr.gray0.tobytes()); f.write(fr.gray1.tobytes())
            want.append(fe.step(fr)["features"])
    subprocess.check_call([exe, str(tmp_path / "cfg.yaml"), str(tmp_path / "frames.bin"), str(n_frames), "1",
                           str(tmp_path / "out")])
    for k in range(n_frames):
        lines = open(tmp_path / f"out_{k}_point.txt").read().strip().split("\n")
        assert len(lines) == len(want[k])
        for ln, (fid, obs) in zip(lines, want[k].items()):
            tok = ln.split()
            assert int(tok[1]) == fid and int(tok[0]) == (1 if len(obs) == 2 else 0)
            vals = np.array([float(x) for x in tok[2:]])
            ref = np.concatenate([o[1] for o in obs])
            assert len(vals) == len(ref) and np.abs(vals[3:5] - ref[3:5]).max() <= POS_TOL


def test_cpp_batch_tracker_pipelined(tmp_path):
    """dynamic_vins::BatchFeatureTracker (C++ host side for many cameras: TrackImageAsync / Wait / Features, 2 stream groups)
    gives, value for value, the records of the python-driven tracker on the same frames"""
    import subprocess
    from conftest import ROOT
    name, B, T = "c2_kitti_stereo", 3, 5
    c = synth.CONFIGS[name]
    for i, cam in enumerate((c["cam0"], c["cam1"])):
        (tmp_path / f"cam{i}.yaml").write_text(
            "%YAML:1.0\n---\nmodel_type: PINHOLE\ncamera_name: camera\n"
            f"image_width: {c['width']}\nimage_height: {c['height']}\ndistortion_parameters:\n"
            f"   k1: {cam['k1']!r}\n   k2: {cam['k2']!r}\n   p1: {cam['p1']!r}\n   p2: {cam['p2']!r}\n"
            f"projection_parameters:\n   fx: {cam['fx']!r}\n   fy: {cam['fy']!r}\n   cx: {cam['cx']!r}\n   cy: {cam['cy']!r}\n")
    (tmp_path / "cfg.yaml").write_text(
        "%YAML:1.0\n\nnum_of_cam: 2\nslam_type: \"raw\"\n"
        f"image_width: {c['width']}\nimage_height: {c['height']}\ncam0_calib: \"cam0.yaml\"\ncam1_calib: \"cam1.yaml\"\n"
        f"max_cnt: {c['max_cnt']}\nmin_dist: {c['min_dist']}\nF_threshold: 1.0\nshow_track: 0\nflow_back: 1\n"
        "min_dynamic_dist: 5\nmax_dynamic_cnt: 50\nuse_mask_morphology: 0\nmask_morphology_size: 5\n")
    exe = str(tmp_path / "test_feature_tracker")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "test_feature_tracker.cpp"),
                           "-L" + os.path.join(ROOT, "dynamic_vins_b200"), "-ldvfe",
                           "-Wl,-rpath," + os.path.join(ROOT, "dynamic_vins_b200"), "-o", exe])
    streams = [synth.make_stream(name, 20 + s) for s in range(B)]
    ref = BatchTracker(cfg_of(name, n_streams=B))
    want = []
    with open(tmp_path / "frames.bin", "wb") as f:
        for k in range(T):
            frs = [s.frame(k) for s in streams]
            tm = np.array([fr.time0 + 0.002 * i for i, fr in enumerate(frs)], np.float64)
            L = np.stack([fr.gray0 for fr in frs]); R = np.stack([fr.gray1 for fr in frs])
            f.write(tm.tobytes()); f.write(L.tobytes()); f.write(R.tobytes())
            ref.track_image(L, R, tm)
            want.append([obs_to_map(ref.features(s)) for s in range(B)])
    ref.close()
    subprocess.check_call([exe, "batch", str(tmp_path / "cfg.yaml"), str(tmp_path / "frames.bin"), str(T), str(B),
                           str(tmp_path / "out"), "2"])
    for k in range(T):
        for s in range(B):
            lines = open(tmp_path / f"out_s{s}_{k}_point.txt").read().strip().split("\n")
            assert len(lines) == len(want[k][s]) > 50
            for ln, (fid, obs) in zip(lines, want[k][s].items()):
                tok = ln.split()
                assert int(tok[1]) == fid and int(tok[0]) == (1 if len(obs) == 2 else 0)
                vals = np.array([float(x) for x in tok[2:]])
                assert np.array_equal(vals, np.concatenate([o[1] for o in obs]))


def test_pipelined_async_equals_sync():
    """dvfe_track_image_async / dvfe_wait (two steps in flight, upload overlapping compute) gives the same records
    as the synchronous call, step by step"""
    name, B, T = "c2_kitti_stereo", 2, 7
    streams = [synth.make_stream(name, 10 + s) for s in range(B)]
    frames = [[s.frame(k) for s in streams] for k in range(T)]
    L = [np.stack([f.gray0 for f in fr]) for fr in frames]
    R = [np.stack([f.gray1 for f in fr]) for fr in frames]
    sync, pipe = BatchTracker(cfg_of(name, n_streams=B)), BatchTracker(cfg_of(name, n_streams=B))
    want = []
    for k in range(T):
        sync.track_image(L[k], R[k], frames[k][0].time0)
        want.append([sync.features(s).tobytes() for s in range(B)])
    got = []
    pipe.track_image_async(L[0], R[0], frames[0][0].time0)
    for k in range(1, T):
        pipe.track_image_async(L[k], R[k], frames[k][0].time0)
        pipe.wait()                                     # step k-1
        got.append([pipe.features(s).tobytes() for s in range(B)])
    pipe.wait()
    got.append([pipe.features(s).tobytes() for s in range(B)])
    assert got == want
    sync.close(); pipe.close()


@pytest.mark.parametrize("name,n_frames", [("c1_euroc_mono", 30), ("c2_kitti_stereo", 30), ("c4_hd_stereo", 8)])
def test_free_running_long_sequence_vs_oracle(name, n_frames):
    """SURVEY.md Appendix D, free-running mode: both trackers start cold and run the whole sequence on their own
    state.  Reported: first frame whose id set differs, fraction of common ids, position error on common ids.
    Bar (north_star): corner-set agreement >= 99 %, positions within 0.02 px on common features."""
    c = synth.CONFIGS[name]
    st = synth.make_stream(name, 7)
    fe = cvfe.FrontEnd(params_of(name), c["cam0"], c["cam1"], "raw")
    trk = BatchTracker(cfg_of(name))
    first_div, worst_px, min_agree = None, 0.0, 1.0
    for k in range(n_frames):
        fr = st.frame(k)
        want = fe.step(fr)["features"]
        trk.track_image(fr.gray0, fr.gray1, fr.time0)
        got = obs_to_map(trk.features(0))
        common = set(got) & set(want)
        agree = len(common) / max(1, len(set(got) | set(want)))
        min_agree = min(min_agree, agree)
        if first_div is None and set(got) != set(want):
            first_div = k
        for fid in common:
            if [c0 for c0, _ in got[fid]] != [c0 for c0, _ in want[fid]]:
                continue
            for (_, a), (_, b) in zip(got[fid], want[fid]):
                if np.array_equal(np.rint(a[3:5]), np.rint(b[3:5])) or first_div is None:
                    worst_px = max(worst_px, float(np.abs(a[3:5] - b[3:5]).max()))
    print(f"{name}: first id-set divergence at frame {first_div}, min corner-set agreement {min_agree:.4f}, "
          f"worst position error {worst_px:.2e} px")
    assert min_agree >= 0.99
    if first_div is None:
        assert worst_px <= POS_TOL
    trk.close()


def test_64_stream_batch_full_size():
    """BASELINE.json config 5 shape: 64 x 1280x720 stereo streams in one tracker.  8 distinct streams are replicated
    8x; every replica must produce the same bytes as a single-stream tracker fed the same frames."""
    name, B, T = "c5_zed_streams", 64, 3
    src = [synth.make_stream(name, s) for s in range(8)]
    frames = [[s.frame(k) for s in src] for k in range(T)]
    batch = BatchTracker(cfg_of(name, n_streams=B))
    singles = [BatchTracker(cfg_of(name)) for _ in range(2)]
    for k in range(T):
        L = np.stack([frames[k][s % 8].gray0 for s in range(B)])
        R = np.stack([frames[k][s % 8].gray1 for s in range(B)])
        batch.track_image(L, R, frames[k][0].time0)
        ref = {}
        for j, s in enumerate((0, 5)):
            singles[j].track_image(frames[k][s].gray0, frames[k][s].gray1, frames[k][s].time0)
            ref[s] = singles[j].features(0).tobytes()
        for s in range(B):
            rec = batch.features(s)
            if s % 8 in ref:
                assert rec.tobytes() == ref[s % 8]
            left = rec[rec["cam"] == 0]
            assert len(left) == 400 and len(np.unique(left["id"])) == 400
            if k == 0:
                p = left["v"][:, 3:5]
                d2 = ((p[:, None, :] - p[None, :, :]) ** 2).sum(-1) + np.eye(len(p)) * 1e9
                assert d2.min() >= 25 * 25
    batch.close()
    [s.close() for s in singles]


def test_lk_mode_cuda_call_pattern_vs_oracle():
    """the FeatureTrackByLKGpu call pattern (backward pass over all 4 levels, FB threshold 1.0 px,
    front_end/feature_utils.cpp:83-163) with the CPU arithmetic, against the oracle run with the same parameters"""
    name = "c2_kitti_stereo"
    c = synth.CONFIGS[name]
    st = synth.make_stream(name, 9)
    P = params_of(name)
    P.lk_back_max_level, P.fb_threshold = 3, 1.0
    fe = cvfe.FrontEnd(P, c["cam0"], c["cam1"], "raw")
    trk = BatchTracker(cfg_of(name))
    trk.set_lk_mode(3, 1.0)
    for k in range(5):
        fr = st.frame(k)
        want = fe.step(fr)["features"]
        trk.track_image(fr.gray0, fr.gray1, fr.time0)
        ids, cams, v = feature_map_arrays(want)
        compare_records(trk.features(0), ids, cams, v, c["cam0"])
    with pytest.raises(dv.DvfeError):
        trk.set_lk_mode(9, 1.0)
    trk.close()


def test_track_image_naive_vs_oracle():
    """FeatureTracker::TrackImageNaive (front_end/background_tracker.cpp:400-516) against the oracle's restatement of it: eroded
    inverse mask, cv::cuda LK call pattern for the temporal and the stereo call, cv::cuda detector threshold; two streams, the
    second without instances on some frames"""
    name = "c3_zed_dynamic"
    c = synth.CONFIGS[name]
    B = 2
    sts = [synth.make_stream(name, 50 + s) for s in range(B)]
    fes = [cvfe.FrontEnd(params_of(name), c["cam0"], c["cam1"], "naive") for _ in range(B)]
    trk = BatchTracker(cfg_of(name, n_streams=B))
    for k in range(6):
        frs = [st.frame(k) for st in sts]
        if k in (2, 3):
            frs[1].exist_inst, frs[1].inv_merge_mask = False, None
        masks = [fr.inv_merge_mask if fr.inv_merge_mask is not None else np.full(fr.gray0.shape, 255, np.uint8) for fr in frs]
        trk.track_image_naive(np.stack([fr.gray0 for fr in frs]), np.stack([fr.gray1 for fr in frs]), np.stack(masks),
                              [1 if fr.exist_inst else 0 for fr in frs], [fr.time0 for fr in frs])
        for s in range(B):
            want = fes[s].step(frs[s])["features"]
            ids, cams, v = feature_map_arrays(want)
            compare_records(trk.features(s), ids, cams, v, c["cam0"])
    trk.close()


def test_dynamic_mode_batched_streams_equal_single():
    """dvfe_insts_track_batch over B streams == per-stream dvfe_insts_track on single-stream trackers, bit for bit
    (including a stream that has no detections in one frame)"""
    name, B, T = "c3_zed_dynamic", 2, 5
    streams = [synth.make_stream(name, s) for s in range(B)]
    batch = BatchTracker(cfg_of(name, n_streams=B, max_instances=8))
    singles = [BatchTracker(cfg_of(name, max_instances=8)) for _ in range(B)]
    for k in range(T):
        frs = [s.frame(k) for s in streams]
        if k == 2:                                   # stream 1 loses all detections for one frame
            frs[1].boxes, frs[1].exist_inst = [], False
            frs[1].inv_merge_mask = np.full_like(frs[1].inv_merge_mask, 255)
        L = np.stack([f.gray0 for f in frs]); R = np.stack([f.gray1 for f in frs])
        M = np.stack([f.inv_merge_mask for f in frs])
        batch.track_semantic_image(L, R, M, [int(f.exist_inst) for f in frs], [f.time0 for f in frs])
        batch.insts_track_batch([f.boxes for f in frs], [f.time0 for f in frs])
        for s in range(B):
            singles[s].track_semantic_image(frs[s].gray0, frs[s].gray1, frs[s].inv_merge_mask, int(frs[s].exist_inst), frs[s].time0)
            singles[s].insts_track(0, frs[s].boxes, frs[s].time0)
            assert batch.features(s).tobytes() == singles[s].features(0).tobytes()
            a, b = batch.insts_output(s), singles[s].insts_output(0)
            assert a.tobytes() == b.tobytes()
            if k == 2 and s == 1:
                assert len(a) == 0
            else:
                assert len(a) > 300
    batch.close()
    [s.close() for s in singles]


def test_stream_groups_equal_ungrouped():
    """dvfe_config::n_groups splits the streams over G leaf trackers on their own CUDA streams; every record is the
    same as without groups, through the synchronous, pipelined, state and lk-mode entry points"""
    name, B, T = "c2_kitti_stereo", 5, 5
    streams = [synth.make_stream(name, s) for s in range(B)]
    plain, grouped = BatchTracker(cfg_of(name, n_streams=B)), BatchTracker(cfg_of(name, n_streams=B, n_groups=2))
    outs = {0: [], 1: []}
    for k in range(T):
        frs = [s.frame(k) for s in streams]
        L = np.stack([f.gray0 for f in frs]); R = np.stack([f.gray1 for f in frs])
        tm = [f.time0 + 0.001 * i for i, f in enumerate(frs)]
        if k < 2:
            plain.track_image(L, R, tm); grouped.track_image(L, R, tm)
            for s in range(B):
                assert plain.features(s).tobytes() == grouped.features(s).tobytes()
        else:                                              # pipelined: results of step k-1 after wait()
            for j, t in enumerate((plain, grouped)):
                t.track_image_async(L, R, tm)
                if k > 2:
                    t.wait()
                    outs[j].append([t.features(s).tobytes() for s in range(B)])
    for j, t in enumerate((plain, grouped)):
        t.wait()
        outs[j].append([t.features(s).tobytes() for s in range(B)])
    assert outs[0] == outs[1] and len(outs[0]) == T - 2
    st = plain.get_state(3)
    sg = grouped.get_state(3)                              # stream 3 = group 1, local stream 0
    assert all(np.array_equal(st[k], sg[k]) for k in st)
    grouped.set_state(4, st)
    assert all(np.array_equal(grouped.get_state(4)[k], st[k]) for k in st)
    with pytest.raises(Exception):
        grouped.features(B)
    plain.close(); grouped.close()


def test_stream_groups_dynamic_mode():
    name, B, T = "c3_zed_dynamic", 3, 3
    streams = [synth.make_stream(name, s) for s in range(B)]
    plain = BatchTracker(cfg_of(name, n_streams=B, max_instances=8))
    grouped = BatchTracker(cfg_of(name, n_streams=B, max_instances=8, n_groups=2))
    for k in range(T):
        frs = [s.frame(k) for s in streams]
        L = np.stack([f.gray0 for f in frs]); R = np.stack([f.gray1 for f in frs])
        M = np.stack([f.inv_merge_mask for f in frs])
        for t in (plain, grouped):
            t.track_semantic_image(L, R, M, [int(f.exist_inst) for f in frs], [f.time0 for f in frs])
            if k == 1:
                for s in range(B):
                    t.insts_track(s, frs[s].boxes, frs[s].time0)
            else:
                t.insts_track_batch([f.boxes for f in frs], [f.time0 for f in frs])
        for s in range(B):
            assert plain.features(s).tobytes() == grouped.features(s).tobytes()
            a, b = plain.insts_output(s), grouped.insts_output(s)
            assert len(a) > 300 and a.tobytes() == b.tobytes()
    plain.close(); grouped.close()


@pytest.mark.parametrize("name,B,groups", [("c1_euroc_mono", 1, 1), ("c2_kitti_stereo", 3, 1), ("c2_kitti_stereo", 4, 2)])
def test_color_and_undistort_ingest_equals_prepared_gray(name, B, groups):
    """dvfe_set_input(3) + dvfe_set_undistort_maps: BGR frames remapped + converted on the device give byte-identical
    records to a tracker fed with ImageProcessor::Run's host result (cv2.remap + cvtColor), through the synchronous,
    pipelined and device-pointer entry points"""
    from oracle import image_process as ip
    c = synth.CONFIGS[name]
    W, H, stereo = c["width"], c["height"], bool(c["stereo"])
    maps0 = ip.undistort_maps(c["cam0"] if name == "c1_euroc_mono" else synth.EUROC_CAM0 | {"cx": W / 2, "cy": H / 2}, W, H)
    maps1 = ip.undistort_maps(synth.EUROC_CAM0 | {"cx": W / 2 + 3, "cy": H / 2 - 2, "k1": -0.2}, W, H) if stereo else None
    streams = [synth.make_stream(name, s) for s in range(B)]
    ref = BatchTracker(cfg_of(name, n_streams=B))
    trk = BatchTracker(cfg_of(name, n_streams=B, n_groups=groups))
    trk.set_input(3)
    trk.set_undistort_maps(0, maps0[0], maps0[1])
    if stereo:
        trk.set_undistort_maps(1, maps1[0], maps1[1])
    import torch
    for k in range(5):
        frs = [s.frame(k) for s in streams]
        c0 = np.stack([synth.colorize(f.gray0) for f in frs])
        c1 = np.stack([synth.colorize(f.gray1) for f in frs]) if stereo else None
        prepared = [ip.run(c0[s], c1[s] if stereo else None, maps0[:2], maps1[:2] if stereo else None) for s in range(B)]
        g0 = np.stack([p[0] for p in prepared])
        g1 = np.stack([p[1] for p in prepared]) if stereo else None
        tm = [f.time0 for f in frs]
        ref.track_image(g0, g1, tm)
        if k < 2:
            trk.track_image(c0, c1, tm)
        elif k < 4:
            trk.track_image_async(c0, c1, tm); trk.wait()
        else:
            d0 = torch.from_numpy(c0).cuda()
            d1 = torch.from_numpy(c1).cuda() if stereo else None
            trk.track_image_device(d0.data_ptr(), d1.data_ptr() if stereo else 0, H * W * 3, W * 3, tm)
        for s in range(B):
            a, b = trk.features(s), ref.features(s)
            assert len(a) > 50 and a.tobytes() == b.tobytes(), (k, s)
    # switching back to gray input, maps cleared: plain path again
    trk.set_input(1)
    trk.set_undistort_maps(0, None, None)
    trk.set_undistort_maps(1, None, None)
    frs = [s.frame(5) for s in streams]
    g0 = np.stack([f.gray0 for f in frs]); g1 = np.stack([f.gray1 for f in frs]) if stereo else None
    ref.track_image(g0, g1, [f.time0 for f in frs]); trk.track_image(g0, g1, [f.time0 for f in frs])
    assert all(trk.features(s).tobytes() == ref.features(s).tobytes() for s in range(B))
    ref.close(); trk.close()


def test_color_ingest_semantic_path():
    """BGR input through dvfe_track_semantic_image (the region mask stays one byte per pixel) + instances"""
    name = "c3_zed_dynamic"
    st = synth.make_stream(name, 0)
    ref = BatchTracker(cfg_of(name, max_instances=8))
    trk = BatchTracker(cfg_of(name, max_instances=8))
    trk.set_input(3)
    for k in range(3):
        f = st.frame(k)
        ref.track_semantic_image(f.gray0, f.gray1, f.inv_merge_mask, int(f.exist_inst), f.time0)
        ref.insts_track(0, f.boxes, f.time0)
        # colorize() keeps B = gray; use a colour image whose BGR2GRAY is exactly the gray frame: B = G = R = gray
        c0, c1 = np.repeat(f.gray0[..., None], 3, -1), np.repeat(f.gray1[..., None], 3, -1)
        assert np.array_equal(cv2.cvtColor(c0, cv2.COLOR_BGR2GRAY), f.gray0)
        trk.track_semantic_image(c0, c1, f.inv_merge_mask, int(f.exist_inst), f.time0)
        trk.insts_track(0, f.boxes, f.time0)
        assert trk.features(0).tobytes() == ref.features(0).tobytes()
        assert trk.insts_output(0).tobytes() == ref.insts_output(0).tobytes()
    ref.close(); trk.close()


@pytest.mark.parametrize("groups", [1, 2])
def test_dynamic_pipelined_equals_synchronous(groups):
    """dvfe_track_dynamic_async + dvfe_wait (two frames in flight) == dvfe_track_semantic_image + dvfe_insts_track_batch,
    record for record, including a stream that loses its detections for one frame and a frame without any"""
    name, B, T = "c3_zed_dynamic", 3, 7
    streams = [synth.make_stream(name, s) for s in range(B)]
    sync = BatchTracker(cfg_of(name, n_streams=B, max_instances=8))
    pipe = BatchTracker(cfg_of(name, n_streams=B, max_instances=8, n_groups=groups))
    want, got = [], []
    for k in range(T):
        frs = [s.frame(k) for s in streams]
        if k == 2:
            frs[1].boxes, frs[1].exist_inst = [], False
            frs[1].inv_merge_mask = np.full_like(frs[1].inv_merge_mask, 255)
        if k == 4:
            for f in frs:
                f.boxes, f.exist_inst = [], False
                f.inv_merge_mask = np.full_like(f.inv_merge_mask, 255)
        L = np.stack([f.gray0 for f in frs]); R = np.stack([f.gray1 for f in frs]); M = np.stack([f.inv_merge_mask for f in frs])
        ex = [int(f.exist_inst) for f in frs]; tm = [f.time0 for f in frs]
        sync.track_semantic_image(L, R, M, ex, tm)
        sync.insts_track_batch([f.boxes for f in frs], tm)
        want.append([(sync.features(s).tobytes(), sync.insts_output(s).tobytes()) for s in range(B)])
        pipe.track_dynamic_async(L, R, M, ex, [f.boxes for f in frs], tm)
        if k > 0:
            pipe.wait()
            got.append([(pipe.features(s).tobytes(), pipe.insts_output(s).tobytes()) for s in range(B)])
    pipe.wait()
    got.append([(pipe.features(s).tobytes(), pipe.insts_output(s).tobytes()) for s in range(B)])
    assert len(got) == T
    for k in range(T):
        for s in range(B):
            assert got[k][s][0] == want[k][s][0], ("features", k, s)
            assert got[k][s][1] == want[k][s][1], ("instances", k, s)
    assert len(want[3][0][1]) > 0 and len(want[4][0][1]) == 0
    # a synchronous call after pipelined ones continues the same state
    frs = [s.frame(T) for s in streams]
    L = np.stack([f.gray0 for f in frs]); R = np.stack([f.gray1 for f in frs]); M = np.stack([f.inv_merge_mask for f in frs])
    for t in (sync, pipe):
        t.track_semantic_image(L, R, M, [int(f.exist_inst) for f in frs], [f.time0 for f in frs])
        t.insts_track_batch([f.boxes for f in frs], [f.time0 for f in frs])
    for s in range(B):
        assert sync.features(s).tobytes() == pipe.features(s).tobytes()
        assert sync.insts_output(s).tobytes() == pipe.insts_output(s).tobytes()
    sync.close(); pipe.close()


def test_graph_replay_equals_plain_launches_with_staged_upload():
    """The frame step replayed as a CUDA graph must order the right image's level-0 copy behind the upload of the tracker's
    staging buffer (widths whose rows are not a multiple of 64 bytes go through it).  24 C2-shaped streams make the upload
    long enough (22 MB) that a copy kernel that does not wait for it reads an unwritten buffer: the first frames after a cold
    start then lose every stereo match.  Graph replay (default) vs plain launches (DVFE_GRAPHS=0), pipelined, byte for byte."""
    name, B, T = "c2_kitti_stereo", 24, 4
    src = [synth.make_stream(name, 70 + s) for s in range(4)]
    frames = [[s.frame(k) for s in src] for k in range(T)]
    L = [np.stack([frames[k][s % 4].gray0 for s in range(B)]) for k in range(T)]
    R = [np.stack([frames[k][s % 4].gray1 for s in range(B)]) for k in range(T)]
    out = {}
    for graphs in ("0", "1"):
        os.environ["DVFE_GRAPHS"] = graphs
        try:
            trk = BatchTracker(cfg_of(name, n_streams=B, n_groups=2))
        finally:
            del os.environ["DVFE_GRAPHS"]
        rec = []
        trk.track_image_async(L[0], R[0], frames[0][0].time0)
        for k in range(1, T):
            trk.track_image_async(L[k], R[k], frames[k][0].time0)
            trk.wait()
            rec.append([trk.features(s).tobytes() for s in range(B)])
        trk.wait()
        rec.append([trk.features(s).tobytes() for s in range(B)])
        trk.close()
        out[graphs] = rec
    assert out["0"] == out["1"]
    from dynamic_vins_b200 import _lib
    n_right = int((np.frombuffer(out["1"][1][0], dtype=_lib.OBS_DTYPE)["cam"] == 1).sum())
    assert n_right > 100, "stereo matches of the second frame are missing"
