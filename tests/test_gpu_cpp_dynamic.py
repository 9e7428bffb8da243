"""The C++ mirror of the reference's front-end classes (include/dvfe/feature_tracker.hpp), compiled into a driver that runs
dynamic mode exactly like FeatureTrack() in system/main.cpp:193-254 does -- reset visibility, AddViodeInstances,
TrackSemanticImage, InstsTrack, Output -- on C3-shaped frames (1280x720 stereo, 8 instance masks), compared with the oracle;
and BatchFeatureTracker::TrackDynamicAsync (pipelined, several cameras) compared with the synchronous calls."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT
from dynamic_vins_b200 import BatchTracker, obs_to_map, synth
from oracle import cv_front_end as cvfe
from test_gpu_tracker import POS_TOL, cfg_of, params_of

pytestmark = pytest.mark.gpu


def write_config(tmp_path, name, slam_type="dynamic"):
    c = synth.CONFIGS[name]
    for i, cam in enumerate((c["cam0"], c["cam1"])):
        (tmp_path / f"cam{i}.yaml").write_text(
            "%YAML:1.0\n---\nmodel_type: PINHOLE\ncamera_name: camera\n"
            f"image_width: {c['width']}\nimage_height: {c['height']}\ndistortion_parameters:\n"
            f"   k1: {cam['k1']!r}\n   k2: {cam['k2']!r}\n   p1: {cam['p1']!r}\n   p2: {cam['p2']!r}\n"
            f"projection_parameters:\n   fx: {cam['fx']!r}\n   fy: {cam['fy']!r}\n   cx: {cam['cx']!r}\n   cy: {cam['cy']!r}\n")
    (tmp_path / "cfg.yaml").write_text(
        f"%YAML:1.0\n\nnum_of_cam: 2\nslam_type: \"{slam_type}\"\n"
        f"image_width: {c['width']}\nimage_height: {c['height']}\ncam0_calib: \"cam0.yaml\"\ncam1_calib: \"cam1.yaml\"\n"
        f"max_cnt: {c['max_cnt']}\nmin_dist: {c['min_dist']}\nF_threshold: 1.0\nshow_track: 0\nflow_back: 1\n"
        f"min_dynamic_dist: {c.get('min_dynamic_dist', 5)}\nmax_dynamic_cnt: {c.get('max_dynamic_cnt', 50)}\n"
        f"use_mask_morphology: {c.get('use_mask_morphology', 0)}\nmask_morphology_size: {c.get('mask_morphology_size', 5)}\n")
    return str(tmp_path / "cfg.yaml")


def build_driver(tmp_path):
    exe = str(tmp_path / "test_feature_tracker")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "test_feature_tracker.cpp"),
                           "-L" + os.path.join(ROOT, "dynamic_vins_b200"), "-ldvfe",
                           "-Wl,-rpath," + os.path.join(ROOT, "dynamic_vins_b200"), "-o", exe])
    return exe


def write_dyn_frame(f, fr):
    f.write(np.float64(fr.time0).tobytes()); f.write(fr.gray0.tobytes()); f.write(fr.gray1.tobytes())
    inv = fr.inv_merge_mask if fr.inv_merge_mask is not None else np.full(fr.gray0.shape, 255, np.uint8)
    f.write(np.ascontiguousarray(inv).tobytes())
    f.write(np.int32(1 if fr.exist_inst else 0).tobytes()); f.write(np.int32(len(fr.boxes)).tobytes())
    for b in fr.boxes:
        x, y, w, h = b["rect"]
        f.write(np.array([b["track_id"], x, y, w, h], np.int32).tobytes())
        f.write(np.ascontiguousarray(b["mask"], np.uint8).tobytes())


def read_points(path):
    out = {}
    for ln in open(path).read().strip().split("\n"):
        if not ln:
            continue
        tok = ln.split()
        vals = np.array([float(x) for x in tok[2:]])
        out[int(tok[1])] = vals.reshape(-1, 7)
    return out


def read_instances(path):
    out = {}
    txt = open(path).read().strip()
    for ln in txt.split("\n") if txt else []:
        t = ln.split()
        v = [float(x) for x in t[3:14]]
        out.setdefault(int(t[0]), {})[int(t[1])] = dict(is_stereo=bool(int(t[2])), point=np.array(v[0:3]), vel=np.array(v[3:5]),
                                                        point_right=np.array(v[5:8]), vel_right=np.array(v[8:10]), disp=v[10],
                                                        box_track_id=int(t[14]))
    return out


def test_cpp_dynamic_mode_reference_shaped_api(tmp_path):
    name, n_frames = "c3_zed_dynamic", 7
    c = synth.CONFIGS[name]
    cfg = write_config(tmp_path, name)
    exe = build_driver(tmp_path)
    st = synth.make_stream(name, 41)
    fe = cvfe.FrontEnd(params_of(name), c["cam0"], c["cam1"], "dynamic")
    H, W = c["height"], c["width"]
    disp = -(np.arange(H * W, dtype=np.float32) % 4099).reshape(H, W) - 1.0        # the map the driver builds
    drop = {2: {4}, 3: {4}, 4: {4}, 5: {4}, 6: {4}}                                  # instance 4 disappears and is erased
    want = []
    with open(tmp_path / "frames.bin", "wb") as f:
        for k in range(n_frames):
            fr = st.frame(k)
            if k in drop:
                fr.boxes = [b for b in fr.boxes if b["track_id"] not in drop[k]]
            write_dyn_frame(f, fr)
            res = fe.step(fr, disp=disp)
            table = [(key, inst.lost_num, int(inst.is_curr_visible)) for key, inst in sorted(fe.insts.instances.items())]
            want.append((res, table))
    subprocess.check_call([exe, "dynamic", cfg, str(tmp_path / "frames.bin"), str(n_frames), str(tmp_path / "out"), "12"])
    for k, (res, table) in enumerate(want):
        got = read_points(tmp_path / f"out_{k}_point.txt")
        assert sorted(got) == sorted(res["features"]), f"frame {k}: background ids"
        for fid, obs in res["features"].items():
            ref = np.stack([o[1] for o in obs])
            assert got[fid].shape == ref.shape, f"frame {k}: camera list of id {fid}"
            assert np.abs(got[fid][:, 3:5] - ref[:, 3:5]).max() <= POS_TOL
        gi = read_instances(tmp_path / f"out_{k}_inst.txt")
        wi = {key: v for key, v in res["instances"].items() if v["features"]}
        assert sorted(gi) == sorted(wi), f"frame {k}: instances with features"
        for key in wi:
            assert sorted(gi[key]) == sorted(wi[key]["features"]), f"frame {k}: feature ids of instance {key}"
            for fid, f in wi[key]["features"].items():
                g = gi[key][fid]
                assert g["is_stereo"] == f["is_stereo"] and g["box_track_id"] == key
                assert np.abs(g["point"][:2] - f["point"][:2]).max() <= 1.5 * POS_TOL / c["cam0"]["fx"]
                # the ROI-local disparity lookup (reference quirk Q8) lands on the same pixel unless the position sits within
                # the tolerance of a rounding boundary
                u, v = f["uv"]
                if min(abs(u - np.floor(u) - 0.5), abs(v - np.floor(v) - 0.5)) > POS_TOL:
                    assert g["disp"] == f["disp"]
        got_table = [tuple(int(x) for x in ln.split()) for ln in open(tmp_path / f"out_{k}_table.txt").read().strip().split("\n") if ln]
        assert got_table == table, f"frame {k}: instance table (key, lost_num, is_curr_visible)"


def test_cpp_batch_dynamic_async_equals_synchronous(tmp_path):
    """BatchFeatureTracker::TrackDynamicAsync (two frames in flight, 2 stream groups) == the synchronous python-driven calls"""
    name, B, T = "c3_zed_dynamic", 2, 5
    cfg = write_config(tmp_path, name)
    exe = build_driver(tmp_path)
    streams = [synth.make_stream(name, 50 + s) for s in range(B)]
    ref = BatchTracker(cfg_of(name, n_streams=B, max_instances=12))
    want = []
    with open(tmp_path / "frames.bin", "wb") as f:
        for k in range(T):
            frs = [s.frame(k) for s in streams]
            if k == 3:
                frs[1].boxes = frs[1].boxes[:3]
            for fr in frs:
                write_dyn_frame(f, fr)
            L = np.stack([fr.gray0 for fr in frs]); R = np.stack([fr.gray1 for fr in frs])
            M = np.stack([fr.inv_merge_mask for fr in frs])
            ref.track_semantic_image(L, R, M, [fr.exist_inst for fr in frs], [fr.time0 for fr in frs])
            ref.insts_track_batch([fr.boxes for fr in frs], [fr.time0 for fr in frs])
            want.append([(obs_to_map(ref.features(s)), ref.insts_output(s)) for s in range(B)])
    ref.close()
    subprocess.check_call([exe, "dynbatch", cfg, str(tmp_path / "frames.bin"), str(T), str(B), str(tmp_path / "out"), "12"])
    for k in range(T):
        for s in range(B):
            feats, inst = want[k][s]
            got = read_points(tmp_path / f"out_s{s}_{k}_point.txt")
            assert sorted(got) == sorted(feats)
            for fid, obs in feats.items():
                assert np.array_equal(got[fid], np.stack([o[1] for o in obs]))
            gi = read_instances(tmp_path / f"out_s{s}_{k}_inst.txt")
            assert sum(len(v) for v in gi.values()) == len(inst)
            for row in inst:
                g = gi[int(row["inst_id"])][int(row["id"])]
                assert g["is_stereo"] == bool(row["is_stereo"])
                assert np.array_equal(g["point"], row["point"]) and np.array_equal(g["vel"], row["vel"])
                assert np.array_equal(g["point_right"], row["point_right"]) and np.array_equal(g["vel_right"], row["vel_right"])


def test_label_image_path_equals_mask_path_and_oracle():
    """SetMaskAndRoi on the device.  One label image per stream (bit b = instance b) replaces the inv_merge_mask
    upload and every per-box ROI mask: the records must equal, byte for byte, those of the host-mask path, from host memory and
    from device memory, and agree with the oracle's add_instances + TrackSemanticImage + InstsTrack."""
    import torch
    name, B, T = "c3_zed_dynamic", 2, 6
    c = synth.CONFIGS[name]
    streams = [synth.make_stream(name, 60 + s) for s in range(B)]
    fe = cvfe.FrontEnd(params_of(name), c["cam0"], c["cam1"], "dynamic")          # stream 0 against the oracle
    ref = BatchTracker(cfg_of(name, n_streams=B, max_instances=12))
    lab_h = BatchTracker(cfg_of(name, n_streams=B, max_instances=12))
    lab_d = BatchTracker(cfg_of(name, n_streams=B, max_instances=12))
    keep = []
    for k in range(T):
        frs = [s.frame(k) for s in streams]
        if k >= 3:
            frs[0].boxes = [b for b in frs[0].boxes if b["track_id"] != 2]         # an instance disappears
            frs[0].merge_mask[:] = 0
            for b in frs[0].boxes:
                x, y, w, h = b["rect"]
                frs[0].merge_mask[y:y + h, x:x + w] |= b["mask"]
            frs[0].inv_merge_mask = (255 - frs[0].merge_mask).astype(np.uint8)
        if k == 5:
            frs[1].boxes, frs[1].exist_inst = [], False                              # a frame without instances
            frs[1].inv_merge_mask[:] = 255
        L = np.stack([fr.gray0 for fr in frs]); R = np.stack([fr.gray1 for fr in frs])
        M = np.stack([fr.inv_merge_mask for fr in frs])
        tm = [fr.time0 for fr in frs]
        ref.track_dynamic_async(L, R, M, [fr.exist_inst for fr in frs], [fr.boxes for fr in frs], tm)
        ref.wait()
        labs = [synth.label_image(fr) for fr in frs]
        LAB = np.stack([lb for lb, _ in labs])
        lab_h.track_dynamic_labels_async(L, R, LAB, [bx for _, bx in labs], tm)
        lab_h.wait()
        dl, dr, dm = torch.from_numpy(L).cuda(), torch.from_numpy(R).cuda(), torch.from_numpy(LAB).cuda()
        keep.append((dl, dr, dm))
        torch.cuda.synchronize()
        lab_d.track_dynamic_labels_async(dl.data_ptr(), dr.data_ptr(), dm.data_ptr(), [bx for _, bx in labs], tm, device=True)
        lab_d.wait()
        for s in range(B):
            a = ref.features(s).tobytes(), ref.insts_output(s).tobytes()
            assert (lab_h.features(s).tobytes(), lab_h.insts_output(s).tobytes()) == a, f"frame {k} stream {s}: host label path"
            assert (lab_d.features(s).tobytes(), lab_d.insts_output(s).tobytes()) == a, f"frame {k} stream {s}: device label path"
            assert ref.insts_table(s) == lab_h.insts_table(s) == lab_d.insts_table(s)
        want = fe.step(frs[0])
        got = obs_to_map(lab_d.features(0))
        assert sorted(got) == sorted(want["features"])
        rec = lab_d.insts_output(0)
        assert sorted(set(int(x) for x in rec["inst_id"])) == sorted(key for key, v in want["instances"].items() if v["features"])
        for key, v in want["instances"].items():
            r = rec[rec["inst_id"] == key]
            assert [int(x) for x in r["id"]] == sorted(v["features"])
            for row in r:
                assert np.abs(row["uv"] - v["features"][int(row["id"])]["uv"]).max() <= POS_TOL
    ref.close(); lab_h.close(); lab_d.close()


def test_cpp_feature_track_frame_and_queue(tmp_path):
    """include/dvfe/frontend_io.hpp: FeatureTrackFrame (one iteration of FeatureTrack(), system/main.cpp:178-330) as the producer,
    a consumer thread taking the FrontendFeature frames from the FeatureQueue and writing them with SerializePointFeature: the
    files must be identical to those of the hand-written call sequence of test_feature_tracker.cpp (compared with the oracle in
    test_cpp_dynamic_mode_reference_shaped_api)."""
    name, n_frames = "c3_zed_dynamic", 5
    cfg = write_config(tmp_path, name)
    exe_a = build_driver(tmp_path)
    exe_b = str(tmp_path / "test_frontend_io")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "oracle", "shim"),
                           os.path.join(ROOT, "tests", "cpp", "test_frontend_io.cpp"),
                           "-L" + os.path.join(ROOT, "dynamic_vins_b200"), "-ldvfe", "-lpthread",
                           "-Wl,-rpath," + os.path.join(ROOT, "dynamic_vins_b200"), "-o", exe_b])
    st = synth.make_stream(name, 43)
    with open(tmp_path / "frames.bin", "wb") as f:
        for k in range(n_frames):
            write_dyn_frame(f, st.frame(k))
    subprocess.check_call([exe_a, "dynamic", cfg, str(tmp_path / "frames.bin"), str(n_frames), str(tmp_path / "a"), "12"])
    subprocess.check_call([exe_b, "dynamic", cfg, str(tmp_path / "frames.bin"), str(n_frames), str(tmp_path / "b"), "12"])
    for k in range(n_frames):
        for kind in ("point", "inst"):
            a = open(tmp_path / f"a_{k}_{kind}.txt").read()
            b = open(tmp_path / f"b_{k}_{kind}.txt").read()
            assert a == b and (kind == "inst" or len(a) > 0), f"frame {k}: {kind} file"
