"""CPU tests of the host-side logic: synthetic generator determinism, ping-pong playback, record -> map
conversion, serialisation format, and the multi-rank stream sharding over a world-size-2 gloo group."""
import os
import socket

import numpy as np
import pytest

from conftest import ROOT, crc
import dynamic_vins_b200 as dv
from dynamic_vins_b200 import shard, synth
from dynamic_vins_b200._lib import OBS_DTYPE


def test_synth_is_deterministic_and_pure():
    a = synth.make_stream("c1_euroc_mono", 3)
    b = synth.make_stream("c1_euroc_mono", 3)
    f5a, f2a = a.frame(5), a.frame(2)
    f2b, f5b = b.frame(2), b.frame(5)
    assert crc(f5a.gray0) == crc(f5b.gray0) and crc(f2a.gray0) == crc(f2b.gray0)
    assert f5a.gray1 is None and f5a.time0 == 0.25
    c = synth.make_stream("c1_euroc_mono", 4)
    assert crc(c.frame(2).gray0) != crc(f2a.gray0)


def test_synth_dynamic_frame_shape():
    fr = synth.make_stream("c3_zed_dynamic", 0).frame(1)
    assert fr.gray0.shape == (720, 1280) and fr.gray1.shape == (720, 1280) and fr.exist_inst and len(fr.boxes) == 8
    merge = np.zeros_like(fr.merge_mask)
    for b in fr.boxes:
        x, y, w, h = b["rect"]
        assert b["mask"].shape == (h, w) and set(np.unique(b["mask"])) <= {0, 255}
        assert b["mask"][0].any() and b["mask"][-1].any() and b["mask"][:, 0].any() and b["mask"][:, -1].any()  # tight box
        merge[y:y + h, x:x + w] |= b["mask"]
    assert np.array_equal(merge, fr.merge_mask) and np.array_equal(255 - merge, fr.inv_merge_mask)


def test_pingpong_positions():
    assert synth.pingpong_positions(4, 10) == [0, 1, 2, 3, 2, 1, 0, 1, 2, 3]
    assert synth.pingpong_positions(1, 3) == [0, 0, 0]


def test_obs_to_map_and_serialisation():
    rec = np.zeros(3, dtype=OBS_DTYPE)
    rec["id"] = [7, 7, 9]
    rec["cam"] = [0, 1, 0]
    rec["v"][:, 2] = 1.0
    rec["v"][0, 3] = 12.5
    m = dv.obs_to_map(rec)
    assert list(m) == [7, 9] and [c for c, _ in m[7]] == [0, 1] and m[7][0][1][3] == 12.5
    txt = dv.tracker.serialize_point_features(m)
    lines = txt.strip().split("\n")
    assert lines[0].startswith("1 7 ") and len(lines[0].split()) == 16      # stereo observation: 2 + 14 fields
    assert lines[1].startswith("0 9 ") and len(lines[1].split()) == 9


def test_sharding_covers_all_streams():
    for world in (1, 2, 4, 8):
        ids = sorted(s for r in range(world) for s in shard.shard_fixed_total(64, world, r))
        assert ids == list(range(64))
        ids = sorted(s for r in range(world) for s in shard.stream_ids_for_rank(64, r))
        assert ids == list(range(64 * world))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = shard.stream_ids_for_rank(4, rank)
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    ms = shard.reduce_max([10.0 + rank, 5.0 - rank], dist)      # device times: the slowest rank counts
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, gathered, ms))


def test_two_rank_gloo_sharding_and_max_reduce():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = [q.get(timeout=120) for _ in range(2)]
    [p.join(timeout=30) for p in procs]
    for rank, gathered, ms in res:
        assert gathered == [[0, 1, 2, 3], [4, 5, 6, 7]]         # disjoint, every stream owned exactly once
        assert ms == [11.0, 5.0]


def test_point_feature_wire_format_round_trip():
    """SerializePointFeature / DeserializePointFeature text format (utils/io/feature_serialization.cpp:26-70)"""
    rng = np.random.default_rng(0)
    pts = {}
    for fid in (3, 8, 21):
        obs = [(0, np.r_[rng.standard_normal(2), 1.0, rng.uniform(0, 700, 2), rng.standard_normal(2)])]
        if fid != 8:
            obs.append((1, np.r_[rng.standard_normal(2), 1.0, rng.uniform(0, 700, 2), rng.standard_normal(2)]))
        pts[fid] = obs
    txt = dv.tracker.serialize_point_features(pts)
    back = dv.tracker.deserialize_point_features(txt)
    assert list(back) == [3, 8, 21]
    for fid in pts:
        assert [c for c, _ in back[fid]] == [c for c, _ in pts[fid]]
        for (_, a), (_, b) in zip(back[fid], pts[fid]):
            assert np.array_equal(a, b)          # repr() round-trips doubles exactly
    from oracle import cv_front_end as cvfe
    assert cvfe.serialize_point_features(pts) == txt


def test_cpp_frontend_hand_off_header(tmp_path):
    """include/dvfe/frontend_io.hpp compiled by g++ and run without a device: FeatureQueue semantics (basic/feature_queue.h:19-71),
    the text round trip of the reference's point-feature format, and the estimator's map type instantiated with an Eigen
    7-vector (the stand-in Eigen of oracle/shim; test code may use oracle/)."""
    import subprocess
    exe = str(tmp_path / "test_frontend_io")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "oracle", "shim"),
                           os.path.join(ROOT, "tests", "cpp", "test_frontend_io.cpp"),
                           "-L" + os.path.join(ROOT, "dynamic_vins_b200"), "-ldvfe", "-lpthread",
                           "-Wl,-rpath," + os.path.join(ROOT, "dynamic_vins_b200"), "-o", exe])
    out = subprocess.run([exe, "host", str(tmp_path)], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    assert "frontend_io host checks ok" in out.stdout
