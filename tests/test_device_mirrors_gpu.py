"""Device-resident mirrors of the reference's other two call sites and the sharded keyframe DB,
against golden outputs of the UNMODIFIED reference functions:
  Point3D.find_2D_and_3D_correspondenses (Point3D.py:33-54)  -> slammatch.find_2d_3d_device
  keypoint.track_keypoints_left_to_right_new (keypoint.py:35-57, up to findFundamentalMat) -> slammatch.stereo_matches_device
"""
import os

import numpy as np
import pytest

import slammatch
from slammatch import _lib, synth
from oracle import oracle as orc
from conftest import GOLDEN

pytestmark = pytest.mark.gpu


def test_find_2d_3d_device_equals_reference_outputs():
    import torch
    z = np.load(os.path.join(GOLDEN, "reference_functions.npz"))
    for name in z["names"]:
        name = str(name)
        des1, des2 = torch.from_numpy(z[name + "/q"]).cuda(), torch.from_numpy(z[name + "/t"]).cuda()
        p1, p2 = torch.from_numpy(z[name + "/p1"]).cuda(), torch.from_numpy(z[name + "/p2"]).cuda()
        pts3d = torch.from_numpy(z[name + "/pts3d"]).cuda()
        q2, Q1, q1 = slammatch.find_2d_3d_device(des1, p1, p2, des2, pts3d, max_distance=500)
        assert np.array_equal(q2.cpu().numpy().astype(np.float64).reshape(-1), z[name + "/p3_q2"].reshape(-1)), name
        assert np.array_equal(Q1.cpu().numpy().reshape(-1), z[name + "/p3_Q1"].reshape(-1)), name
        assert np.array_equal(q1.cpu().numpy().astype(np.float64).reshape(-1), z[name + "/p3_q1"].reshape(-1)), name
    # the distance filter is not vacuous and is strict
    q, t = synth.planted(400, 500, 9)
    pts = np.zeros((400, 3))
    pts[::2, 1] = 500.0                      # |Y| == max_Distance is rejected (strict '<', Point3D.py:45)
    pts[1::4, 2] = -499.999
    r_q2, r_Q1, r_q1 = orc.find_2d_3d_restated(q, np.zeros((400, 2), np.float32), np.zeros((500, 2), np.float32), t, pts, 500,
                                               knn=orc.c_knn2)
    g = slammatch.find_2d_3d_device(torch.from_numpy(q).cuda(), torch.zeros((400, 2), device="cuda"),
                                    torch.zeros((500, 2), device="cuda"), torch.from_numpy(t).cuda(),
                                    torch.from_numpy(pts).cuda(), max_distance=500)
    assert g[1].shape[0] == np.asarray(r_Q1).reshape(-1, 3).shape[0] > 0
    assert np.array_equal(g[1].cpu().numpy(), np.asarray(r_Q1).reshape(-1, 3))


def test_stereo_matches_device_equals_reference_gathers():
    import torch
    z = np.load(os.path.join(GOLDEN, "reference_stereo.npz"))
    for name in z["names"]:
        name = str(name)
        dl, dr = torch.from_numpy(z[name + "/q"]).cuda(), torch.from_numpy(z[name + "/t"]).cuda()
        p1, p2 = torch.from_numpy(z[name + "/p1"]).cuda(), torch.from_numpy(z[name + "/p2"]).cuda()
        pts_l, pts_r, des_l, des_r = slammatch.stereo_matches_device(p1, dl, p2, dr)
        assert np.array_equal(pts_l.cpu().numpy().astype(np.float64), z[name + "/pts_left"]), name
        assert np.array_equal(pts_r.cpu().numpy().astype(np.float64), z[name + "/pts_right"]), name
        assert np.array_equal(des_l.cpu().numpy(), z[name + "/des_left"]), name
        assert np.array_equal(des_r.cpu().numpy(), z[name + "/des_right"]), name


def test_sharded_keyframe_db_routes_keyframes_and_equals_the_flat_collection():
    """ShardedKeyframeDB with 3 'ranks' on one GPU (rank / world given explicitly): round-robin keyframe routing, keys
    rebased from local to global rows, merged by slm_merge_top2 == the single-GPU KeyframeDB == the oracle."""
    import torch
    world = 3
    frames = [synth.uniform(n, 800 + i) for i, n in enumerate((300, 1, 257, 1000, 64, 0, 511, 90))]
    frames[3][10] = frames[0][5]                      # duplicates across ranks: the lowest GLOBAL row must win
    frames[6][7] = frames[0][5]
    flat = np.concatenate(frames)
    dbs = [slammatch.ShardedKeyframeDB(capacity=128, rank=r, world=world) for r in range(world)]
    single = slammatch.KeyframeDB(capacity=128)
    for f in frames:
        ids = {db.add(f) for db in dbs}
        assert len(ids) == 1
        single.add(f)
    assert sum(db.n_local_rows for db in dbs) == flat.shape[0] and all(db.n_rows == flat.shape[0] for db in dbs)
    assert dbs[0].n_local_rows == 300 + 1000 + 511 and dbs[2].n_local_rows == 257 + 0
    q = np.concatenate([flat[[5, 301, 400, 1600, 2100]] ^ np.uint8(1), synth.uniform(200, 77)])
    qd = torch.from_numpy(q).cuda()
    keys = torch.stack([db.local_keys(qd) for db in dbs])
    nq = q.shape[0]
    idx = torch.empty((nq, 2), dtype=torch.int32, device="cuda")
    dist = torch.empty((nq, 2), dtype=torch.int32, device="cuda")
    acc = torch.empty((nq,), dtype=torch.uint8, device="cuda")
    ctx = slammatch.context(0)
    _lib.check(ctx.lib.slm_merge_top2(ctx.handle, keys.data_ptr(), world, nq, 7, 10, idx.data_ptr(), dist.data_ptr(),
                                      acc.data_ptr(), None))
    torch.cuda.synchronize()
    oi, od = orc.c_knn2(q, flat)
    assert np.array_equal(idx.cpu().numpy(), oi) and np.array_equal(dist.cpu().numpy(), od)
    assert np.array_equal(acc.cpu().numpy(), orc.c_ratio(od, 7, 10))
    si, sd, sa = single.query(q)
    assert np.array_equal(si, oi) and np.array_equal(sd, od)
    kf, loc = dbs[1].locate(oi[:5, 0])
    assert kf.tolist() == [0, 2, 2, 4, 6] and loc[0] == 5
    # world == 1 degenerates to the plain collection
    one = slammatch.ShardedKeyframeDB(capacity=64)
    for f in frames:
        one.add(f)
    i1, d1, a1 = one.query(q)
    assert np.array_equal(i1, oi) and np.array_equal(d1, od) and np.array_equal(a1, orc.c_ratio(od, 7, 10))
