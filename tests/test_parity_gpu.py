"""Parity tests proper: the CUDA path, called through the C-ABI, against the oracle and the committed
golden vectors.  Bit-exact: same best / second-best train indices, same integer distances, lowest train
index on ties, same ratio and cross-check verdicts.  Needs a B200 (``-m gpu``)."""
import ctypes
import os

import numpy as np
import pytest

import slammatch
from slammatch import synth
from oracle import oracle as orc
from conftest import GOLDEN

pytestmark = pytest.mark.gpu

VARIANTS = ["popc", "tensor", "tensor4", "bmma", "auto"]


def _variant_available(v):
    ctx = slammatch.context(0)
    try:
        ctx.set_variant(v)
        q = synth.uniform(4, 1)
        slammatch.knn2(q, q, variant=v)
        return True
    except slammatch.SlamMatchError as e:
        if e.code == -4:
            return False
        raise
    finally:
        ctx.set_variant("auto")


@pytest.fixture(params=VARIANTS)
def variant(request):
    if not _variant_available(request.param):
        pytest.skip(f"variant {request.param} not built")
    yield request.param
    slammatch.context(0).set_variant("auto")


def test_native_library_is_loaded_and_launches_kernels():
    ctx = slammatch.context(0)
    before = ctx.launch_count()
    q = synth.uniform(10, 1)
    slammatch.knn2(q, q)
    assert ctx.launch_count() > before
    with open("/proc/self/maps") as fh:
        assert "libslammatch.so" in fh.read()


def test_golden_bfmatcher_vectors_host_path(knn2_golden, variant):
    for name, q, t, idx, dist, cross in knn2_golden:
        i, d, _ = slammatch.knn2(q, t, ratio=None, variant=variant)
        assert np.array_equal(i, idx), (name, variant)
        assert np.array_equal(d, dist), (name, variant)


def test_golden_cross_check(knn2_golden, variant):
    for name, q, t, idx, dist, cross in knn2_golden:
        i, d, acc = slammatch.knn2(q, t, ratio=None, cross_check=True, variant=variant)
        rows = np.nonzero(acc)[0]
        got = np.stack([rows, i[rows, 0], d[rows, 0]], axis=1).astype(np.int32).reshape(-1, 3)
        assert np.array_equal(got, cross), (name, variant)


@pytest.mark.parametrize("ratio", [(7, 10), (3, 4)])
def test_ratio_verdicts_equal_oracle(ratio, variant):
    for nq, nt, seed in ((1000, 1000, 1001), (257, 4099, 1002), (64, 2, 1003), (33, 1, 1004)):
        q, t = synth.planted(nq, nt, seed)
        i, d, acc = slammatch.knn2(q, t, ratio=ratio, variant=variant)
        oi, od = orc.c_knn2(q, t)
        assert np.array_equal(i, oi) and np.array_equal(d, od)
        want = orc.c_ratio(od, *ratio)
        assert np.array_equal(acc, want)
        if nt >= 1000:
            assert 0.3 * nq < acc.sum() < 0.7 * nq  # planted data: the test is not vacuous


def test_kNN2_plus_cross_check_and_ratio_combined(variant):
    q, t = synth.planted(2000, 20000, 2001)   # BASELINE config 2 shape
    i, d, acc = slammatch.knn2(q, t, ratio=(7, 10), cross_check=True, variant=variant)
    oi, od = orc.c_knn2(q, t)
    want = orc.c_ratio(od, 7, 10) & orc.c_cross_check(q, t, oi)
    assert np.array_equal(i, oi) and np.array_equal(d, od) and np.array_equal(acc, want)


@pytest.mark.parametrize("nq,nt", [(1, 1), (1, 2), (7, 3), (129, 127), (128, 128), (513, 255), (2000, 257),
                                   (5, 70001), (1025, 1023)])
def test_ragged_sizes_against_oracle(nq, nt, variant):
    q = synth.heavy_ties(nq, nq * 7 + nt) if (nq + nt) % 2 else synth.uniform(nq, nq + nt)
    t = synth.heavy_ties(nt, nq + 3 * nt) if (nq + nt) % 2 else synth.uniform(nt, nt + 9)
    i, d, _ = slammatch.knn2(q, t, ratio=None, variant=variant)
    oi, od = orc.c_knn2(q, t)
    assert np.array_equal(i, oi) and np.array_equal(d, od)


def test_empty_inputs(variant):
    e = np.zeros((0, 32), np.uint8)
    q = synth.uniform(5, 3)
    i, d, a = slammatch.knn2(q, e, variant=variant)
    assert (i == -1).all() and (d == -1).all() and (a == 0).all()
    i, d, a = slammatch.knn2(e, q, variant=variant)
    assert i.shape == (0, 2) and a.shape == (0,)


def test_device_fast_path_with_torch_tensors_and_index_base(variant):
    import torch
    q, t = synth.planted(700, 3001, 77)
    qd = torch.from_numpy(q).cuda()
    td = torch.from_numpy(t).cuda().view(torch.int32).view(-1, 8)   # packed uint32x8 view
    i, d, acc = slammatch.knn2(qd, td, ratio=(7, 10), train_index_base=5000, variant=variant)
    torch.cuda.synchronize()
    oi, od = orc.c_knn2(q, t, train_index_base=5000)
    assert np.array_equal(i.cpu().numpy(), oi) and np.array_equal(d.cpu().numpy(), od)
    assert np.array_equal(acc.cpu().numpy(), orc.c_ratio(od, 7, 10))


def test_shard_count_invariance_through_merge(variant):
    """Sharded train set on ONE GPU (shards looped serially): keys -> gather -> slm_merge_top2 gives
    byte-identical output for 1/2/4/8 shards, with exact duplicates across shard boundaries."""
    import torch
    q, t = synth.planted(500, 6000, 91)
    t = synth.with_duplicates(t, 92, 0.4)
    oi, od = orc.c_knn2(q, t)
    ctx = slammatch.context(0)
    ctx.set_variant(variant)
    qd = torch.from_numpy(q).cuda()
    for shards in (1, 2, 4, 8):
        bounds = np.linspace(0, t.shape[0], shards + 1).astype(int)
        keys = torch.empty((shards, 500, 2), dtype=torch.int64, device="cuda")
        for s, (a, b) in enumerate(zip(bounds[:-1], bounds[1:])):
            td = torch.from_numpy(t[a:b]).cuda()
            slammatch._lib.check(ctx.lib.slm_knn2_keys(ctx.handle, qd.data_ptr(), 500, td.data_ptr(), int(b - a),
                                                       int(a), keys[s].data_ptr(), None))
            torch.cuda.synchronize()
        idx = torch.empty((500, 2), dtype=torch.int32, device="cuda")
        dist = torch.empty((500, 2), dtype=torch.int32, device="cuda")
        acc = torch.empty((500,), dtype=torch.uint8, device="cuda")
        slammatch._lib.check(ctx.lib.slm_merge_top2(ctx.handle, keys.data_ptr(), shards, 500, 7, 10, idx.data_ptr(),
                                                    dist.data_ptr(), acc.data_ptr(), None))
        torch.cuda.synchronize()
        assert np.array_equal(idx.cpu().numpy(), oi) and np.array_equal(dist.cpu().numpy(), od), shards
        assert np.array_equal(acc.cpu().numpy(), orc.c_ratio(od, 7, 10)), shards
        # the packed keys themselves equal the oracle's
        ok = np.stack([orc.np_knn2_keys(q, t[a:b], int(a)) for a, b in zip(bounds[:-1], bounds[1:])])
        assert np.array_equal(keys.cpu().numpy().view(np.uint64), ok), shards


@pytest.mark.parametrize("n", [300, 777])
def test_batched_pairs_config3_shape(n, variant):
    """Config 3 in miniature: all unordered frame pairs (i<j) of a keyframe batch in one call."""
    import torch
    if variant == "bmma":
        pytest.skip("batched entry point runs the popc or tensor variant")
    frames = 6
    rng = np.random.default_rng(5)
    base = synth.uniform(n, 50)
    desc = np.stack([base ^ np.packbits(rng.random((n, 256)) < 0.05 * (f + 1), axis=1, bitorder="little")
                     for f in range(frames)])
    pairs = np.array([(i, j) for i in range(frames) for j in range(i + 1, frames)], dtype=np.int32)
    dd = torch.from_numpy(desc).cuda()
    P = pairs.shape[0]
    idx = torch.empty((P, n, 2), dtype=torch.int32, device="cuda")
    dist = torch.empty((P, n, 2), dtype=torch.int32, device="cuda")
    acc = torch.empty((P, n), dtype=torch.uint8, device="cuda")
    ctx = slammatch.context(0)
    ctx.set_variant(variant)
    slammatch._lib.check(ctx.lib.slm_knn2_batched(ctx.handle, dd.data_ptr(), frames, n, pairs.ctypes.data, P, 7, 10,
                                                  idx.data_ptr(), dist.data_ptr(), acc.data_ptr(), None))
    torch.cuda.synchronize()
    for p, (a, b) in enumerate(pairs):
        oi, od = orc.c_knn2(desc[a], desc[b])
        assert np.array_equal(idx[p].cpu().numpy(), oi) and np.array_equal(dist[p].cpu().numpy(), od), p
        assert np.array_equal(acc[p].cpu().numpy(), orc.c_ratio(od, 7, 10)), p
    assert acc.sum().item() > 0


def test_compaction_equals_reference_good_list():
    import torch
    ctx = slammatch.context(0)
    for nq, nt, seed in ((800, 900, 5), (50, 1, 6), (300, 2, 7)):
        q, t = synth.planted(nq, nt, seed)
        i, d, acc = slammatch.knn2(q, t, ratio=(7, 10))
        idd, ddd, add = (torch.from_numpy(x).cuda() for x in (i, d, acc))
        out = torch.full((nq, 3), -7, dtype=torch.int32, device="cuda")
        cnt = torch.zeros(1, dtype=torch.int32, device="cuda")
        slammatch._lib.check(ctx.lib.slm_compact_matches(ctx.handle, idd.data_ptr(), ddd.data_ptr(), add.data_ptr(),
                                                         nq, 1, out.data_ptr(), cnt.data_ptr(), None))
        torch.cuda.synchronize()
        rows = orc.good_rows(i, d)
        m = int(cnt.item())
        assert m == rows.size
        got = out[:m].cpu().numpy()
        assert np.array_equal(got[:, 0], rows) and np.array_equal(got[:, 1], i[rows, 0]) and np.array_equal(got[:, 2], d[rows, 0])


def test_matcher_object_is_a_drop_in_for_the_reference_loop():
    """The reference's own loop (tracking.py:24-33), written out verbatim against slammatch.Matcher, against
    the golden outputs of the real tracking.get_matches / Point3D.find_2D_and_3D_correspondenses."""
    z = np.load(os.path.join(GOLDEN, "reference_functions.npz"))
    for name in z["names"]:
        name = str(name)
        des1, des2 = z[name + "/q"], z[name + "/t"]
        p1, p2, pts3d = z[name + "/p1"], z[name + "/p2"], z[name + "/pts3d"]
        flann = slammatch.Matcher(indexParams=dict(algorithm=6, table_number=6, key_size=12, multi_probe_level=1),
                                  searchParams=dict(checks=50))
        matches = flann.knnMatch(des1, des2, k=2)
        good = []
        try:
            for m, n in matches:
                if m.distance < 0.7 * n.distance:
                    good.append(m)
        except ValueError:
            pass
        q1 = np.float32([p1[m.queryIdx] for m in good])
        q2 = np.float32([p2[m.trainIdx] for m in good])
        assert np.array_equal(q1.reshape(-1), z[name + "/gm_q1"].reshape(-1)), name
        assert np.array_equal(q2.reshape(-1), z[name + "/gm_q2"].reshape(-1)), name
        good3 = [m for m in good if abs(pts3d[m.queryIdx, 0]) < 500 and abs(pts3d[m.queryIdx, 1]) < 500
                 and abs(pts3d[m.queryIdx, 2]) < 500]
        Q1 = np.asarray([pts3d[m.queryIdx] for m in good3])
        assert np.array_equal(Q1.reshape(-1), z[name + "/p3_Q1"].reshape(-1)), name
        # array-level mirror gives the same thing without DMatch objects
        a1, a2 = slammatch.get_matches(p1, des1, p2, des2)
        assert np.array_equal(a1.reshape(-1), z[name + "/gm_q1"].reshape(-1)), name
        assert np.array_equal(a2.reshape(-1), z[name + "/gm_q2"].reshape(-1)), name


def test_matcher_equals_cv2_bfmatcher_object_for_object():
    cv2 = pytest.importorskip("cv2")
    q, t = synth.planted(400, 500, 41)
    want = cv2.BFMatcher(cv2.NORM_HAMMING).knnMatch(q, t, k=2)
    got = slammatch.Matcher().knnMatch(q, t, k=2)
    assert len(want) == len(got)
    for rw, rg in zip(want, got):
        assert [(m.queryIdx, m.trainIdx, m.imgIdx, m.distance) for m in rw] == \
               [(m.queryIdx, m.trainIdx, m.imgIdx, m.distance) for m in rg]
    # multi-image collection: add([...]) + knnMatch(q, k); order (distance, imgIdx, trainIdx)
    parts = [t[:100], t[100:350], t[350:]]
    bf = cv2.BFMatcher(cv2.NORM_HAMMING)
    bf.add(parts)
    mine = slammatch.Matcher()
    mine.add(parts)
    for rw, rg in zip(bf.knnMatch(q, k=2), mine.knnMatch(q, k=2)):
        assert [(m.queryIdx, m.trainIdx, m.imgIdx, m.distance) for m in rw] == \
               [(m.queryIdx, m.trainIdx, m.imgIdx, m.distance) for m in rg]
    # crossCheck=True match()
    w = sorted((m.queryIdx, m.trainIdx, m.distance) for m in cv2.BFMatcher(cv2.NORM_HAMMING, crossCheck=True).match(q, t))
    g = sorted((m.queryIdx, m.trainIdx, m.distance) for m in slammatch.Matcher(crossCheck=True).match(q, t))
    assert w == g


def test_install_seam_runs_opencv_style_caller_unmodified():
    cv2 = pytest.importorskip("cv2")
    q, t = synth.planted(300, 310, 43)
    slammatch.install(cv2)
    try:
        flann = cv2.FlannBasedMatcher(indexParams=dict(algorithm=6, table_number=6, key_size=12, multi_probe_level=1),
                                      searchParams=dict(checks=50))
        rows = flann.knnMatch(q, t, k=2)
    finally:
        slammatch.uninstall(cv2)
    oi, od = orc.c_knn2(q, t)
    assert [r[0].trainIdx for r in rows] == oi[:, 0].tolist()
    assert [int(r[1].distance) for r in rows] == od[:, 1].tolist()


def test_full_size_properties_config5_slice(variant):
    """Large-shape check through size-independent properties (the oracle would take minutes):
    planted neighbours are found at their planted index, distances are self-consistent, and the result is
    invariant under splitting the train set (idempotent merge)."""
    import torch
    nq, nt = 2000, 1_000_000
    rng = np.random.default_rng(123)
    t = synth.uniform(nt, 124)
    src = rng.choice(nt, size=nq, replace=False)
    noise = np.packbits(rng.random((nq, 256)) < 0.04, axis=1, bitorder="little")
    q = t[src] ^ noise
    i, d, acc = slammatch.knn2(q, t, ratio=(7, 10), variant=variant)
    assert np.array_equal(i[:, 0], src.astype(np.int32))
    true_d = np.bitwise_count((q ^ t[src]).view(np.uint64)).sum(axis=1)
    assert np.array_equal(d[:, 0], true_d.astype(np.int32))
    d2 = np.bitwise_count((q ^ t[i[:, 1]]).view(np.uint64)).sum(axis=1)
    assert np.array_equal(d[:, 1], d2.astype(np.int32)) and (d[:, 1] >= d[:, 0]).all()
    assert acc.all()
    # a 64-query slice against the C oracle, full width
    oi, od = orc.c_knn2(q[:64], t)
    assert np.array_equal(i[:64], oi) and np.array_equal(d[:64], od)


def test_host_path_chunked_copy_pipeline():
    """slm_knn2_host cuts train sets of >= 3M rows into 1M-row chunks (H2D on a second stream, every chunk
    searched as it lands, chunk results merged by global index): same bytes as the oracle, with duplicates
    planted across chunk boundaries."""
    nq, nt = 96, 3_300_000
    t = synth.uniform(nt, 501)
    rng = np.random.default_rng(502)
    q = synth.uniform(nq, 503)
    src = rng.choice(nt, size=nq // 2, replace=False)
    q[: nq // 2] = t[src] ^ np.packbits(rng.random((nq // 2, 256)) < 0.05, axis=1, bitorder="little")
    # exact duplicates of a few matched rows in other chunks, at higher AND lower indices
    t[(src[:8] + 1_048_576) % nt] = t[src[:8]]
    t[(src[8:16] + 2_200_000) % nt] = t[src[8:16]]
    i, d, acc = slammatch.knn2(q, t, ratio=(7, 10))
    oi, od = orc.c_knn2(q, t)
    assert np.array_equal(i, oi) and np.array_equal(d, od)
    assert np.array_equal(acc, orc.c_ratio(od, 7, 10))


def test_candidate_epochs_of_resident_clusters():
    """A resident cluster that walks more than 4096 tiles (1M rows) flushes its candidates in epochs.  The
    limit is lowered through SLM_TC_EPOCH_TILES in a fresh process so a 40k-row train set exercises it."""
    import subprocess
    import sys
    code = r"""
import sys, numpy as np
sys.path.insert(0, 'slam-1_b200'); sys.path.insert(0, '.')
import slammatch
from slammatch import synth
from oracle import oracle as orc
q, t = synth.planted(300, 40000, 321)
t = synth.with_duplicates(t, 322, 0.3)
i, d, a = slammatch.knn2(q, t, ratio=(7, 10), variant='tensor')
oi, od = orc.c_knn2(q, t)
assert np.array_equal(i, oi) and np.array_equal(d, od)
assert np.array_equal(a, orc.c_ratio(od, 7, 10))
print('EPOCH_OK')
"""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, SLM_TC_EPOCH_TILES="8", SLM_TC_MAX_CPG="3")
    r = subprocess.run([sys.executable, "-c", code], cwd=root, env=env, capture_output=True, text=True, timeout=300)
    assert "EPOCH_OK" in r.stdout, r.stdout + r.stderr


def test_device_resident_get_matches_equals_reference_outputs():
    """get_matches_device: search + ratio + compaction + gathers on the GPU against the golden outputs of the real
    tracking.get_matches (tracking.py:12-34)."""
    import torch
    z = np.load(os.path.join(GOLDEN, "reference_functions.npz"))
    for name in z["names"]:
        name = str(name)
        des1, des2 = torch.from_numpy(z[name + "/q"]).cuda(), torch.from_numpy(z[name + "/t"]).cuda()
        p1, p2 = torch.from_numpy(z[name + "/p1"]).cuda(), torch.from_numpy(z[name + "/p2"]).cuda()
        q1, q2 = slammatch.get_matches_device(p1, des1, p2, des2)
        assert np.array_equal(q1.cpu().numpy().reshape(-1), z[name + "/gm_q1"].reshape(-1)), name
        assert np.array_equal(q2.cpu().numpy().reshape(-1), z[name + "/gm_q2"].reshape(-1)), name


def test_keyframe_db_is_an_append_only_resident_collection():
    """KeyframeDB == OpenCV's add([...]) + knnMatch(q, k): global row order, (keyframe, local row) mapping,
    growth of the device array, queries from host and from device memory."""
    import torch
    db = slammatch.KeyframeDB(capacity=256)
    frames = [synth.uniform(n, 700 + i) for i, n in enumerate((300, 1, 257, 1000, 64))]
    frames[3][10] = frames[0][5]                      # duplicate descriptor in a later keyframe
    for i, f in enumerate(frames):
        assert db.add(f) == i
    flat = np.concatenate(frames)
    assert len(db) == 5 and db.n_rows == flat.shape[0]
    assert np.array_equal(db.rows().cpu().numpy(), flat)
    q = flat[[5, 301, 400, 1600]] ^ np.uint8(1)
    i, d, a = db.query(q, ratio=(7, 10))
    oi, od = orc.c_knn2(q, flat)
    assert np.array_equal(i, oi) and np.array_equal(d, od) and np.array_equal(a, orc.c_ratio(od, 7, 10))
    kf, loc = db.locate(i[:, 0])
    assert kf.tolist() == [0, 2, 2, 4] and loc[0] == 5          # lowest global row wins the planted duplicate
    i2, d2, a2 = db.query(torch.from_numpy(q).cuda(), ratio=(7, 10))
    assert np.array_equal(i2.cpu().numpy(), oi)


@pytest.mark.parametrize("seed", range(6))
def test_random_shapes_fuzz(seed, variant):
    """Random (nq, nt) pairs around tile / chunk / warp boundaries, mixed generators, every variant."""
    rng = np.random.default_rng(4242 + seed)
    for _ in range(6):
        nq = int(rng.choice([1, 2, 7, 8, 9, 31, 33, 127, 128, 129, 255, 257, 383, 385, 511, 513, 1023, 1500]))
        nt = int(rng.choice([1, 2, 31, 32, 33, 63, 255, 256, 257, 511, 513, 767, 1025, 2047, 4097, 9999, 33000]))
        kind = int(rng.integers(0, 3))
        if kind == 0:
            q, t = synth.uniform(nq, seed * 100 + nq), synth.uniform(nt, seed * 100 + nt + 1)
        elif kind == 1:
            q, t = synth.heavy_ties(nq, seed * 100 + nq), synth.heavy_ties(nt, seed * 100 + nt + 1)
        else:
            q, t = synth.planted(nq, nt, seed * 100 + nq + nt)
            t = synth.with_duplicates(t, seed, 0.3)
        i, d, a = slammatch.knn2(q, t, ratio=(7, 10), cross_check=bool(rng.integers(0, 2)), variant=variant)
        oi, od = orc.c_knn2(q, t)
        assert np.array_equal(i, oi) and np.array_equal(d, od), (variant, nq, nt, kind)
