"""Bag-of-words follow-on: word histogram + chi-square scan + argmin against the
reference's own numpy expressions (bag_of_words.py:23-42), bit-exact including the float64 sums."""
import numpy as np
import pytest

from oracle import oracle as orc
from slammatch import synth


def test_histogram_binning_of_the_reference_is_the_identity():
    # np.histogram(labels, bins=k, range=(0, k-1)) == bincount for integer labels in [0, k)  (CPU only)
    rng = np.random.default_rng(1)
    for k in (2, 50, 64, 1000):
        labels = rng.integers(0, k, 5000)
        assert np.array_equal(orc.np_bow_hist(labels, k), np.bincount(labels, minlength=k))


@pytest.mark.gpu
def test_bow_hist_and_scan_match_reference_arithmetic():
    from slammatch.bow import BoW
    for k, n_img, n_desc, seed in ((50, 40, 100, 1), (64, 130, 300, 2), (300, 25, 500, 3), (1000, 12, 800, 4)):
        vocab = synth.uniform(k, 900 + seed)
        imgs = [synth.planted(n_desc, k, 1000 * seed + i)[0] for i in range(n_img)]
        bow = BoW(vocab, capacity=8)                       # small capacity: exercises the growth path
        bow.train(imgs)
        db = bow.db
        assert db.shape == (n_img, k)
        for i in (0, n_img // 2, n_img - 1):
            words = orc.c_knn2(imgs[i], vocab)[0][:, 0]
            want = orc.np_bow_hist(words, k)
            assert np.array_equal(db[i], want), (k, i)
            assert np.array_equal(bow.hist(imgs[i]), want)
        q = synth.planted(n_desc, k, 77 + seed)[0]
        h = orc.np_bow_hist(orc.c_knn2(q, vocab)[0][:, 0], k)
        for img_index, thr in ((n_img - 1, 0), (n_img - 1, 5), (10, 3), (2, 5)):
            want = orc.np_predict_previous(h, db, img_index, thr)
            got = bow.predict_previous(q, img_index, thr)
            assert got[0] == want[0], (k, img_index, thr)
            assert got[1] == want[1], (k, img_index, thr, got, want)      # bit-exact float64
        i, v = bow.predict(q)
        dist = [orc.np_chi2(h, e) for e in db]
        assert (i, v) == (int(np.argmin(dist)), float(np.min(dist)))
        # an image that is in the db scores exactly 0 against itself and wins with the lowest index
        i, v = bow.predict(imgs[3])
        assert v == 0.0 and i == 3


@pytest.mark.gpu
def test_chi2_scan_first_minimum_wins_and_sum_order_is_numpys():
    import torch
    from slammatch import _lib
    ctx = _lib.context(0)
    rng = np.random.default_rng(9)
    for k in (5, 8, 9, 50, 127, 128, 129, 300, 1000):
        n_db = 700
        db = rng.integers(0, 40, (n_db, k)).astype(np.int32)
        db[500] = db[20]                                    # exact tie: argmin must report 20
        h = db[20].copy(); h[:3] += 1
        want = np.array([orc.np_chi2(h.astype(np.int64), r.astype(np.int64)) for r in db])
        hd, dd = torch.from_numpy(h).cuda(), torch.from_numpy(db).cuda()
        dist = torch.empty(n_db, dtype=torch.float64, device="cuda")
        bi = torch.empty(1, dtype=torch.int32, device="cuda"); bv = torch.empty(1, dtype=torch.float64, device="cuda")
        _lib.check(ctx.lib.slm_chi2_scan(ctx.handle, hd.data_ptr(), dd.data_ptr(), n_db, k, dist.data_ptr(),
                                         bi.data_ptr(), bv.data_ptr(), None))
        torch.cuda.synchronize()
        assert np.array_equal(dist.cpu().numpy(), want), k            # every distance bit-exact
        assert int(bi.item()) == int(np.argmin(want)) == 20 and float(bv.item()) == float(np.min(want))


@pytest.mark.gpu
def test_chi2_scan_wide_vocabularies_are_bit_exact_too():
    """Above 128 words (config 4's vocabulary has 65 536) the scan takes one block per stored histogram: the leaves of
    numpy's pairwise-sum tree by 8-lane groups, then the tree itself -- still every distance bit-identical with np.sum; k
    changes between calls (the cached leaf table is rebuilt), odd sizes, and the sizes between 8192 and 12288 words at which
    the first version of the scan (one thread per histogram, recursive sum) outgrew the per-thread stack."""
    import torch
    from slammatch import _lib
    ctx = _lib.context(0)
    rng = np.random.default_rng(10)
    for k, n_db in ((1024, 300), (1025, 50), (4099, 30), (12288, 30), (12289, 40), (65536, 60), (65536, 7), (100003, 12),
                    (1024, 5), (500000, 3)):
        db = rng.integers(0, 6, (n_db, k)).astype(np.int32)
        db[n_db - 1] = db[1]                                   # exact tie: argmin must report 1
        h = db[1].copy(); h[::977] += 2
        want = np.array([orc.np_chi2(h.astype(np.int64), r.astype(np.int64)) for r in db])
        hd, dd = torch.from_numpy(h).cuda(), torch.from_numpy(db).cuda()
        dist = torch.empty(n_db, dtype=torch.float64, device="cuda")
        bi = torch.empty(1, dtype=torch.int32, device="cuda"); bv = torch.empty(1, dtype=torch.float64, device="cuda")
        _lib.check(ctx.lib.slm_chi2_scan(ctx.handle, hd.data_ptr(), dd.data_ptr(), n_db, k, dist.data_ptr(),
                                         bi.data_ptr(), bv.data_ptr(), None))
        torch.cuda.synchronize()
        assert np.array_equal(dist.cpu().numpy(), want), k
        assert int(bi.item()) == int(np.argmin(want)) == 1 and float(bv.item()) == float(np.min(want))
    assert ctx.lib.slm_chi2_scan(ctx.handle, hd.data_ptr(), dd.data_ptr(), 1, (1 << 19) + 1, dist.data_ptr(), bi.data_ptr(),
                                 bv.data_ptr(), None) == -1


@pytest.mark.gpu
def test_chi2_term_count_ranges_zero_one_and_beyond_16_bits():
    """The per-word term takes a short path for the counts a real histogram holds (zero numerator, denominator 0 or 1, counts
    below 32768) and 64-bit arithmetic beyond; both must give numpy's int64 / float64 value bit for bit, on every scan kernel
    (narrow: k <= 128, wide: one block per histogram)."""
    import torch
    from slammatch import _lib
    ctx = _lib.context(0)
    rng = np.random.default_rng(11)
    pool = np.array([0, 0, 0, 1, 1, 2, 3, 7, 255, 32766, 32767, 32768, 32769, 40000, 65535, 65536, 100000, 1 << 20, 1 << 30],
                    dtype=np.int32)
    for k, n_db in ((7, 400), (128, 300), (129, 200), (1000, 64), (4099, 20)):
        db = pool[rng.integers(0, len(pool), (n_db, k))]
        h = pool[rng.integers(0, len(pool), k)]
        db[n_db - 1] = db[2] = h                                # exact tie at distance 0: argmin must report 2
        want = np.array([orc.np_chi2(h.astype(np.int64), r.astype(np.int64)) for r in db])
        hd, dd = torch.from_numpy(h).cuda(), torch.from_numpy(db).cuda()
        dist = torch.empty(n_db, dtype=torch.float64, device="cuda")
        bi = torch.empty(1, dtype=torch.int32, device="cuda"); bv = torch.empty(1, dtype=torch.float64, device="cuda")
        _lib.check(ctx.lib.slm_chi2_scan(ctx.handle, hd.data_ptr(), dd.data_ptr(), n_db, k, dist.data_ptr(),
                                         bi.data_ptr(), bv.data_ptr(), None))
        torch.cuda.synchronize()
        assert np.array_equal(dist.cpu().numpy(), want), k
        assert int(bi.item()) == 2 and float(bv.item()) == 0.0


@pytest.mark.gpu
def test_bow_predict_with_a_64k_word_vocabulary():
    """BoW.predict / predict_previous at config 4's vocabulary size (was rejected: > 12288 words)."""
    from slammatch.bow import BoW
    k, n_img, n_desc = 65536, 9, 2000
    vocab = synth.uniform(k, 31)
    imgs = [synth.planted(n_desc, k, 500 + i)[0] for i in range(n_img)]
    bow = BoW(vocab, capacity=4)
    bow.train(imgs)
    db = bow.db
    q = synth.planted(n_desc, k, 600)[0]
    h = orc.np_bow_hist(orc.c_knn2(q, vocab)[0][:, 0], k)
    assert np.array_equal(bow.hist(q), h)
    want = orc.np_predict_previous(h, db, n_img - 1, 2)
    assert bow.predict_previous(q, n_img - 1, 2) == want
    i, v = bow.predict(imgs[4])
    assert (i, v) == (4, 0.0)


def test_vocab_update_rule_majority_ties_and_empty_words():
    """The k-majority update rule of the oracle itself on a hand-made case (CPU only)."""
    desc = np.zeros((5, 32), np.uint8)
    desc[0, 0] = 0b0000_0111
    desc[1, 0] = 0b0000_0011
    desc[2, 0] = 0b0000_0001          # word 0 members: bit0 3/3, bit1 2/3, bit2 1/3
    desc[3, 1] = 0xFF
    desc[4, 1] = 0x00                 # word 1 members: every bit of byte 1 tied 1/2
    vocab = np.zeros((3, 32), np.uint8)
    vocab[1, 1] = 0b1010_1010         # ties keep these bits
    vocab[2, 5] = 0x5A                # no members: stays
    new, counts, changed = orc.np_vocab_update(desc, np.array([0, 0, 0, 1, 1]), vocab)
    assert counts.tolist() == [3, 2, 0]
    assert new[0, 0] == 0b0000_0011 and new[1, 1] == 0b1010_1010 and new[2, 5] == 0x5A
    assert changed == 1


@pytest.mark.gpu
def test_vocab_update_and_training_equal_the_oracle():
    """slm_vocab_update against np_vocab_update (random assignments incl. empty and tied words), then whole training
    runs: same initial words, same iterations, bit-identical vocabulary."""
    import torch
    import slammatch
    from slammatch import _lib
    from slammatch.bow import train_vocabulary, BoW
    ctx = slammatch.context(0)
    rng = np.random.default_rng(11)
    for n, k in ((1, 1), (2, 1), (500, 7), (4000, 50), (3000, 1500), (20000, 300)):
        desc = synth.uniform(n, 50 + n)
        vocab = synth.uniform(k, 60 + k)
        words = rng.integers(-1, k + 1, size=(n, 2)).astype(np.int32)     # -1 and k are out of range: skipped
        words[: n // 3, 0] = rng.integers(0, max(1, k // 4), size=n // 3)  # some crowded words
        want, want_counts, want_changed = orc.np_vocab_update(desc, words[:, 0], vocab)
        dd, wd, vd = torch.from_numpy(desc).cuda(), torch.from_numpy(words).cuda(), torch.from_numpy(vocab).cuda()
        counts = torch.full((k,), -1, dtype=torch.int32, device="cuda")
        changed = torch.full((1,), -1, dtype=torch.int32, device="cuda")
        _lib.check(ctx.lib.slm_vocab_update(ctx.handle, dd.data_ptr(), n, wd.data_ptr(), 2, vd.data_ptr(), k,
                                            counts.data_ptr(), changed.data_ptr(), None))
        torch.cuda.synchronize()
        assert np.array_equal(counts.cpu().numpy(), want_counts), (n, k)
        assert np.array_equal(vd.cpu().numpy(), want), (n, k)
        assert int(changed.item()) == want_changed, (n, k)
    # clustered data: 40 prototypes x noisy copies; the trained words must equal the oracle's, bit for bit
    protos = synth.uniform(40, 5)
    noise = np.packbits(rng.random((6000, 256)) < 0.08, axis=1, bitorder="little")
    pool = protos[rng.integers(0, 40, 6000)] ^ noise
    for k, iters in ((40, 8), (64, 5), (9, 6)):
        want, want_it = orc.np_train_vocabulary(pool, k, iters=iters, seed=3, knn=orc.c_knn2)
        got, got_it = train_vocabulary(pool, k, iters=iters, seed=3)
        assert got_it == want_it and np.array_equal(got, want), (k, iters)
    # the reference's BoW(n_clusters).train(...) flow from per-image descriptors
    imgs = [pool[i * 100:(i + 1) * 100] for i in range(20)]
    bow = BoW.fit(imgs, n_clusters=50, iters=4, seed=1)
    vocab50, _ = orc.np_train_vocabulary(np.concatenate(imgs), 50, iters=4, seed=1, knn=orc.c_knn2)
    for i in (0, 7, 19):
        assert np.array_equal(bow.db[i], orc.np_bow_hist(orc.c_knn2(imgs[i], vocab50)[0][:, 0], 50))
