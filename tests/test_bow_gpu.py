"""GPU parity: DBoW2 transform on the device-resident vocabulary and match::bow_tree through the C ABI vs the oracle."""
import numpy as np
import pytest

import bow_data
import synth

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("k,L,levelsup", [(10, 3, 1), (4, 5, 4), (20, 2, 4), (10, 4, 2), (32, 2, 1), (3, 6, 4)])
def test_bow_transform(ctx, orc, plp, tmp_path, k, L, levelsup):
    vocab = bow_data.make_vocab(k * 100 + L, k=k, L=L)
    rng = np.random.default_rng(L + 1)
    leaves = vocab["desc"][vocab["is_leaf"] > 0]
    n = 1221   # a real frame: not a multiple of the group size
    desc = np.concatenate([synth.rand_desc(rng, 300),
                           synth.flip_bits(rng, leaves[rng.integers(0, len(leaves), n - 300)], rng.integers(0, 30, n - 300))])
    ov = orc.bow_vocab_create(k, L, vocab["parent"], vocab["desc"], vocab["weight"], vocab["is_leaf"])
    want = orc.bow_transform(ov, desc, levelsup)
    gv = plp.BowVocabulary(ctx, k=k, L=L, parent=vocab["parent"], desc=vocab["desc"], weight=vocab["weight"],
                           is_leaf=vocab["is_leaf"])
    assert gv.info() == orc.bow_vocab_info(ov)
    got = gv.transform(desc, levelsup)
    for a, b in zip(got, want):
        assert np.array_equal(a, b)
    assert (want[2] == 0).any() and (want[2] > 0).any()      # stop words and real words both occur
    # the .dbow2 loader gives the same device vocabulary
    path = tmp_path / "v.dbow2"
    bow_data.write_dbow2(path, vocab)
    gv2 = plp.BowVocabulary(ctx, path=path)
    assert gv2.info() == gv.info()
    for a, b in zip(gv2.transform(desc[:77], levelsup), want):
        assert np.array_equal(a, b[:77])
    # the folded maps (bow_vec_, bow_feat_vec_) agree as well
    fg, fo = plp.capi.fold_bow(*got), plp.capi.fold_bow(*want)
    assert np.array_equal(fg[0], fo[0]) and np.array_equal(fg[1], fo[1])
    assert all(np.array_equal(x, y) for x, y in zip(fg[2], fo[2]))
    gv.close()
    gv2.close()
    orc.bow_vocab_destroy(ov)


def test_bow_vocab_rejects_inconsistent_trees(ctx, plp, tmp_path):
    vocab = bow_data.make_vocab(1, k=4, L=3)
    bad = dict(vocab)
    bad["is_leaf"] = vocab["is_leaf"].copy()
    bad["is_leaf"][0] ^= 1
    with pytest.raises(plp.PlpError):
        plp.BowVocabulary(ctx, k=4, L=3, parent=bad["parent"], desc=bad["desc"], weight=bad["weight"], is_leaf=bad["is_leaf"])
    bad = dict(vocab)
    bad["parent"] = vocab["parent"].copy()
    bad["parent"][5] = 9
    with pytest.raises(plp.PlpError):
        plp.BowVocabulary(ctx, k=4, L=3, parent=bad["parent"], desc=bad["desc"], weight=bad["weight"], is_leaf=bad["is_leaf"])
    p = tmp_path / "trunc.dbow2"
    bow_data.write_dbow2(p, vocab)
    p.write_bytes(p.read_bytes()[:-10])
    with pytest.raises(plp.PlpError):
        plp.BowVocabulary(ctx, path=p)
    with pytest.raises(plp.PlpError):
        plp.BowVocabulary(ctx, path=tmp_path / "missing.dbow2")


@pytest.mark.parametrize("seed", range(5))
def test_match_bow_tree_pairs(ctx, orc, seed):
    s1, s2, _ = bow_data.make_bow_sides(seed, n1=900 + 40 * seed, n2=1000, num_nodes=90)
    frame = dict(s2)
    frame.pop("valid")
    # frame_tracker.cc:130 (0.7), relocalizer (bow_match_lowe_ratio 0.75), loop_detector.cc:343 (0.75)
    for ratio, check, b in [(0.7, True, frame), (0.75, True, s2), (0.9, False, s2), (0.75, False, frame)]:
        want = orc.bow_tree_match(s1, b, ratio, check)
        got = ctx.match_bow_tree([(s1, b)], ratio, check)[0]
        assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1]) and got[2] == want[2]
        assert want[2] > 100


def test_match_bow_tree_batch_one_frame_many_keyframes(ctx, orc):
    """relocalizer.cc:79: the current frame against every candidate keyframe, one call."""
    frame = None
    pairs = []
    for seed in range(12):
        s1, s2, _ = bow_data.make_bow_sides(100 + seed, n1=700 + 25 * seed, n2=1000, num_nodes=80)
        if frame is None:
            frame = dict(s2)
            frame.pop("valid")
        pairs.append((s1, frame))       # the SAME frame object on side 2 of every pair
    got = ctx.match_bow_tree(pairs, 0.75, True)
    for (s1, f), g in zip(pairs, got):
        want = orc.bow_tree_match(s1, f, 0.75, True)
        assert np.array_equal(g[0], want[0]) and np.array_equal(g[1], want[1]) and g[2] == want[2]
    # loop_detector.cc:356: the current keyframe on side 1 of every pair
    cur = pairs[0][0]
    kk = [(cur, dict(bow_data.make_bow_sides(200 + s, n1=len(cur["desc"]), n2=900, num_nodes=80)[1])) for s in range(5)]
    got = ctx.match_bow_tree(kk, 0.75, True)
    for (a, b), g in zip(kk, got):
        want = orc.bow_tree_match(a, b, 0.75, True)
        assert np.array_equal(g[0], want[0]) and np.array_equal(g[1], want[1]) and g[2] == want[2]


def test_match_bow_tree_edge_cases(ctx, orc, plp):
    s1, s2, _ = bow_data.make_bow_sides(3, n1=200, n2=220, num_nodes=20)
    empty_fv = (np.zeros(0, np.uint32), np.zeros(1, np.int32), np.zeros(0, np.uint32))
    a = dict(s1, fv=empty_fv)
    got = ctx.match_bow_tree([(a, s2), (s1, s2)], 0.75, True)
    assert got[0][2] == 0 and np.all(got[0][0] == -1) and np.all(got[0][1] == -1)
    want = orc.bow_tree_match(s1, s2, 0.75, True)
    assert np.array_equal(got[1][0], want[0]) and got[1][2] == want[2]
    none = dict(desc=np.zeros((0, 32), np.uint8), angle=np.zeros(0, np.float32), fv=empty_fv)
    got = ctx.match_bow_tree([(none, s2), (s1, none)], 0.75, True)
    assert got[0][2] == 0 and got[1][2] == 0 and np.all(got[1][0] == -1)
    # all side-1 keypoints invalid -> nothing matches
    dead = dict(s1, valid=np.zeros(len(s1["desc"]), np.uint8))
    assert ctx.match_bow_tree([(dead, s2)], 0.75, True)[0][2] == 0
    # a keypoint listed in two nodes is rejected (the node-parallel formulation relies on it)
    ids, off, idx = s1["fv"]
    dup = idx.copy()
    dup[-1] = dup[0]
    with pytest.raises(plp.PlpError):
        ctx.match_bow_tree([(dict(s1, fv=(ids, off, dup)), s2)], 0.75, True)
