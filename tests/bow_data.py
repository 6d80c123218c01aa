"""Seeded synthetic DBoW2-style vocabularies and BoW match scenes (no reference data needed)."""
from __future__ import annotations

import struct

import numpy as np

import synth


def make_vocab(seed, k=10, L=3, early_leaf_frac=0.05, zero_weight_frac=0.1, min_children=2):
    """A random vocabulary tree in the record layout of the .dbow2 binary file: arrays over nodes 1..N-1 (BFS order, so
    parent < id and the children of a node are contiguous, like the shipped orb_vocab.dbow2).  Children descriptors are
    their parent's with random bit flips (a descent is meaningful); some leaves sit above level L, some words have
    weight 0 (stop words), some nodes have fewer than k children."""
    rng = np.random.default_rng(seed)
    parent, desc, weight, leaf, level = [], [], [], [], []
    frontier = [(0, np.zeros(32, np.uint8), 0)]   # (id, desc, level)
    next_id = 1
    while frontier:
        new_frontier = []
        for pid, pdesc, plevel in frontier:
            nc = k if rng.random() < 0.7 else int(rng.integers(min_children, k + 1))
            for _ in range(nc):
                d = synth.flip_bits(rng, pdesc[None, :] if plevel else synth.rand_desc(rng, 1), 40 if plevel else 0)[0]
                lvl = plevel + 1
                is_leaf = lvl == L or (lvl >= 2 and rng.random() < early_leaf_frac)
                parent.append(pid)
                desc.append(d)
                leaf.append(1 if is_leaf else 0)
                weight.append(0.0 if (not is_leaf or rng.random() < zero_weight_frac) else float(rng.uniform(0.1, 9.7)))
                level.append(lvl)
                if not is_leaf:
                    new_frontier.append((next_id, d, lvl))
                next_id += 1
        frontier = new_frontier
    return dict(k=k, L=L, parent=np.array(parent, np.int32), desc=np.array(desc, np.uint8).reshape(-1, 32),
                weight=np.array(weight, np.float32), is_leaf=np.array(leaf, np.uint8), level=np.array(level, np.int32))


def write_dbow2(path, vocab, scoring=0, weighting=0):
    """The binary layout bow_vocab_->loadFromBinaryFile reads (system.cc:82)."""
    n = len(vocab["parent"])
    with open(path, "wb") as f:
        f.write(struct.pack("<IIiiii", n + 1, 41, vocab["k"], vocab["L"], scoring, weighting))
        rec = np.zeros((n, 41), np.uint8)
        rec[:, 0:4] = vocab["parent"].astype("<i4").view(np.uint8).reshape(n, 4)
        rec[:, 4:36] = vocab["desc"]
        rec[:, 36:40] = vocab["weight"].astype("<f4").view(np.uint8).reshape(n, 4)
        rec[:, 40] = vocab["is_leaf"]
        f.write(rec.tobytes())


def transform_numpy(vocab, desc, levelsup):
    """Independent restatement of TemplatedVocabulary::transform(feature, ...) on the record arrays."""
    parent = vocab["parent"]
    n_nodes = len(parent) + 1
    children = [[] for _ in range(n_nodes)]
    for i, p in enumerate(parent):
        children[p].append(i + 1)
    bits = np.unpackbits(vocab["desc"], axis=1)
    word_of = -np.ones(n_nodes, np.int64)
    word_of[1:][vocab["is_leaf"] > 0] = np.arange(int(vocab["is_leaf"].sum()))
    nid_level = vocab["L"] - levelsup
    out = []
    for d in np.asarray(desc, np.uint8).reshape(-1, 32):
        db = np.unpackbits(d)
        cur, lvl, nid = 0, 0, 0
        while children[cur]:
            lvl += 1
            ch = children[cur]
            dist = [(int((bits[c - 1] != db).sum()), j) for j, c in enumerate(ch)]
            cur = ch[min(dist)[1]]
            if lvl == nid_level:
                nid = cur
        out.append((int(word_of[cur]), nid, float(vocab["weight"][cur - 1])))
    return out


def make_bow_sides(seed, n1=900, n2=1000, num_nodes=90, stray=0.1):
    """Two keypoint sets whose descriptors are noisy copies of each other, bucketed into `num_nodes` shared vocabulary
    nodes (plus nodes only one side has).  Mimics what transform(levelsup = 4) yields on 1000 ORB features (~100 nodes)."""
    rng = np.random.default_rng(seed)
    n_common = min(n1, n2) * 3 // 4
    base = synth.rand_desc(rng, n_common)
    d1 = np.concatenate([synth.flip_bits(rng, base, rng.integers(0, 30, n_common)), synth.rand_desc(rng, n1 - n_common)])
    d2 = np.concatenate([synth.flip_bits(rng, base, rng.integers(0, 45, n_common)), synth.rand_desc(rng, n2 - n_common)])
    # near-duplicates on side 2 (ratio test / ties) -- copy some side-2 descriptors with 0..2 bit flips
    dup = rng.choice(n_common, n_common // 6, replace=False)
    tgt = rng.choice(np.arange(n_common, n2), min(len(dup), n2 - n_common), replace=False)
    d2[tgt] = synth.flip_bits(rng, d2[dup[:len(tgt)]], rng.integers(0, 3, len(tgt)))
    node_common = rng.integers(0, num_nodes, n_common)
    nid1 = np.concatenate([node_common, rng.integers(0, num_nodes + 20, n1 - n_common)])
    nid2 = np.concatenate([node_common, rng.integers(0, num_nodes + 20, n2 - n_common)])
    nid2[tgt] = node_common[dup[:len(tgt)]]
    move = rng.random(n_common) < stray            # a match whose two ends fell into different nodes: never found
    nid2[:n_common][move] = rng.integers(0, num_nodes, move.sum())
    ids = np.sort(rng.choice(200000, num_nodes + 20, replace=False)).astype(np.uint32)
    p1, p2 = rng.permutation(n1), rng.permutation(n2)
    inv2 = np.argsort(p2)
    a1 = rng.uniform(0, 360, n1).astype(np.float32)
    a2 = rng.uniform(0, 360, n2).astype(np.float32)
    rot = np.float32(rng.uniform(0, 360))
    a2[:n_common] = (a1[:n_common] - rot + rng.normal(0, 5, n_common).astype(np.float32)) % np.float32(360)
    wild = rng.random(n_common) < 0.1
    a2[:n_common][wild] = rng.uniform(0, 360, wild.sum()).astype(np.float32)

    def side(d, nid, perm, ang, valid_frac):
        d, nid, ang = d[perm], nid[perm], ang[perm]
        order = rng.permutation(len(d))            # list order inside a node is arbitrary but fixed
        nodes = sorted(set(nid.tolist()))
        offsets, indices = [0], []
        for nd in nodes:
            idx = [int(i) for i in order if nid[i] == nd]
            indices += idx
            offsets.append(len(indices))
        fv = (ids[np.array(nodes)], np.array(offsets, np.int32), np.array(indices, np.uint32))
        return dict(desc=d, angle=ang, valid=(rng.random(len(d)) < valid_frac).astype(np.uint8), fv=fv)
    s1 = side(d1, nid1, p1, a1, 0.85)
    s2 = side(d2, nid2, p2, a2, 0.9)
    truth = np.full(n1, -1)
    truth[np.argsort(p1)[:n_common]] = inv2[:n_common]
    return s1, s2, truth


def bow_tree_match_python(side1, side2, lowe_ratio, check_orientation=True):
    """Line-by-line Python restatement of match/bow_tree.cc:41-165 / :167-305 (independent of the C++ oracle)."""
    n1, n2 = len(side1["desc"]), len(side2["desc"])
    b1, b2 = np.unpackbits(side1["desc"], axis=1), np.unpackbits(side2["desc"], axis=1)
    v1, v2 = side1.get("valid"), side2.get("valid")
    m21, m12 = np.full(n1, -1, np.int64), np.full(n2, -1, np.int64)
    f1, f2 = side1["fv"], side2["fv"]
    hist = [[] for _ in range(30)]
    a = b = 0
    while a < len(f1[0]) and b < len(f2[0]):
        if f1[0][a] == f2[0][b]:
            for i1 in f1[2][f1[1][a]:f1[1][a + 1]]:
                if v1 is not None and not v1[i1]:
                    continue
                best, second, best_j = 256, 256, -1
                for j in f2[2][f2[1][b]:f2[1][b + 1]]:
                    if v2 is not None and not v2[j]:
                        continue
                    if m12[j] >= 0:
                        continue
                    d = int((b1[i1] != b2[j]).sum())
                    if d < best:
                        second, best, best_j = best, d, int(j)
                    elif d < second:
                        second = d
                if 50 < best:
                    continue
                if np.float32(lowe_ratio) * np.float32(second) < np.float32(best):
                    continue
                m21[i1], m12[best_j] = best_j, i1
                if check_orientation:
                    delta = np.float32(side1["angle"][i1]) - np.float32(side2["angle"][best_j])
                    if delta < 0.0:
                        delta = np.float32(np.float64(delta) + 360.0)
                    if 360.0 <= delta:
                        delta = np.float32(np.float64(delta) - 360.0)
                    hist[int(np.rint(np.float32(delta * np.float32(1.0 / 30))))].append(int(i1))
            a += 1
            b += 1
        elif f1[0][a] < f2[0][b]:
            a += 1
        else:
            b += 1
    if check_orientation:
        order = sorted(range(30), key=lambda h: (-len(hist[h]), h))
        for h in order[3:]:
            for i1 in hist[h]:
                m12[m21[i1]] = -1
                m21[i1] = -1
    return m21, m12, int((m21 >= 0).sum())
