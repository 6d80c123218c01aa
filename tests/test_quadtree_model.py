"""The data-parallel formulation of the keypoint quadtree (tools/quadtree_parallel_model.py, transcribed into
csrc/orb.cu) must equal the sequential std::list restatement in the oracle."""
import sys
from pathlib import Path

import numpy as np
import pytest

import oracle_api
import synth

sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tools"))
import quadtree_parallel_model as qm  # noqa: E402


def _check(orc, xs, ys, resp, min_x, max_x, min_y, max_y, budget):
    p = oracle_api.orb_params()
    cands = np.zeros(len(xs), oracle_api.KP_DTYPE)
    cands["x"], cands["y"], cands["response"] = xs, ys, resp
    ref = orc.orb_distribute(p, cands, min_x, max_x, min_y, max_y, budget)
    idx = qm.distribute(xs, ys, resp, min_x, max_x, min_y, max_y, budget)
    assert len(idx) == len(ref)
    assert np.array_equal(xs[idx], ref["x"]) and np.array_equal(ys[idx], ref["y"])
    assert np.array_equal(resp[idx], ref["response"])


@pytest.mark.parametrize("seed", range(25))
def test_random_candidates(orc, seed):
    rng = np.random.default_rng(seed)
    w, h = [(602, 442), (714, 442), (141, 96), (300, 700), (495, 362)][seed % 5]
    n = int(rng.integers(1, 3000))
    budget = int(rng.choice([5, 60, 217, 1000]))
    xs = rng.integers(0, w, n).astype(np.float32)
    ys = rng.integers(0, h, n).astype(np.float32)
    if seed % 4 == 0:  # heavy duplication -> many equal-count leaves, exercises the tie-break
        xs = (xs // 16 * 16).astype(np.float32)
        ys = (ys // 16 * 16).astype(np.float32)
    resp = rng.integers(7, 255, n).astype(np.float32)
    _check(orc, xs, ys, resp, 19, 19 + w, 19, 19 + h, budget)


def test_real_fast_candidates(orc):
    img = synth.make_texture(99)
    p = oracle_api.orb_params()
    r = orc.orb_extract(p, img, debug=True)
    w, h = orc.orb_level_sizes(p, *img.shape)
    t = orc.orb_tables(p)
    off = 0
    for l in range(8):
        c = r["cands"][off: off + r["cands_per_level"][l]]
        off += r["cands_per_level"][l]
        _check(orc, c["x"].copy(), c["y"].copy(), c["response"].copy(), 19, int(w[l]) - 19, 19, int(h[l]) - 19,
               int(t["num_keypts_per_level"][l]))
