"""Array ("data-parallel") formulation of orb_extractor::distribute_keypoints_via_tree.

The reference (feature/orb_extractor.cc:468-685) manipulates a std::list of nodes sequentially.  The CUDA
kernel in csrc/orb.cu uses the equivalent array formulation below: every sweep divides all non-leaf nodes at
once, children positions in the new list follow from suffix sums, keypoints are stably 4-way partitioned
inside their node's segment.  This file is the executable specification of that formulation; it is checked
against the sequential oracle by tests/test_quadtree_model.py and then transcribed 1:1 into the kernel.
"""
from __future__ import annotations

import math

import numpy as np


def cv_ceil(v: float) -> int:
    i = int(v)
    return i + (1 if i < v else 0)


class Nodes:
    """Nodes in list order (struct of arrays)."""

    def __init__(self, n=0):
        self.bx = np.zeros(n, np.int64)
        self.by = np.zeros(n, np.int64)
        self.ex = np.zeros(n, np.int64)
        self.ey = np.zeros(n, np.int64)
        self.start = np.zeros(n, np.int64)
        self.cnt = np.zeros(n, np.int64)
        self.leaf = np.zeros(n, bool)

    def __len__(self):
        return len(self.bx)


def _classify(nodes: Nodes, p: int, x: float, y: float) -> int:
    half_x = cv_ceil((nodes.ex[p] - nodes.bx[p]) / 2.0)
    half_y = cv_ceil((nodes.ey[p] - nodes.by[p]) / 2.0)
    q = 0
    if np.float32(nodes.bx[p] + half_x) <= np.float32(x):
        q += 1
    if np.float32(nodes.by[p] + half_y) <= np.float32(y):
        q += 2
    return q


def _child_rect(nodes: Nodes, p: int, q: int):
    bx, by, ex, ey = nodes.bx[p], nodes.by[p], nodes.ex[p], nodes.ey[p]
    half_x = cv_ceil((ex - bx) / 2.0)
    half_y = cv_ceil((ey - by) / 2.0)
    if q == 0:
        return bx, by, bx + half_x, by + half_y
    if q == 1:
        return bx + half_x, by, ex, by + half_y
    if q == 2:
        return bx, by + half_y, bx + half_x, ey
    return bx + half_x, by + half_y, ex, ey


def _divide(nodes: Nodes, perm: np.ndarray, xs, ys, sel: np.ndarray):
    """Class totals [len(nodes),4] and the stably partitioned permutation for the selected nodes."""
    n = len(nodes)
    tot = np.zeros((n, 4), np.int64)
    new_perm = perm.copy()
    cls = {}
    for p in range(n):
        if not sel[p]:
            continue
        seg = perm[nodes.start[p]: nodes.start[p] + nodes.cnt[p]]
        q = np.array([_classify(nodes, p, xs[i], ys[i]) for i in seg], np.int64)
        for c in range(4):
            tot[p, c] = int((q == c).sum())
        cls[p] = q
    return tot, cls


def distribute(xs, ys, resp, min_x, max_x, min_y, max_y, num_keypts):
    """Returns the indices (into the candidate arrays) of the selected keypoints, in output order."""
    n = len(xs)
    if n == 0:
        return np.zeros(0, np.int64)
    # ---- initialize_nodes (orb_extractor.cc:557-637)
    ratio = float(max_x - min_x) / (max_y - min_y)
    if ratio > 1:
        gx, gy = int(round(ratio)), 1
        dx, dy = float(max_x - min_x) / gx, float(max_y - min_y)
    else:
        gx, gy = 1, int(round(1 / ratio))
        dx, dy = float(max_x - min_y), float(max_y - min_y) / gy
    g = gx * gy
    node_of = np.array([int(xs[i] / dx) + int(ys[i] / dy) * gx for i in range(n)], np.int64)
    perm = np.argsort(node_of, kind="stable")
    counts = np.bincount(node_of, minlength=g)
    nodes = Nodes(0)
    lst = []
    off = 0
    for i in range(g):
        if counts[i] > 0:
            ix, iy = i % gx, i // gx
            lst.append((int(dx * ix), int(dy * iy), int(dx * (ix + 1)), int(dy * (iy + 1)), off, counts[i], counts[i] == 1))
        off += counts[i]
    nodes = _from_list(lst)

    pool = np.zeros(0, np.int64)  # list positions, in creation order
    filled = False
    # ---- phase 1 (orb_extractor.cc:482-518): whole-list sweeps
    while True:
        prev = len(nodes)
        sel = ~nodes.leaf
        tot, cls = _divide(nodes, perm, xs, ys, sel)
        nch = np.where(sel, (tot > 0).sum(1), 1)
        proc = np.nonzero(sel)[0]
        keep = np.nonzero(~sel)[0]
        total_children = int(nch[proc].sum())
        new = []
        new_pos_of_child = {}
        # suffix sums over processed nodes: children of later nodes come first
        suffix = 0
        offsets = {}
        for p in proc[::-1]:
            offsets[p] = suffix
            suffix += nch[p]
        new_nodes = [None] * (total_children + len(keep))
        new_perm = perm.copy()
        pool_list = []
        for p in proc:  # processing order == creation order
            nonempty = [c for c in range(4) if tot[p, c] > 0]
            seg_start = nodes.start[p]
            seg = perm[seg_start: seg_start + nodes.cnt[p]]
            q = cls[p]
            o = seg_start
            for r, c in enumerate(nonempty):
                pos = offsets[p] + (nch[p] - 1 - r)
                members = seg[q == c]
                new_perm[o: o + len(members)] = members
                new_nodes[pos] = (*_child_rect(nodes, p, c), o, len(members), False)
                if len(members) > 1:
                    pool_list.append(pos)
                o += len(members)
        for r, p in enumerate(keep):
            new_nodes[total_children + r] = (nodes.bx[p], nodes.by[p], nodes.ex[p], nodes.ey[p], nodes.start[p],
                                             nodes.cnt[p], True)
        nodes = _from_list(new_nodes)
        perm = new_perm
        pool = np.array(pool_list, np.int64)
        if num_keypts <= len(nodes) or len(nodes) == prev:
            filled = True
            break
        if num_keypts < len(nodes) + len(pool):
            break
    # ---- phase 2 (orb_extractor.cc:520-552): densest leaves first
    while not filled:
        prev = len(nodes)
        # sort pool by (cnt desc, creation desc); pool index == creation rank
        order = sorted(range(len(pool)), key=lambda k: (-nodes.cnt[pool[k]], -k))
        sel = np.zeros(len(nodes), bool)
        sel[pool] = True
        tot, cls = _divide(nodes, perm, xs, ys, sel)
        nch = (tot > 0).sum(1)
        size = prev
        t = len(order)
        for r, k in enumerate(order):
            size += nch[pool[k]] - 1
            if num_keypts <= size:
                t = r + 1
                filled = True
                break
        processed = [pool[k] for k in order[:t]]
        proc_set = set(processed)
        total_children = int(sum(nch[p] for p in processed))
        rest = [p for p in range(len(nodes)) if p not in proc_set]
        new_nodes = [None] * (total_children + len(rest))
        new_perm = perm.copy()
        pool_list = []
        suffix = 0
        offsets = {}
        for p in processed[::-1]:
            offsets[p] = suffix
            suffix += nch[p]
        for p in processed:
            nonempty = [c for c in range(4) if tot[p, c] > 0]
            seg_start = nodes.start[p]
            seg = perm[seg_start: seg_start + nodes.cnt[p]]
            q = cls[p]
            o = seg_start
            for r, c in enumerate(nonempty):
                pos = offsets[p] + (nch[p] - 1 - r)
                members = seg[q == c]
                new_perm[o: o + len(members)] = members
                new_nodes[pos] = (*_child_rect(nodes, p, c), o, len(members), False)
                if len(members) > 1:
                    pool_list.append(pos)
                o += len(members)
        remap = {}
        for r, p in enumerate(rest):
            new_nodes[total_children + r] = (nodes.bx[p], nodes.by[p], nodes.ex[p], nodes.ey[p], nodes.start[p],
                                             nodes.cnt[p], bool(nodes.leaf[p]))
            remap[p] = total_children + r
        nodes = _from_list(new_nodes)
        perm = new_perm
        pool = np.array(pool_list, np.int64)
        if filled or num_keypts <= len(nodes) or len(nodes) == prev:
            break
    # ---- find_keypoints_with_max_response (orb_extractor.cc:659-685): first maximum wins
    out = []
    for p in range(len(nodes)):
        seg = perm[nodes.start[p]: nodes.start[p] + nodes.cnt[p]]
        best = seg[0]
        for i in seg[1:]:
            if resp[i] > resp[best]:
                best = i
        out.append(best)
    return np.array(out, np.int64)


def _from_list(lst):
    nodes = Nodes(len(lst))
    for i, (bx, by, ex, ey, st, cnt, leaf) in enumerate(lst):
        nodes.bx[i], nodes.by[i], nodes.ex[i], nodes.ey[i] = bx, by, ex, ey
        nodes.start[i], nodes.cnt[i], nodes.leaf[i] = st, cnt, leaf
    return nodes
