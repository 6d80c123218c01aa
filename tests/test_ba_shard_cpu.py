"""CPU (gloo, world_size 2) coverage of the host-side multi-GPU logic of local BA: landmark blocks partition the
problem exactly, are balanced by edge count, and the per-rank packed vectors sum to the global one under the
same all-reduce the GPU path uses (here a numpy stand-in of the packed vector: edge-count histograms)."""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

sys.path.insert(0, str(Path(__file__).resolve().parent))


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import ba_data
    import conftest
    plp = conftest.load_package()
    from plpslam_b200.ba import shard_boundaries, shard_edges
    prob = ba_data.make_ba_problem(11, n_local=6, n_fixed=3, n_points=500, n_lines=90, n_plane_pts=40)
    sub = prob.shard(world, rank, shard_boundaries, shard_edges)
    # every rank replicates all keyframes
    assert np.array_equal(sub.kf_pose_cw, prob.kf_pose_cw) and np.array_equal(sub.kf_fixed, prob.kf_fixed)
    # edges stay grouped by ascending local landmark index (what plp_ba_create requires)
    assert np.all(np.diff(sub.pt_edge_lm) >= 0) and np.all(np.diff(sub.line_edge_lm) >= 0)
    assert sub.pt_edge_lm.max() < len(sub.pt_pos_w) and (len(sub.line_edge_lm) == 0 or sub.line_edge_lm.max() < len(sub.line_plucker))
    # "packed vector" stand-in: per-keyframe edge counts; the all-reduce(sum) must reproduce the global histogram
    n_kf = len(prob.kf_fixed)
    local = torch.tensor(np.concatenate([np.bincount(sub.pt_edge_kf, minlength=n_kf), np.bincount(sub.line_edge_kf, minlength=n_kf),
                                         [len(sub.pt_pos_w), len(sub.line_plucker), len(sub.plane_edge_lm)]]), dtype=torch.float64)
    dist.all_reduce(local, op=dist.ReduceOp.SUM)
    glob = np.concatenate([np.bincount(prob.pt_edge_kf, minlength=n_kf), np.bincount(prob.line_edge_kf, minlength=n_kf),
                           [len(prob.pt_pos_w), len(prob.line_plucker), len(prob.plane_edge_lm)]])
    assert np.array_equal(local.numpy(), glob.astype(np.float64))
    # balance: no rank holds more than 65 % of the edges at world_size 2
    share = (len(sub.pt_edge_kf) + len(sub.line_edge_kf)) / (len(prob.pt_edge_kf) + len(prob.line_edge_kf))
    assert 0.35 < share < 0.65, share
    blocks = [None] * world
    dist.all_gather_object(blocks, sub.block["pts"] + sub.block["lines"])
    if rank == 0:
        assert blocks[0][0] == 0 and blocks[0][1] == blocks[1][0] and blocks[1][1] == len(prob.pt_pos_w)
        assert blocks[0][2] == 0 and blocks[0][3] == blocks[1][2] and blocks[1][3] == len(prob.line_plucker)
        ret.put("ok")
    dist.destroy_process_group()


def test_landmark_sharding_world_size_2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 400)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=180)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    assert q.get(timeout=5) == "ok"


def test_front_end_frame_sharding_is_a_partition():
    """bench.py shards frames over ranks with no collective: rank r builds problems from seed + 1000 r."""
    seeds = {1234 + 1000 * r for r in range(8)}
    assert len(seeds) == 8
