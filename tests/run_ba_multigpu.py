"""Multi-GPU local BA check, launched with torchrun (one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tests/run_ba_multigpu.py

Every rank owns a contiguous block of landmarks; the ranks exchange only the packed reduced camera system
(one-shot all-reduce kernel over NVLink peer memory, or ncclAllReduce).  The merged result must equal the single-GPU / oracle result (1e-4 relative)."""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parent))
import ba_data  # noqa: E402
import conftest  # noqa: E402
import oracle_api  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    torch.cuda.set_device(local)
    plp = conftest.load_package()
    from plpslam_b200.ba import BaComm, LocalBA, shard_boundaries, shard_edges
    ctx = plp.Context(local)
    uid = [BaComm.unique_id(ctx) if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    comm = BaComm(ctx, uid[0], world, rank)
    ok = True
    for seed, kw in [(5, dict(n_local=8, n_fixed=4, n_points=600, n_lines=100, n_plane_pts=30)), (42, dict()),
                     # > 32 non-fixed keyframes: reduced system dense in HBM, 1 all-reduce of 0.5 MB per try
                     (71, dict(n_local=40, n_fixed=8, n_points=1500, n_lines=200, n_plane_pts=50))]:
        prob = ba_data.make_ba_problem(seed, **kw)
        sub = prob.shard(world, rank, shard_boundaries, shard_edges)
        st = sub.struct()
        ba = LocalBA(ctx, st, (len(sub.kf_fixed), len(sub.pt_pos_w), len(sub.line_plucker), len(sub.pt_edge_kf),
                               len(sub.line_edge_kf)), comm=comm)
        out = ba.solve()
        ba.close()
        gathered = [None] * world
        dist.all_gather_object(gathered, dict(block=sub.block, pts=out["pt_pos_w"], lines=out["line_plucker"],
                                              pt_out=out["pt_edge_outlier"], ln_out=out["line_edge_outlier"],
                                              poses=out["kf_pose_cw"], it=(out["iters_first"], out["iters_second"],
                                                                           out["lm_tries"])))
        if rank == 0:
            orc = oracle_api.Oracle()
            o = ba_data.oracle_local_ba(orc, prob)
            pts = np.zeros_like(o.pt_pos_w)
            lines = np.zeros_like(o.line_plucker)
            pt_out = np.zeros_like(o.pt_edge_outlier)
            ln_out = np.zeros_like(o.line_edge_outlier)
            for g in gathered:
                p0, p1 = g["block"]["pts"]
                l0, l1 = g["block"]["lines"]
                pts[p0:p1] = g["pts"]
                lines[l0:l1] = g["lines"]
                pt_out[g["block"]["pt_edges"]] = g["pt_out"]
                ln_out[g["block"]["line_edges"]] = g["ln_out"]
                assert np.array_equal(g["poses"], gathered[0]["poses"]), "ranks disagree on the poses"
                assert g["it"] == gathered[0]["it"]
            rel_pose = np.linalg.norm(gathered[0]["poses"] - o.kf_pose_cw) / np.linalg.norm(o.kf_pose_cw)
            rel_pts = np.quantile(np.linalg.norm(pts - o.pt_pos_w, axis=1) / np.linalg.norm(o.pt_pos_w, axis=1), 0.999)
            mism = int((pt_out != o.pt_edge_outlier).sum()) + int((ln_out != o.line_edge_outlier).sum())
            good = rel_pose < 1e-4 and rel_pts < 1e-4 and mism <= 1e-3 * (len(pt_out) + len(ln_out)) and \
                gathered[0]["it"] == (o.iters_first, o.iters_second, o.lm_tries)
            print(f"[ba multi-gpu world={world}] seed {seed}: rel_pose {rel_pose:.2e} rel_pts {rel_pts:.2e} "
                  f"flag mismatches {mism} iters {gathered[0]['it']} vs oracle {(o.iters_first, o.iters_second, o.lm_tries)} "
                  f"-> {'OK' if good else 'FAIL'}")
            ok = ok and good
    if rank == 0:
        print(f"[ba multi-gpu world={world}] collective: {'NVLink peer-memory one-shot kernel' if comm.peer_active() else 'ncclAllReduce'}"
              f" ({comm.peer_count()} of {comm.allreduce_count()} all-reduces on the peer path)")
    comm.close()
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0 and not ok:
        sys.exit(1)


if __name__ == "__main__":
    main()
