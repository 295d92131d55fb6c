"""Worker for the multi-process tests (CPU gloo logic test and 2-GPU parity test).
Launched by torch.distributed.run; writes a JSON result per rank into the directory given as argv[2]."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ljpkg  # noqa: E402


def main():
    mode, outdir = sys.argv[1], sys.argv[2]
    pkg = ljpkg.load()
    D = pkg.dist
    if mode == "cpu":
        rank, world, _ = D.init("gloo")
        uid = D.share_unique_id(lambda: bytes(range(128)))
        N = 1000003
        lo, hi = D.shard_bounds(N, rank, world)
        plan = pkg.ljmd.plan(N, rank, world, 148)
        slowest = D.max_over_ranks(10.0 + rank)
        D.barrier()
        res = dict(rank=rank, world=world, uid_ok=(uid == bytes(range(128))), lo=lo, hi=hi,
                   plan_lo=plan["i_begin"], plan_hi=plan["i_end"], slowest=slowest)
    else:
        import torch
        rank, world, local_rank = D.init("nccl")
        uid = D.share_unique_id(pkg.ljmd.LJSystem.nccl_unique_id)
        name, canonical, bc = sys.argv[3], int(sys.argv[4]), int(sys.argv[5])
        cfg = dict(N=int(sys.argv[6]), T=1.0, rho=float(sys.argv[7]))
        pos = pkg.snapshots.lattice(cfg["N"], cfg["rho"], jitter=0.05, seed=21)
        vel = pkg.snapshots.velocities(cfg["N"], cfg["T"], seed=21)
        s = pkg.ljmd.LJSystem(cfg["N"], T0=cfg["T"], rho=cfg["rho"], canonical=canonical, bc=bc, device=local_rank,
                              rank=rank, world=world, nccl_unique_id=uid)
        fabric = D.connect_fabric(s)
        s.set_state(pos, vel)
        _, _, f0 = s.get_state()
        sc0 = s.scalars()
        rdf0 = s.rdf_counts()
        s.step(0.004, 5, rdf_every=5)
        p1, v1, f1 = s.get_state()
        sc1 = s.scalars()
        rdf1, nacc = s.rdf_accum()
        vh = s.velocity_histogram(0.12, 101)
        # observation trace and sub-volume counters on the sharded state (counts are all-reduced on read-out)
        sub = s.subvolume_counts(3, 0.05)
        s.trace_begin([(0, 0.05), (3, 0.1), (5, 0.05, 3.0)], 8)
        s.step(0.004, 3)
        tr = s.trace_read()
        s.trace_end()
        np.savez(os.path.join(outdir, f"{name}_rank{rank}.npz"), f0=f0, rdf0=rdf0, p1=p1, v1=v1, f1=f1, rdf1=rdf1,
                 sub=sub, tr_counts=tr["counts"], tr_scal=tr["scalars"], tr_mv=tr["mean_velocity"],
                 vh=vh, sc0=np.array([sc0[k] for k in sorted(sc0)]), sc1=np.array([sc1[k] for k in sorted(sc1)]))
        res = dict(rank=rank, world=world, nacc=nacc, info=s.launch_info(), fabric=bool(fabric))
        s.close()
    with open(os.path.join(outdir, f"{mode}_rank{rank}.json"), "w") as fh:
        json.dump(res, fh)
    D.finalize()


if __name__ == "__main__":
    main()
