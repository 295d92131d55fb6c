"""CPU, world_size 2 over gloo: the host-side logic of the one-process-per-GPU layout — unique-id
broadcast, shard bounds agreeing with the library's plan, barrier, max-over-ranks."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def test_two_rank_plumbing_over_gloo(tmp_path):
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tests", "dist_worker.py"), "cpu", str(tmp_path)]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env)
    assert out.returncode == 0, out.stderr[-3000:]
    res = [json.load(open(os.path.join(tmp_path, f"cpu_rank{r}.json"))) for r in range(2)]
    assert [r["rank"] for r in res] == [0, 1] and all(r["world"] == 2 for r in res)
    assert all(r["uid_ok"] for r in res)
    assert res[0]["lo"] == 0 and res[0]["hi"] == res[1]["lo"] and res[1]["hi"] == 1000003
    for r in res:
        assert (r["lo"], r["hi"]) == (r["plan_lo"], r["plan_hi"])
        assert r["slowest"] == 11.0


def test_shard_bounds_cover_every_particle(pkg):
    for N in (2, 7, 400, 65536, 1000003):
        for world in (1, 2, 3, 4, 8):
            prev = 0
            for r in range(world):
                lo, hi = pkg.dist.shard_bounds(N, r, world)
                assert lo == prev and hi >= lo
                prev = hi
            assert prev == N
