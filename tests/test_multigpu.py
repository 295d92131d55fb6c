"""GPU, several devices: the i-sharded step against the single-GPU step on the same snapshot — forces equal up to
the order of the partial sums, RDF and speed histograms bit-identical, positions after 5 steps equal to float
rounding.  Two ways to shard: one process per GPU (torchrun, NCCL or CUDA-IPC peer windows) and ONE process
driving all GPUs through ljmd_create_multi (peer pointers, no NCCL).  World sizes 2, 4 and 8 as far as the box has
devices; skipped with fewer."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu

# LJMD_TEST_WORLDS="4,8" restricts the world sizes (an 8-GPU box is charged 8x: run what the 2-GPU box cannot)
_ONLY = [int(w) for w in os.environ.get("LJMD_TEST_WORLDS", "").split(",") if w.strip()]


def need(gpu_lib, world):
    if _ONLY and world not in _ONLY:
        pytest.skip(f"LJMD_TEST_WORLDS excludes world {world}")
    if gpu_lib.ljmd_device_count() < world:
        pytest.skip(f"needs {world} GPUs")


CASES = [("tvn_periodic", 1, 0, 6000, 0.8), ("evn_hardwall", 0, 1, 5003, 0.05), ("tvn_periodic_sym", 1, 0, 20000, 0.5)]
# (world, comm, case): every transport and kernel on 2 GPUs; the Newton-3 and hard-wall cases again on 4 and 8
SHARDED = [(2, c, k) for c in ("p2p", "nccl") for k in CASES] + \
          [(w, c, k) for w in (4, 8) for c, k in (("p2p", CASES[2]), ("nccl", CASES[2]), ("p2p", ("evn_hardwall_w", 0, 1, 20011, 0.05)))]


@pytest.mark.parametrize("world,comm,case", SHARDED, ids=[f"w{w}-{c}-{k[0]}" for w, c, k in SHARDED])
def test_sharded_step_matches_single_gpu(pkg, gpu_lib, tmp_path, world, comm, case):
    """One process per GPU.  comm = p2p: per-step exchange over the CUDA-IPC peer windows (fabric); nccl: NCCL."""
    name, canonical, bc, N, rho = case
    need(gpu_lib, world)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", LJMD_COMM=comm)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr",
           "127.0.0.1", "--master-port", "29544", os.path.join(ROOT, "tests", "dist_worker.py"), "gpu", str(tmp_path),
           name, str(canonical), str(bc), str(N), str(rho)]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0, out.stderr[-3000:]
    import json
    meta = json.load(open(os.path.join(tmp_path, "gpu_rank0.json")))
    assert meta["fabric"] == (comm == "p2p")
    r0 = np.load(os.path.join(tmp_path, f"{name}_rank0.npz"))
    for r in range(1, world):
        r1 = np.load(os.path.join(tmp_path, f"{name}_rank{r}.npz"))
        for k in r0.files:                       # every rank reports the same global state
            assert np.array_equal(r0[k], r1[k]), (r, k)
    compare_with_single_gpu(pkg, dict((k, r0[k]) for k in r0.files), canonical, bc, N, rho)


def observe(s, pos, vel):
    """The observation sequence of tests/dist_worker.py on system `s`."""
    s.set_state(pos, vel)
    _, _, f0 = s.get_state()
    sc0 = s.scalars()
    rdf0 = s.rdf_counts()
    s.step(0.004, 5, rdf_every=5)
    p1, v1, f1 = s.get_state()
    sc1 = s.scalars()
    rdf1, _ = s.rdf_accum()
    vh = s.velocity_histogram(0.12, 101)
    sub = s.subvolume_counts(3, 0.05)
    s.trace_begin([(0, 0.05), (3, 0.1), (5, 0.05, 3.0)], 8)
    s.step(0.004, 3)
    tr = s.trace_read()
    s.trace_end()
    return dict(f0=f0, rdf0=rdf0, p1=p1, v1=v1, f1=f1, rdf1=rdf1, vh=vh, sub=sub, tr_counts=tr["counts"],
                tr_scal=tr["scalars"], tr_mv=tr["mean_velocity"], sc0=np.array([sc0[k] for k in sorted(sc0)]),
                sc1=np.array([sc1[k] for k in sorted(sc1)]))


def compare_with_single_gpu(pkg, r0, canonical, bc, N, rho):
    pos = pkg.snapshots.lattice(N, rho, jitter=0.05, seed=21)
    vel = pkg.snapshots.velocities(N, 1.0, seed=21)
    with pkg.ljmd.LJSystem(N, T0=1.0, rho=rho, canonical=canonical, bc=bc) as s:
        one = observe(s, pos, vel)
        L = s.L
    # sharded counters and trace rows: the same occupancies up to a particle within 1e-6 of a sub-volume face
    assert r0["sub"].shape == one["sub"].shape and np.abs(r0["sub"] - one["sub"]).max() <= 2
    assert r0["tr_counts"].shape == one["tr_counts"].shape == (3, 19 + 9 + 20)
    assert np.abs(r0["tr_counts"] - one["tr_counts"]).max() <= 2
    assert (r0["tr_counts"] <= N).all() and (r0["tr_counts"][:, -1] > 0.99 * N).all()      # |vy| < 3 sigma: 99.7 %
    assert np.allclose(r0["tr_scal"], one["tr_scal"], rtol=1e-5, atol=1e-7)
    assert np.allclose(r0["tr_mv"], one["tr_mv"], rtol=0, atol=1e-6)
    fscale = np.abs(one["f0"][:, :3]).max()
    assert np.abs(r0["f0"][:, :3] - one["f0"][:, :3]).max() <= 2e-6 * fscale
    assert np.array_equal(r0["rdf0"], one["rdf0"])
    assert np.abs(r0["p1"][:, :3] - one["p1"][:, :3]).max() <= 1e-6 * max(1.0, L)
    assert np.abs(r0["v1"][:, :3] - one["v1"][:, :3]).max() <= 1e-5 * np.abs(one["v1"][:, :3]).max()
    assert np.abs(r0["rdf1"] - one["rdf1"]).sum() <= max(4, 1e-5 * one["rdf1"].sum())   # a pair may cross a bin edge after 5 steps
    assert np.abs(r0["vh"] - one["vh"]).sum() <= 2
    assert np.allclose(r0["sc0"], one["sc0"], rtol=1e-6, atol=1e-9)
    assert np.allclose(r0["sc1"], one["sc1"], rtol=1e-5, atol=1e-7)


# ------------------------------------------------------------------ one process, one handle, several GPUs
@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("case", CASES + [("evn_hardwall_w", 0, 1, 20011, 0.05)], ids=lambda c: c[0])
def test_multi_handle_matches_single_gpu(pkg, gpu_lib, world, case):
    """ljmd_create_multi: the same observation sequence through ONE handle that drives `world` devices from the
    calling thread (worker threads inside the library, peer pointers over NVLink, no NCCL, no launcher)."""
    name, canonical, bc, N, rho = case
    need(gpu_lib, world)
    nblk = -(-N // 512)
    if (world - 1) * -(-nblk // world) >= nblk:
        pytest.skip("shards are whole 512-particle blocks: the last device would get none")
    pos = pkg.snapshots.lattice(N, rho, jitter=0.05, seed=21)
    vel = pkg.snapshots.velocities(N, 1.0, seed=21)
    with pkg.ljmd.LJSystem(N, T0=1.0, rho=rho, canonical=canonical, bc=bc, devices=list(range(world))) as s:
        assert s.launch_info()["world"] == world
        got = observe(s, pos, vel)
        # the drop-in call with host buffers: every device moves its shard of the caller's arrays
        p, v, _ = s.get_state()
        hp, hv, hf = p.copy(), v.copy(), np.zeros_like(p)
        s.integrate_host(0.004, hp, hv, hf)
        p2, v2, f2 = s.get_state()
        assert np.array_equal(hp, p2) and np.array_equal(hv, v2) and np.array_equal(hf, f2)
        assert not np.array_equal(hp, p)
    compare_with_single_gpu(pkg, got, canonical, bc, N, rho)


@pytest.mark.parametrize("world", [2, 4, 8])
def test_device_init_is_identical_on_any_gpu_count(pkg, gpu_lib, world):
    """ljmd_init_state: Philox draws are counted by the global particle index and the two global sums are integer,
    so the sampled state is bit-identical on 1 and on `world` GPUs; P_xy on demand agrees too."""
    need(gpu_lib, world)
    N, T0, rho = 20011, 1.2, 0.4
    with pkg.ljmd.LJSystem(N, T0=T0, rho=rho, canonical=True, bc=0) as one:
        one.init_state(77)
        p1, v1, f1 = one.get_state()
        ps1 = one.pshear()
    with pkg.ljmd.LJSystem(N, T0=T0, rho=rho, canonical=True, bc=0, devices=list(range(world))) as many:
        many.init_state(77)
        p2, v2, f2 = many.get_state()
        ps2 = many.pshear()
    assert np.array_equal(p1, p2) and np.array_equal(v1, v2)
    assert np.abs(f1[:, :3] - f2[:, :3]).max() <= 2e-6 * np.abs(f1[:, :3]).max()
    assert abs(ps1 - ps2) <= 1e-9 * max(1.0, abs(ps1))


# ------------------------------------------------------------------ several ranks on ONE device
@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("case", [CASES[0], CASES[2], ("evn_hardwall_w", 0, 1, 20011, 0.05)], ids=lambda c: c[0])
def test_ranks_sharing_one_device_match_single_gpu(pkg, gpu_lib, tmp_path, world, case):
    """LJMD_SHARE_DEVICES=1 lets ljmd_create_multi put several ranks on one GPU: the complete sharded data path —
    shard plan, peer windows, barriers, reaction exchange, shard-only host I/O — runs at world 2, 4 and 8 on a box
    with a single device, against the plain single-GPU handle."""
    name, canonical, bc, N, rho = case
    nblk = -(-N // 512)
    if (world - 1) * -(-nblk // world) >= nblk:
        pytest.skip("shards are whole 512-particle blocks: the last rank would get none")
    out = os.path.join(tmp_path, "shared.npz")
    # ranks that share a device wait for each other INSIDE kernels (the barrier), so their kernels must be able to
    # run side by side: one hardware queue per stream, and no lazy module loading (the first launch of a kernel would
    # synchronise the context behind a peer's spinning barrier)
    env = dict(os.environ, LJMD_SHARE_DEVICES="1", CUDA_DEVICE_MAX_CONNECTIONS="32", CUDA_MODULE_LOADING="EAGER")
    run = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "shared_device_worker.py"), out, str(world),
                          str(canonical), str(bc), str(N), str(rho)], capture_output=True, text=True, timeout=600, env=env)
    assert run.returncode == 0, run.stderr[-3000:]
    z = np.load(out)
    got = {k: z[k] for k in z.files}
    compare_with_single_gpu(pkg, got, canonical, bc, N, rho)
    with pkg.ljmd.LJSystem(N, T0=1.0, rho=rho, canonical=canonical, bc=bc) as one:
        one.init_state(77)
        p1, v1, _ = one.get_state()
    assert np.array_equal(got["init_pos"], p1) and np.array_equal(got["init_vel"], v1)


def test_multi_handle_argument_errors(pkg, gpu_lib, monkeypatch):
    monkeypatch.delenv("LJMD_SHARE_DEVICES", raising=False)
    with pytest.raises(pkg.ljmd.LJMDError):
        pkg.ljmd.LJSystem(4096, devices=[0, 0])                       # a device listed twice
    with pytest.raises(pkg.ljmd.LJMDError):
        pkg.ljmd.LJSystem(4096, devices=[0, 999])                     # no such device
    monkeypatch.setenv("LJMD_SHARE_DEVICES", "1")
    if os.environ.get("CUDA_MODULE_LOADING") != "EAGER":
        with pytest.raises(pkg.ljmd.LJMDError, match="CUDA_MODULE_LOADING=EAGER"):
            pkg.ljmd.LJSystem(4096, devices=[0, 0])                   # sharing needs eager module loading: say so at once
    monkeypatch.delenv("LJMD_SHARE_DEVICES")
    with pkg.ljmd.LJSystem(2048, T0=1.0, rho=0.5, devices=[0]) as s:  # one device: plain ljmd_create
        assert s.launch_info()["world"] == 1
    if gpu_lib.ljmd_device_count() >= 2:
        with pytest.raises(pkg.ljmd.LJMDError):
            pkg.ljmd.LJSystem(400, T0=1.0, rho=0.5, devices=[0, 1])   # one 512-block cannot be split over two devices


REF_DIR = os.path.join(ROOT, "oracle", "_ref")


@pytest.mark.skipif(not os.path.exists(os.path.join(REF_DIR, "run-fluctuations")), reason="drop-in tasks not built")
def test_reference_driver_on_two_gpus_prints_the_same_table(gpu_lib, tmp_path):
    """The reference's own run-fluctuations driver (unmodified source, compiled against host/MDSystem.h): with
    LJMD_DEVICES=0,1 the MDSystem class shards the system over two GPUs and the driver prints the table it
    prints on one (same seed; trajectories differ only by the summation order of the forces)."""
    if gpu_lib.ljmd_device_count() < 2:
        pytest.skip("needs 2 GPUs")
    exe = os.path.join(REF_DIR, "run-fluctuations")
    inp = os.path.join(ROOT, "tests", "data", "N2048.short.input")
    tables = []
    for devs in (None, "0,1"):
        env = dict(os.environ, LJMD_SEED="2024")
        env.pop("LJMD_DEVICES", None)
        if devs:
            env["LJMD_DEVICES"] = devs
        wd = tmp_path / ("multi" if devs else "single")
        wd.mkdir()
        out = subprocess.run([exe, inp], cwd=wd, capture_output=True, text=True, timeout=600, env=env)
        assert out.returncode == 0, out.stderr[-2000:]
        rows = [list(map(float, ln.split())) for ln in out.stdout.splitlines()
                if len(ln.split()) == 8 and ln.split()[0][0].isdigit()]
        assert len(rows) >= 1, out.stdout[-2000:]
        tables.append(np.array(rows))
    one, two = tables
    assert one.shape == two.shape
    assert np.allclose(one[:, 0], two[:, 0])                        # time column
    assert np.allclose(one[:, 1:7], two[:, 1:7], rtol=2e-3, atol=2e-3), (one, two)
