#!/usr/bin/env python
"""Benchmark of the Lennard-Jones MD step (BASELINE.json: pair interactions/s and MD steps/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config C5] [--impl reference]

A "step" is one MDSystem::Integrate(dt): drift, all-pairs force/potential/virial, EVN half-kick or TVN
rescale, boundary conditions, parameters.  N(N-1) ordered pair interactions per step.  With N > 1 ranks
(torchrun, one process per GPU) the i-particles are sharded: strong scaling on the same workload.

Prints ONE JSON line on rank 0.  See DESIGN.md §Measurement for every field.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import ljpkg  # noqa: E402

DT = 0.004                      # GUI/task default (reference src/gui/mainwindow.cpp:150, input/*)
FLOP_PER_PAIR = {0: 37, 1: 25, 2: 25}   # flops of one ORDERED pair evaluation, periodic / open
REACTION_FLOP = 6                        # Newton-3 kernel: 3 more FMAs per unordered pair put -f_ij on particle j
# FP32-pipe lane-instructions this library's kernels execute per pair evaluation (SASS count):
# 7 packed FFMA2-class instructions = 14 lanes periodic, 8 = 16 open, + 3 scalar FFMA of reaction (Newton-3)
FP32_LANE_INSTR = {0: 14, 1: 16, 2: 16}
SM_COUNT, FP32_LANES = 148, 128
# dram__bytes_read.sum + dram__bytes_write.sum of the force kernel per launch on ONE GPU, from `ncu --set full`
# captures of this command (bench.py cannot run under a profiler: a number taken under ncu is never a bench
# value).  Keyed by configuration; `source` names the committed summary the figure comes from.
TRAFFIC_NCU = {}
try:
    TRAFFIC_NCU = json.load(open(os.path.join(ROOT, "profiles", "traffic_ncu.json")))
except Exception:
    pass


_REAL_STDOUT = None


def emit(line):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def measured_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


class ClockSampler:
    """SM clock + throttle reasons sampled during the timed region (pynvml, else nvidia-smi)."""

    def __init__(self, index=0, period=0.1):
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        self._mode = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
            self._mode = "nvml"
        except Exception:
            self._mode = "smi"

    def _reason_names(self, mask):
        nv = self._nv
        table = [("hw_slowdown", "nvmlClocksThrottleReasonHwSlowdown"),
                 ("hw_thermal_slowdown", "nvmlClocksThrottleReasonHwThermalSlowdown"),
                 ("sw_thermal_slowdown", "nvmlClocksThrottleReasonSwThermalSlowdown"),
                 ("sw_power_cap", "nvmlClocksThrottleReasonSwPowerCap"),
                 ("hw_power_brake", "nvmlClocksThrottleReasonHwPowerBrakeSlowdown"),
                 ("sync_boost", "nvmlClocksThrottleReasonSyncBoost"),
                 ("app_clocks", "nvmlClocksThrottleReasonApplicationsClocksSetting")]
        out = []
        for name, attr in table:
            bit = getattr(nv, attr, None)
            if bit is not None and (mask & bit):
                out.append(name)
        return out

    def _run_nvml(self):
        nv = self._nv
        while not self._stop.is_set():
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self._h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
                self.reasons.update(self._reason_names(mask))
            except Exception:
                pass
            self._stop.wait(self.period)

    def _run_smi(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5)
                f = [x.strip() for x in out.stdout.strip().splitlines()[0].split(",")]
                self.samples.append(float(f[0]))
                self.max_mhz = float(f[1])
                for n, v in zip(names, f[2:]):
                    if v.lower().startswith("active"):
                        self.reasons.add(n)
            except Exception:
                pass
            self._stop.wait(max(self.period, 0.2))

    def start(self):
        self._thread = threading.Thread(target=self._run_nvml if self._mode == "nvml" else self._run_smi, daemon=True)
        self._thread.start()

    def stop(self):
        self._stop.set()
        if self._thread:
            self._thread.join(timeout=5)
        if not self.samples:
            return None
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples), "source": self._mode}


def workload(pkg, name):
    cfg = dict(pkg.snapshots.CONFIGS[name])
    pos, vel = pkg.snapshots.make(name)
    return cfg, pos, vel


def describe(name, cfg, world, transport="none"):
    ens = "TVN" if cfg["canonical"] else "EVN"
    bc = {0: "periodic", 1: "hard-wall", 2: "open"}[cfg["bc"]]
    return {"workload": f"{name}: N={cfg['N']} T*={cfg['T']} rho*={cfg['rho']} {bc} {ens} dt*={DT}"
                        + (f" RDF every {cfg['rdf_every']} steps" if cfg["rdf_every"] else ""),
            "N": cfg["N"], "rho": cfg["rho"], "T": cfg["T"], "boundary": bc, "ensemble": ens,
            "rdf_every": cfg["rdf_every"], "dt": DT,
            "parallelism": f"i-shards x{world}" if world > 1 else "single GPU",
            "transport": transport,
            "init": "reference start lattice (MDSystem.cpp:147-168) + 5% jitter, seeded Gaussian velocities",
            "l2": "192 MiB scratch overwritten before every timed step (L2 flush, outside the per-step event pairs); "
                  "inputs themselves fit in L2 by design"}


# ------------------------------------------------------------------------------------- reference arm
def reference_sample_n(cfg, steps, warmup, budget_s=150.0):
    for n in (16384, 8192, 4096, 2048, 1024):
        if n <= cfg["N"] and (steps + warmup + 2) * 4.0e-8 * n * n <= budget_s:
            return n
    return min(cfg["N"], 1024)


def run_reference_sample(pkg, cfg, n_s, steps, warmup):
    """Time the reference's own CPU implementation (oracle/_ref when built, else the C restatement) on a
    bounded sample: the same density / temperature / boundary / ensemble at a smaller N.  The reference
    cost is exactly N(N-1) pair evaluations per step on one thread, so pairs/s transfers."""
    from oracle.oracle import Oracle, Reference, reference_available
    sub = dict(cfg, N=n_s)
    pos = pkg.snapshots.lattice(n_s, cfg["rho"], jitter=0.05)
    vel = pkg.snapshots.velocities(n_s, cfg["T"])
    if reference_available():
        kind = "reference"
        ref = Reference(n_s, cfg["T"], cfg["rho"], cfg["canonical"], cfg["bc"])
        ref.set_state(pos, vel)
        for _ in range(warmup):
            ref.integrate(DT, 1)
        t0 = time.perf_counter()
        for _ in range(steps):
            ref.integrate(DT, 1)
        dt_s = time.perf_counter() - t0
        ref.close()
    else:
        kind = "port"
        o = Oracle()
        L, dr2 = o.box_length(n_s, cfg["rho"]), o.rdf_dr2(n_s)
        frc, _, _ = o.forces(pos, L, cfg["bc"], dr2)
        p, v, f = pos, vel, frc
        for _ in range(warmup):
            p, v, f, _, _ = o.integrate(n_s, cfg["rho"], cfg["T"], cfg["canonical"], cfg["bc"], DT, p, v, f)
        t0 = time.perf_counter()
        for _ in range(steps):
            p, v, f, _, _ = o.integrate(n_s, cfg["rho"], cfg["T"], cfg["canonical"], cfg["bc"], DT, p, v, f)
        dt_s = time.perf_counter() - t0
    pairs = float(n_s) * (n_s - 1) * steps
    return dict(value=pairs / dt_s, ms_per_step=1e3 * dt_s / steps, kind=kind, n=sub["N"], steps=steps)


def host_description():
    model = "unknown"
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                model = line.split(":", 1)[1].strip()
                break
    except Exception:
        pass
    return f"{model}, {os.cpu_count()} logical CPUs; the reference CPU path is single-threaded"


def main_reference(args, pkg):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cfg = dict(pkg.snapshots.CONFIGS[args.config])
    n_s = reference_sample_n(cfg, args.steps, args.warmup)
    r = run_reference_sample(pkg, cfg, n_s, args.steps, args.warmup)
    full_pairs = float(cfg["N"]) * (cfg["N"] - 1)
    extrapolated = n_s != cfg["N"]
    sample = (f"{args.steps} x Integrate(dt) at N={n_s} (same rho*, T*, boundary, ensemble as the workload; "
              f"the reference cost is N(N-1) pair evaluations per step on one thread, so the pair rate transfers); "
              f"{host_description()}")
    config = describe(args.config, cfg, 1)
    # the workload named is the GPU arm's; what was TIMED is a bounded sample of it at a smaller N (a real C5 step
    # takes the reference ~7 h): say so in the config itself, and give ms_per_step for the named N by extrapolation
    config["N_sampled"] = n_s
    config["extrapolated"] = extrapolated
    config["workload"] += f" [reference CPU arm: timed at N={n_s}, pair rate extrapolated to N={cfg['N']}]" if extrapolated else ""
    line = {
        "impl": "reference", "metric": "pair_interactions_per_s", "value": r["value"], "unit": "pairs/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * full_pairs / r["value"],
        "ms_per_step_is": f"extrapolated to N={cfg['N']} from the measured pair rate" if extrapolated else "measured",
        "measured": {"N": n_s, "ms_per_step": r["ms_per_step"], "steps": args.steps},
        "extrapolated": extrapolated, "same_config": not extrapolated,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32/f64 mixed (reference CPU)",
        "data": "synthetic", "config": config,
        "md_steps_per_s_extrapolated": r["value"] / full_pairs,
        "cpu_baseline": {"value": r["value"], "unit": "pairs/s", "cores": 1, "kind": r["kind"], "sample": sample},
        "e2e": {"value": r["value"], "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


def timed_steps(sysm, steps, warmup, rdf_every, barrier, max_over_ranks, sampler=None):
    """W warm-up steps, then exactly K timed steps (CUDA events on the handle's stream, barrier + sync on both
    sides, max over ranks).  Returns timings of the K steps."""
    sysm.step(DT, warmup, rdf_every)
    barrier()
    if sampler:
        sampler.start()
    l0 = sysm.launch_count()
    t0 = time.perf_counter()
    sysm.step(DT, steps, rdf_every)
    barrier()
    wall_s = time.perf_counter() - t0
    launches = sysm.launch_count() - l0
    tim = sysm.last_step_timing()
    clocks = sampler.stop() if sampler else None
    return dict(dev_ms=max_over_ranks(tim["total_ms"]),
                force_ms=max_over_ranks(tim["force_ms"] / max(1, tim["force_launches"])),
                wall_s=wall_s, launches=launches, clocks=clocks)


def force_roofline(cfg, info, world, force_ms, dev_ms, steps, fp32_nominal, fp32_measured, clk_mhz):
    """FP32 roofline of the force kernel: algorithmic flops of one launch / its measured duration."""
    N = cfg["N"]
    pairs_per_step = float(N) * (N - 1)
    # The Newton-3 kernel evaluates every UNORDERED pair once (37 + 6 flops periodic) where the reference's
    # double loop evaluates both orders (2 x 37): `achieved` counts the flops the kernel's own algorithm
    # executes; `ordered_pair_equivalent` is the same launch priced at the reference's algorithm.
    newton3 = bool(info.get("newton3"))
    fpp = FLOP_PER_PAIR[cfg["bc"]]
    evals_per_launch = pairs_per_step / world / (2.0 if newton3 else 1.0)
    flops_per_launch = (fpp + (REACTION_FLOP if newton3 else 0)) * evals_per_launch
    achieved = flops_per_launch / (force_ms * 1e-3) / 1e12
    ordered_equiv = fpp * pairs_per_step / world / (force_ms * 1e-3) / 1e12
    return {
        "bound": "fp32",
        "kernel": ("k_force_sym (all-pairs LJ force/potential/virial, each unordered pair once)" if newton3
                   else "k_force (all-pairs LJ force/potential/virial, ordered pairs)"),
        "achieved": achieved, "peak": fp32_nominal, "unit": "TFLOP/s", "frac": achieved / fp32_nominal,
        "peak_source": "nominal: 148 SMs x 128 FP32 lanes x 2 x sm_max_mhz of MEASURED_PEAKS.json (that file holds no "
                       "FP32 CUDA-core figure; the kernel is FP32-issue bound, not HBM or tensor bound)",
        "peak_measured": fp32_measured,
        "frac_of_measured_peak": (achieved / fp32_measured) if fp32_measured else None,
        "peak_measured_source": "ljmd_fp32_peak_probe: stream of independent packed FMAs on this GPU, CUDA events, in this run",
        "flop_per_pair_evaluation": fpp + (REACTION_FLOP if newton3 else 0),
        "pair_evaluations_per_launch": evals_per_launch,
        "ordered_pair_equivalent": {"tflops": ordered_equiv, "frac": ordered_equiv / fp32_nominal,
                                    "note": "same launch priced as the reference's ordered double loop (37/25 flop x N(N-1))"},
        "kernel_ms": force_ms, "kernel_share_of_step": force_ms * steps / dev_ms,
        "fp32_pipe_util": (FP32_LANE_INSTR[cfg["bc"]] + (3 if newton3 else 0)) * evals_per_launch / (force_ms * 1e-3)
                          / (SM_COUNT * FP32_LANES * clk_mhz * 1e6),
    }


def parity_check(pkg, sysm, cfg, rank, rdf_initial, nsample=512):
    """After the timed regions: re-evaluate the forces of the evolved state (at this world size) and compare
    `nsample` seeded particles with the FP64 arbiter over all N partners (oracle, host threads, rank 0).  The
    RDF hash is of the INITIAL snapshot (identical input on every GPU count, so the hash must be too)."""
    import hashlib
    from oracle.oracle import Oracle
    sysm.compute_forces()
    pos, _, frc = sysm.get_state(vel=False)
    out = None
    if rank == 0:
        o = Oracle()
        N = cfg["N"]
        idx = np.sort(np.random.default_rng(99).choice(N, size=min(nsample, N), replace=False)).astype(np.int32)
        t0 = time.perf_counter()
        f64, fterm, _, _ = o.forces_f64_subset(pos, sysm.L, cfg["bc"], idx)
        err = np.abs(frc[idx, :3].astype(np.float64) - f64).max(axis=1) / fterm
        out = {"max_err": float(err.max()), "n_sampled": int(idx.size), "tolerance": 1e-5, "ok": bool(err.max() <= 1e-5),
               "scale": "max-norm force error / sum_j(|repulsive| + |attractive|) pair terms, FP64 arbiter over all N partners",
               "state": "evolved state after the timed steps, forces re-evaluated at this GPU count",
               "rdf_sha": hashlib.sha256(np.asarray(rdf_initial, dtype=np.int64).tobytes()).hexdigest()[:16],
               "rdf_of": "initial snapshot (bit-exact bins: the hash is the same on 1, 2, 4 and 8 GPUs)",
               "arbiter_s": time.perf_counter() - t0}
    return out


def reference_gpu_leg():
    """The reference's own CUDA force kernel (MDSystem.cu, unmodified, recompiled for sm_100a: oracle/_ref/
    ref_gpu_bench, built where /root/reference exists) on this GPU at N = 65 536: the a-3' bar."""
    exe = os.path.join(ROOT, "oracle", "_ref", "ref_gpu_bench")
    if not os.path.exists(exe):
        return {"unavailable": "oracle/_ref/ref_gpu_bench not built"}
    try:
        out = subprocess.run([exe, "65536", "1.1", "2", "1"], capture_output=True, text=True, timeout=120)
        return json.loads(out.stdout.strip().splitlines()[-1])
    except Exception as e:   # noqa: BLE001
        return {"unavailable": f"{type(e).__name__}: {e}"}


# ------------------------------------------------------------------------------------------ our arm
def main_ours(args, pkg):
    import torch

    D = pkg.dist
    rank, world, local_rank = D.env_rank_world()
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback "
                         "(use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local_rank)
    ljmd = pkg.ljmd
    D.init("nccl")
    barrier, max_over_ranks = D.barrier, D.max_over_ranks

    def make_system(name):
        cfg, pos, vel = workload(pkg, name)
        uid = D.share_unique_id(ljmd.LJSystem.nccl_unique_id) if world > 1 else None
        sysm = ljmd.LJSystem(cfg["N"], T0=cfg["T"], rho=cfg["rho"], canonical=cfg["canonical"], bc=cfg["bc"],
                             device=local_rank, rank=rank, world=world, nccl_unique_id=uid)
        fabric = D.connect_fabric(sysm) if world > 1 else False
        sysm.set_state(pos, vel)
        return cfg, sysm, fabric

    peaks = measured_peaks()
    sm_max = float(peaks.get("sm_max_mhz", 1965.0))
    fp32_nominal = SM_COUNT * FP32_LANES * 2 * sm_max * 1e6 / 1e12      # TFLOP/s at max clock
    fp32_measured = ljmd.fp32_peak_probe(local_rank)

    cfg, sysm, fabric = make_system(args.config)
    N = cfg["N"]
    rdf_initial = sysm.rdf_counts()
    sysm.set_l2_flush(192 << 20)
    sysm.set_event_timing(True)
    info = sysm.launch_info()
    rdf_every = cfg["rdf_every"]

    # ---- device-resident: W warm-up steps, then exactly K timed steps
    sampler = ClockSampler(local_rank) if rank == 0 else None
    tm = timed_steps(sysm, args.steps, args.warmup, rdf_every, barrier, max_over_ranks, sampler)
    dev_ms, force_ms, clocks = tm["dev_ms"], tm["force_ms"], tm["clocks"]
    gat = sysm.last_gather_timing()
    red = sysm.last_reduce_timing()
    sc = sysm.scalars()
    pairs_per_step = float(N) * (N - 1)
    value = pairs_per_step * args.steps / (dev_ms * 1e-3)

    # ---- end to end through the drop-in call with pinned HOST buffers (upload + step + download each step)
    sysm.set_event_timing(False)
    hp = torch.empty((N, 4), dtype=torch.float32).pin_memory()
    hv = torch.empty((N, 4), dtype=torch.float32).pin_memory()
    hp_np, hv_np = hp.numpy(), hv.numpy()
    p_now, v_now, _ = sysm.get_state(force=False)
    hp_np[...] = p_now
    hv_np[...] = v_now
    e2e_steps = max(3, min(args.steps, 20))
    for _ in range(2):
        sysm.integrate_host(DT, hp_np, hv_np)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        sysm.integrate_host(DT, hp_np, hv_np)
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_value = pairs_per_step * e2e_steps / e2e_s
    h2d = 2 * 16 * N                       # every rank uploads its shard of pos and vel: N records in total
    d2h = 2 * 16 * N + 8 * 20 * world      # and reads its shard back, plus the scalar block per rank

    # ---- rooflines
    clk = (clocks or {}).get("sm_mhz") or sm_max
    roofline = force_roofline(cfg, info, world, force_ms, dev_ms, args.steps, fp32_nominal, fp32_measured, clk)
    tr = TRAFFIC_NCU.get(args.config) if world == 1 else None
    roofline["traffic"] = tr["bytes"] if tr else None
    roofline["traffic_source"] = tr["source"] if tr else "not captured for this configuration / GPU count"
    # second roofline: the HBM-bound part of the step against the measured copy bandwidth of MEASURED_PEAKS.json.
    # One GPU: k_gather (sums the partial-force rows and the reaction blocks, finishes the velocities).  Sharded:
    # k_reduce_reaction does the block reads (k_gather then pulls one pre-reduced record per rank over NVLink).
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    if world == 1 or red["launches"] == 0:
        k_ms = max_over_ranks(gat["gather_ms"] / max(1, gat["launches"]))
        k_bytes, k_name = gat["bytes_per_launch"], "k_gather (sum partial-force rows and reaction blocks, finish velocities)"
    else:
        k_ms = max_over_ranks(red["reduce_ms"] / max(1, red["launches"]))
        k_bytes, k_name = red["bytes_per_launch"], "k_reduce_reaction (column sums of this rank's reaction blocks)"
    roofline_hbm = {
        "bound": "hbm", "kernel": k_name,
        "achieved": k_bytes / (k_ms * 1e-3) / 1e9 if k_ms > 0 else None,
        "peak": hbm_peak, "unit": "GB/s",
        "frac": (k_bytes / (k_ms * 1e-3) / 1e9 / hbm_peak) if k_ms > 0 else None,
        "peak_source": "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)",
        "bytes_per_launch": k_bytes, "kernel_ms": k_ms,
        "kernel_share_of_step": k_ms * args.steps / dev_ms,
        "note": "blocks written by the force kernel a moment earlier are partly served from the 126 MB L2",
    }

    # ---- parity of this run (after the timed regions; at this GPU count)
    parity = None if args.no_parity else parity_check(pkg, sysm, cfg, rank, rdf_initial)
    sysm.close()

    # ---- the metric is quoted over N = 64k-1M: short device-timed runs of the other large configurations
    sweep = {}
    if args.sweep:
        for name in [c for c in ("C3", "C4") if c != args.config]:
            c2, s2, _ = make_system(name)
            s2.set_l2_flush(192 << 20)
            s2.set_event_timing(True)
            st = 20 if name == "C3" else 8
            t2 = timed_steps(s2, st, 3, c2["rdf_every"], barrier, max_over_ranks)
            i2 = s2.launch_info()
            pp = float(c2["N"]) * (c2["N"] - 1)
            r2 = force_roofline(c2, i2, world, t2["force_ms"], t2["dev_ms"], st, fp32_nominal, fp32_measured, clk)
            g2, d2 = s2.last_gather_timing(), s2.last_reduce_timing()
            g2_ms = max_over_ranks(g2["gather_ms"] / max(1, g2["launches"]))
            d2_ms = max_over_ranks(d2["reduce_ms"] / max(1, d2["launches"])) if world > 1 else 0.0
            sweep[name] = {"workload": describe(name, c2, world)["workload"], "steps": st, "warmup": 3,
                           "ms_per_step": t2["dev_ms"] / st, "value": pp * st / (t2["dev_ms"] * 1e-3), "unit": "pairs/s",
                           "md_steps_per_s": st / (t2["dev_ms"] * 1e-3), "force_kernel_ms": t2["force_ms"],
                           "roofline_frac": r2["frac"], "frac_of_measured_peak": r2["frac_of_measured_peak"],
                           "force_share_of_step": r2["kernel_share_of_step"],
                           # where the rest of the step goes (per launch, max over ranks): k_gather, the sharded
                           # runs' k_reduce_reaction, and what is left (drift/finish kernels, barriers, launch gaps)
                           "gather_kernel_ms": g2_ms, "reduce_reaction_kernel_ms": d2_ms,
                           "other_ms_per_step": t2["dev_ms"] / st - t2["force_ms"] - g2_ms - d2_ms,
                           "launch": i2}
            s2.close()

    if rank == 0:
        line = {
            "metric": "pair_interactions_per_s", "value": value, "unit": "pairs/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": describe(args.config, cfg, world,
                               ("NVLink peer windows (CUDA IPC), fused into drift/gather kernels" if fabric
                                else "NCCL") if world > 1 else "none"),
            "md_steps_per_s": args.steps / (dev_ms * 1e-3),
            "wall_ms_per_step": 1e3 * tm["wall_s"] / args.steps,
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "pairs/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": 1e3 * e2e_s / e2e_steps, "steps": e2e_steps,
                    "call": "ljmd_integrate_host (every rank uploads its shard of pos+vel from pinned host memory, "
                            "Integrate, downloads its shard of pos+vel and the scalars)"},
            "gpu_launches": int(tm["launches"]),
            "roofline": roofline,
            "roofline_hbm": roofline_hbm,
            "parity": parity,
            "launch": info,
            "state": {"U_per_N": sc["U"] / N, "T": sc["T"], "P": sc["P"]},
            "extra": {"sweep": sweep},
        }
        if world == 1 and not args.no_cpu_baseline:
            n_s = 16384 if N >= 16384 else N
            r = run_reference_sample(pkg, cfg, n_s, 1 if n_s >= 8192 else 20, 0)
            line["cpu_baseline"] = {
                "value": r["value"], "unit": "pairs/s", "cores": 1, "kind": r["kind"],
                "sample": f"{r['steps']} x Integrate(dt) at N={n_s} (same rho*, T*, boundary, ensemble; the reference "
                          f"cost is N(N-1) pair evaluations per step); {host_description()}"}
            # the reference's own CUDA kernel on this GPU (the bar row a-3' of SURVEY.md 8 names)
            line["reference_gpu"] = reference_gpu_leg()
        emit(line)
    D.finalize()
    return 0


def main():
    # The contract is ONE JSON line on stdout.  Libraries loaded below write there too (NCCL prints its version
    # banner on stdout at communicator creation), so file descriptor 1 is pointed at stderr for the run and the
    # JSON line goes to the saved descriptor.
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--config", default="C5", choices=["C1", "C2", "C3", "C4", "C5"])
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the post-run FP64-arbiter check")
    ap.add_argument("--sweep", dest="sweep", action="store_true", default=None,
                    help="also time short runs of C3 and C4 (default: on for the default workload C5)")
    ap.add_argument("--no-sweep", dest="sweep", action="store_false")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.sweep is None:
        args.sweep = args.config == "C5"
    pkg = ljpkg.load()
    if args.impl == "reference":
        return main_reference(args, pkg)
    return main_ours(args, pkg)


if __name__ == "__main__":
    sys.exit(main())
