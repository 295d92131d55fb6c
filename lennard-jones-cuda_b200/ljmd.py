"""ctypes binding of the C ABI in include/ljmd.h (lennard-jones-cuda_b200/csrc/libljmd.so).

This is the product path: there is no CPU fallback.  Importing works anywhere (so the host logic can
be tested without a GPU); constructing an `LJSystem` without the CUDA library or without a device
raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libljmd.so")

BC_PERIODIC, BC_HARDWALL, BC_NONE = 0, 1, 2
RDF_BINS = 256
S_U, S_T, S_K, S_V, S_P, S_PVIRIAL, S_TIME, S_L, S_AV_U_TOT, S_AV_T_TOT, S_AV_P_TOT, S_AV_ITERS, S_CHI, \
    S_TKIN_TRIAL = range(14)
S_COUNT = 16

# every symbol include/ljmd.h declares (tests check the library exports all of them)
API_SYMBOLS = [
    "ljmd_last_error", "ljmd_device_count", "ljmd_create", "ljmd_create_multi", "ljmd_create_distributed", "ljmd_nccl_unique_id",
    "ljmd_fabric_export", "ljmd_fabric_connect", "ljmd_destroy", "ljmd_rdf_dr2", "ljmd_set_canonical", "ljmd_set_boundary", "ljmd_set_T0", "ljmd_set_state",
    "ljmd_set_velocities", "ljmd_upload", "ljmd_init_state", "ljmd_get_state", "ljmd_step", "ljmd_integrate_host", "ljmd_compute_forces",
    "ljmd_get_scalars", "ljmd_get_pshear", "ljmd_device_arrays", "ljmd_reset_averaging", "ljmd_get_rdf", "ljmd_get_rdf_accum", "ljmd_velocity_histogram", "ljmd_subvolume_counts", "ljmd_velocity_subvolume_counts",
    "ljmd_trace_begin", "ljmd_trace_row_length", "ljmd_trace_read", "ljmd_trace_end",
    "ljmd_launch_count", "ljmd_set_event_timing", "ljmd_last_step_timing", "ljmd_last_gather_timing", "ljmd_last_reduce_timing", "ljmd_get_launch_info",
    "ljmd_image_threshold", "ljmd_plan", "ljmd_plan_newton3", "ljmd_set_l2_flush", "ljmd_fp32_peak_probe",
    # legacy seam (MDSystem.cpp:9-25)
    "allocateArray", "deleteArray", "copyArrayToDevice", "copyArrayFromDevice", "calculateNForces", "threadExit",
    "allocateNBodyArrays", "deleteNBodyArrays", "registerGLBufferObject", "unregisterGLBufferObject", "threadSync",
]

_lib = None


class LJMDError(RuntimeError):
    pass


def _preload_bundled_nccl():
    """libljmd.so needs `libnccl.so.2`; PyTorch ships its own, newer one under the same soname.  Whichever is
    loaded first serves both, and torch does not import against the system's older NCCL — so a process that loads
    this library before `import torch` would break torch.  Loading torch's copy first (when there is one) makes
    the import order irrelevant; the collectives this library calls exist in both."""
    import sys
    for base in sys.path:
        cand = os.path.join(base, "nvidia", "nccl", "lib", "libnccl.so.2")
        if os.path.exists(cand):
            try:
                C.CDLL(cand, mode=C.RTLD_GLOBAL)
            except OSError:
                pass
            return


def load_library(path=None):
    """Load libljmd.so; raise (never fall back) when it is missing."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise LJMDError(f"{p} not found: build it with __graft_entry__.build() (make -C lennard-jones-cuda_b200/csrc); "
                        "there is no CPU fallback")
    _preload_bundled_nccl()
    lib = C.CDLL(p, mode=C.RTLD_GLOBAL)
    vp, ip, dp, fp = C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_double), C.POINTER(C.c_float)
    lib.ljmd_last_error.restype = C.c_char_p
    lib.ljmd_rdf_dr2.restype = C.c_float
    lib.ljmd_rdf_dr2.argtypes = [C.c_int]
    lib.ljmd_launch_count.restype = C.c_longlong
    lib.ljmd_launch_count.argtypes = [vp]
    lib.ljmd_create.argtypes = [C.POINTER(vp), C.c_int, C.c_double, C.c_double, C.c_int, C.c_int, C.c_float, C.c_int]
    lib.ljmd_create_distributed.argtypes = lib.ljmd_create.argtypes + [C.c_int, C.c_int, vp]
    lib.ljmd_create_multi.argtypes = [C.POINTER(vp), C.c_int, C.c_double, C.c_double, C.c_int, C.c_int, C.c_float, ip, C.c_int]
    lib.ljmd_nccl_unique_id.argtypes = [vp]
    lib.ljmd_destroy.argtypes = [vp]
    lib.ljmd_fabric_export.argtypes = [vp, vp]
    lib.ljmd_fabric_connect.argtypes = [vp, vp]
    lib.ljmd_set_canonical.argtypes = [vp, C.c_int]
    lib.ljmd_set_boundary.argtypes = [vp, C.c_int]
    lib.ljmd_set_T0.argtypes = [vp, C.c_double]
    lib.ljmd_set_state.argtypes = [vp, vp, vp]
    lib.ljmd_set_velocities.argtypes = [vp, vp]
    lib.ljmd_init_state.argtypes = [vp, C.c_ulonglong]
    lib.ljmd_upload.argtypes = [vp, vp, vp]
    lib.ljmd_get_state.argtypes = [vp, vp, vp, vp]
    lib.ljmd_step.argtypes = [vp, C.c_double, C.c_int, C.c_int]
    lib.ljmd_integrate_host.argtypes = [vp, C.c_double, vp, vp, vp]
    lib.ljmd_compute_forces.argtypes = [vp, C.c_int]
    lib.ljmd_get_scalars.argtypes = [vp, dp]
    lib.ljmd_reset_averaging.argtypes = [vp]
    lib.ljmd_get_pshear.argtypes = [vp, dp]
    lib.ljmd_device_arrays.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]
    lib.ljmd_get_rdf.argtypes = [vp, ip]
    lib.ljmd_get_rdf_accum.argtypes = [vp, C.POINTER(C.c_longlong), ip, C.c_int]
    lib.ljmd_velocity_histogram.argtypes = [vp, C.c_double, C.c_int, ip]
    lib.ljmd_subvolume_counts.argtypes = [vp, C.c_int, C.c_double, ip, C.c_int, ip]
    lib.ljmd_velocity_subvolume_counts.argtypes = [vp, C.c_int, C.c_double, C.c_double, ip, C.c_int, ip]
    lib.ljmd_trace_begin.argtypes = [vp, C.c_int, ip, dp, dp, C.c_int]
    lib.ljmd_trace_row_length.argtypes = [vp, ip]
    lib.ljmd_trace_read.argtypes = [vp, C.c_int, ip, dp, C.POINTER(C.c_longlong), dp]
    lib.ljmd_trace_end.argtypes = [vp]
    lib.ljmd_set_event_timing.argtypes = [vp, C.c_int]
    lib.ljmd_last_step_timing.argtypes = [vp, dp, dp, ip]
    lib.ljmd_get_launch_info.argtypes = [vp, ip]
    lib.ljmd_last_gather_timing.argtypes = [vp, dp, ip, dp]
    lib.ljmd_last_reduce_timing.argtypes = [vp, dp, ip, dp]
    lib.ljmd_set_l2_flush.argtypes = [vp, C.c_longlong]
    lib.ljmd_image_threshold.restype = C.c_float
    lib.ljmd_image_threshold.argtypes = [C.c_double, C.c_int]
    lib.ljmd_plan.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, ip]
    lib.ljmd_plan_newton3.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, ip]
    lib.calculateNForces.argtypes = [vp, vp, fp, C.c_int, C.c_float, C.c_int, ip, C.c_float, C.c_int, C.c_int]
    lib.calculateNForces.restype = None
    lib.allocateArray.argtypes = [C.POINTER(vp), C.c_int]
    lib.allocateArray.restype = None
    lib.deleteArray.argtypes = [vp]
    lib.deleteArray.restype = None
    lib.copyArrayToDevice.argtypes = [vp, vp, C.c_int]
    lib.copyArrayToDevice.restype = None
    lib.copyArrayFromDevice.argtypes = [vp, vp, C.c_uint, C.c_int]
    lib.copyArrayFromDevice.restype = None
    lib.threadExit.restype = None
    if path is None:
        _lib = lib
    return lib


def _f4(a, N, name):
    a = np.ascontiguousarray(a, dtype=np.float32)
    if a.size != 4 * N:
        raise ValueError(f"{name} must hold 4*N = {4 * N} floats, got {a.size}")
    return a


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class LJSystem:
    """One MD system on one GPU (or one rank of a sharded system).

    Mirrors the hot-path surface of the reference's `MDSystem` (MDSystem.h:104-163): construct from
    (N, T0, rho, canonical, boundaryConditions), `set_state` a snapshot, `step`/`integrate`, read the
    scalars U, T, K, V, P and the RDF / speed histograms.
    """

    def __init__(self, N, T0=1.5, rho=0.2, canonical=False, bc=BC_PERIODIC, device=0, rank=0, world=1,
                 nccl_unique_id=None, rdf_dr2=None, devices=None):
        """devices=[d0, d1, ...]: ONE handle in ONE process drives all listed GPUs (ljmd_create_multi: i-shards
        over the devices, exchange over NVLink peer memory, no NCCL and no launcher).  rank/world/nccl_unique_id:
        one process per GPU (ljmd_create_distributed)."""
        self._lib = load_library()
        self._h = C.c_void_p()
        self.N, self.T0, self.rho, self.bc, self.canonical = int(N), float(T0), float(rho), int(bc), bool(canonical)
        self.rank, self.world = rank, world
        dr2 = self._lib.ljmd_rdf_dr2(self.N) if rdf_dr2 is None else rdf_dr2
        self.rdf_dr2 = float(dr2)
        self.devices = list(devices) if devices is not None else None
        if self.devices is not None:
            if world != 1:
                raise ValueError("devices=[...] (one process, many GPUs) excludes rank/world (one process per GPU)")
            arr = (C.c_int * len(self.devices))(*self.devices)
            rc = self._lib.ljmd_create_multi(C.byref(self._h), self.N, self.rho, self.T0, int(self.canonical), self.bc,
                                             C.c_float(dr2), arr, len(self.devices))
        elif world == 1:
            rc = self._lib.ljmd_create(C.byref(self._h), self.N, self.rho, self.T0, int(self.canonical), self.bc,
                                       C.c_float(dr2), device)
        else:
            uid = C.create_string_buffer(bytes(nccl_unique_id), 128)
            rc = self._lib.ljmd_create_distributed(C.byref(self._h), self.N, self.rho, self.T0, int(self.canonical),
                                                   self.bc, C.c_float(dr2), device, rank, world, uid)
        self._check(rc)
        self.L = self.scalars()["L"]

    # -- plumbing
    def _check(self, rc):
        if rc != 0:
            raise LJMDError(f"ljmd error {rc}: {self._lib.ljmd_last_error().decode()}")

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.ljmd_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    @staticmethod
    def nccl_unique_id():
        lib = load_library()
        buf = C.create_string_buffer(128)
        rc = lib.ljmd_nccl_unique_id(buf)
        if rc != 0:
            raise LJMDError(lib.ljmd_last_error().decode())
        return buf.raw

    # -- intra-node fabric (peer windows over NVLink instead of NCCL for the per-step exchange)
    def fabric_handle(self):
        buf = C.create_string_buffer(64)
        self._check(self._lib.ljmd_fabric_export(self._h, buf))
        return buf.raw

    def fabric_connect(self, handles):
        """handles: list of `world` 64-byte handles in rank order (every rank's fabric_handle())."""
        blob = b"".join(bytes(h) for h in handles)
        if len(blob) != 64 * self.world:
            raise ValueError("need one 64-byte handle per rank")
        self._check(self._lib.ljmd_fabric_connect(self._h, C.create_string_buffer(blob, len(blob))))

    # -- state
    def set_state(self, pos, vel):
        pos, vel = _f4(pos, self.N, "pos"), _f4(vel, self.N, "vel")
        self._check(self._lib.ljmd_set_state(self._h, _ptr(pos), _ptr(vel)))

    def init_state(self, seed):
        """Device-side SampleInitialConditions (start lattice + seeded Philox velocities at T0) and evaluation."""
        self._check(self._lib.ljmd_init_state(self._h, int(seed)))

    def upload(self, pos=None, vel=None):
        pos = _f4(pos, self.N, "pos") if pos is not None else None
        vel = _f4(vel, self.N, "vel") if vel is not None else None
        self._check(self._lib.ljmd_upload(self._h, _ptr(pos) if pos is not None else None,
                                          _ptr(vel) if vel is not None else None))

    def set_velocities(self, vel):
        vel = _f4(vel, self.N, "vel")
        self._check(self._lib.ljmd_set_velocities(self._h, _ptr(vel)))

    def get_state(self, pos=True, vel=True, force=True):
        out = []
        bufs = []
        for want in (pos, vel, force):
            bufs.append(np.empty((self.N, 4), dtype=np.float32) if want else None)
        self._check(self._lib.ljmd_get_state(self._h, *[(_ptr(b) if b is not None else None) for b in bufs]))
        out = tuple(b for b in bufs)
        return out

    def set_canonical(self, canonical):
        self.canonical = bool(canonical)
        self._check(self._lib.ljmd_set_canonical(self._h, int(self.canonical)))

    def set_boundary(self, bc):
        self.bc = int(bc)
        self._check(self._lib.ljmd_set_boundary(self._h, self.bc))

    def set_T0(self, T0):
        self.T0 = float(T0)
        self._check(self._lib.ljmd_set_T0(self._h, self.T0))

    # -- stepping
    def step(self, dt, nsteps=1, rdf_every=0):
        self._check(self._lib.ljmd_step(self._h, float(dt), int(nsteps), int(rdf_every)))

    def integrate_host(self, dt, pos, vel, force=None):
        """Drop-in single Integrate(dt) on caller-owned host arrays (updated in place)."""
        for name, a in (("pos", pos), ("vel", vel), ("force", force)):
            if a is None and name == "force":
                continue
            # raw pointers cross the C ABI: anything but a writable C-contiguous float32 [N,4] would corrupt memory
            if not (isinstance(a, np.ndarray) and a.dtype == np.float32 and a.flags.c_contiguous and a.flags.writeable
                    and a.size == 4 * self.N):
                raise ValueError(f"{name} must be a writable C-contiguous float32 array of {self.N} x 4 values")
        self._check(self._lib.ljmd_integrate_host(self._h, float(dt), _ptr(pos), _ptr(vel),
                                                  _ptr(force) if force is not None else None))

    def compute_forces(self, with_rdf=False):
        self._check(self._lib.ljmd_compute_forces(self._h, int(with_rdf)))

    # -- observables
    def scalars(self):
        buf = (C.c_double * S_COUNT)()
        self._check(self._lib.ljmd_get_scalars(self._h, buf))
        names = ["U", "T", "K", "V", "P", "Pvirial", "t", "L", "av_U_tot", "av_T_tot", "av_p_tot", "av_iters", "chi",
                 "Tkin_trial"]
        return {n: buf[i] for i, n in enumerate(names)}

    def pshear(self):
        """Shear stress P_xy (MDSystem.cpp:299,309,335,353), evaluated on demand."""
        out = C.c_double(0.0)
        self._check(self._lib.ljmd_get_pshear(self._h, C.byref(out)))
        return float(out.value)

    def device_arrays(self):
        """Device pointers (ints) of pos / vel / force (float4[N]) and the library's CUDA stream."""
        p, v, f, st = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p()
        self._check(self._lib.ljmd_device_arrays(self._h, C.byref(p), C.byref(v), C.byref(f), C.byref(st)))
        return dict(pos=p.value, vel=v.value, force=f.value, stream=st.value or 0)

    def reset_averaging(self):
        self._check(self._lib.ljmd_reset_averaging(self._h))

    def rdf_counts(self):
        out = np.zeros(RDF_BINS, dtype=np.int32)
        self._check(self._lib.ljmd_get_rdf(self._h, out.ctypes.data_as(C.POINTER(C.c_int))))
        return out

    def rdf_accum(self, reset=False):
        out = np.zeros(RDF_BINS, dtype=np.int64)
        n = C.c_int(0)
        self._check(self._lib.ljmd_get_rdf_accum(self._h, out.ctypes.data_as(C.POINTER(C.c_longlong)), C.byref(n),
                                                 int(reset)))
        return out, n.value

    def velocity_histogram(self, step=0.12, nbins=101):
        out = np.zeros(nbins, dtype=np.int32)
        self._check(self._lib.ljmd_velocity_histogram(self._h, float(step), int(nbins),
                                                      out.ctypes.data_as(C.POINTER(C.c_int))))
        return out

    def subvolume_counts(self, type=3, alpha_step=0.05):
        """Cumulative counts of GetNSubsystemBatch (type 0/1/2 slab, 3 cube)."""
        out = np.zeros(128, dtype=np.int32)
        n = C.c_int(0)
        self._check(self._lib.ljmd_subvolume_counts(self._h, int(type), float(alpha_step),
                                                    out.ctypes.data_as(C.POINTER(C.c_int)), 128, C.byref(n)))
        return out[:n.value].copy()

    def velocity_subvolume_counts(self, type=2, vcut_max=3.0, alpha_step=0.05):
        """Cumulative counts of GetNsubVzBatch."""
        out = np.zeros(128, dtype=np.int32)
        n = C.c_int(0)
        self._check(self._lib.ljmd_velocity_subvolume_counts(self._h, int(type), float(vcut_max), float(alpha_step),
                                                             out.ctypes.data_as(C.POINTER(C.c_int)), 128, C.byref(n)))
        return out[:n.value].copy()

    # -- observation trace (per-step counters and scalars recorded on the device)
    def trace_begin(self, counters, capacity_steps):
        """counters: list of (kind, alpha_step[, vcut_max]); kind 0-2 slab x/y/z, 3 cube, 4-6 |vx|,|vy|,|vz|."""
        n = len(counters)
        kinds = (C.c_int * max(1, n))(*[int(c[0]) for c in counters])
        steps = (C.c_double * max(1, n))(*[float(c[1]) for c in counters])
        vcut = (C.c_double * max(1, n))(*[float(c[2]) if len(c) > 2 else 0.0 for c in counters])
        self._check(self._lib.ljmd_trace_begin(self._h, n, kinds, steps, vcut, int(capacity_steps)))
        self._trace_cap = int(capacity_steps)

    def trace_read(self):
        """-> dict(scalars [n,8] = t,U,T,P,K,V,Pvirial,0; counts [n,row] cumulative per counter; mean_velocity [n,3])."""
        row = C.c_int(0)
        self._check(self._lib.ljmd_trace_row_length(self._h, C.byref(row)))
        cap = self._trace_cap
        scal = np.zeros((cap, 8), dtype=np.float64)
        counts = np.zeros((cap, max(1, row.value)), dtype=np.int64)
        mv = np.zeros((cap, 3), dtype=np.float64)
        n = C.c_int(0)
        self._check(self._lib.ljmd_trace_read(self._h, cap, C.byref(n), scal.ctypes.data_as(C.POINTER(C.c_double)),
                                              counts.ctypes.data_as(C.POINTER(C.c_longlong)),
                                              mv.ctypes.data_as(C.POINTER(C.c_double))))
        k = n.value
        return dict(scalars=scal[:k].copy(), counts=counts[:k, :row.value].copy(), mean_velocity=mv[:k].copy())

    def trace_end(self):
        self._check(self._lib.ljmd_trace_end(self._h))

    # -- instrumentation
    def launch_count(self):
        return int(self._lib.ljmd_launch_count(self._h))

    def set_event_timing(self, on=True):
        self._check(self._lib.ljmd_set_event_timing(self._h, int(on)))

    def set_l2_flush(self, nbytes):
        self._check(self._lib.ljmd_set_l2_flush(self._h, int(nbytes)))

    def last_step_timing(self):
        f, t, n = C.c_double(0), C.c_double(0), C.c_int(0)
        self._check(self._lib.ljmd_last_step_timing(self._h, C.byref(f), C.byref(t), C.byref(n)))
        return dict(force_ms=f.value, total_ms=t.value, force_launches=n.value)

    def last_gather_timing(self):
        ms, n, b = C.c_double(0), C.c_int(0), C.c_double(0)
        self._check(self._lib.ljmd_last_gather_timing(self._h, C.byref(ms), C.byref(n), C.byref(b)))
        return dict(gather_ms=ms.value, launches=n.value, bytes_per_launch=b.value)

    def last_reduce_timing(self):
        ms, n, b = C.c_double(0), C.c_int(0), C.c_double(0)
        self._check(self._lib.ljmd_last_reduce_timing(self._h, C.byref(ms), C.byref(n), C.byref(b)))
        return dict(reduce_ms=ms.value, launches=n.value, bytes_per_launch=b.value)

    def launch_info(self):
        buf = (C.c_int * 8)()
        self._check(self._lib.ljmd_get_launch_info(self._h, buf))
        return dict(num_sms=buf[0], i_tile=buf[1], j_splits=buf[2], force_ctas=buf[3], world=buf[4], n_local=buf[5],
                    newton3=bool(buf[6]))


def image_threshold(L, k=1):
    return float(load_library().ljmd_image_threshold(float(L), int(k)))


def plan(N, rank=0, world=1, num_sms=148):
    """Launch plan of the force kernel (host logic only; no device needed)."""
    lib = load_library()
    buf = (C.c_int * 8)()
    rc = lib.ljmd_plan(int(N), int(rank), int(world), int(num_sms), buf)
    if rc != 0:
        raise LJMDError(lib.ljmd_last_error().decode())
    return dict(i_begin=buf[0], i_end=buf[1], i_tiles=buf[2], j_splits=buf[3], force_ctas=buf[4], i_tile=buf[5],
                newton3=bool(buf[6]), partner_blocks=buf[7])


def fp32_peak_probe(device=0):
    """Measured FP32 CUDA-core throughput (TFLOP/s) of `device`: a stream of independent packed FMAs."""
    lib = load_library()
    out = C.c_double(0.0)
    lib.ljmd_fp32_peak_probe.argtypes = [C.c_int, C.POINTER(C.c_double)]
    rc = lib.ljmd_fp32_peak_probe(int(device), C.byref(out))
    if rc != 0:
        raise LJMDError(lib.ljmd_last_error().decode())
    return float(out.value)


def plan_newton3(N, rank=0, world=1, num_sms=148):
    """Super-tile geometry of the Newton-3 kernel (host logic only; no device needed)."""
    lib = load_library()
    buf = (C.c_int * 8)()
    rc = lib.ljmd_plan_newton3(int(N), int(rank), int(world), int(num_sms), buf)
    if rc != 0:
        raise LJMDError(lib.ljmd_last_error().decode())
    return dict(bj=buf[0], mi=buf[1], mju=buf[2], nwin=buf[3], n_super=buf[4], win_shift=buf[5], nblk=buf[6], blk0=buf[7])


def rdf_curve(N, L, dr2, counts):
    """g(r) points from the r^2 histogram (MDSystem::RDF, MDSystem.cpp:633-649)."""
    n0 = N / L / L / L
    ir = np.arange(RDF_BINS)
    r = np.sqrt((ir + 0.5) * np.float64(np.float32(dr2)))
    g = (np.asarray(counts, dtype=np.float32) / np.float32(dr2)).astype(np.float64) / 2.0 / np.pi / r / n0 / float(N)
    return r, g
