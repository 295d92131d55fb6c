"""TEST INFRASTRUCTURE ONLY — ctypes wrappers of the CPU oracle.

  Oracle      the C restatement (oracle/ljmd_oracle.c -> oracle/libljmd_oracle.so)
  Reference   the UNMODIFIED reference CPU path (oracle/ref_shim.cpp + /root/reference sources ->
              oracle/_ref/libljmd_ref.so; built here, travels to the GPU box as a binary)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
The product path (lennard-jones-cuda_b200/) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_LIB = os.path.join(HERE, "libljmd_oracle.so")
REF_LIB = os.path.join(HERE, "_ref", "libljmd_ref.so")
REF_LEGACY_LIB = os.path.join(HERE, "_ref", "libljmd_ref_legacy.so")
HOST_SHIM_LIB = os.path.join(HERE, "libljmd_host_shim.so")   # the same shim over the PRODUCT's MDSystem class
RDF_BINS = 256


def build(quiet=True):
    """make -C oracle (compiles the restatement; and the reference when /root/reference exists)."""
    subprocess.run(["make", "-C", HERE], check=True, stdout=subprocess.DEVNULL if quiet else None)


def _f4(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class Oracle:
    """Stateless functions of the C restatement."""

    def __init__(self):
        if not os.path.exists(ORACLE_LIB):
            build()
        lib = C.CDLL(ORACLE_LIB)
        vp = C.c_void_p
        lib.ljo_rdf_dr2.restype = C.c_float
        lib.ljo_rdf_dr2.argtypes = [C.c_int]
        lib.ljo_box_length.restype = C.c_double
        lib.ljo_box_length.argtypes = [C.c_int, C.c_double]
        lib.ljo_forces.argtypes = [C.c_int, vp, C.c_double, C.c_int, C.c_float, vp, vp, vp]
        lib.ljo_forces_f64.argtypes = [C.c_int, vp, C.c_double, C.c_int, vp, vp, vp]
        lib.ljo_forces_f64_subset.argtypes = [C.c_int, vp, C.c_double, C.c_int, C.c_int, vp, C.c_int, vp, vp, vp]
        lib.ljo_kinetic_temperature.restype = C.c_double
        lib.ljo_kinetic_temperature.argtypes = [C.c_int, vp]
        lib.ljo_apply_boundary.argtypes = [C.c_int, vp, vp, C.c_double, C.c_int]
        lib.ljo_parameters.argtypes = [C.c_int, C.c_double, vp, C.c_double, C.c_double, C.c_double, vp]
        lib.ljo_integrate.argtypes = [C.c_int, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int, C.c_float,
                                      C.c_double, vp, vp, vp, vp, vp]
        lib.ljo_velocity_histogram.argtypes = [C.c_int, vp, C.c_double, C.c_int, vp]
        lib.ljo_rdf_curve.argtypes = [C.c_int, C.c_double, C.c_float, vp, vp, vp]
        lib.ljo_lattice.argtypes = [C.c_int, C.c_double, vp]
        self.lib = lib

    def rdf_dr2(self, N):
        return float(self.lib.ljo_rdf_dr2(N))

    def box_length(self, N, rho):
        return float(self.lib.ljo_box_length(N, rho))

    def forces(self, pos, L, bc, dr2):
        """-> frc[N,4] float32 (w=0), dict(V, Pvirial, Pshear_conf), rdf[256] int32"""
        pos = _f4(pos)
        N = pos.size // 4
        frc = np.zeros((N, 4), dtype=np.float32)
        scal = np.zeros(3, dtype=np.float64)
        rdf = np.zeros(RDF_BINS, dtype=np.int32)
        self.lib.ljo_forces(N, _p(pos), L, bc, C.c_float(dr2), _p(frc), _p(scal), _p(rdf))
        return frc, dict(V=scal[0], Pvirial=scal[1], Pshear_conf=scal[2]), rdf

    def forces_f64(self, pos, L, bc):
        """FP64 arbiter -> frc[N,3] float64, fabs_sum[N] float64 (sum_j |f_ij|), dict(V, Pvirial, Vabs, Pabs,
        fterm_sum[N] = sum_j (|repulsive| + |attractive| pair force))"""
        pos = _f4(pos)
        N = pos.size // 4
        frc = np.zeros((N, 3), dtype=np.float64)
        fa = np.zeros(2 * N, dtype=np.float64)
        scal = np.zeros(4, dtype=np.float64)
        self.lib.ljo_forces_f64(N, _p(pos), L, bc, _p(frc), _p(fa), _p(scal))
        return frc, fa[:N].copy(), dict(V=scal[0], Pvirial=scal[1], Vabs=scal[2], Pabs=scal[3], fterm_sum=fa[N:].copy())

    def forces_f64_subset(self, pos, L, bc, idx, threads=None):
        """FP64 arbiter for the particles `idx` only (O(len(idx) * N), host threads): -> frc[n,3] float64,
        fterm[n] (tolerance scale: 4 sum_j (12 r^-13 + 6 r^-7)), pe[n] = sum_j (r^-12 - r^-6), peabs[n] = sum_j (r^-12 + r^-6)."""
        pos = _f4(pos)
        N = pos.size // 4
        idx = np.ascontiguousarray(idx, dtype=np.int32)
        n = idx.size
        frc = np.zeros((n, 3), dtype=np.float64)
        fterm = np.zeros(n, dtype=np.float64)
        pe = np.zeros((n, 2), dtype=np.float64)
        if threads is None:
            threads = max(1, min(64, len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else os.cpu_count() or 1))
        self.lib.ljo_forces_f64_subset(N, _p(pos), L, bc, n, _p(idx), int(threads), _p(frc), _p(fterm), _p(pe))
        return frc, fterm, pe[:, 0].copy(), pe[:, 1].copy()

    def parameters(self, N, rho, vel, V, Pvirial, Pshear_conf=0.0):
        vel = _f4(vel)
        out = np.zeros(6, dtype=np.float64)
        self.lib.ljo_parameters(N, rho, _p(vel), V, Pvirial, Pshear_conf, _p(out))
        return dict(U=out[0], T=out[1], K=out[2], V=out[3], P=out[4], Pshear=out[5])

    def integrate(self, N, rho, T0, canonical, bc, dt, pos, vel, frc, nsteps=1, dr2=None):
        """nsteps x Integrate(dt) in place on copies; -> pos, vel, frc, scalars dict, rdf"""
        L = self.box_length(N, rho)
        dr2 = self.rdf_dr2(N) if dr2 is None else dr2
        pos, vel, frc = _f4(pos).copy(), _f4(vel).copy(), _f4(frc).copy()
        scal = np.zeros(6, dtype=np.float64)
        rdf = np.zeros(RDF_BINS, dtype=np.int32)
        for _ in range(nsteps):
            self.lib.ljo_integrate(N, rho, L, T0, int(canonical), bc, C.c_float(dr2), dt, _p(pos), _p(vel), _p(frc),
                                   _p(scal), _p(rdf))
        return pos, vel, frc, dict(U=scal[0], T=scal[1], K=scal[2], V=scal[3], P=scal[4], Pshear=scal[5]), rdf

    def velocity_histogram(self, vel, step=0.12, nbins=101):
        vel = _f4(vel)
        out = np.zeros(nbins, dtype=np.int32)
        self.lib.ljo_velocity_histogram(vel.size // 4, _p(vel), step, nbins, _p(out))
        return out

    def rdf_curve(self, N, L, dr2, rdf):
        rdf = np.ascontiguousarray(rdf, dtype=np.int32)
        r = np.zeros(RDF_BINS)
        g = np.zeros(RDF_BINS)
        self.lib.ljo_rdf_curve(N, L, C.c_float(dr2), _p(rdf), _p(r), _p(g))
        return r, g

    def lattice(self, N, L):
        pos = np.zeros((N, 4), dtype=np.float32)
        self.lib.ljo_lattice(N, L, _p(pos))
        return pos


def reference_available():
    return os.path.exists(REF_LIB)


class Reference:
    """The unmodified reference MDSystem behind oracle/ref_shim.cpp.

    legacy=False: its CPU path (the oracle).  legacy=True: the same host code built with
    -DUSE_CUDA_TOOLKIT and linked against the product's legacy C seam, i.e. the reference host layer
    driving the new kernels (needs a GPU; used by the drop-in tests only).
    """

    _libs = {}

    @classmethod
    def lib(cls, legacy=False):
        """legacy: False = reference CPU path, True = reference host layer on the product's legacy seam,
        "product" = the product's own MDSystem class (libljmd_host.so) behind the same shim."""
        if legacy not in cls._libs:
            path = HOST_SHIM_LIB if legacy == "product" else (REF_LEGACY_LIB if legacy else REF_LIB)
            if not os.path.exists(path):
                raise RuntimeError(f"{path} missing: run `make -C oracle ref ref_legacy` where /root/reference exists")
            lib = C.CDLL(path)
            vp = C.c_void_p
            lib.ljref_create.restype = vp
            lib.ljref_create.argtypes = [C.c_int, C.c_double, C.c_double, C.c_int, C.c_int]
            lib.ljref_destroy.argtypes = [vp]
            lib.ljref_set_state.argtypes = [vp, vp, vp]
            lib.ljref_get_state.argtypes = [vp, vp, vp, vp]
            lib.ljref_set_canonical.argtypes = [vp, C.c_int]
            lib.ljref_set_boundary.argtypes = [vp, C.c_int]
            lib.ljref_set_T0.argtypes = [vp, C.c_double]
            lib.ljref_integrate.argtypes = [vp, C.c_double, C.c_int]
            lib.ljref_calculate_forces.argtypes = [vp]
            lib.ljref_get_scalars.argtypes = [vp, vp]
            lib.ljref_get_rdf.restype = C.c_float
            lib.ljref_get_rdf.argtypes = [vp, vp]
            lib.ljref_rdf_curve.argtypes = [vp, vp, vp]
            lib.ljref_initvelo.argtypes = [vp, C.c_double, C.c_double]
            lib.ljref_updatevelo.argtypes = [vp]
            lib.ljref_getvelo.argtypes = [vp, vp, vp, C.c_int]
            lib.ljref_renormalize_to_energy.argtypes = [vp, C.c_double]
            lib.ljref_renormalize_velocities.argtypes = [vp, C.c_int]
            lib.ljref_kinetic_temperature.restype = C.c_double
            lib.ljref_kinetic_temperature.argtypes = [vp]
            if hasattr(lib, "ljref_correct_total_momentum"):
                lib.ljref_correct_total_momentum.argtypes = [vp]
                lib.ljref_poke_velocities.argtypes = [vp, vp]
            if hasattr(lib, "ljref_subsystem_batch"):
                lib.ljref_subsystem_batch.argtypes = [vp, C.c_double, C.c_int, vp, C.c_int]
                lib.ljref_velocity_batch.argtypes = [vp, C.c_double, C.c_double, C.c_int, vp, C.c_int]
                lib.ljref_subsystem.argtypes = [vp, C.c_double, C.c_int]
                lib.ljref_velocity_subsystem.argtypes = [vp, C.c_double, C.c_int]
            cls._libs[legacy] = lib
        return cls._libs[legacy]

    def __init__(self, N, T0, rho, canonical, bc, legacy=False):
        self.N = N
        self._l = self.lib(legacy)
        self._h = self._l.ljref_create(N, T0, rho, int(canonical), bc)

    def close(self):
        if self._h:
            self._l.ljref_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_state(self, pos, vel):
        pos, vel = _f4(pos), _f4(vel)
        self._l.ljref_set_state(self._h, _p(pos), _p(vel))

    def get_state(self):
        pos = np.empty((self.N, 4), dtype=np.float32)
        vel = np.empty((self.N, 4), dtype=np.float32)
        frc = np.empty((self.N, 4), dtype=np.float32)
        self._l.ljref_get_state(self._h, _p(pos), _p(vel), _p(frc))
        return pos, vel, frc

    def set_canonical(self, c):
        self._l.ljref_set_canonical(self._h, int(c))

    def set_boundary(self, bc):
        self._l.ljref_set_boundary(self._h, bc)

    def integrate(self, dt, nsteps=1):
        self._l.ljref_integrate(self._h, dt, nsteps)

    def calculate_forces(self):
        self._l.ljref_calculate_forces(self._h)

    def scalars(self):
        out = np.zeros(12)
        self._l.ljref_get_scalars(self._h, _p(out))
        names = ["U", "T", "K", "V", "P", "Pshear", "t", "L", "av_U_tot", "av_T_tot", "av_p_tot", "av_iters"]
        return dict(zip(names, out))

    def rdf_counts(self):
        out = np.zeros(RDF_BINS, dtype=np.int32)
        dr2 = self._l.ljref_get_rdf(self._h, _p(out))
        return out, float(dr2)

    def rdf_curve(self):
        r = np.zeros(RDF_BINS)
        g = np.zeros(RDF_BINS)
        n = self._l.ljref_rdf_curve(self._h, _p(r), _p(g))
        return r[:n], g[:n]

    def velocity_histogram(self, vmax=12.0, step=0.12):
        """initvelo(vmax, step) then read back: density values (counts / step / N)."""
        self._l.ljref_initvelo(self._h, vmax, step)
        v = np.zeros(4096)
        d = np.zeros(4096)
        n = self._l.ljref_getvelo(self._h, _p(v), _p(d), 4096)
        return v[:n], d[:n]

    def updatevelo(self):
        self._l.ljref_updatevelo(self._h)

    def _getvelo(self):
        v = np.zeros(4096)
        d = np.zeros(4096)
        n = self._l.ljref_getvelo(self._h, _p(v), _p(d), 4096)
        return v[:n], d[:n]

    def subsystem_batch(self, alpha_step=0.05, type=3):
        """The reference task helper GetNSubsystemBatch on this system's h_Pos."""
        out = np.zeros(256, dtype=np.int32)
        n = self._l.ljref_subsystem_batch(self._h, alpha_step, type, _p(out), 256)
        return out[:n].copy()

    def velocity_batch(self, vcut_max=3.0, alpha_step=0.05, type=2):
        """The reference task helper GetNsubVzBatch on this system's h_Vel."""
        out = np.zeros(256, dtype=np.int32)
        n = self._l.ljref_velocity_batch(self._h, vcut_max, alpha_step, type, _p(out), 256)
        return out[:n].copy()

    def renormalize_to_energy(self, ust):
        self._l.ljref_renormalize_to_energy(self._h, ust)

    def renormalize_velocities(self, recalculate=True):
        self._l.ljref_renormalize_velocities(self._h, int(recalculate))

    def correct_total_momentum(self):
        self._l.ljref_correct_total_momentum(self._h)

    def poke_velocities(self, vel):
        vel = _f4(vel)
        self._l.ljref_poke_velocities(self._h, _p(vel))

    def kinetic_temperature(self):
        return float(self._l.ljref_kinetic_temperature(self._h))
