import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, ljpkg
from oracle.oracle import Oracle
from scipy.spatial import cKDTree
pkg=ljpkg.load(); o=Oracle()
for seed in (2,3):
    rng = np.random.Generator(np.random.PCG64(1000 + seed))
    N = int(rng.integers(2, 2600)); rho = float(10 ** rng.uniform(-2.3, 0.0)); bc = int(rng.integers(0, 3)); canonical = int(rng.integers(0, 2)); T = float(rng.uniform(0.5, 2.5))
    pos = pkg.snapshots.random_gas(N, rho, min_sep=0.85, seed=seed, periodic=bc == 0)
    vel = pkg.snapshots.velocities(N, T, seed=seed)
    with pkg.ljmd.LJSystem(N, T0=T, rho=rho, canonical=canonical, bc=bc) as s:
        s.set_state(pos, vel)
        p0,v0,f0=s.get_state()
        f64,fa,sc64=o.forces_f64(pos,s.L,bc); fr,scr,rdfr=o.forces(pos,s.L,bc,s.rdf_dr2)
        fg=f0[:,:3].astype(np.float64)
        ea=(np.abs(fg-f64).max(axis=1)/fa)
        for i in np.argsort(ea)[-3:]:
            d=pos[:,:3].astype(np.float64)-pos[i,:3].astype(np.float64)
            if bc==0: d-=s.L*np.rint(d/s.L)
            r=np.sqrt((d*d).sum(1)); r[i]=1e9
            print(seed,'i',i,'err',ea[i],'fa',fa[i],'F64',f64[i],'Fgpu',fg[i],'Fref',fr[i,:3],'rmin',np.sort(r)[:3],'pos',pos[i,:3])
