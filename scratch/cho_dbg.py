import sys
sys.path.insert(0,'tests'); sys.path.insert(0,'.')
import numpy as np
from xyst_b200 import hostapi as H
for n in (8, 40, 120, 203):
    kw = dict(solver="chocg", ncomp=3, cfl=0.9, flux="damp4", mu=0.01, p_iter=40, p_tol=1.0e-3, p_pc="jacobi",
              p_hydrostat=0, problem="userdef", noslip=(1, 2, 3, 5, 6), dir_=((4, 2, 2, 2),),
              dirval=((4, 1.0, 0.0, 0.0),), nstep=2)
    s=H.Solver.box(H.make_cfg(**kw), n,n,n); s.prepare(); s.attach(0); s.setup()
    u=s.get("u"); y=s.get("y"); lid=np.isclose(y,1.0)
    print(n, "after setup: lid u sum", u[lid].sum(axis=0), "all abs max", np.abs(u).max(), flush=True)
    rows=s.step(2)
    u=s.get("u")
    print(n, "after 2 steps: lid u sum", u[lid].sum(axis=0), "nlid", lid.sum(), "rows", rows[:, :8], "pit", s.scalar("pit"), flush=True)
    s.close()
