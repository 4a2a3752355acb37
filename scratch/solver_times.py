"""Wall-clock per-step time of the three solvers through the host mirror on a Sedov box."""
import sys, time
from xyst_b200 import hostapi as H
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
for solver in ("riecg", "zalcg", "kozcg"):
    cfg = H.make_cfg(problem="sedov", gamma=5.0/3.0, p0=4.86e3, cfl=0.5, nstep=1000, term=1.0, sym=(1, 2, 3),
                     solver=solver, fctsys=(1, 2, 3, 4, 5), diag_iter=1000)
    s = H.Solver.box(cfg, n, n, n)
    t0 = time.time(); s.prepare(); s.attach(0); s.setup(); t1 = time.time()
    s.step(3)
    t2 = time.time(); s.step(20); t3 = time.time()
    print(solver, "n", n, "setup %.1fs" % (t1-t0), "ms/step %.3f" % ((t3-t2)/20*1e3), flush=True)
    del s
