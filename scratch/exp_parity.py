import sys, os
sys.path.insert(0,'tests'); sys.path.insert(0,'.')
import numpy as np
import oraclelib as O
from xyst_b200 import hostapi as H
from host_common import fixture_to_host_mesh
np.set_printoptions(linewidth=250, precision=2)
for case in O.CASES:
    gold = O.load_golden_diag(case); nsteps = int(gold[-1,0])
    for exact in (True, False):
        kw = dict(O.CASES[case], exact_muscl=exact)
        hm = fixture_to_host_mesh(O.load_mesh(case))
        s = H.Solver.mesh(H.make_cfg(**kw), hm["coord"], hm["tets"], hm["set_id"], hm["set_off"], hm["set_tri"])
        s.prepare(); s.attach(0); s.setup()
        rows = s.step(nsteps)
        o = O.Oracle(O.load_mesh(case), O.make_cfg(**O.CASES[case]), "port"); o.step(nsteps); d = o.diag()
        err = np.abs(rows-d).max(axis=0)/np.maximum(np.abs(d).max(axis=0),1e-300)
        U, Uo = s.get("u"), o.get("u")
        pw = np.abs(U-Uo).max(axis=0)/np.abs(Uo).max(axis=0)
        print(os.environ.get("XYST_B200_LIB","default"), case, "exact" if exact else "fast", "diag relerr", err[1:14], "pointwise", pw)
