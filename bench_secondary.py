#!/usr/bin/env python
"""Secondary measurements of SURVEY.md section 8(d), items 3 and 4 -- NOT the driver's bench
contract (that is bench.py): the limiter-heavy FCT solvers and the projection solver on the
50M-tet box (n = 203), one B200, timed on the device with CUDA events after warm-up, inputs
larger than L2. One JSON line per workload; results are kept under profiles/.

  ZalCG / KozCG : edge-updates/s (one stage per step) against the 223 B/edge-update model
  ChoCG         : step time, CG iterations/s, SpMV GB/s against 12 nnz + 20 N bytes per product
  LohCG         : step time, edge-updates/s (--only lohcg,lohcg_damp4)

    python bench_secondary.py [--n 203] [--steps 20]
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6540.8


def timed(stream, fn, reps):
    import torch
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(stream)
    for _ in range(reps):
        fn()
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def fct(solver, n, steps):
    import torch
    import bench
    from xyst_b200 import hostapi as H
    h = 1.2 / n
    cfg = bench.sedov_cfg(H.make_cfg, h, solver=solver, fctsys=(1, 2, 3, 4, 5))
    s = H.Solver.box(cfg, n, n, n, 1.2, 1.2, 1.2)
    s.prepare(); s.attach(0); ctx = s.ctx(); s.setup()
    stream = torch.cuda.Stream(); ctx.set_stream(stream.cuda_stream)      # a non-blocking stream, not the legacy default one
    s.step(3, want_diag=False)
    ms = timed(stream, lambda: s.step(1, want_diag=False), steps)
    E = bench.box_edges(n, n, n)
    v = E / (ms * 1e-3)
    model = 223.0
    return {"workload": "%s Sedov, %d^3-cell box = %d tets, %d edges, fct on, fctsys 1-5" % (solver, n, 6 * n ** 3, E),
            "metric": "edge-updates/sec (one stage per step)", "value": v, "ms_per_step": ms, "steps": steps,
            "roofline": {"bound": "hbm", "model_bytes_per_edge_update": model, "achieved": v * model / 1e9,
                         "peak": peak(), "unit": "GB/s", "frac": v * model / 1e9 / peak()},
            "finite": bool(np.isfinite(s.get("u")).all())}


def chocg(n, steps):
    import torch
    from xyst_b200 import hostapi as H
    kw = dict(solver="chocg", ncomp=3, cfl=0.5, flux="damp2", rk=1, mu=0.0, p_iter=100, p_tol=1.0e-6, p_pc="jacobi",
              p_hydrostat=0, problem="userdef", sym=(1, 2, 3, 4, 5, 6), nstep=10 ** 9)
    s = H.Solver.box(H.make_cfg(**kw), n, n, n)
    s.prepare(); s.host_setup()
    x, y = s.get("x"), s.get("y")
    u0 = np.stack([np.sin(np.pi * x) * np.cos(np.pi * y), -np.cos(np.pi * x) * np.sin(np.pi * y), np.zeros_like(x)], 1)
    s.set_u0(u0)
    s.attach(0); ctx = s.ctx(); s.setup()
    stream = torch.cuda.Stream(); ctx.set_stream(stream.cuda_stream)
    s.step(1, want_diag=False)
    ctx.kernel_time("spmv", reset=True)
    its = []
    def one():
        s.step(1, want_diag=False); its.append(int(s.scalar("pit")))
    ms = timed(stream, one, steps)
    sp_ms, sp_n = ctx.kernel_time("spmv")
    npoin = int(s.scalar("npoin")); nnz = len(s.get("plhs_ja"))
    spmv_bytes = 12.0 * nnz + 20.0 * npoin
    spmv_gbs = spmv_bytes / (sp_ms / max(sp_n, 1) * 1e-3) / 1e9
    return {"workload": "ChoCG, %d^3-cell box = %d tets, %d nodes, nnz %d; Taylor-Green-like velocity injected, "
                        "symmetry on all sides, hydrostat node 0, damp2, rk 1, CG jacobi tol 1e-6 max 100 it"
                        % (n, 6 * n ** 3, npoin, nnz),
            "ms_per_step": ms, "steps": steps, "cg_iterations_per_step": float(np.mean(its)),
            "cg_iterations_per_s": float(np.sum(its)) / (ms * steps * 1e-3),
            "spmv": {"avg_ms": sp_ms / max(sp_n, 1), "launches": sp_n, "bytes_per_product": spmv_bytes,
                     "achieved_gbs": spmv_gbs, "peak": peak(), "frac": spmv_gbs / peak()},
            "finite": bool(np.isfinite(s.get("u")).all())}


def lohcg(n, steps, flux):
    import torch
    import bench
    from xyst_b200 import hostapi as H
    rk = 2
    kw = dict(solver="lohcg", ncomp=4, cfl=0.3, flux=flux, rk=rk, mu=0.01, soundspeed=10.0, p_iter=40, p_tol=1.0e-3,
              p_pc="jacobi", p_hydrostat=0, problem="userdef", sym=(1, 2, 3, 4, 5, 6), nstep=10 ** 9)
    s = H.Solver.box(H.make_cfg(**kw), n, n, n)
    s.prepare(); s.host_setup()
    x, y = s.get("x"), s.get("y")
    u0 = np.stack([np.zeros_like(x), np.sin(np.pi * x) * np.cos(np.pi * y), -np.cos(np.pi * x) * np.sin(np.pi * y),
                   np.zeros_like(x)], 1)
    s.set_u0(u0)
    s.attach(0); ctx = s.ctx(); s.setup()
    stream = torch.cuda.Stream(); ctx.set_stream(stream.cuda_stream)
    s.step(2, want_diag=False)
    ctx.kernel_time("loh_rhs", reset=True)
    ms = timed(stream, lambda: s.step(1, want_diag=False), steps)
    r_ms, r_n = ctx.kernel_time("loh_rhs")
    E = bench.box_edges(n, n, n)
    return {"workload": "LohCG, %d^3-cell box = %d tets, %d edges; Taylor-Green-like velocity, symmetry on all sides, "
                        "%s, rk %d, soundspeed 10, mu 0.01" % (n, 6 * n ** 3, E, flux, rk),
            "metric": "edge-updates/sec (rk stages per step)", "value": E * rk / (ms * 1e-3), "ms_per_step": ms, "steps": steps,
            "rhs_kernel_avg_ms": r_ms / max(r_n, 1), "rhs_launches": r_n,
            "finite": bool(np.isfinite(s.get("u")).all())}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=203)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--only", default="zalcg,kozcg,chocg")
    a = ap.parse_args()
    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench_secondary.py: no CUDA device (the B200 path has no CPU fallback)")
    for w in a.only.split(","):
        if w == "chocg":
            r = chocg(a.n, max(2, a.steps // 4))
        elif w.startswith("lohcg"):
            r = lohcg(a.n, a.steps, "damp4" if w.endswith("damp4") else "damp2")
        else:
            r = fct(w, a.n, a.steps)
        print(json.dumps(r), flush=True)
