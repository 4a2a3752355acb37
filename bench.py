#!/usr/bin/env python3
"""Benchmark of the B200-native RieCG hot path: edge-updates/s (fp64 RHS + RK stage).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--n CELLS]

Workload (BASELINE.json configs[1]): RieCG Sedov blast on a synthetic structured box of
n^3 hexahedra split into 6 tets each (n=150: 20.25M tets, 3.44M nodes, 23.8M edges), fp64,
Rusanov flux, symmetry BCs on the three planes through the origin, cfl 0.5. With N>1 GPUs
(torchrun, one rank per GPU) every rank owns one n^3 box of a (2n,n,n)/(2n,2n,n)/(2n,2n,2n)
box -- weak scaling -- with NCCL halo exchange of shared-node partial sums.

One "step" = one 3-stage Runge-Kutta time step = 3 edge-updates per mesh edge
(edge_updates/s = E * nstage * nsteps / time). One JSON line is printed by
rank 0. `--impl reference` times the reference's own CPU implementation of the path
(oracle/_ref = its unmodified Physics sources under the serial driver restatement, or the
oracle port where _ref is unavailable) on all host cores, on a bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

B_STAGE = 138.0          # algorithmic bytes per edge-update, SURVEY.md 8(d): (12c+5)*8*N/E + 64
GAMMA = 5.0 / 3.0


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=150, help="box cells per side per GPU")
    ap.add_argument("--cpu-n", type=int, default=40, help="box cells per side of the CPU sample")
    ap.add_argument("--cpu-steps", type=int, default=6)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--exact-muscl", action="store_true")
    return ap.parse_args()


def box_dims(n, ngpu):
    m = {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}[ngpu]
    return n * m[0], n * m[1], n * m[2]


def box_edges(nx, ny, nz):
    return 7 * nx * ny * nz + 3 * (nx * ny + ny * nz + nx * nz) + nx + ny + nz


def sedov_cfg(make_cfg, h, **extra):
    # p0 such that p0 * V(origin node) = 4.13e-2 as in the reference's Sedov mesh series
    # (tests/regression/inciter/RieCG/Sedov/sedov.q comments); V(origin) = h^3/4 on a Kuhn box
    p0 = 4.13e-2 / (h ** 3 / 4.0)
    return make_cfg(problem="sedov", gamma=GAMMA, p0=p0, cfl=0.5, sym=(1, 3, 5), diag_iter=10 ** 9,
                    **extra)


# --------------------------------------------------------------------------------------
# clocks during the timed region (B200_PROFILING.md recipe)
# --------------------------------------------------------------------------------------
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown," \
        "clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index = index; self.rows = []; self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "20"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True); self.t.start()
        except Exception:
            self.p = None

    def _read(self):
        for line in self.p.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.p:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=2)
        except Exception:
            self.p.kill()
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i] == "Active" for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# --------------------------------------------------------------------------------------
# CPU arm: the oracle on host cores (test infrastructure used as the measured baseline;
# the only place outside tests/ and smoke() where oracle/ is executed)
# --------------------------------------------------------------------------------------
def _cpu_worker(args):
    n, steps, flavour = args
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oraclelib as O
    from host_common import host_mesh_to_oracle
    from xyst_b200 import hostapi as H
    L = 1.2 * n / 150.0
    m = H.box_mesh(n, n, n, L, L, L)                        # mesh generation only (host C++)
    cfg = sedov_cfg(O.make_cfg, L / n)
    o = O.Oracle(host_mesh_to_oracle(m), cfg, flavour)
    o.step(1)                                               # warm-up
    t0 = time.perf_counter()
    o.step(steps)
    return time.perf_counter() - t0


def cpu_arm(n, steps, procs):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oraclelib as O
    flavour = "reference" if O.lib("reference") is not None else "port"
    O.lib(flavour)
    if procs == 1:
        ts = [_cpu_worker((n, steps, flavour))]
    else:
        import multiprocessing as mp
        with mp.get_context("spawn").Pool(procs) as pool:
            ts = pool.map(_cpu_worker, [(n, steps, flavour)] * procs)
    E = box_edges(n, n, n)
    value = procs * E * 3 * steps / max(ts)
    return {"value": value, "unit": "edge-updates/s", "cores": procs, "kind": flavour,
            "sample": "RieCG Sedov, %d^3-cell box (%d tets, %d edges) per core, %d steps, "
                      "%d independent partition(s)" % (n, 6 * n ** 3, E, steps, procs),
            "seconds": max(ts)}


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    procs = max(1, len(os.sched_getaffinity(0)))
    t0 = time.perf_counter()
    vals = []
    for _ in range(max(1, min(a.warmup, 1)) + 1):          # one warm pass + one measured pass
        cb = cpu_arm(a.cpu_n, a.cpu_steps, procs)
        vals.append(cb)
    cb = vals[-1]
    nx, ny, nz = box_dims(a.n, a.gpus)
    line = {"impl": "reference", "metric": "edge-updates/sec (fp64 RHS+RK stage)", "value": cb["value"],
            "unit": "edge-updates/s", "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": 1e3 * cb["seconds"] / a.cpu_steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "RieCG Sedov blast, structured box %dx%dx%d cells (bounded CPU "
                                   "sample: %d^3 cells per core)" % (nx, ny, nz, a.cpu_n)},
            "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": "edge-updates/s", "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "wall_s": time.perf_counter() - t0}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------
def run_ours(a):
    import numpy as np
    import torch
    import torch.distributed as dist
    from xyst_b200 import hostapi as H, capi

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != a.gpus:
        if world == 1 and a.gpus > 1:
            raise SystemExit("launch with torchrun --nproc-per-node %d for --gpus %d" % (a.gpus, a.gpus))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the B200 path has no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ncclid = None
    if world > 1:
        idb = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            buf = (C.c_char * 128)()
            if capi.lib().xyst_comm_unique_id(buf) != 0:
                raise SystemExit(capi.lib().xyst_last_error().decode())
            idb = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).clone()
        idb = idb.cuda(); dist.broadcast(idb, 0)
        ncclid = bytes(idb.cpu().numpy().tobytes())

    n = a.n
    nx, ny, nz = box_dims(n, world)
    h = 1.2 / 150.0                                        # the 20M-tet box is [0,1.2]^3 with n=150
    cfg = sedov_cfg(H.make_cfg, h, exact_muscl=a.exact_muscl)
    t0 = time.perf_counter()
    s = H.Solver.box(cfg, nx, ny, nz, nx * h, ny * h, nz * h, nparts=world, part=rank)
    s.prepare()
    t_prep = time.perf_counter() - t0
    s.attach(local, world, rank, ncclid)
    ctx = s.ctx()
    s.setup()
    t_setup = time.perf_counter() - t0
    stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)   # a non-blocking stream (not the legacy default one)
    ctx.set_stream(stream.cuda_stream)                     # so that torch events see our kernels
    npoin = int(s.scalar("npoin")); nedge_local = ctx.nedge()
    E = box_edges(nx, ny, nz)                              # unique edges of the whole box

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def maxover(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident timing (the `value`) ----
    for _ in range(a.warmup):
        s.step(1, want_diag=False)
    for k in ("grad", "flux", "update"):
        ctx.kernel_time(k, reset=True)                     # switches per-kernel events on
    barrier()
    clocks = ClockSampler(local); clocks.start()
    l0 = ctx.launch_count()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(a.steps):
        s.step(1, want_diag=False)
    e1.record(stream)
    barrier()
    ms = maxover(e0.elapsed_time(e1))
    launches = ctx.launch_count() - l0
    clk = clocks.stop()
    kt = {k: ctx.kernel_time(k) for k in ("grad", "flux", "update")}
    value = E * 3 * a.steps / (ms * 1e-3)
    finite = bool(np.isfinite(s.get("u")).all())

    # ---- end to end through the C ABI with host buffers (the `e2e`) ----
    e2e = None
    if not a.no_e2e:
        U = torch.empty((npoin, 5), dtype=torch.float64, pin_memory=True)
        L = capi.lib()
        L.xyst_state_get(ctx.h, C.c_void_p(U.data_ptr()))
        nb = U.numel() * 8
        ke = max(2, min(a.steps, 5))
        barrier()
        e0.record(stream)
        for _ in range(ke):
            L.xyst_state_set(ctx.h, C.c_void_p(U.data_ptr()))     # host -> device, pinned
            s.step(1, want_diag=False)
            L.xyst_state_get(ctx.h, C.c_void_p(U.data_ptr()))     # device -> host
        e1.record(stream)
        barrier()
        ms2 = maxover(e0.elapsed_time(e1))
        e2e = {"value": E * 3 * ke / (ms2 * 1e-3), "unit": "edge-updates/s",
               "h2d_bytes_per_step": nb, "d2h_bytes_per_step": nb, "steps": ke,
               "ms_per_step": ms2 / ke,
               "call": "xyst_state_set(host U) + RieCG::step + xyst_state_get(host U)"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s (of fallback)"
    # dominant kernel: per-edge MUSCL+Riemann flux. Algorithmic bytes per launch (SURVEY.md 8d,
    # rhs row: read U c, G 3c, coord 3, write R c per node; normals 24 B + ids 8 B per edge)
    flux_ms, flux_n = kt["flux"]
    alg_flux = (5 + 15 + 3 + 5) * 8 * npoin + 32 * nedge_local
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get("flux_bytes_per_launch_n%d" % n)
    except Exception:
        pass
    ach = alg_flux / (flux_ms / max(flux_n, 1) * 1e-3) / 1e9 if flux_n else None
    roofline = {"bound": "hbm", "kernel": "k_flux_edge", "achieved": ach, "peak": peak, "unit": "GB/s",
                "frac": (ach / peak) if ach else None, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_flux, "avg_launch_ms": flux_ms / max(flux_n, 1),
                "launches_timed": flux_n}
    stage_gbs = value / world * B_STAGE / 1e9
    line = {"metric": "edge-updates/sec (fp64 RHS+RK stage)", "value": value, "unit": "edge-updates/s",
            "n_gpus": world, "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms / a.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": "RieCG Sedov blast, structured box %dx%dx%d cells = %d tets, %d nodes, "
                                   "%d edges; fp64, rusanov, sym BC on 3 planes, cfl 0.5"
                                   % (nx, ny, nz, 6 * nx * ny * nz, (nx + 1) * (ny + 1) * (nz + 1), E),
                       "per_gpu_cells": n, "partition": "%d x RCB box part" % world,
                       "l2_policy": "inputs larger than L2 (%.1f GB touched per stage)" % (B_STAGE * E / world / 1e9),
                       "muscl": "exact" if a.exact_muscl else "2-reciprocal form"},
            "roofline": roofline,
            "roofline_stage": {"model_bytes_per_edge_update": B_STAGE, "achieved_gbs_per_gpu": stage_gbs,
                               "frac_of_peak": stage_gbs / peak,
                               "kernel_ms_per_stage": {k: (v[0] / v[1] if v[1] else None) for k, v in kt.items()}},
            "clocks": clk, "e2e": e2e, "gpu_launches": launches,
            "setup_s": {"host_prepare": t_prep, "total": t_setup}, "finite": finite}
    if world == 1 and not a.no_cpu_baseline:
        line["cpu_baseline"] = cpu_arm(a.cpu_n, a.cpu_steps, 1)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


import ctypes as C  # noqa: E402

if __name__ == "__main__":
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
