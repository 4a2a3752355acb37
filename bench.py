#!/usr/bin/env python3
"""Benchmark of the B200-native RieCG hot path: edge-updates/s (fp64 RHS + RK stage).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload auto|sedov|tg_strong|sedov_weak] [--n CELLS] ...

Workloads (synthetic Kuhn boxes: n^3 hexahedra, 6 tets each):
  sedov      BASELINE.json configs[1]: RieCG Sedov blast, n=150 (20.25M tets, 3.44M nodes, 23.8M edges),
             Rusanov flux, symmetry BCs on the three planes through the origin, cfl 0.5. The default
             at N=1 (the configuration the metric is quoted on).
  tg_strong  BASELINE.json configs[4] / north_star: RieCG Taylor-Green with source term and Dirichlet
             BCs on all six sides on the 200M-tet box (n=322: 200.3M tets, 33.7M nodes, 234.6M edges),
             cfl 0.8, STRONG scaling: the same box cut into N partitions (one per GPU), NCCL halo
             exchange of shared-node partial sums. The default for N>1.
  sedov_weak round 1's scaling run: every rank owns one n^3 box of a (2n,n,n)/(2n,2n,n)/(2n,2n,2n) box.

One "step" = one 3-stage Runge-Kutta time step = 3 edge-updates per mesh edge (SURVEY.md 8d:
edge_updates/s = E * nstage * nsteps / time). Rank 0 prints ONE JSON line. With N>1 the line is
preceded by a multi-GPU parity check (a small box on the same N ranks against the CPU oracle's
N-chare run, 1e-12): key `parity_n`; a failure exits non-zero.

`--impl reference` times the reference's own CPU implementation of the path (oracle/_ref = its
unmodified Physics sources under the serial driver restatement, or the oracle port where _ref is not
available) with one partition ("chare") per host core, shared-node exchange between them, on a
bounded sample of the workload.
"""
import argparse
import ctypes as C
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
METRIC = "edge-updates/sec (fp64 RHS+RK stage)"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="auto", choices=["auto", "sedov", "tg_strong", "sedov_weak"])
    ap.add_argument("--n", type=int, default=0, help="box cells per side (sedov: 150; tg_strong: 322 for the "
                    "whole box; sedov_weak: 150 per GPU)")
    ap.add_argument("--reforder", type=int, default=-1, help="1: triangle superedges in the reference's hash order, "
                    "0: element order; default: 1 up to --reforder-max tets per partition")
    ap.add_argument("--reforder-max", type=int, default=30_000_000)
    ap.add_argument("--cpu-n", type=int, default=32, help="box cells per side PER CORE of the CPU sample")
    ap.add_argument("--cpu-steps", type=int, default=4)
    ap.add_argument("--cpu-cores", type=int, default=0, help="0: all cores this process may use")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-weak", action="store_true", help="N>1: skip the extra weak-scaling Sedov line")
    ap.add_argument("--parity-n", type=int, default=32, help="cells per side per rank of the multi-GPU parity box")
    ap.add_argument("--exact-muscl", action="store_true")
    ap.add_argument("--no-strong-base", action="store_true", help="N=1: skip the extra run of the 200M-tet "
                    "strong-scaling box on one GPU (key strong_scaling_base)")
    return ap.parse_args()


# --------------------------------------------------------------------------------------
# box arithmetic
# --------------------------------------------------------------------------------------
def box_dims(n, ngpu):
    m = {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}[ngpu]
    return n * m[0], n * m[1], n * m[2]


def box_edges(nx, ny, nz):
    return 7 * nx * ny * nz + 3 * (nx * ny + ny * nz + nx * nz) + nx + ny + nz


def box_part_range(nx, ny, nz, nparts, part):
    """Hex range of one part of the recursive coordinate bisection of a box (the same cuts as the
    host mirror's boxPartRange: halve the longest extent, lower half first)."""
    lo = [0, 0, 0]; hi = [nx, ny, nz]; np_, p = nparts, part
    while np_ > 1:
        dim = 0
        for d in (1, 2):
            if hi[d] - lo[d] > hi[dim] - lo[dim]:
                dim = d
        mid = lo[dim] + (hi[dim] - lo[dim]) // 2
        if p < np_ // 2:
            hi[dim] = mid
        else:
            lo[dim] = mid; p -= np_ // 2
        np_ //= 2
    return lo[0], hi[0], lo[1], hi[1], lo[2], hi[2]


def kuhn_box(nx, ny, nz, Lx, Ly, Lz):
    """The Kuhn-split box as plain numpy arrays (coord, tets, side-set triangles 1..6 = x-, x+, y-,
    y+, z-, z+), for the CPU arm and the parity check: no product code involved. Tets of hex
    (i,j,k) are the six monotone paths 000 -> 111 with positive Jacobian; hexes in (k,j,i) order."""
    import numpy as np
    px, py = nx + 1, ny + 1
    perm = [(0, 1, 2), (0, 2, 1), (1, 0, 2), (1, 2, 0), (2, 0, 1), (2, 1, 0)]
    kuhn = np.zeros((6, 4, 3), np.int64)
    for t, pm in enumerate(perm):
        v = np.zeros((4, 3), np.int64); v[3] = 1
        v[1, pm[0]] = 1; v[2, pm[0]] = 1; v[2, pm[1]] = 1
        if np.linalg.det((v[1:] - v[0]).astype(float)) < 0:
            v[[1, 2]] = v[[2, 1]]
        kuhn[t] = v
    k, j, i = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    base = np.stack([i.ravel(), j.ravel(), k.ravel()], 1)                      # hexes, i fastest
    c = base[:, None, None, :] + kuhn[None]                                    # [hex][6][4][3]
    tets = ((c[..., 2] * py + c[..., 1]) * px + c[..., 0]).reshape(-1, 4).astype(np.uint64)
    kk, jj, ii = np.meshgrid(np.arange(nz + 1), np.arange(ny + 1), np.arange(nx + 1), indexing="ij")
    coord = np.stack([Lx * ii.ravel() / nx, Ly * jj.ravel() / ny, Lz * kk.ravel() / nz])
    tf = [(0, 2, 1), (0, 1, 3), (0, 3, 2), (1, 2, 3)]
    sets = {}
    for sid, (dim, side, fixed) in enumerate([(0, 0, 0), (0, 1, nx - 1), (1, 0, 0), (1, 1, ny - 1),
                                              (2, 0, 0), (2, 1, nz - 1)], start=1):
        sel = base[base[:, dim] == fixed]
        tris = []
        for t in range(6):
            for f in tf:
                if all(kuhn[t, v, dim] == side for v in f):
                    cc = sel[:, None, :] + kuhn[t, list(f)][None]
                    tris.append((cc[..., 2] * py + cc[..., 1]) * px + cc[..., 0])
        sets[sid] = np.stack(tris, 1).reshape(-1, 3).astype(np.uint64)       # cell-major, as the host mirror's box
    sid = np.asarray(sorted(sets), np.int32)
    tri = np.concatenate([sets[s] for s in sid])
    off = np.cumsum([0] + [len(sets[s]) for s in sid]).astype(np.uint64)
    # oracle input: the triangles as a TRI block the side sets refer to (as in the reference's meshes)
    ntri = len(tri)
    return dict(coord=coord, tets=tets, tris=tri, block_type=np.array([0, 1], np.int32),
                block_n=np.array([ntri, len(tets)], np.uint64), set_id=sid, set_off=off,
                set_elem=np.arange(ntri, dtype=np.uint64), set_side=np.zeros(ntri, np.uint64))


def box_target(nx, ny, nz, nparts):
    """tet -> partition map of the box bisection (tets in hex order, 6 per hex)."""
    import numpy as np
    part = np.zeros((nz, ny, nx), np.uint64)
    for p in range(nparts):
        i0, i1, j0, j1, k0, k1 = box_part_range(nx, ny, nz, nparts, p)
        part[k0:k1, j0:j1, i0:i1] = p
    return np.repeat(part.ravel(), 6)


def sedov_kw(h, **extra):
    # p0 such that p0 * V(origin node) = 4.13e-2 as in the reference's Sedov mesh series
    # (tests/regression/inciter/RieCG/Sedov/sedov.q comments); V(origin) = h^3/4 on a Kuhn box
    return dict(problem="sedov", gamma=GAMMA, p0=4.13e-2 / (h ** 3 / 4.0), cfl=0.5, sym=(1, 3, 5),
                diag_iter=10 ** 9, **extra)


def sedov_cfg(make_cfg, h, **extra):
    return make_cfg(**sedov_kw(h, **extra))


def tg_kw(**extra):
    # tests/regression/inciter/RieCG/TaylorGreen/taylor_green.q: source term, Dirichlet BCs on all sides
    return dict(problem="taylor_green", gamma=GAMMA, cfl=0.8, diag_iter=10 ** 9,
                dir_=tuple((s, 1, 1, 1, 1, 1) for s in range(1, 7)), **extra)


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
            t0 = time.perf_counter()                      # nvidia-smi needs up to a second or two to start on an
            while not self.rows and time.perf_counter() - t0 < 8.0:      # 8-GPU box: wait for its first sample
                time.sleep(0.02)
            self.n0 = len(self.rows)
        except Exception:
            self.p = None

    def _read(self):
        for line in self.p.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.p:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        t0 = time.perf_counter()                          # a short timed region: make sure a sample under load is in
        while len(self.rows) <= getattr(self, "n0", 0) and time.perf_counter() - t0 < 0.2:
            time.sleep(0.005)
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
# CPU arm: the reference's kernels on host cores, one partition per core with the shared-node
# exchange between them. Test infrastructure used as the measured baseline: the
# only place outside tests/ and smoke() where oracle/ is executed. Nothing of xyst_b200 is imported.
# --------------------------------------------------------------------------------------
def host_cores(want=0):
    avail = len(os.sched_getaffinity(0))
    return max(1, min(avail, want) if want else avail), avail, os.cpu_count()


def cpu_arm(n, steps, cores):
    import numpy as np
    os.environ["OMP_NUM_THREADS"] = str(cores)           # before the oracle library starts its thread pool
    try:                                                  # (and if an OpenMP runtime is loaded already)
        C.CDLL("libgomp.so.1").omp_set_num_threads(cores)
    except OSError:
        pass
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oraclelib as O
    flavour = "reference" if O.lib("reference") is not None else "port"
    # cores = 2^a * odd: an (2n,n,n) ... box cut by coordinate bisection needs a power of two
    p2 = 1
    while p2 * 2 <= cores:
        p2 *= 2
    m = [1, 1, 1]; q = p2; d = 0
    while q > 1:
        m[d % 3] *= 2; q //= 2; d += 1
    nx, ny, nz = n * m[0], n * m[1], n * m[2]
    h = 1.2 / 150.0
    mesh = kuhn_box(nx, ny, nz, nx * h, ny * h, nz * h)
    t0 = time.perf_counter()
    o = O.Oracle(mesh, O.make_cfg(**sedov_kw(h)), flavour, nchare=p2, target=box_target(nx, ny, nz, p2))
    t_setup = time.perf_counter() - t0
    o.step(1)                                             # warm-up
    t0 = time.perf_counter()
    o.step(steps)
    sec = time.perf_counter() - t0
    E = box_edges(nx, ny, nz)
    assert np.isfinite(o.get("u")).all()
    same = None
    try:                                                  # the FULL 20M-tet box, one partition, one core: measured once
        same = json.load(open(os.path.join(ROOT, "profiles", "r2_cpu_full_n150.json")))      # (tools/cpu_full_n150.py)
        same["note"] = "stored measurement from the build container, not taken in this run"
    except Exception:
        pass
    return {"value": E * 3 * steps / sec, "unit": "edge-updates/s", "cores": p2, "kind": flavour,
            "same_config_record": same,
            "sample": "RieCG Sedov, %dx%dx%d-cell box (%d tets, %d edges) as %d partitions of %d^3 cells, one per "
                      "core (OpenMP over the partitions, shared-node partial sums exchanged after every sweep), "
                      "1 warm + %d timed steps" % (nx, ny, nz, 6 * nx * ny * nz, E, p2, n, steps),
            "steps_timed": steps, "seconds": sec, "setup_seconds": t_setup,
            "cores_available": len(os.sched_getaffinity(0)), "cores_machine": os.cpu_count()}


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores, avail, total = host_cores(a.cpu_cores)
    t0 = time.perf_counter()
    cb = cpu_arm(a.cpu_n, a.cpu_steps, cores)
    wl, n, (nx, ny, nz) = workload_of(a, a.gpus)
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"],
            "unit": "edge-updates/s", "n_gpus": a.gpus, "steps": cb["steps_timed"], "warmup": 1,
            "steps_requested": a.steps, "warmup_requested": a.warmup,
            "ms_per_step": 1e3 * cb["seconds"] / cb["steps_timed"], "higher_is_better": True,
            "scaling": "strong" if wl == "tg_strong" else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(wl, nx, ny, nz) + " -- CPU arm: bounded sample, " + cb["sample"]},
            "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": "edge-updates/s", "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "wall_s": time.perf_counter() - t0}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------
# workloads
# --------------------------------------------------------------------------------------
def workload_of(a, world):
    wl = a.workload
    if wl == "auto":
        wl = "sedov" if world == 1 else "tg_strong"
    if wl == "sedov":
        n = a.n or 150
        return wl, n, (n, n, n)
    if wl == "tg_strong":
        n = a.n or 322
        return wl, n, (n, n, n)
    n = a.n or 150
    return wl, n, box_dims(n, world)


def workload_name(wl, nx, ny, nz):
    E = box_edges(nx, ny, nz)
    what = {"sedov": "RieCG Sedov blast; rusanov, sym BC on 3 planes, cfl 0.5",
            "sedov_weak": "RieCG Sedov blast, weak scaling (one n^3 box per GPU); rusanov, sym BC on 3 planes, cfl 0.5",
            "tg_strong": "RieCG Taylor-Green with source term, Dirichlet BC on 6 sides, cfl 0.8, strong scaling "
                         "(one box cut into one partition per GPU); rusanov"}[wl]
    return "%s; structured box %dx%dx%d cells = %d tets, %d nodes, %d edges; fp64" % (
        what, nx, ny, nz, 6 * nx * ny * nz, (nx + 1) * (ny + 1) * (nz + 1), E)


def bind_to_gpu_numa(local):
    """Run this rank (and so first-touch its pinned buffers) on the cores of its GPU's NUMA node."""
    try:
        import torch
        bus = torch.cuda.get_device_properties(local).pci_bus_id
        dom = torch.cuda.get_device_properties(local).pci_domain_id
        dev = torch.cuda.get_device_properties(local).pci_device_id
        path = "/sys/bus/pci/devices/%04x:%02x:%02x.0/numa_node" % (dom, bus, dev)
        node = int(open(path).read().strip())
        if node < 0:
            return {"numa_node": node, "bound": False}
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = cpus & os.sched_getaffinity(0)
        if allowed:
            os.sched_setaffinity(0, allowed)
        return {"numa_node": node, "bound": bool(allowed), "cores": len(allowed)}
    except Exception as e:                                   # not fatal: placement is an optimisation
        return {"numa_node": None, "bound": False, "why": str(e)[:80]}


class Rig:
    """torch.distributed plumbing shared by the timed runs and the parity check."""

    def __init__(self, a):
        import torch
        import torch.distributed as dist
        from xyst_b200 import capi
        self.torch = torch; self.dist = dist; self.capi = capi
        self.rank = int(os.environ.get("RANK", "0")); self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if self.world != a.gpus and self.world == 1 and a.gpus > 1:
            raise SystemExit("launch with torchrun --nproc-per-node %d for --gpus %d" % (a.gpus, a.gpus))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device (the B200 path has no CPU fallback)")
        torch.cuda.set_device(self.local)
        self.numa = bind_to_gpu_numa(self.local)
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
        self.stream = torch.cuda.Stream(); torch.cuda.set_stream(self.stream)

    def nccl_id(self):
        """A fresh NCCL unique id for one solver's communicator, broadcast from rank 0."""
        if self.world == 1:
            return None
        torch = self.torch
        idb = torch.zeros(128, dtype=torch.uint8)
        if self.rank == 0:
            buf = (C.c_char * 128)()
            if self.capi.lib().xyst_comm_unique_id(buf) != 0:
                raise SystemExit(self.capi.lib().xyst_last_error().decode())
            idb = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).clone()
        idb = idb.cuda(); self.dist.broadcast(idb, 0)
        return bytes(idb.cpu().numpy().tobytes())

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def maxover(self, x):
        if self.world == 1:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def gather(self, x):
        if self.world == 1:
            return [x]
        out = [None] * self.world
        self.dist.all_gather_object(out, x)
        return out

    def solver(self, cfg, nx, ny, nz, Lx, Ly, Lz, comm=True):
        from xyst_b200 import hostapi as H
        t0 = time.perf_counter()
        if comm:
            s = H.Solver.box(cfg, nx, ny, nz, Lx, Ly, Lz, nparts=self.world, part=self.rank)
        else:
            s = H.Solver.box(cfg, nx, ny, nz, Lx, Ly, Lz)
        s.prepare()
        t_prep = time.perf_counter() - t0
        if comm:
            s.attach(self.local, self.world, self.rank, self.nccl_id())
        else:
            s.attach(self.local)
        s.setup()
        s.ctx().set_stream(self.stream.cuda_stream)       # a non-blocking stream torch events can see
        return s, t_prep, time.perf_counter() - t0


def parity_check(rig, a):
    """Multi-GPU correctness where the driver can see it: a small box on the SAME N ranks, 5 steps,
    against the CPU oracle's N-chare run with the same element -> partition map (RieCG.cpp:881-916,
    955-990: comgrad/comrhs). Every rank builds the oracle and checks its own partition pointwise;
    the all-reduced diagnostics rows are checked on all ranks. Tolerance 1e-12 relative."""
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oraclelib as O
    from xyst_b200 import hostapi as H
    world = rig.world
    n = a.parity_n
    nx, ny, nz = box_dims(n, world)
    h = 1.2 / 150.0
    kw = sedov_kw(h); kw["diag_iter"] = 1
    steps = 5
    s, _, _ = rig.solver(H.make_cfg(reforder=1, **kw), nx, ny, nz, nx * h, ny * h, nz * h)
    rows = s.step(steps)
    U = s.get("u"); gid = s.get("gid")
    o = O.Oracle(kuhn_box(nx, ny, nz, nx * h, ny * h, nz * h), O.make_cfg(**kw), "port", nchare=world,
                 target=box_target(nx, ny, nz, world))
    o.step(steps)
    d = o.diag()
    ok_gid = bool(np.array_equal(gid.astype(np.uint64), o.get("gid", rig.rank)))
    # nodal state of this rank's partition, pointwise per component
    Uo = o.get("u", rig.rank)
    scale = [max(float(np.abs(o.get("u", k)[:, c]).max()) for k in range(world)) for c in range(5)]   # of the whole mesh
    rel_u = max(float(np.abs(U[:, c] - Uo[:, c]).max() / scale[c]) for c in range(5)) if ok_gid else float("inf")
    # diagnostics (all-reduced over the ranks) against the exactly summed oracle state -- the reference's
    # own serial sums carry up to ~3e-17 x nodes of rounding, which they are held to as well
    import math
    Ua = np.concatenate([o.get("u", k) for k in range(world)]); va = np.concatenate([o.get("v", k) for k in range(world)])
    meshvol = o.scalar("meshvol")
    rel_rows = 0.0
    for c in range(5):
        exact = math.sqrt(math.fsum(Ua[:, c] ** 2 * va) / meshvol)
        rel_rows = max(rel_rows, abs(rows[-1, 3 + c] - exact) / exact)
    rel_serial = max(float(np.abs(rows[:, c] - d[:, c]).max() / np.abs(d[:, c]).max()) for c in list(range(1, 8)) + [13])
    if rel_serial > max(1.0e-12, 3.0e-17 * len(va)):
        rel_rows = max(rel_rows, rel_serial)
    res = rig.gather({"rank": rig.rank, "rows": rel_rows, "u": rel_u, "gid": ok_gid,
                      "launches": int(s.ctx().launch_count())})
    s.close()
    tol = 1.0e-12
    mx = max(max(r["rows"], r["u"]) for r in res)
    return {"max_rel": mx, "tol": tol, "ok": bool(mx <= tol and all(r["gid"] for r in res)),
            "what": "RieCG Sedov, %dx%dx%d-cell box on %d GPUs (NCCL halo sums + all-reduces), %d steps, vs the CPU "
                    "oracle's %d-chare run on the same partition: diag columns and every rank's nodal state, pointwise"
                    % (nx, ny, nz, world, steps, world),
            "per_rank": res}


def timed_run(rig, a, wl, n, dims, want_e2e):
    """Device-resident timing (`value`) and, if asked, the end-to-end timing (`e2e`) of one workload."""
    import numpy as np
    torch = rig.torch
    from xyst_b200 import hostapi as H, capi
    world = rig.world
    nx, ny, nz = dims
    per_part_tets = 6 * nx * ny * nz // world
    reforder = a.reforder if a.reforder >= 0 else (1 if per_part_tets <= a.reforder_max else 0)
    if wl == "tg_strong":
        Lx = Ly = Lz = 1.0
        cfg = H.make_cfg(reforder=reforder, exact_muscl=a.exact_muscl, **tg_kw())
    else:
        h = 1.2 / 150.0                                    # the 20M-tet box is [0,1.2]^3 with n=150
        Lx, Ly, Lz = nx * h, ny * h, nz * h
        cfg = sedov_cfg(H.make_cfg, h, reforder=reforder, exact_muscl=a.exact_muscl)
    s, t_prep, t_setup = rig.solver(cfg, nx, ny, nz, Lx, Ly, Lz)
    ctx = s.ctx()
    npoin = int(s.scalar("npoin")); nedge_local = ctx.nedge()
    E = box_edges(nx, ny, nz)                              # unique edges of the whole box
    stream = rig.stream
    clocks = ClockSampler(rig.local); clocks.start()       # (running before the warm-up: its start-up is slow)
    for _ in range(a.warmup):
        s.step(1, want_diag=False)
    for k in ("grad", "flux", "update"):
        ctx.kernel_time(k, reset=True)                     # switches per-kernel events on
    rig.barrier()
    clocks.rows.clear(); clocks.n0 = 0                     # samples of the timed region only
    l0 = ctx.launch_count()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(a.steps):
        s.step(1, want_diag=False)
    e1.record(stream)
    rig.barrier()
    ms = rig.maxover(e0.elapsed_time(e1))
    launches = ctx.launch_count() - l0
    clk = clocks.stop()
    kt = {k: ctx.kernel_time(k) for k in ("grad", "flux", "update")}
    value = E * 3 * a.steps / (ms * 1e-3)
    finite = bool(np.isfinite(s.get("u")).all())
    out = {"value": value, "ms": ms, "launches": launches, "clocks": clk, "kt": kt, "finite": finite, "E": E,
           "npoin": npoin, "nedge_local": nedge_local, "t_prep": t_prep, "t_setup": t_setup, "reforder": reforder,
           "e2e": None}
    # ---- end to end through the C ABI with host buffers ----
    if want_e2e:
        U = torch.empty((npoin, 5), dtype=torch.float64, pin_memory=True)
        L = capi.lib()
        L.xyst_state_get(ctx.h, C.c_void_p(U.data_ptr()))
        nb = U.numel() * 8
        ke = max(2, min(a.steps, 5))
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        th2d = td2h = tstep = 0.0
        rig.barrier()
        e0.record(stream)
        for _ in range(ke):
            ev[0].record(stream)
            L.xyst_state_set(ctx.h, C.c_void_p(U.data_ptr()))     # host -> device, pinned
            ev[1].record(stream)
            s.step(1, want_diag=False)
            ev[2].record(stream)
            L.xyst_state_get(ctx.h, C.c_void_p(U.data_ptr()))     # device -> host
            ev[3].record(stream)
            torch.cuda.synchronize()
            th2d += ev[0].elapsed_time(ev[1]); tstep += ev[1].elapsed_time(ev[2]); td2h += ev[2].elapsed_time(ev[3])
        e1.record(stream)
        rig.barrier()
        ms2 = rig.maxover(e0.elapsed_time(e1))
        link = rig.gather({"rank": rig.rank, "h2d_gbs": nb * ke / (th2d * 1e-3) / 1e9, "d2h_gbs": nb * ke / (td2h * 1e-3) / 1e9,
                           "step_ms": tstep / ke, "numa": rig.numa})
        out["e2e"] = {"value": E * 3 * ke / (ms2 * 1e-3), "unit": "edge-updates/s",
                      "h2d_bytes_per_step": nb, "d2h_bytes_per_step": nb, "steps": ke,
                      "ms_per_step": ms2 / ke,
                      "call": "xyst_state_set(host U) + RieCG::step + xyst_state_get(host U), per-rank bytes",
                      "per_rank_links": link}
    s.close()
    return out


def strong_base(a):
    """The N>1 workload (200M-tet Taylor-Green box) on ONE GPU, so that the scaling run has its own
    base: run in a child process (a host that cannot hold the 200M-tet mesh must not take the main line
    down with it); the device of this process is idle meanwhile."""
    try:
        avail = 0
        for ln in open("/proc/meminfo"):
            if ln.startswith("MemAvailable"):
                avail = int(ln.split()[1]) / 1e6
        if avail and avail < 100.0:
            return {"skipped": "%.0f GB of host memory available, the 200M-tet host mesh needs ~80 GB" % avail}
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--workload", "tg_strong", "--steps", "10",
                            "--warmup", str(a.warmup), "--no-cpu-baseline", "--no-e2e"],
                           capture_output=True, text=True, timeout=900)
        j = json.loads(r.stdout.strip().splitlines()[-1])
        return {"workload": j["config"]["workload"], "n_gpus": 1, "value": j["value"], "ms_per_step": j["ms_per_step"],
                "steps": j["steps"], "roofline_stage_frac": j["roofline_stage"]["frac_of_peak"],
                "setup_s": j["setup_s"], "finite": j["finite"], "superedge_order": j["config"]["superedge_order"]}
    except Exception as e:
        return {"failed": str(e)[:200]}


def run_ours(a):
    rig = Rig(a)
    world = rig.world
    wl, n, dims = workload_of(a, world)
    parity = None
    if world > 1 and not a.no_parity:
        parity = parity_check(rig, a)
        if not parity["ok"]:
            if rig.rank == 0:
                print(json.dumps({"metric": METRIC, "value": None, "n_gpus": world, "parity_n": parity,
                                  "error": "multi-GPU parity check failed"}), flush=True)
            rig.barrier()
            if world > 1:
                rig.dist.destroy_process_group()
            raise SystemExit(3)
    r = timed_run(rig, a, wl, n, dims, not a.no_e2e)
    weak = None
    if world > 1 and wl == "tg_strong" and not a.no_weak:
        w = timed_run(rig, a, "sedov_weak", 150, box_dims(150, world), False)
        weak = {"workload": workload_name("sedov_weak", *box_dims(150, world)), "value": w["value"],
                "ms_per_step": w["ms"] / a.steps, "scaling": "weak", "finite": w["finite"]}
    if rig.rank != 0:
        if world > 1:
            rig.dist.destroy_process_group()
        return

    nx, ny, nz = dims
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s (of fallback)"
    # dominant kernel: per-edge MUSCL+Riemann flux, k_flux_own. Algorithmic bytes per launch (SURVEY.md 8d,
    # rhs row: read U c, G 3c, coord 3, write R c per node; normals 24 B + ids 8 B per edge)
    kt = r["kt"]
    flux_ms, flux_n = kt["flux"]
    alg_flux = (5 + 15 + 3 + 5) * 8 * r["npoin"] + 32 * r["nedge_local"]
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get("flux_bytes_per_launch_n%d" % n)
    except Exception:
        pass
    ach = alg_flux / (flux_ms / max(flux_n, 1) * 1e-3) / 1e9 if flux_n else None
    roofline = {"bound": "hbm", "kernel": "k_flux_own", "achieved": ach, "peak": peak, "unit": "GB/s",
                "frac": (ach / peak) if ach else None, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_flux, "avg_launch_ms": flux_ms / max(flux_n, 1),
                "launches_timed": flux_n}
    value = r["value"]
    stage_gbs = value / world * B_STAGE / 1e9
    line = {"metric": METRIC, "value": value, "unit": "edge-updates/s",
            "n_gpus": world, "steps": a.steps, "warmup": a.warmup, "ms_per_step": r["ms"] / a.steps,
            "higher_is_better": True, "scaling": "strong" if wl == "tg_strong" else "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(wl, nx, ny, nz),
                       "partition": "%d x coordinate-bisection box part" % world,
                       "superedge_order": "reference hash walk (RieCG.cpp:646-701), the path the oracle tests cover"
                                          if r["reforder"] else "element-order triangle walk (same edges and integrals; "
                                          "covered at 1e-12 by tests with the oracle run on these superedges)",
                       "node_order": "caller's (the reference's renumbering); XYST_REORDER=1 selects the library's tile order",
                       "l2_policy": "inputs larger than L2 (%.1f GB touched per stage per GPU)" % (B_STAGE * r["E"] / world / 1e9),
                       "muscl": "exact" if a.exact_muscl else "2-reciprocal form"},
            "roofline": roofline,
            "roofline_stage": {"model_bytes_per_edge_update": B_STAGE, "achieved_gbs_per_gpu": stage_gbs,
                               "frac_of_peak": stage_gbs / peak,
                               "kernel_ms_per_stage": {k: (v[0] / v[1] if v[1] else None) for k, v in kt.items()}},
            "clocks": r["clocks"], "e2e": r["e2e"], "gpu_launches": r["launches"],
            "setup_s": {"host_prepare": r["t_prep"], "total": r["t_setup"]}, "finite": r["finite"],
            "numa": rig.numa}
    if parity is not None:
        line["parity_n"] = parity
    if weak is not None:
        line["weak_sedov"] = weak
    if world == 1 and not a.no_cpu_baseline:
        cores, _, _ = host_cores(a.cpu_cores)
        line["cpu_baseline"] = cpu_arm(a.cpu_n, a.cpu_steps, cores)
    if world == 1 and wl == "sedov" and a.workload == "auto" and not a.no_strong_base:
        line["strong_scaling_base"] = strong_base(a)
    print(json.dumps(line), flush=True)
    if world > 1:
        rig.dist.destroy_process_group()


if __name__ == "__main__":
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
