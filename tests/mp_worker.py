"""Multi-process worker for the partitioned-path tests (launched by torch.distributed.run).

  mode=host : gloo, CPU only -- partition, comm maps, volume / boundary-normal exchange
              through the host mirror's communication hooks, compared with the oracle's
              multi-chare serial run (same element->partition map)
  mode=gpu  : nccl, one GPU per rank -- the full RieCG time loop with device-side packing +
              NCCL send/recv halo sums and all-reduces, compared with the same oracle run
"""
import os
import sys
import json
import numpy as np
import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE); sys.path.insert(0, os.path.dirname(HERE))
import oraclelib as O                                     # noqa: E402
from xyst_b200 import hostapi as H, capi                 # noqa: E402
from host_common import fixture_to_host_mesh             # noqa: E402


def gloo_comm(solver, rank, world):
    """Communication hooks on torch.distributed (any backend with all_gather_object)."""
    commmap = solver.get("commmap"); gid = solver.get("gid")
    shared = solver.get("shared")
    sh_gid = gid[shared.astype(np.int64)]

    def fn(op, w, n, a):
        if op == 0:
            vals = a.reshape(n, w).copy()
            out = [None] * world
            dist.all_gather_object(out, (sh_gid.tolist(), vals.tolist()))
            tot = vals.copy()
            # fixed order: ascending rank of the contributing partition
            for r in range(world):
                if r == rank:
                    continue
                g, v = out[r]
                pos = {x: i for i, x in enumerate(g)}
                for i, x in enumerate(sh_gid.tolist()):
                    j = pos.get(x)
                    if j is not None and _shares(commmap, r, x):
                        tot[i] += np.asarray(v[j])
            a[:] = tot.reshape(-1)
        else:
            t = torch.from_numpy(a.copy())
            dist.all_reduce(t, op=dist.ReduceOp.SUM if op == 1 else dist.ReduceOp.MIN)
            a[:] = t.numpy()
    return fn


def _shares(commmap, r, g):
    i = 0
    f = commmap.astype(np.int64)
    while i < len(f):
        b, n = int(f[i]), int(f[i + 1])
        if b == r:
            return g in set(f[i + 2:i + 2 + n].tolist())
        i += 2 + n
    return False


def cg_main(out, rank, world, local):
    """2-partition CG of the reference's unit test (TestConjugateGradients.cpp:236-355) on 2 GPUs."""
    import ctypes as C
    import xyst_b200
    sys.path.insert(0, os.path.join(HERE, "golden"))
    import cg_cube as K
    ncomp = int(sys.argv[2])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    P = K.PART[rank]
    o = O.CGOracle("port")
    o.add(P["inpoel"], len(P["gid"]), ncomp, P["gid"], P["comm"])
    o.laplacian(0, P["inpoel"], P["coord"])
    npn = len(P["gid"])
    o.set(0, x=np.zeros(npn * ncomp), b=np.ones(npn * ncomp))
    for c in range(ncomp):
        o.dirichlet(0, int(np.where(P["gid"] == 0)[0][0]), 0.0, c)       # CSR::dirichlet with 1/count diagonal
    ctx = xyst_b200.Context(device=local)
    idb = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        buf = (C.c_char * 128)()
        assert capi.lib().xyst_comm_unique_id(buf) == 0
        idb = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).cuda()
    dist.broadcast(idb, 0)
    L = capi.lib()
    assert L.xyst_comm_init(ctx.h, world, rank, bytes(idb.cpu().numpy().tobytes())) == 0
    other = 1 - rank
    sh_g = np.asarray(P["comm"][other], np.uint64)                       # ascending global ids
    lid = {int(g): i for i, g in enumerate(P["gid"])}
    sh = np.asarray([lid[int(g)] for g in sh_g], np.uint64)
    nr = np.asarray([other], np.int32); off = np.asarray([0, len(sh)], np.uint64)
    assert L.xyst_halo_upload(ctx.h, 1, nr.ctypes.data, off.ctypes.data, sh.ctypes.data) == 0
    ctx.csr_upload(o.get(0, "ia"), o.get(0, "ja"), o.get(0, "a"), ncomp)
    shared = set(int(g) for g in sh_g)
    slave = np.asarray([1 if (int(g) in shared and other < rank) else 0 for g in P["gid"]], np.uint8)
    count = np.asarray([2.0 if int(g) in shared else 1.0 for g in P["gid"]])
    k = K.KAT[ncomp]
    normb = ctx.cg_setup(np.zeros(npn * ncomp), o.get(0, "b"), "none", slave, count)
    res, it = ctx.cg_solve(k["maxit"], k["tol"])
    json.dump({"rank": rank, "normb": normb, "res": res, "it": it, "x": ctx.cg_x().tolist(),
               "gid": P["gid"].tolist()}, open("%s.%d.json" % (out, rank), "w"))
    dist.barrier()
    dist.destroy_process_group()


def nccl_id(rank):
    idb = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        import ctypes as C
        buf = (C.c_char * 128)()
        assert capi.lib().xyst_comm_unique_id(buf) == 0
        idb = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).cuda()
    dist.broadcast(idb, 0)
    return bytes(idb.cpu().numpy().tobytes())


def proj_main(cases, nsteps, out, rank, world, local):
    """ChoCG / LohCG on one GPU per partition: several cases in one launch (one NCCL communicator each). Per
    case and rank an .npz with the state after the start-up projection and after the last step, the
    diagnostics rows and the iteration counts of the linear solves of every step."""
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    tab = {**O.CCASES, **O.ICASES, **O.HCASES, **O.SCASES, "chocg_sphere_point_src": O.SPHERE_SRC}
    for case in cases.split(","):
        kw = tab[case]
        hm = fixture_to_host_mesh(O.load_mesh(kw.get("mesh", case)))
        part = H.rcb(hm["coord"], hm["tets"], world)
        s = H.Solver.mesh(H.make_cfg(**kw), hm["coord"], hm["tets"], hm["set_id"], hm["set_off"],
                          hm["set_tri"], nparts=world, part=rank, tetpart=part)
        s.prepare()
        s.attach(local, world, rank, nccl_id(rank))
        s.setup()
        cho = kw["solver"] == "chocg"
        res = {"part": np.asarray(part), "gid": s.get("gid"), "u0": s.get("u")}
        if cho:
            res["pr0"] = s.get("pr")
        rows, its = [], []
        for _ in range(nsteps if nsteps > 0 else int(kw["nstep"])):
            r = s.step(1)
            if len(r):
                rows.append(r[0])
            its.append([s.scalar("pit"), s.scalar("mit")])
        res["rows"] = np.asarray(rows); res["its"] = np.asarray(its); res["u"] = s.get("u")
        if cho:
            res["pr"] = s.get("pr")
        res["launches"] = np.asarray(s.ctx().launch_count())
        np.savez("%s.%s.%d.npz" % (out, case, rank), **res)
        del s
        dist.barrier()
    dist.destroy_process_group()


def main():
    mode, case, nsteps, out = sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4]
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    if mode == "cg":
        return cg_main(out, rank, world, local)
    if mode == "proj":
        return proj_main(case, nsteps, out, rank, world, local)
    kw = {**O.CASES, **O.LCASES, **O.CCASES, **O.HCASES, **O.ZCASES, **O.KCASES, **O.SCASES}[case]
    mesh = O.load_mesh(kw.get("mesh", case))
    hm = fixture_to_host_mesh(mesh)
    part = H.rcb(hm["coord"], hm["tets"], world)
    if mode == "gpu":
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        dist.init_process_group("gloo")
    s = H.Solver.mesh(H.make_cfg(**kw), hm["coord"], hm["tets"], hm["set_id"], hm["set_off"],
                      hm["set_tri"], nparts=world, part=rank, tetpart=part)
    s.prepare()
    res = {"rank": rank}
    if mode == "host":
        s.set_comm(gloo_comm(s, rank, world), world, rank)
        s.host_setup()
    else:
        idb = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            import ctypes as C
            buf = (C.c_char * 128)()
            assert capi.lib().xyst_comm_unique_id(buf) == 0
            idb = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).cuda()
        dist.broadcast(idb, 0)
        s.attach(local, world, rank, bytes(idb.cpu().numpy().tobytes()))
        s.setup()
        rows = s.step(nsteps)
        res["rows"] = rows.tolist()
        res["u"] = s.get("u").tolist()
        res["launches"] = s.ctx().launch_count()
    for n in ("gid", "vol", "v", "symbcnodes", "symbcnorms", "dsupint0", "dsupedge0"):
        res[n] = s.get(n).tolist()
    if kw.get("solver") in ("chocg", "lohcg"):        # partition-level pieces of the projection solvers
        for n in ("dsupint2", "dsupedge2", "dirbcmasks", "dirbcval", "dirbcmaskp", "dirbcvalp", "noslipbcnodes",
                  "plhs_ia", "plhs_ja", "plhs_a"):
            res[n] = s.get(n).tolist()
        if mode == "host":       # Dirichlet rows of the two linear solvers after the union over the sharers
            res["pbc"] = s.get("pbc").tolist()
            if kw.get("solver") == "chocg":
                res["mbcrows"] = s.get("mbcrows").tolist()
    res["meshvol"] = s.scalar("meshvol")
    res["part"] = part.tolist() if rank == 0 else None
    json.dump(res, open("%s.%d.json" % (out, rank), "w"))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
