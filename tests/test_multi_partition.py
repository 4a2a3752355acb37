"""The N>1 path: element partition + shared-node partial-sum exchange.

CPU (gloo, world_size 2): host-side logic -- partitions, symmetric comm maps, the volume and
boundary-normal exchanges -- against the oracle's serial multi-chare run with the same
element->partition map. GPU (nccl, 2 GPUs): the whole time loop with device packing + NCCL
send/recv + all-reduces against the same oracle run."""
import json
import os
import subprocess
import sys
import numpy as np
import pytest
import oraclelib as O

HERE = os.path.dirname(os.path.abspath(__file__))


def launch(mode, case, nsteps, tmp_path, world=2):
    out = str(tmp_path / "res")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="2")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(29400 + os.getpid() % 500),
           os.path.join(HERE, "mp_worker.py"), mode, case, str(nsteps), out]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    return [json.load(open("%s.%d.json" % (out, k))) for k in range(world)]


def oracle_for(case, part, world):
    kw = {**O.CASES, **O.LCASES, **O.CCASES, **O.HCASES, **O.ZCASES, **O.KCASES}[case]
    return O.Oracle(O.load_mesh(kw.get("mesh", case)), O.make_cfg(**kw), "port", nchare=world,
                    target=np.asarray(part, np.uint64))


def check_setup(res, o, world):
    for k in range(world):
        r = res[k]
        assert np.array_equal(np.asarray(r["gid"], np.uint64), o.get("gid", k))
        assert np.array_equal(np.asarray(r["v"]), o.get("v", k))                   # own volumes, bitwise
        # full nodal volumes: sums of the sharers' partial volumes
        assert np.abs(np.asarray(r["vol"]) - o.get("vol", k)).max() <= 4e-16 * o.get("vol", k).max()
        assert np.array_equal(np.asarray(r["dsupedge0"], np.uint64), o.get("dsupedge0", k))
        assert np.array_equal(np.asarray(r["dsupint0"]), o.get("dsupint0", k))     # partial integrals, bitwise
        assert np.array_equal(np.asarray(r["symbcnodes"], np.uint64), o.get("symbcnodes", k))
        sn = np.asarray(r["symbcnorms"]); so = o.get("symbcnorms", k)
        assert sn.shape == so.shape and (len(so) == 0 or np.abs(sn - so).max() < 1e-15)
        assert abs(r["meshvol"] - o.scalar("meshvol")) <= 1e-15 * o.scalar("meshvol")


@pytest.mark.parametrize("case", ["riecg_sod", "riecg_taylor_green", "zalcg_sod"])
def test_two_partitions_host_logic_gloo(case, tmp_path):
    res = launch("host", case, 0, tmp_path)
    o = oracle_for(case, res[0]["part"], 2)
    assert o.scalar("nchare") == 2
    assert len(o.get("commmap", 0)) > 2           # the partitions do share nodes
    check_setup(res, o, 2)


@pytest.mark.parametrize("case", ["chocg_poiseuille_damp2", "chocg_ldc", "lohcg_poiseuille_damp4"])
def test_two_partitions_projection_solver_setup_gloo(case, tmp_path):
    """The partition-level setup of the projection solvers (groundwork for their multi-GPU path; the device
    time loop of ChoCG/LohCG is single-partition so far): per-partition stride-5 / stride-4 edge integrals
    incl. the Laplacian term (partial sums on shared edges), Dirichlet masks and values of velocity and
    pressure, no-slip nodes and the partition's part of the pressure Poisson matrix -- bitwise equal to the
    oracle's chares on the same element partition."""
    res = launch("host", case, 0, tmp_path)
    o = oracle_for(case, res[0]["part"], 2)
    assert o.scalar("nchare") == 2 and len(o.get("commmap", 0)) > 2
    check_setup(res, o, 2)
    st = 4 if case.startswith("lohcg") else 5
    for k in range(2):
        r = res[k]
        def singles(e, d):
            e = np.asarray(e).reshape(-1, 2); d = np.asarray(d).reshape(-1, st)
            return sorted((tuple(a), tuple(b)) for a, b in zip(e.tolist(), d.tolist()))
        assert singles(r["dsupedge2"], r["dsupint2"]) == singles(o.get("dsupedge2", k), o.get("dsupint2", k))
        for n in ("plhs_ia", "plhs_ja", "noslipbcnodes"):
            assert np.array_equal(np.asarray(r[n], np.uint64), o.get(n, k)), n
        assert np.array_equal(np.asarray(r["plhs_a"]), o.get("plhs_a", k))
        for masks, vals, w in (("dirbcmasks", "dirbcval", (3 if st == 5 else 4) + 1), ("dirbcmaskp", "dirbcvalp", 2)):
            mo, ms = o.get(masks, k).reshape(-1, w), np.asarray(r[masks], np.uint64).reshape(-1, w)
            assert np.array_equal(mo[np.argsort(mo[:, 0])], ms[np.argsort(ms[:, 0])]), masks
            vo, vs = o.get(vals, k).reshape(-1, w), np.asarray(r[vals]).reshape(-1, w)
            assert np.array_equal(vo[np.argsort(vo[:, 0])], vs[np.argsort(vs[:, 0])]), vals


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["riecg_sod", "riecg_sedov", "riecg_taylor_green"])
def test_two_gpus_match_oracle_two_chares(case, tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    nsteps = 10
    res = launch("gpu", case, nsteps, tmp_path)
    o = oracle_for(case, res[0]["part"], 2)
    check_setup(res, o, 2)
    o.step(nsteps)
    d = o.diag()
    rows = np.asarray(res[0]["rows"])
    assert rows.shape == d.shape
    for c in range(1, d.shape[1]):
        assert np.abs(rows[:, c] - d[:, c]).max() <= 1e-12 * np.abs(d[:, c]).max(), c
    assert np.array_equal(np.asarray(res[1]["rows"]), rows)      # all ranks see the same reductions
    for k in range(2):
        U = np.asarray(res[k]["u"]); Uo = o.get("u", k)
        assert np.abs(U - Uo).max() <= 1e-12 * np.abs(Uo).max()
        assert res[k]["launches"] > 0


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["zalcg_sod", "zalcg_sedov", "kozcg_sod", "kozcg_taylor_green"])
def test_two_gpus_zalcg_match_oracle_two_chares(case, tmp_path):
    """ZalCG and KozCG on 2 GPUs: after each FCT pass the shared nodes' own sums travel over NCCL -- rhs and
    antidiffusive sums P+/- (summed; ZalCG::comrhs/comaec), allowed bounds Q+/- (max / min; comalw,
    ZalCG.cpp:1316-1325), limited sums (summed; comlim) -- against the oracle's 2-chare run."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    nsteps = 10
    res = launch("gpu", case, nsteps, tmp_path)
    o = oracle_for(case, res[0]["part"], 2)
    o.step(nsteps)
    d = o.diag()
    rows = np.asarray(res[0]["rows"])
    assert rows.shape == d.shape
    for c in range(1, d.shape[1]):
        assert np.abs(rows[:, c] - d[:, c]).max() <= 1e-12 * np.abs(d[:, c]).max() + 1e-300, c
    for k in range(2):
        U = np.asarray(res[k]["u"]); Uo = o.get("u", k)
        assert np.abs(U - Uo).max() <= 1e-12 * np.abs(Uo).max()


@pytest.mark.gpu
def test_two_gpus_laxcg_match_oracle_two_chares(tmp_path):
    """LaxCG (preconditioned update reads the primitives of the stage: shared nodes must be
    updated from the complete sums only) with steady-state local time stepping on 2 GPUs."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    case = "laxcg_bump"; nsteps = 10
    res = launch("gpu", case, nsteps, tmp_path)
    o = oracle_for(case, res[0]["part"], 2)
    o.step(nsteps)
    d = o.diag()
    rows = np.asarray(res[0]["rows"])
    assert rows.shape == d.shape
    for c in range(1, 8):
        assert np.abs(rows[:, c] - d[:, c]).max() <= 1e-11 * np.abs(d[:, c]).max(), c
    for k in range(2):
        U = np.asarray(res[k]["u"]); Uo = o.get("u", k)
        assert np.abs(U - Uo).max() <= 1e-11 * np.abs(Uo).max()


@pytest.mark.gpu
@pytest.mark.parametrize("ncomp", [1, 3])
def test_two_gpus_cg_known_answer(ncomp, tmp_path):
    """TestConjugateGradients.cpp tests 2 and 4 (2 PEs, 1 and 3 DOFs) with the SpMV halo sum,
    masked dots and shared-row averaging done on two GPUs over NCCL."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    sys.path.insert(0, os.path.join(HERE, "golden"))
    import cg_cube as K
    res = launch("cg", str(ncomp), 0, tmp_path)
    k = K.KAT[ncomp]
    # serial oracle solution for comparison
    o = O.CGOracle("port")
    o.add(K.INPOEL, 14, ncomp); o.laplacian(0, K.INPOEL, K.COORD)
    o.set(0, x=np.zeros(14 * ncomp), b=np.ones(14 * ncomp))
    for c in range(ncomp):
        o.dirichlet(0, 0, 0.0, c)
    o.setup(); o.solve(k["maxit"], k["tol"])
    xs = o.get(0, "x").reshape(14, ncomp)
    for r in res:
        assert abs(r["normb"] - k["normb"]) < 1e-12
        assert abs(r["res"] - k["normres"]) < 1e-12
        x = np.asarray(r["x"]).reshape(-1, ncomp)
        assert np.abs(x - xs[np.asarray(r["gid"], np.int64)]).max() < 1e-12
    assert res[0]["it"] == res[1]["it"]
