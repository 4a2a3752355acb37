"""bench.py's own (numpy) box mesh and partition arithmetic -- used by the CPU arm and by the multi-GPU
parity check so that neither depends on product code -- against the host mirror's box generator."""
import numpy as np
import pytest
import bench
from xyst_b200 import hostapi as H
from host_common import host_mesh_to_oracle


@pytest.mark.parametrize("dims", [(3, 3, 3), (4, 2, 3), (6, 5, 2)])
def test_numpy_kuhn_box_is_the_host_mirrors_box(dims):
    ref = host_mesh_to_oracle(H.box_mesh(*dims, 1.2, 0.8, 1.0))
    mine = bench.kuhn_box(*dims, 1.2, 0.8, 1.0)
    for k in ref:
        a = np.asarray(ref[k]); b = np.asarray(mine[k]).reshape(a.shape)
        assert np.array_equal(a, b), k


def test_box_partition_ranges_and_target_map():
    for dims in ((64, 32, 32), (64, 64, 32), (64, 64, 64), (322, 322, 322), (300, 150, 150)):
        for n in (1, 2, 4, 8):
            for p in range(n):
                assert tuple(int(x) for x in H.box_part_range(*dims, n, p)) == bench.box_part_range(*dims, n, p)
    t = bench.box_target(8, 4, 4, 4)
    assert len(t) == 6 * 8 * 4 * 4 and sorted(set(t.tolist())) == [0, 1, 2, 3]
    assert np.bincount(t.astype(np.int64)).tolist() == [6 * 32] * 4       # equal parts
    assert bench.box_edges(150, 150, 150) == 23827950 and bench.box_edges(322, 322, 322) == 234637858


def test_workload_selection():
    a = bench.parse.__globals__["argparse"].Namespace(workload="auto", n=0)
    assert bench.workload_of(a, 1) == ("sedov", 150, (150, 150, 150))
    assert bench.workload_of(a, 8) == ("tg_strong", 322, (322, 322, 322))
    a.workload = "sedov_weak"
    assert bench.workload_of(a, 4) == ("sedov_weak", 150, (300, 300, 150))


@pytest.mark.parametrize("nparts", [2, 4, 8])
def test_parity_check_partition_equals_the_oracles_chares(nparts):
    """bench.py's multi-GPU parity check feeds the oracle the tet -> partition map of the box bisection:
    every rank's host-mirror partition must then hold exactly the nodes of the oracle's chare of the
    same number (CPU only: partitions built one after the other, no device, no communication)."""
    import oraclelib as O
    n = 4
    nx, ny, nz = bench.box_dims(n, nparts)
    h = 1.2 / 150.0
    kw = bench.sedov_kw(h)
    o = O.Oracle(bench.kuhn_box(nx, ny, nz, nx * h, ny * h, nz * h), O.make_cfg(**kw), "port", nchare=nparts,
                 target=bench.box_target(nx, ny, nz, nparts))
    assert o.scalar("nchare") == nparts
    for r in range(nparts):
        s = H.Solver.box(H.make_cfg(reforder=1, **kw), nx, ny, nz, nx * h, ny * h, nz * h, nparts=nparts, part=r)
        s.prepare()
        assert np.array_equal(s.get("gid").astype(np.uint64), o.get("gid", r)), r


def test_reference_flavour_holds_the_compiled_shim_and_it_refuses_to_run_without_a_gpu():
    """include/xyst_shim.hpp is compiled against the reference's headers into oracle/_ref; on a CPU-only
    machine the wrappers must fail loudly (no fallback), not compute."""
    import oraclelib as O
    import torch
    if O.lib("reference") is None:
        pytest.skip("oracle/_ref not built")
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: covered by tests/test_gpu_shim.py")
    kw = O.CASES["riecg_sod"]
    o = O.Oracle(O.load_mesh("riecg_sod"), O.make_cfg(**kw), "reference")
    with pytest.raises(RuntimeError, match="no CUDA device|not implemented|unknown"):
        o.kernel("shim_grad")
