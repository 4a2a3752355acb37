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
