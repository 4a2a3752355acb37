"""Pin the linear-solver oracle (oracle/cg_port.hpp) against the reference's own unit-test
known answers: tk::CSR::mult (TestCSR.cpp:297-364) and the ConjugateGradients Laplace
solves on 1 and 2 partitions with 1 and 3 DOFs (TestConjugateGradients.cpp:120-532), with the
reference's tolerances; and, where oracle/_ref is available, against the reference's own
tk::CSR class bit for bit."""
import sys, os
import numpy as np
import pytest
import oraclelib as O
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import cg_cube as K

FLAVOURS = ["port"] + (["reference"] if O.lib("reference") is not None else [])


def serial(flavour, ncomp):
    s = O.CGOracle(flavour)
    s.add(K.INPOEL, 14, ncomp)
    s.laplacian(0, K.INPOEL, K.COORD)
    return s


@pytest.mark.parametrize("flavour", FLAVOURS)
def test_csr_mult_known_answer(flavour):
    s = O.CGOracle(flavour)
    s.add(K.CSR_INPOEL, 14, 1)
    s.laplacian(0, K.CSR_INPOEL, K.CSR_COORD)
    r = s.mult(0, np.arange(14.0))
    assert np.abs(r - K.MULT_IOTA).max() <= np.finfo(float).eps * 100      # TestCSR.cpp:355
    ia = s.get(0, "ia")
    assert ia[0] == 1 and len(ia) == 15 and len(s.get(0, "ja")) == ia[-1] - 1
    ja = s.get(0, "ja")
    for i in range(14):
        row = ja[int(ia[i]) - 1:int(ia[i + 1]) - 1]
        assert (np.diff(row.astype(np.int64)) > 0).all() and (i + 1) in row   # sorted, has diagonal


@pytest.mark.parametrize("flavour", FLAVOURS)
@pytest.mark.parametrize("ncomp", [1, 3])
def test_cg_serial_known_answer(flavour, ncomp):
    s = serial(flavour, ncomp)
    s.set(0, x=np.zeros(14 * ncomp), b=np.ones(14 * ncomp))
    for c in range(ncomp):
        s.dirichlet(0, 0, 0.0, c)
    k = K.KAT[ncomp]
    assert abs(s.setup() - k["normb"]) < 1e-12
    res, it = s.solve(k["maxit"], k["tol"])
    assert abs(res - k["normres"]) < 1e-12
    x = s.get(0, "x").reshape(14, ncomp)
    assert np.isfinite(x).all() and np.abs(x[0] - 1.0).max() < 1e-12      # identity row: x0 = b0
    assert np.abs(s.mult(0, x.reshape(-1)) - s.get(0, "b")).max() < 1e-12     # A x = b


@pytest.mark.parametrize("ncomp", [1, 3])
def test_cg_two_partitions_known_answer(ncomp):
    s = O.CGOracle("port")
    for m, P in enumerate(K.PART):
        s.add(P["inpoel"], len(P["gid"]), ncomp, P["gid"], P["comm"])
        s.laplacian(m, P["inpoel"], P["coord"])
        s.set(m, x=np.zeros(len(P["gid"]) * ncomp), b=np.ones(len(P["gid"]) * ncomp))
        for c in range(ncomp):
            s.dirichlet(m, int(np.where(P["gid"] == 0)[0][0]), 0.0, c)
    k = K.KAT[ncomp]
    assert abs(s.setup() - k["normb"]) < 1e-12
    res, it = s.solve(k["maxit"], k["tol"])
    assert abs(res - k["normres"]) < 1e-12
    # the partitioned solution equals the serial one on the shared numbering
    ref = serial("port", ncomp)
    ref.set(0, x=np.zeros(14 * ncomp), b=np.ones(14 * ncomp))
    for c in range(ncomp):
        ref.dirichlet(0, 0, 0.0, c)
    ref.setup(); ref.solve(k["maxit"], k["tol"])
    xs = ref.get(0, "x").reshape(14, ncomp)
    for m, P in enumerate(K.PART):
        xm = s.get(m, "x").reshape(-1, ncomp)
        assert np.abs(xm - xs[P["gid"].astype(np.int64)]).max() < 1e-12


@pytest.mark.skipif(O.lib("reference") is None, reason="oracle/_ref not available")
def test_port_csr_bit_identical_to_reference_class():
    a = serial("port", 3); b = serial("reference", 3)
    assert O.lib("reference").orc_cg_backend() == b"reference"
    assert np.array_equal(a.get(0, "a"), b.get(0, "a"))
    x = np.linspace(-1, 2, 42)
    assert np.array_equal(a.mult(0, x), b.mult(0, x))
    for s in (a, b):
        s.set(0, x=np.zeros(42), b=np.ones(42))
        for c in range(3):
            s.dirichlet(0, 0, 0.0, c)
        s.setup(); s.solve(1000, 1e-3)
    assert np.array_equal(a.get(0, "x"), b.get(0, "x"))
