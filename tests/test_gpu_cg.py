"""GPU parity of the linear-solver path (CSR SpMV, masked dots, axpys = ConjugateGradients)
through the C ABI, against the reference's unit-test known answers and the CG oracle."""
import os
import sys
import numpy as np
import pytest
import oraclelib as O
import xyst_b200
from xyst_b200 import hostapi as H
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import cg_cube as K

pytestmark = pytest.mark.gpu


def device_from(o, part=0, ncomp=1):
    ctx = xyst_b200.Context()
    ctx.csr_upload(o.get(part, "ia"), o.get(part, "ja"), o.get(part, "a"), ncomp)
    return ctx


def test_csr_mult_known_answer():
    o = O.CGOracle("port")
    o.add(K.CSR_INPOEL, 14, 1)
    o.laplacian(0, K.CSR_INPOEL, K.CSR_COORD)
    ctx = device_from(o)
    r = ctx.csr_mult(np.arange(14.0))
    assert np.abs(r - K.MULT_IOTA).max() <= np.finfo(float).eps * 100          # TestCSR.cpp:355
    x = np.linspace(-3, 5, 14)
    assert np.abs(ctx.csr_mult(x) - o.mult(0, x)).max() <= 1e-14


@pytest.mark.parametrize("ncomp", [1, 3])
def test_cg_serial_known_answer(ncomp):
    o = O.CGOracle("port")
    o.add(K.INPOEL, 14, ncomp)
    o.laplacian(0, K.INPOEL, K.COORD)
    o.set(0, x=np.zeros(14 * ncomp), b=np.ones(14 * ncomp))
    for c in range(ncomp):
        o.dirichlet(0, 0, 0.0, c)
    k = K.KAT[ncomp]
    ctx = device_from(o, 0, ncomp)
    normb = ctx.cg_setup(np.zeros(14 * ncomp), o.get(0, "b"))
    assert abs(normb - k["normb"]) < 1e-12                                     # reference tolerance
    res, it = ctx.cg_solve(k["maxit"], k["tol"])
    assert abs(res - k["normres"]) < 1e-12
    o.setup(); ro, ito = o.solve(k["maxit"], k["tol"])
    assert it == ito
    assert np.abs(ctx.cg_x() - o.get(0, "x")).max() < 1e-12


@pytest.mark.parametrize("pc", ["none", "jacobi"])
def test_cg_box_laplacian_matches_oracle(pc):
    n = 10
    m = H.box_mesh(n, n, n)
    npn = m["coord"].shape[1]
    o = O.CGOracle("port", pc)
    o.add(m["tets"], npn, 1)
    o.laplacian(0, m["tets"], m["coord"])
    x, y, z = m["coord"]
    b = np.sin(np.pi * x) * np.cos(2 * np.pi * y) + z
    o.set(0, x=np.zeros(npn), b=b)
    for node in (0, npn // 2, npn - 1):                      # a few Dirichlet rows
        o.dirichlet(0, node, 0.25, 0)
    ctx = device_from(o)
    bb = o.get(0, "b")
    nb = ctx.cg_setup(np.zeros(npn), bb, pc)
    assert abs(nb - o.setup()) <= 1e-13 * nb
    res, it = ctx.cg_solve(500, 1e-10)
    ro, ito = o.solve(500, 1e-10)
    assert it == ito and it < 500
    xo = o.get(0, "x")
    assert np.abs(ctx.cg_x() - xo).max() <= 1e-10 * np.abs(xo).max()
    assert np.abs(ctx.csr_mult(ctx.cg_x()) - bb).max() <= 1e-8 * np.abs(bb).max()
