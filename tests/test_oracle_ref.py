"""Pin the oracle restatement against the reference's OWN objects (no GPU).

oracle/_ref/liboracle_ref.so is the same serial driver but running the reference's
unmodified Physics/{Riemann,BC,Problems}.cpp and Mesh/{DerivedData,Reorder}.cpp
compiled in place from /root/reference (oracle/Makefile target `ref`), and the
reference's own SipHash-keyed containers. The self-contained restatement must be
BIT-identical to it: connectivity, superedges, integrals, and every nodal value
after the full regression run. Skipped where neither /root/reference nor a prebuilt
oracle/_ref exists.
"""
import numpy as np
import pytest
import oraclelib as O

ARRAYS = ["gid", "inpoel", "x", "y", "z", "vol", "v", "triinpoel", "besym", "bface",
          "dsupedge0", "dsupedge1", "dsupedge2", "dsupint0", "dsupint1", "dsupint2",
          "symbcnodes", "symbcnorms", "dirbcmasks", "u"]

needs_ref = pytest.mark.skipif(O.lib("reference") is None, reason="oracle/_ref not available")


@needs_ref
@pytest.mark.parametrize("case", list(O.CASES))
def test_port_is_bit_identical_to_reference_objects(case):
    mesh = O.load_mesh(case)
    gold = O.load_golden_diag(case)
    a = O.Oracle(mesh, O.make_cfg(**O.CASES[case]), "port")
    b = O.Oracle(mesh, O.make_cfg(**O.CASES[case]), "reference")
    assert O.lib("reference").orc_backend() == b"reference"
    for n in ARRAYS:
        assert np.array_equal(a.get(n), b.get(n)), n
    # one gradient + rhs evaluation, then the whole run
    for o in (a, b):
        o.kernel("grad")
    assert np.array_equal(a.get("grad"), b.get("grad"))
    for o in (a, b):
        o.kernel("rhs", 0, 0.0)
    assert np.array_equal(a.get("rhs"), b.get("rhs"))
    n = int(gold[-1, 0])
    a.step(n); b.step(n)
    assert np.array_equal(a.diag(), b.diag())
    assert np.array_equal(a.get("u"), b.get("u"))


@needs_ref
@pytest.mark.parametrize("flux", ["rusanov", "hllc"])
@pytest.mark.parametrize("stab2", [False, True])
def test_flux_variants_bit_identical(flux, stab2):
    case = "riecg_sod"
    mesh = O.load_mesh(case)
    kw = dict(O.CASES[case], flux=flux, stab2=stab2)
    a = O.Oracle(mesh, O.make_cfg(**kw), "port")
    b = O.Oracle(mesh, O.make_cfg(**kw), "reference")
    a.step(5); b.step(5)
    assert np.array_equal(a.get("u"), b.get("u"))
    assert np.isfinite(a.get("u")).all()


@needs_ref
def test_point_source_port_is_bit_identical_to_reference_objects():
    """problems::PHYS_SRC (point_src::src) from the reference's Problems.cpp vs the restatement."""
    kw = O.CANYON
    mesh = O.load_mesh(kw["mesh"])
    a = O.Oracle(mesh, O.make_cfg(**kw), "port")
    b = O.Oracle(mesh, O.make_cfg(**kw), "reference")
    a.step(20); b.step(20)
    assert np.array_equal(a.diag(), b.diag())
    assert np.array_equal(a.get("u"), b.get("u"))
    assert a.get("u").reshape(-1, 6)[:, 5].max() == 1.0          # the source has been applied
    kw = O.SPHERE_SRC
    mesh = O.load_mesh(kw["mesh"])
    a = O.Oracle(mesh, O.make_cfg(**kw), "port")
    b = O.Oracle(mesh, O.make_cfg(**kw), "reference")
    a.step(10); b.step(10)
    assert np.array_equal(a.diag(), b.diag())
    assert np.array_equal(a.get("u"), b.get("u"))


@needs_ref
@pytest.mark.parametrize("case", list(O.SCASES))
def test_scalar_transport_port_is_bit_identical_to_reference_objects(case):
    """problems::slot_cyl and the scalar parts of the Riemann/Zalesak/Kozak/Chorin/Lohner operators from the
    reference's own translation units vs the restatement, under the same drivers."""
    kw = O.SCASES[case]
    mesh = O.load_mesh(kw["mesh"])
    a = O.Oracle(mesh, O.make_cfg(**kw), "port")
    b = O.Oracle(mesh, O.make_cfg(**kw), "reference")
    assert np.array_equal(a.get("u"), b.get("u"))
    a.step(10); b.step(10)
    assert np.array_equal(a.diag(), b.diag())
    assert np.array_equal(a.get("u"), b.get("u"))


@needs_ref
@pytest.mark.parametrize("case", list(O.ZCASES) + list(O.ZSCASES))
def test_zalcg_port_is_bit_identical_to_reference_objects(case):
    """zalesak::rhs from the reference's own Zalesak.cpp vs the restatement, under the same
    serial FCT driver: bitwise equal states after the full regression runs."""
    kw = {**O.ZCASES, **O.ZSCASES}[case]
    gold = O.load_golden_diag(case)
    mesh = O.load_mesh(kw["mesh"])
    a = O.Oracle(mesh, O.make_cfg(**kw), "port")
    b = O.Oracle(mesh, O.make_cfg(**kw), "reference")
    for n in ["gid", "inpoel", "dsupedge0", "dsupint0", "dsupedge1", "dsupint1", "dsupedge2", "dsupint2"]:
        assert np.array_equal(a.get(n), b.get(n)), n
    assert len(a.get("dsupint0")) == len(a.get("dsupedge0")) // 4 * 6 * 4      # stride 4
    for o in (a, b):
        o.kernel("zrhs", 0, 0.0, 1.0e-3)
    assert np.array_equal(a.get("rhs"), b.get("rhs"))
    n = int(gold[-1, 0])
    a.step(n); b.step(n)
    assert np.array_equal(a.diag(), b.diag())
    assert np.array_equal(a.get("u"), b.get("u"))


@needs_ref
@pytest.mark.parametrize("case", list(O.KCASES) + list(O.KTCASES))
def test_kozcg_port_is_bit_identical_to_reference_objects(case):
    """kozak::rhs from the reference's own Kozak.cpp vs the restatement under the same driver."""
    kw = {**O.KCASES, **O.KTCASES}[case]
    gold = O.load_golden_diag(case)
    mesh = O.load_mesh(kw["mesh"])
    a = O.Oracle(mesh, O.make_cfg(**kw), "port")
    b = O.Oracle(mesh, O.make_cfg(**kw), "reference")
    for o in (a, b):
        o.kernel("krhs", 0, 0.0, 1.0e-3)
    assert np.array_equal(a.get("rhs"), b.get("rhs"))
    n = int(gold[-1, 0])
    a.step(n); b.step(n)
    assert np.array_equal(a.diag(), b.diag())
    assert np.array_equal(a.get("u"), b.get("u"))


@needs_ref
@pytest.mark.parametrize("case", list(O.LCASES))
def test_laxcg_port_is_bit_identical_to_reference_objects(case):
    """lax::grad/rhs/refvel from the reference's own Lax.cpp vs the restatement under the same driver."""
    kw = O.LCASES[case]
    mesh = O.load_mesh(kw["mesh"])
    a = O.Oracle(mesh, O.make_cfg(**kw), "port")
    b = O.Oracle(mesh, O.make_cfg(**kw), "reference")
    for o in (a, b):
        o.kernel("mindt")
        o.kernel("lgrad")
        o.kernel("lrhs", 0, 0.0)
    assert np.array_equal(a.get("grad"), b.get("grad"))
    assert np.array_equal(a.get("rhs"), b.get("rhs"))
    # lgrad left u in (p,u,v,w,T) form: step fresh instances
    a = O.Oracle(mesh, O.make_cfg(**kw), "port")
    b = O.Oracle(mesh, O.make_cfg(**kw), "reference")
    a.step(5); b.step(5)
    assert np.isfinite(a.diag()).all()
    assert np.array_equal(a.diag(), b.diag())
    assert np.array_equal(a.get("u"), b.get("u"))


@needs_ref
@pytest.mark.parametrize("case", ["chocg_poisson_neumann", "chocg_poiseuille_rk3", "chocg_poiseuille_damp4", "chocg_ldc",
                                  "chocg_poiseuille_theta"])
def test_chocg_port_is_bit_identical_to_reference_objects(case):
    """chorin::div/grad/vgrad/flux/rhs, tk::CSR (scalar and, for the semi-implicit momentum solve,
    3-component block rows) and the problem functions from the reference's own translation units vs
    the restatement under the same ChoCG driver."""
    kw = {**O.CCASES, **O.ICASES}[case]
    mesh = O.load_mesh(kw["mesh"])
    a = O.Oracle(mesh, O.make_cfg(**kw), "port")
    b = O.Oracle(mesh, O.make_cfg(**kw), "reference")
    n = min(kw["nstep"], 5)
    a.step(n); b.step(n)
    assert np.array_equal(a.diag(), b.diag())
    for name in ("u", "pr", "pgrad", "grad", "div"):
        assert np.array_equal(a.get(name), b.get(name)), name


@needs_ref
@pytest.mark.parametrize("case", ["riecg_vortical_flow_hllc_stab2", "riecg_vortical_flow_steady", "riecg_nleg",
                                  "riecg_rayleigh_taylor"])
def test_manufactured_problems_port_is_bit_identical_to_reference_objects(case):
    """problems::vortical_flow / nonlinear_energy_growth / rayleigh_taylor ic and src from the
    reference's Problems.cpp vs the restatement."""
    kw = {**O.VCASES, **O.TCASES}[case]
    mesh = O.load_mesh(kw["mesh"])
    a = O.Oracle(mesh, O.make_cfg(**kw), "port")
    b = O.Oracle(mesh, O.make_cfg(**kw), "reference")
    assert np.array_equal(a.get("u"), b.get("u"))
    a.step(10); b.step(10)
    assert np.array_equal(a.diag(), b.diag())
    assert np.array_equal(a.get("u"), b.get("u"))


@needs_ref
@pytest.mark.parametrize("case", list(O.HCASES))
def test_lohcg_port_is_bit_identical_to_reference_objects(case):
    """lohner::div/grad/vgrad/flux/rhs and physics::dirbcp from the reference's own translation units
    vs the restatement under the same LohCG driver."""
    kw = O.HCASES[case]
    mesh = O.load_mesh(kw["mesh"])
    a = O.Oracle(mesh, O.make_cfg(**kw), "port")
    b = O.Oracle(mesh, O.make_cfg(**kw), "reference")
    assert np.array_equal(a.get("u"), b.get("u"))
    assert np.array_equal(a.get("dsupint0"), b.get("dsupint0"))
    a.step(20); b.step(20)
    assert np.array_equal(a.diag(), b.diag())
    assert np.array_equal(a.get("u"), b.get("u"))
