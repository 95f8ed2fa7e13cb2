"""The oracle pinned: reference known answers -> dgtd_ref (MFEM + reference integrators) -> golden vectors -> numpy oracle."""
import json
import os
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, golden_cases, load_golden, rel_l2
from oracle.dgtd_oracle import PEC, HesthavenOracle, build_ref_element, gll01

REF_BIN = os.path.join(ROOT, "oracle", "_ref", "dgtd_ref")
REF_MESH = "/root/reference/testData/mfemMeshes/2D/Maxwell2D_K2.mesh"


@pytest.mark.skipif(not (os.path.exists(REF_BIN) and os.path.exists(REF_MESH)), reason="needs oracle/_ref and /root/reference (build container)")
def test_reference_known_answer_matrices():
    """The nine M^-1*flux blocks of test/hesthavenComparison/Hesthaven2DTest.cpp:234-553, tolerance 1e-8 as there."""
    r = subprocess.run([REF_BIN, "known-answers", os.path.join(GOLDEN, "ref_known_answers.txt"), REF_MESH], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "9 checked, 0 failed" in r.stdout


def test_known_answer_fixture_is_complete():
    ka = json.load(open(os.path.join(GOLDEN, "ref_known_answers.json")))
    assert len(ka) == 9
    for name, rec in ka.items():
        m = np.array(rec["matrix"])
        assert m.shape == (6, 6) and rec["source"].startswith("test/hesthavenComparison/Hesthaven2DTest.cpp:")


def test_numpy_oracle_reproduces_reference_flux_blocks():
    """Same nine blocks from the matrix-free restatement: apply unit vectors through the flux terms only.

    M^-1 * ZeroNormal, OneNormal{x}, TwoNormal{x,y} are recovered from mult() on the 2-triangle mesh by linearity:
    with alpha = 1 and PEC, rhs(E_z <- H_x ...) columns are combinations of those blocks; here we check the three
    independent combinations that appear in the 2-D TM system against the literals.
    """
    from oracle.dgtd_oracle import Problem
    ka = json.load(open(os.path.join(GOLDEN, "ref_known_answers.json")))
    msh = json.load(open(os.path.join(GOLDEN, "Maxwell2D_K2.json")))
    verts = np.zeros((4, 3)); verts[:, :2] = np.array(msh["vertices"])
    pb = Problem(verts=verts, elems=np.array([e[2] for e in msh["elements"]]), elem_attr=np.array([1, 1]),
                 bdr=np.array([b[2] for b in msh["boundary"]]), bdr_attr=np.array([b[0] for b in msh["boundary"]]),
                 order=1, alpha=1.0, bdr_cond={2: PEC})
    O = HesthavenOracle(pb)
    N = O.N

    def block(row_c, col_c, alpha):
        pb.alpha = alpha
        B = np.zeros((N, N))
        for j in range(N):
            x = np.zeros(6 * N); x[col_c * N + j] = 1.0
            B[:, j] = O.mult(0.0, x)[row_c * N:(row_c + 1) * N]
        return B

    # rows E_z (c=2) <- H_x (c=3): directional(-D_y) + one-normal;  difference between alpha=1 and alpha=0 isolates penalties
    # E_z <- E_z : -(ZeroNormal) + TwoNormal_zz(=0 in 2-D)  => block = -M^-1 ZeroNormal_E
    ZN = -(block(2, 2, 1.0))
    assert np.abs(ZN - np.array(ka["2D_Operator_ZeroNormal_PEC"]["matrix"])).max() < 1e-8
    # H_x <- H_x : -ZeroNormal_H + TwoNormal_H{x,x};  H_x <- H_y : TwoNormal_H{x,y}
    TNxy = block(3, 4, 1.0)
    assert np.abs(TNxy - np.array(ka["2D_Operator_TwoNormal_nxHXny_HY_PEC"]["matrix"])).max() < 1e-8
    TNyx = block(4, 3, 1.0)
    assert np.abs(TNyx - np.array(ka["2D_Operator_TwoNormal_nyHYnx_HY_PEC"]["matrix"])).max() < 1e-8


@pytest.mark.parametrize("name", golden_cases())
def test_numpy_oracle_matches_reference_vectors(name):
    pb, dat = load_golden(name)
    O = HesthavenOracle(pb)
    meta = dat["meta"]
    assert np.abs(O.xyz.reshape(-1, 3) - dat["nodes_f64"].reshape(-1, 3)).max() < 1e-14
    k = O.mult(meta["t0"], dat["x0_f64"])
    assert rel_l2(k, dat["k0_f64"]) < 1e-12
    x, t = dat["x0_f64"].copy(), meta["t0"]
    for _ in range(meta["steps"]):
        x = O.rk4_step(x, t, meta["dt"]); t += meta["dt"]
    assert rel_l2(x, dat["x_final_f64"]) < 1e-12


def test_tfsf_fixtures_exercise_both_branches():
    assert load_golden("tfsf3d_p2_on")[1]["meta"]["tfsf_applied"] > 0
    m = load_golden("tfsf3d_p2_skipped")[1]["meta"]
    assert m["tfsf_applied"] == 0 and m["tfsf_skipped"] > 0


@pytest.mark.parametrize("dim,p", [(1, 1), (1, 4), (2, 2), (2, 5), (3, 1), (3, 3), (3, 4)])
def test_reference_element_identities(dim, p):
    r = build_ref_element(dim, p)
    x = r.nodes
    # derivatives exact on P_p, mass integrates to the simplex volume, face mass to the face measure
    for d in range(dim):
        f = np.prod(x ** np.arange(1, dim + 1)[None, :] % (p + 1), axis=1) if False else x[:, d] ** p
        assert np.abs(r.D[d] @ f - p * x[:, d] ** (p - 1)).max() < 1e-11
    M = np.linalg.inv(r.Minv)
    vol = {1: 1.0, 2: 0.5, 3: 1.0 / 6.0}[dim]
    assert abs(M.sum() - vol) < 1e-13
    fm = {1: 1.0, 2: 1.0, 3: 0.5}[dim]
    for f in range(dim + 1):
        assert abs((M @ r.lift[f]).sum() - fm) < 1e-12


def test_gll_points():
    g = gll01(3)
    assert np.allclose(g, [0.0, 0.5 - np.sqrt(5) / 10, 0.5 + np.sqrt(5) / 10, 1.0], atol=1e-15)


@pytest.mark.parametrize("name,alpha", [("box3d_p3_pec_upwind", 1.0), ("box3d_p2_mixed_centered", 0.0)])
def test_energy_properties(name, alpha):
    """Upwind PEC/PMC is dissipative, the centred flux conserves the discrete energy x^T M f(x) (M = |J| M_ref)."""
    pb, dat = load_golden(name)
    O = HesthavenOracle(pb)
    rng = np.random.default_rng(0)
    x = rng.standard_normal(6 * O.N)
    f = O.mult(0.0, x)
    M = np.linalg.inv(O.ref.Minv)
    xe, fe = x.reshape(6, O.NE, O.Np), f.reshape(6, O.NE, O.Np)
    e = np.einsum("cei,ij,cej,e->", xe, M, fe, O.detJ)
    scale = np.einsum("cei,ij,cej,e->", xe, M, xe, O.detJ)
    if alpha == 0.0:
        assert abs(e) < 1e-11 * scale * 100
    else:
        assert e < 0


@pytest.mark.parametrize("name", __import__("conftest").config_cases())
def test_numpy_oracle_matches_reference_on_baseline_configs(name):
    """BASELINE.json configs 3 and 4 on the reference's own meshes (22 400 / 15 886 tetrahedra): the portable oracle against
    every 61st entry of the reference operator's Mult output and of the state after two mfem::RK4Solver steps."""
    from conftest import initial_state, load_config_case
    pb, meta, smp = load_config_case(name)
    O = HesthavenOracle(pb)
    assert O.N == meta["n"]
    st = meta["stride"]
    x0 = initial_state(meta, O.xyz.reshape(-1, 3))
    assert rel_l2(x0[::st], smp["x0_sample_f64"]) < 1e-14
    k = O.mult(meta["t0"], x0)
    assert rel_l2(k[::st], smp["k0_sample_f64"]) < 1e-12
    assert abs(np.linalg.norm(k) / meta["k0_norm"] - 1.0) < 1e-12
    x, t = x0, meta["t0"]
    for _ in range(meta["steps"]):
        x = O.rk4_step(x, t, meta["dt"])
        t += meta["dt"]
    assert rel_l2(x[::st], smp["x_final_sample_f64"]) < 1e-12
    assert abs(np.linalg.norm(x) / meta["x_final_norm"] - 1.0) < 1e-12
    if meta.get("tfsf"):
        assert meta["tfsf_applied"] > 0          # the plane wave is on the TF/SF surface at t0


@pytest.mark.parametrize("name", [n for n in golden_cases() if not n.startswith("ibc")])
def test_hesthaven_flavour_equals_global_where_the_reference_says_so(name):
    """the `hesthaven` boundary encodings (HesthavenEvolution.cpp:275-313) and the `global` ones
    (DGOperatorFactory.h:483-568) give the same operator when alpha = 1, or when no SMA face is present; with SMA and
    alpha != 1 they differ by the centred part of the SMA faces scaled 1/alpha (HesthavenEvolution.cpp:292), nowhere else."""
    import dataclasses
    from oracle.dgtd_oracle import SMA
    pb, dat = load_golden(name)
    x = dat["x0_f64"]
    t0 = dat["meta"]["t0"]
    Og = HesthavenOracle(pb)
    has_sma = bool((Og.bc == SMA).any())
    if pb.alpha == 1.0 or not has_sma:
        Oh = HesthavenOracle(dataclasses.replace(pb, flavour="hesthaven"))
        assert rel_l2(Oh.mult(t0, x), Og.mult(t0, x)) < 1e-13
    for alpha in (0.5, 0.25):
        pa = dataclasses.replace(pb, alpha=alpha)
        Og, Oh = HesthavenOracle(pa), HesthavenOracle(dataclasses.replace(pa, flavour="hesthaven"))
        kg, kh = Og.mult(t0, x).reshape(6, Og.NE, Og.Np), Oh.mult(t0, x).reshape(6, Og.NE, Og.Np)
        if not has_sma:
            assert rel_l2(kh, kg) < 1e-13
            continue
        # expected difference: on SMA faces dE_h = -E/alpha instead of -E (same for H); the upwind part alpha*dU is the same
        # (-U in both, since global uses alpha = 1 there), the centred part n x dU picks up (1/alpha - 1) * (-U)
        uM = x.reshape(6, -1)[:, Og.vmapM]                                   # (6, NE, nf, Nfp)
        m = (Og.bc == SMA)[None, :, :, None]
        extra = np.where(m, -(1.0 / alpha - 1.0) * uM, 0.0)
        n = np.transpose(Og.normal, (2, 0, 1))[:, :, :, None]
        cross = lambda a, b: np.stack([a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]])
        sc = 0.5 * Og.fscale[None, :, :, None]
        dE = np.einsum("fij,cefj->cei", Og.ref.lift, cross(n, extra[3:]) * sc) * Og.inv_eps[None, :, None]
        dH = np.einsum("fij,cefj->cei", Og.ref.lift, -cross(n, extra[:3]) * sc) * Og.inv_mu[None, :, None]
        want = kg + np.concatenate([dE, dH])
        assert rel_l2(kh, want) < 1e-12
        untouched = ~(Og.bc == SMA).any(axis=1)
        assert np.abs(kh[:, untouched] - kg[:, untouched]).max() < 1e-12 * max(1.0, np.abs(kg).max())


def test_hesthaven_flavour_interior_boundaries_use_half_coefficients():
    """HesthavenEvolution.cpp:308-310: interior PEC/PMC/SMA jumps are (-1,0)/(0,-1)/(-1/2,-1/2) on both sides, alpha kept;
    `global` puts a true boundary on each side (jumps (-2,0)/(0,-2)/(-1,-1), SMA with alpha = 1).  For PEC/PMC sheets the
    hesthaven face flux is therefore exactly half the global one."""
    import dataclasses
    from oracle.dgtd_oracle import PMC, SMA
    pb, dat = load_golden("ibc3d_p2_pec_box")
    x = dat["x0_f64"]
    Og, Oh = HesthavenOracle(pb), HesthavenOracle(dataclasses.replace(pb, flavour="hesthaven"))
    assert Og.bc_interior.any() and (Og.bc[Og.bc_interior] == PEC).all()
    # reference operator without the sheet's contribution: zero the trace seen by the sheet faces -> flux of those faces only
    kg, kh = Og.mult(0.0, x), Oh.mult(0.0, x)
    pb0 = dataclasses.replace(pb, bdr_cond={k: v for k, v in pb.bdr_cond.items()})
    O0 = HesthavenOracle(pb0)
    O0.bc = O0.bc.copy(); O0.bc[O0.bc_interior] = 0                    # sheet faces: self-neighbour, no condition -> zero jump
    k0 = O0.mult(0.0, x)
    assert rel_l2(kh - k0, 0.5 * (kg - k0)) < 1e-12 and np.abs(kg - k0).max() > 1e-3
