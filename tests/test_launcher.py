"""The JSON launcher (`python -m dgtd_b200.launcher -i case.json -d b200`, mirror of src/launcher/launcher.cpp and the keys of
src/driver/driver.cpp the hot path needs): parsing on the CPU, one run of BASELINE config 1 in the reference's own JSON format on
the GPU — the 1-D PEC cavity returns to its initial state after one period, as test/cases/CasesTest.cpp:15-136 asserts."""
import json
import os

import numpy as np
import pytest

from golden_io import read_fixture, write_mfem_mesh

CASE_1D_PEC = {     # testData/maxwellInputs/1D_PEC/1D_PEC.json with final_time = one period
    "solver_options": {"upwind_alpha": 1.0, "time_step": 0.005, "final_time": 2.0, "order": 3},
    "model": {"filename": "1D_PEC.mesh", "materials": [{"tags": [1], "type": "vacuum"}], "boundaries": [{"tags": [1, 2], "type": "PEC"}]},
    "sources": [{"type": "initial", "field_type": "electric", "center": [0.5], "polarization": [0.0, 1.0, 0.0], "dimension": 1,
                 "magnitude": {"type": "gaussian", "spread": 0.1}}],
}


def _write_case(tmp_path, case, fixture):
    arr, meta = read_fixture(fixture)
    write_mfem_mesh(str(tmp_path / case["model"]["filename"]), arr, meta)
    p = tmp_path / "case.json"
    p.write_text(json.dumps(case))
    return str(p)


def test_case_parsing_follows_the_reference_schema(tmp_path):
    import dgtd_b200 as dg
    from dgtd_b200.launcher import build_case
    path = _write_case(tmp_path, CASE_1D_PEC, "seg1d_config1_pec")
    mesh, kw, dt, t_final, init = build_case(json.load(open(path)), str(tmp_path), dg)
    assert mesh.dim == 1 and mesh.ne == 20 and kw["order"] == 3 and kw["alpha"] == 1.0 and dt == 0.005 and t_final == 2.0
    assert kw["bdr"] == {1: dg.BC_PEC, 2: dg.BC_PEC} and kw["planewave"] is None and kw["tfsf_gate"]
    xyz = np.zeros((5, 3)); xyz[:, 0] = np.linspace(0, 1, 5)
    x0 = init(xyz).reshape(6, 5)
    assert np.allclose(x0[1], np.exp(-(xyz[:, 0] - 0.5) ** 2 / (2 * 0.1 ** 2))) and not x0[[0, 2, 3, 4, 5]].any()
    # plane wave with the automatic delay of driver.cpp:576-589 on the TF/SF box fixture
    case = {"solver_options": {"order": 2, "time_step": 2e-3, "final_time": 0.01, "evolution_operator": "hesthaven"},
            "model": {"filename": "box.mesh", "boundaries": [{"tags": [1, 2, 3, 4, 5], "type": "SMA"}, {"tags": [6], "type": "PEC"}],
                      "materials": [{"tags": [1], "type": "dielectric", "relative_permittivity": 2.0, "bulk_conductivity": 0.1}]},
            "sources": [{"type": "planewave", "polarization": [1, 0, 0], "propagation": [0, 0, 2], "tags": [7], "magnitude": {"spread": 0.15}}]}
    path = _write_case(tmp_path, case, "tfsf3d_p2_on")
    mesh, kw, dt, t_final, init = build_case(case, str(tmp_path), dg)
    assert kw["tfsf"] == (7,) and not kw["tfsf_gate"] and kw["materials"] == {1: (2.0, 1.0, 0.1)}
    assert abs(kw["planewave"].mean1d - (0.25 - 5 * 0.15 * np.sqrt(2.0))) < 1e-12      # the box's upstream face is z = 0.25


@pytest.mark.gpu
def test_config1_json_case_runs_and_returns_after_one_period(tmp_path):
    from dgtd_b200.launcher import main
    path = _write_case(tmp_path, CASE_1D_PEC, "seg1d_config1_pec")
    out = tmp_path / "out"
    assert main(["-i", path, "-d", "b200", "-o", str(out)]) == 0
    x = np.load(out / "final_state_rank0.npy")
    arr, meta = read_fixture("seg1d_config1_pec")
    x0 = arr["x0_f64"].reshape(6, -1)                   # the same Gaussian, projected by the reference-based oracle
    gid = np.load(out / "final_elements_rank0.npy")
    assert np.array_equal(gid, np.arange(20))
    assert np.abs(x[1] - x0[1]).max() < 1e-2 and np.abs(x[[0, 2, 3, 4]]).max() < 1e-12      # E_y back, nothing in the other components
    stats = (out / "SimulationStats" / "statistics_rank0.dat").read_text()
    for key in ("Simulation Run Time:", "Final Time:", "Time Step:", "Number of Mesh Elements: 20", "Number of Local Degrees of Freedom: 80"):
        assert key in stats


def test_automatic_time_step_follows_the_reference_formulas(tmp_path):
    """solver_options.time_step absent or 0 -> estimateTimeStep (Solver.cpp:175-351); known answers worked out by hand from
    those formulas: Gauss-Lobatto points of order 3 on [0,1] are 0, (1 -+ 1/sqrt 5)/2, 1."""
    import dgtd_b200 as dg
    from dgtd_b200.launcher import build_case, estimate_time_step, gauss_lobatto_01
    g3 = (1.0 - 1.0 / np.sqrt(5.0)) / 2.0
    assert np.allclose(gauss_lobatto_01(3), [0.0, g3, 1.0 - g3, 1.0]) and np.allclose(gauss_lobatto_01(2), [0.0, 0.5, 1.0])
    assert np.allclose(gauss_lobatto_01(4), [0.0, (1 - np.sqrt(3 / 7)) / 2, 0.5, (1 + np.sqrt(3 / 7)) / 2, 1.0])
    # 1-D: 20 segments of 0.05, order 3: 0.05 * g3 / 3^1.5 (Solver.cpp:311-320)
    xs = np.zeros((21, 3)); xs[:, 0] = np.linspace(0.0, 1.0, 21)
    seg = np.stack([np.arange(20), np.arange(1, 21)], axis=1)
    assert abs(estimate_time_step(xs, seg, 1, 3) - 0.05 * g3 / 3 ** 1.5) < 1e-15
    assert abs(estimate_time_step(xs, seg, 1, 3, cfl=0.5) - 0.5 * 0.05 * g3 / 3 ** 1.5) < 1e-15
    # 2-D: right triangle with legs 1: area 1/2, perimeter 2 + sqrt 2; rmin = 2 g3
    tri_v = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0.0]])
    want2 = 0.75 * (0.5 / ((2 + np.sqrt(2)) / 2)) * (2 * g3) * 2 / 3
    assert abs(estimate_time_step(tri_v, np.array([[0, 1, 2]]), 2, 3) - want2) < 1e-15
    # 3-D: the Kuhn tetrahedron (0,0,0),(1,0,0),(1,1,0),(1,1,1): volume 1/6, two faces of area 1/2 and two of sqrt(2)/2
    tet_v = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [1, 1, 1.0]])
    tet = np.array([[0, 1, 2, 3]])
    want3 = 0.75 * ((1 / 6) / ((1 + np.sqrt(2)) / 2)) * (2 * g3) * 2 / 3 / 0.8
    assert abs(estimate_time_step(tet_v, tet, 3, 3, operator="global") - want3) < 1e-15
    # hesthaven in 3-D: fscale = 2 |J_f| / |J_e| = 2 (2 A) / (6 V), largest face sqrt(2)/2 -> 2 sqrt 2; dt = cfl / (fscale p^2)
    assert abs(estimate_time_step(tet_v, tet, 3, 3, operator="hesthaven") - 1.0 / (2 * np.sqrt(2) * 9)) < 1e-15
    # through the case parser: config 1 without time_step
    case = json.loads(json.dumps(CASE_1D_PEC)); del case["solver_options"]["time_step"]
    path = _write_case(tmp_path, case, "seg1d_config1_pec")
    _, _, dt, _, _ = build_case(case, str(tmp_path), dg)
    assert abs(dt - 0.05 * g3 / 3 ** 1.5) < 1e-12
