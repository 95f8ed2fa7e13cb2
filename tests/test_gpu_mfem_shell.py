"""The drop-in boundary: B200Evolution (mfem::TimeDependentOperator) and B200RK4Solver (mfem::ODESolver) driven by
MFEM's own RK4Solver next to the reference-based operator, inside one C++ program linked against the reference's MFEM
fork (oracle/_ref/dgtd_ref_shell, built by oracle/ref/Makefile `shell`; the checker, not the product).  The operator is
constructed through B200Adaptor.h with the reference's constructor signature from stand-ins of Model / SourcesManager /
EvolutionOptions that carry the reference's accessor names.  Cases follow the
reference's tests: 1-D PEC cavity (test/cases/CasesTest.cpp:15-136, config 1), 2-D mixed boundaries
(ExtensiveCasesTest.cpp:205-470), 3-D tets order 3 centred/upwind (Solver3DTest.cpp:57-102), TF/SF plane wave."""
import json
import os
import subprocess

import pytest

from conftest import GOLDEN, ROOT

pytestmark = pytest.mark.gpu
EXE = os.path.join(ROOT, "oracle", "_ref", "dgtd_ref_shell")
TFSF = os.path.join(GOLDEN, "tfsf_box.mesh")
SHEETS = os.path.join(GOLDEN, "sheets.mesh")     # interior PEC / PMC / SMA sheets (tests/golden/make_golden.py: sheet_mesh)

CASES = {
    "config1_1d_pec": "--mesh cart1d:20 --order 3 --alpha 1.0 --bdr 1:pec,2:pec --init gauss:E:1:0.1:1:0.5 --dt 5e-3 --steps 10",
    "tri2d_mixed": "--mesh cart2d:3:2 --order 3 --alpha 1.0 --bdr 1:pec,2:pmc,3:sma,4:pec --init random:2 --dt 1e-3 --steps 3",
    "tet_p3_upwind": "--mesh cart3d:2 --order 3 --alpha 1.0 --bdr-all pec --init random:1 --dt 1e-3 --steps 3",
    "tet_p3_centred": "--mesh cart3d:2 --order 3 --alpha 0.0 --bdr 1:pec,2:pmc,3:pec,4:pmc,5:pec,6:pmc --init random:4 --dt 1e-3 --steps 3",
    "tet_p4_sma": "--mesh cart3d:1 --order 4 --alpha 0.7 --bdr 1:sma,2:pec,3:sma,4:pmc,6:sma --init random:5 --dt 5e-4 --steps 2",
    "tfsf_planewave": f"--mesh {TFSF} --order 2 --alpha 1.0 --bdr 1:sma,2:sma,3:sma,4:sma,5:sma,6:pec --tfsf 7 --pw 0.15:0.0:0:1,0,0:0,0,1 --init random:6 --t0 0.4 --dt 2e-3 --steps 3",
    "tfsf_modulated": f"--mesh {TFSF} --order 2 --alpha 1.0 --bdr 1:sma,2:sma,3:sma,4:sma,5:sma,6:pec --tfsf 7 --pw 0.2:0.1:2.5:0,1,-1:1,1,1 --init zero --t0 0.3 --dt 2e-3 --steps 3",
    "interior_sheets": f"--mesh {SHEETS} --order 3 --alpha 0.6 --bdr 1:pec,2:sma,3:pmc,4:pec,5:sma,6:pec,7:pec,8:pmc,9:sma --init random:11 --dt 1e-3 --steps 2",
    "tfsf_skipped": f"--mesh {TFSF} --order 2 --alpha 1.0 --bdr 1:sma,2:sma,3:sma,4:sma,5:sma,6:pec --tfsf 7 --pw 0.05:-3.0:0:0,1,0:1,0,0 --init random:7 --t0 0.0 --dt 2e-3 --steps 2",
}


@pytest.mark.skipif(not os.path.exists(EXE), reason="oracle/_ref/dgtd_ref_shell not built (needs the build container)")
@pytest.mark.parametrize("name", sorted(CASES))
def test_mfem_shells_match_the_reference_operator(name):
    r = subprocess.run([EXE, "shell"] + CASES[name].split(), capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    d = json.loads(r.stdout.strip().splitlines()[-1])
    tol = 1e-10   # north_star: 1e-10 relative L2 per step
    assert d["mult_rel_l2"] < tol and d["mfem_rk4_on_b200_rel_l2"] < tol and d["fused_rk4_rel_l2"] < tol and d["resident_run_rel_l2"] < tol
    # Solver::run's final short step through B200RK4Solver::RunUntil, and a B200Gather snapshot queued before the loop
    assert d["run_until_rel_l2"] < tol and d["gather_abs"] == 0.0
    assert d["generic_fallback_rel_l2"] == 0.0
    if name == "tfsf_planewave":
        assert d["tfsf_applied"] > 0
    if name == "tfsf_skipped":
        assert d["tfsf_applied"] == 0 and d["tfsf_skipped"] > 0
