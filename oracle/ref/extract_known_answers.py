#!/usr/bin/env python3
"""Extract the known-answer matrices the reference's own tests assert for the hot path.

TEST INFRASTRUCTURE.  Reads (never copies source from)
  /root/reference/test/hesthavenComparison/Hesthaven2DTest.cpp:234-553
and writes only the NUMERIC LITERALS of the nine `M^-1 * flux` blocks
(order 1, testData/mfemMeshes/2D/Maxwell2D_K2.mesh, PEC on attribute 2,
tolerance 1e-8 in the reference) to tests/golden/ref_known_answers.{json,txt}.
`oracle/_ref/dgtd_ref known-answers` rebuilds each block with the reference's
integrators and compares.  Run in the build container only (needs /root/reference).
"""
import json, re, sys, pathlib

REF = pathlib.Path(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")
OUT = pathlib.Path(__file__).resolve().parents[2] / "tests" / "golden"
src_path = REF / "test/hesthavenComparison/Hesthaven2DTest.cpp"
text = src_path.read_text()
lines = text.splitlines()

records = {}
for m in re.finditer(r"TEST_F\(MFEMHesthaven2D,\s*(2D_Operator_\w+)\)", text):
    name = m.group(1)
    line = text.count("\n", 0, m.start()) + 1
    body = text[m.end():]
    mm = re.search(r"DynamicMatrix\s+\w+\s*\{(.*?)\};", body, re.S)
    rows = re.findall(r"\{([^{}]*)\}", mm.group(1))
    mat = [[float(v) for v in r.replace(" ", "").strip(",").split(",") if v] for r in rows]
    assert len({len(r) for r in mat}) == 1
    records[name] = {"source": f"test/hesthavenComparison/Hesthaven2DTest.cpp:{line}",
                     "mesh": "testData/mfemMeshes/2D/Maxwell2D_K2.mesh", "order": 1,
                     "bdr": {"2": "PEC"}, "alpha": 1.0, "tol": 1e-8, "matrix": mat}

OUT.mkdir(parents=True, exist_ok=True)
(OUT / "ref_known_answers.json").write_text(json.dumps(records, indent=1))
with open(OUT / "ref_known_answers.txt", "w") as f:
    for name, r in records.items():
        m = r["matrix"]
        f.write(f"{name} {len(m)} {len(m[0])} " + " ".join(repr(v) for row in m for v in row) + "\n")
# the mesh those tests use is 30 lines of public MFEM-format data; restate it as arrays
mesh = {"dimension": 2,
        "elements": [[1, 2, [3, 0, 2]], [1, 2, [2, 0, 1]]],
        "boundary": [[2, 1, [0, 1]], [2, 1, [1, 2]], [2, 1, [2, 3]], [2, 1, [3, 0]]],
        "vertices": [[0, 0], [1, 0], [1, 1], [0, 1]]}
ref_mesh = (REF / "testData/mfemMeshes/2D/Maxwell2D_K2.mesh").read_text()
nums = [l.split() for l in ref_mesh.splitlines() if l and l[0].isdigit()]
assert [int(x) for x in nums[2]] == [1, 2, 3, 0, 2] and [int(x) for x in nums[3]] == [1, 2, 2, 0, 1], nums
(OUT / "Maxwell2D_K2.json").write_text(json.dumps(mesh))
print(f"wrote {len(records)} known-answer matrices to {OUT}")
