"""CPU-only checks of the product's host side: the C ABI loads and exports what include/dgtd_b200.h declares, and the
flat tables a rank uploads (reference element, geometry, vmapP, boundary codes, TF/SF sides, tensor-core plan, halo plan)
agree with the oracle, which is pinned to the reference (tests/test_oracle.py).  No compute entry point is called."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

import dgtd_b200 as dg
from conftest import ROOT, golden_cases, load_golden, product_mesh_and_kwargs
from oracle.dgtd_oracle import HesthavenOracle

FI_TAB = lambda code: (code >> 4) & 0xff


def test_library_exports_every_declared_symbol():
    assert len(dg.HEADER_SYMBOLS) >= 30
    for s in dg.HEADER_SYMBOLS:
        assert hasattr(dg.lib, s), s
    assert b"sm_100a" in dg.lib.dgtd_version()


def test_compute_entry_points_fail_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    pb, _ = load_golden("box3d_p1_gauss")
    mesh, kw = product_mesh_and_kwargs(pb)
    with pytest.raises(dg.DgtdError) as ei:
        dg.Evolution(mesh, **kw)
    assert ei.value.code == -2 and "no CPU fallback" in str(ei.value)


def _q(mesh, kw, name, dtype, **extra):
    k = dict(kw); k.update(extra)
    return dg.setup_query(mesh, name, dtype, **k)


@pytest.mark.parametrize("name", golden_cases())
def test_setup_tables_match_the_oracle(name):
    pb, dat = load_golden(name)
    O = HesthavenOracle(pb)
    mesh, kw = product_mesh_and_kwargs(pb)
    dim, p, Np, Nfp, nf, NE = (int(v) for v in _q(mesh, kw, "dims", np.int32)[:6])
    assert (dim, p, Np, Nfp, nf, NE) == (O.dim, pb.order, O.Np, O.Nfp, O.nf, O.NE)
    # (the product derives its operators in long double, the oracle in double: compare relative to the largest entry)
    assert np.abs(_q(mesh, kw, "D", np.float64).reshape(dim, Np, Np) - O.ref.D).max() < 5e-12 * np.abs(O.ref.D).max()
    assert np.abs(_q(mesh, kw, "lift", np.float64).reshape(nf, Np, Nfp) - O.ref.lift).max() < 5e-12 * np.abs(O.ref.lift).max()
    assert np.array_equal(_q(mesh, kw, "fnodes", np.int32).reshape(nf, Nfp), O.ref.fnodes)
    assert np.abs(_q(mesh, kw, "node_coords", np.float64).reshape(NE, Np, 3) - O.xyz).max() < 1e-14
    gid = _q(mesh, kw, "elem_gid", np.int32)
    assert sorted(gid) == list(range(NE))                      # a permutation (Morton order in 3-D)
    geo = _q(mesh, kw, "geo", np.float64).reshape(NE, 16)
    assert np.abs(geo[:, :9].reshape(NE, 3, 3) - O.Jinv[gid]).max() < 1e-12 * max(1.0, np.abs(O.Jinv).max())
    assert np.abs(geo[:, 9:9 + nf] - O.fscale[gid]).max() < 1e-12 * O.fscale.max()
    assert np.allclose(geo[:, 13], O.inv_eps[gid]) and np.allclose(geo[:, 14], O.inv_mu[gid]) and np.allclose(geo[:, 15], O.sig_eps[gid])
    # vmapP: the neighbour node named by (finfo, ftab) is the oracle's vmapP (and sits at the same point in space)
    finfo = _q(mesh, kw, "finfo", np.int32).reshape(NE, 4, 2)
    ftab = _q(mesh, kw, "ftab", np.uint8).reshape(-1, Nfp)
    l2g = gid
    for le in range(NE):
        e = l2g[le]
        for f in range(nf):
            nb, code = finfo[le, f]
            if O.nbr[e, f] < 0:
                assert nb == -1 and (code & 3) == O.bc[e, f]
                continue
            assert l2g[nb] == O.nbr[e, f]
            assert np.array_equal(l2g[nb] * Np + ftab[FI_TAB(code)], O.vmapP[e, f])
            tf = (code >> 2) & 3
            assert tf == {0: 0, 1: 1, -1: 2}[int(O.tfsf_face[e, f])]
    assert np.array_equal(_q(mesh, kw, "tfsf_side", np.int32), O.tfsf_side[gid])


@pytest.mark.parametrize("name", [n for n in golden_cases() if "3d" in n])
def test_wg_plan_records_are_consistent(name):
    """The warp-per-group plan keeps HostOp's face descriptors (padding elements: boundary, no condition), its geometry
    records satisfy (J/det) det Jinv = I, and its node tables are permutations of the face nodes."""
    pb, _ = load_golden(name)
    O = HesthavenOracle(pb)
    mesh, kw = product_mesh_and_kwargs(pb)
    Np, Nfp, NE = O.Np, O.Nfp, O.NE
    ngroups, NEpad, NT, KSV, nfv, nfl, ntab, GEO = (int(v) for v in _q(mesh, kw, "wg_dims", np.int32))
    assert NEpad == 8 * ngroups >= NE and NT == (Np + 7) // 8 and KSV == (Np + 3) // 4
    finfo = _q(mesh, kw, "finfo", np.int32).reshape(NE, 4, 2)
    desc = _q(mesh, kw, "wg_desc", np.int32).reshape(NEpad, 4, 2)
    assert np.array_equal(desc[:NE, :, 0], finfo[:, :, 0])
    assert np.array_equal(desc[:NE, :, 1] & 0xf, finfo[:, :, 1] & 0xf) and np.array_equal(desc[:NE, :, 1] >> 12, finfo[:, :, 1] >> 12)
    assert (desc[NE:, :, 0] == -1).all() and (desc[NE:, :, 1] == 0).all()
    geo = _q(mesh, kw, "wg_geo", np.float64).reshape(NEpad, GEO)
    Jd, Ji = geo[:, :9].reshape(-1, 3, 3), geo[:, 9:18].reshape(-1, 3, 3)
    det = 1.0 / geo[:, 22]
    assert np.abs(np.einsum("eda,eac->edc", Jd * det[:, None, None], Ji) - np.eye(3)).max() < 1e-12
    assert np.allclose(np.linalg.det(Jd[:NE] * det[:NE, None, None]), det[:NE], rtol=1e-12)
    tab = _q(mesh, kw, "wg_tab", np.uint8).reshape(ntab, 16)
    d2r = _q(mesh, kw, "wg_dev2ref", np.int32)
    fn = _q(mesh, kw, "fnodes", np.int32).reshape(4, Nfp)
    for f in range(4):
        assert sorted(d2r[tab[f, :Nfp]]) == sorted(fn[f])
        assert sorted(tab[4 + f, :Nfp]) == list(range(Nfp))


def test_partitioner_and_halo_plan_two_ranks_in_process():
    """Both sides of every shared face enumerate it identically and ship the nodes the receiver expects."""
    pb, _ = load_golden("box3d_p3_pec_upwind")
    O = HesthavenOracle(pb)
    mesh, kw = product_mesh_and_kwargs(pb)
    part = mesh.partition(2)
    assert set(part) == {0, 1} and abs(int((part == 0).sum()) - int((part == 1).sum())) <= 1
    Np, Nfp = O.Np, O.Nfp
    plan = {}
    for r in range(2):
        gid = _q(mesh, kw, "elem_gid", np.int32, rank=r, nranks=2)
        assert np.array_equal(np.sort(gid), np.nonzero(part == r)[0])
        send = _q(mesh, kw, "send_node", np.int32, rank=r, nranks=2).reshape(-1, Nfp)
        finfo = _q(mesh, kw, "finfo", np.int32, rank=r, nranks=2).reshape(len(gid), 4, 2)
        peers = _q(mesh, kw, "peers", np.int32, rank=r, nranks=2).reshape(-1, 3)
        assert len(peers) == 1 and peers[0, 0] == 1 - r and peers[0, 1] == len(send)
        plan[r] = (gid, send, finfo)
    assert len(plan[0][1]) == len(plan[1][1]) > 0
    xyz = O.xyz.reshape(-1, 3)
    for r in range(2):
        gid, send, finfo = plan[r]
        ogid, osend, _ = plan[1 - r]
        # what the peer sends into my halo slot s must be the points of my face nodes, in my face-node order
        for le in range(len(gid)):
            for f in range(4):
                nb, code = finfo[le, f]
                if nb <= -2:
                    s = -2 - nb
                    mine = xyz[gid[le] * Np + O.ref.fnodes[f]]
                    theirs = xyz[ogid[osend[s] // Np] * Np + osend[s] % Np]
                    assert np.abs(mine - theirs).max() < 1e-14


def test_metis_partition_of_the_dual_graph():
    """dgtd_mesh_partition_metis (the reference partitions with METIS, driver.cpp:1269): every rank owns a balanced,
    non-empty share and the cut is no worse than coordinate bisection's on an unstructured-like case."""
    pb, _, _ = __import__("conftest").load_config_case("config4_rcs_pec_p3")
    mesh, kw = product_mesh_and_kwargs(pb)
    elems = np.asarray(pb.elems)
    faces = {}
    for e, t in enumerate(elems.tolist()):
        for k in range(4):
            faces.setdefault(tuple(sorted(t[:k] + t[k + 1:])), []).append(e)
    pairs = np.array([v for v in faces.values() if len(v) == 2])
    cut = lambda p: int((p[pairs[:, 0]] != p[pairs[:, 1]]).sum())
    for world in (2, 8):
        pm, pr = mesh.partition(world, "metis"), mesh.partition(world, "rcb")
        cnt = np.bincount(pm, minlength=world)
        assert cnt.min() > 0 and cnt.max() <= 1.05 * len(elems) / world
        assert cut(pm) <= cut(pr)


@pytest.mark.parametrize("world,method", [(2, "rcb"), (4, "rcb"), (4, "metis")])
def test_direct_push_plan_is_consistent_across_ranks(world, method):
    """Peer-memory halo path (kernels_wg.cuh: WgP2P): the slot a rank stores a face's traces into on its neighbour must be
    the slot the neighbour reads for that face, in the neighbour's face-node order; flag slots must be mutually consistent."""
    pb, _ = load_golden("box3d_p3_pec_upwind")
    O = HesthavenOracle(pb)
    mesh, kw = product_mesh_and_kwargs(pb)
    Np, Nfp = O.Np, O.Nfp
    xyz = O.xyz.reshape(-1, 3)
    R = {}
    part = mesh.partition(world, method)
    for r in range(world):
        q = lambda name, dt: _q(mesh, kw, name, dt, rank=r, nranks=world, partitioning=part)
        R[r] = dict(gid=q("elem_gid", np.int32), peers=q("peers5", np.int32).reshape(-1, 5), hpush=q("wg_hpush", np.int32).reshape(-1, 2),
                    tab=q("wg_tab", np.uint8).reshape(-1, 16), desc=q("wg_desc", np.int32).reshape(-1, 4, 2), d2r=q("wg_dev2ref", np.int32))
    for r in range(world):
        me = R[r]
        # flag slots: my remote_idx at peer p is my position in p's peer list
        for pi, (prank, nfaces, soff, roff, ridx) in enumerate(me["peers"]):
            theirs = R[int(prank)]["peers"]
            assert theirs[ridx, 0] == r and theirs[ridx, 1] == nfaces and theirs[ridx, 2] == roff
        # halo slot -> (local element, face) on every rank
        slot_face = {}
        for le in range(len(me["gid"])):
            for f in range(4):
                nb = me["desc"][le, f, 0]
                if nb <= -2:
                    slot_face[-2 - nb] = (le, f)
        assert len(slot_face) == len(me["hpush"]) > 0
        for s, (le, f) in slot_face.items():
            pi, row = me["hpush"][s, 0] & 0xff, me["hpush"][s, 0] >> 8
            dst_rank, dst_slot = int(me["peers"][pi, 0]), int(me["hpush"][s, 1])
            there = R[dst_rank]
            # the receiving side: the face behind dst_slot, its canonical face nodes m = 0..Nfp-1
            rle, rf = next((a, b) for a in range(len(there["gid"])) for b in range(4) if there["desc"][a, b, 0] == -2 - dst_slot)
            want = xyz[there["gid"][rle] * Np + O.ref.fnodes[rf]]
            have = xyz[me["gid"][le] * Np + me["d2r"][me["tab"][row, :Nfp]]]
            assert np.abs(want - have).max() < 1e-14


def test_halo_plan_world_size_2_gloo():
    """The same through two processes that exchange their send lists' coordinates over gloo (host logic of the N>1 path)."""
    script = os.path.join(ROOT, "tests", "mp_halo_plan_cpu.py")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29431", script]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=dict(os.environ, OMP_NUM_THREADS="1"))
    assert r.returncode == 0 and r.stdout.count("HALO_PLAN_OK") == 2, r.stdout[-2000:] + r.stderr[-2000:]


def test_boundary_element_listing_for_surface_exports():
    """dgtd_mesh_boundary_elements: every tagged boundary face is reported with the element(s) it belongs to
    (selection step of NearToFarFieldSubMesher, SubMesher.cpp:832-905); interior surfaces list both sides."""
    pb, _ = load_golden("tfsf3d_p2_on")
    mesh, kw = product_mesh_and_kwargs(pb)
    tags = list(pb.tfsf_tags)
    pairs = mesh.boundary_elements(tags)
    nb = int(np.isin(pb.bdr_attr, tags).sum())
    assert nb > 0 and len(pairs) == 2 * nb           # TF/SF surface is interior: two sides per face
    elems = np.asarray(pb.elems)
    bdr = np.asarray(pb.bdr)[np.isin(pb.bdr_attr, tags)]
    want = {tuple(sorted(f)) for f in bdr.tolist()}
    got = {}
    for e, f in pairs.tolist():
        got.setdefault(tuple(sorted(np.delete(elems[e], f).tolist())), []).append(e)
    assert set(got) == want
    assert all(len(v) == 2 and v[0] < v[1] for v in got.values())
    # true boundary: one side
    ext = [a for a in set(np.asarray(pb.bdr_attr).tolist()) if a not in tags]
    p2 = mesh.boundary_elements(ext)
    assert len(p2) == int(np.isin(pb.bdr_attr, ext).sum())


def test_mfem_mesh_reader_reproduces_what_mfem_loads():
    """dgtd_mesh_load on an `MFEM mesh v1.0` file against the arrays MFEM itself produced for the same file (stored in the
    TF/SF fixtures by the reference-based oracle, after Mesh::LoadFromFile(..., fix_orientation = true))."""
    pb, _ = load_golden("tfsf3d_p2_on")
    m = dg.Mesh.load(os.path.join(ROOT, "tests", "golden", "tfsf_box.mesh"))
    v, e, ea, b, ba = m.arrays()
    assert (m.dim, m.ne, m.nbe) == (3, len(pb.elems), len(pb.bdr))
    assert np.array_equal(v, pb.verts) and np.array_equal(ea, pb.elem_attr)
    assert np.array_equal(e, pb.elems)                      # same vertex order as MFEM after its orientation fix
    assert np.array_equal(np.sort(b, axis=1), np.sort(np.asarray(pb.bdr), axis=1)) and np.array_equal(ba, pb.bdr_attr)


def test_gmsh_reader_on_a_hand_written_file(tmp_path):
    """Gmsh 2.2 ASCII: nodes, tetrahedra (type 4) with physical tag -> element attribute, triangles (type 2) -> boundary
    elements; an inverted tetrahedron is re-oriented like MFEM does (mesh.cpp:6437-6493)."""
    p = tmp_path / "two_tets.msh"
    p.write_text("""$MeshFormat
2.2 0 8
$EndMeshFormat
$Nodes
5
1 0 0 0
2 1 0 0
3 0 1 0
4 0 0 1
5 1 1 1
$EndNodes
$Elements
5
1 2 2 7 1 1 2 3
2 2 2 7 1 1 2 4
3 2 2 9 2 2 3 5
4 4 2 1 1 1 2 3 4
5 4 2 2 2 2 4 3 5
$EndElements
""")
    m = dg.Mesh.load(str(p))
    v, e, ea, b, ba = m.arrays()
    assert (m.dim, m.nv, m.ne, m.nbe) == (3, 5, 2, 3)
    assert ea.tolist() == [1, 2] and sorted(ba.tolist()) == [7, 7, 9]
    for t in e:                                             # positive orientation after loading
        J = (v[t[1:]] - v[t[0]]).T
        assert np.linalg.det(J) > 0
    assert sorted(map(sorted, e.tolist())) == [[0, 1, 2, 3], [1, 2, 3, 4]]
    # the operator setup accepts it (shared face found, boundary tags resolved)
    d = dg.setup_query(m, "dims", np.int32, order=2, bdr={7: dg.BC_PEC, 9: dg.BC_SMA})
    assert d[5] == 2
    with pytest.raises(dg.DgtdError):
        dg.Mesh.load(str(tmp_path / "missing.msh"))


@pytest.mark.skipif(not os.path.exists("/root/reference/testData/maxwellInputs/2D_PEC/2D_PEC.msh"), reason="needs /root/reference (build container)")
def test_gmsh_reader_on_the_reference_config2_mesh():
    """The reference's own 2D_PEC.msh through dgtd_mesh_load against MFEM's view of it (config-2 fixture)."""
    pb, _ = load_golden("config2_2d_pec_p3")
    m = dg.Mesh.load("/root/reference/testData/maxwellInputs/2D_PEC/2D_PEC.msh")
    v, e, ea, b, ba = m.arrays()
    assert (m.dim, m.ne, m.nbe) == (2, len(pb.elems), len(pb.bdr))
    # MFEM renumbers the Gmsh nodes, so compare geometry: element k covers the same three points, in the same element order
    canon = lambda pts: np.array(sorted(map(tuple, np.round(pts, 12))))
    mine, theirs = v[e], pb.verts[np.asarray(pb.elems)]
    assert all(np.array_equal(canon(a), canon(c)) for a, c in zip(mine, theirs))
    assert np.array_equal(ea, pb.elem_attr)
    key = lambda V, B, A: sorted((tuple(map(tuple, canon(V[f]))), int(t)) for f, t in zip(B, A))
    assert key(v, b, ba) == key(pb.verts, np.asarray(pb.bdr), pb.bdr_attr)
    # and the operator setup sees the same problem: every element's geometry factors agree with the oracle's on MFEM's arrays
    O = HesthavenOracle(pb)
    d = dg.setup_query(m, "dims", np.int32, order=3, bdr={1: dg.BC_PMC, 2: dg.BC_PEC, 3: dg.BC_PMC, 4: dg.BC_PEC})
    assert d[5] == O.NE


@pytest.mark.parametrize("world", [1, 2, 8])
def test_unstructured_partitions_fit_the_warp_per_group_kernel(world, monkeypatch):
    """BASELINE config 4's sphere mesh: every METIS part must fit the node tables the tensor-core kernels keep in shared
    memory (8 + 4 x 4 x 6 neighbour orientations + 4 x 6 push rows = 128 rows; with 72 rows the parts fell back to the generic
    kernel and NCCL), and the groups of eight elements must share faces (in-group faces well above plain Morton order's 27 %)."""
    from golden_io import product_problem, read_fixture
    arr, meta = read_fixture("config4_rcs_pec_p3")
    mesh, kw = product_problem(arr, meta)
    part = mesh.partition(world, "metis") if world > 1 else None
    monkeypatch.setenv("DGTD_B200_ORDER", "grow")          # the default of multi-rank contexts; a single rank keeps Morton order
    for r in range(world):
        q = lambda name, dt: _q(mesh, kw, name, dt, rank=r, nranks=world, partitioning=part)
        dims = q("wg_dims", np.int32)
        ngroups, NEpad, ntab = int(dims[0]), int(dims[1]), int(dims[6])
        assert ntab <= 128
        nb = q("wg_desc", np.int32).reshape(NEpad, 4, 2)[:, :, 0]
        gid = q("elem_gid", np.int32)
        assert len(np.unique(gid)) == len(gid) and NEpad - len(gid) < 8
        grp = np.arange(NEpad)[:, None] >> 3
        in_group = ((nb >= 0) & ((nb >> 3) == grp)).sum() / max(1, (nb >= 0).sum())
        assert in_group > 0.40, in_group
        # partition-face elements sit at the front of the local numbering
        cut = (nb < -1).any(axis=1)
        if cut.any():
            assert np.nonzero(cut)[0].max() < 8 * ((cut.sum() + 7) // 8) + 8
