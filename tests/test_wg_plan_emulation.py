"""CPU replay of the warp-per-group stage kernel from its PLAN (dgtd_b200/csrc/wgplan.cpp): the operator B-fragments, the
face-step tables, the descriptors, the geometry records and the device node numbering are read through dgtd_setup_query and
interpreted in numpy exactly as kernels_wg.cuh does (DMMA tiles incl. the mixed last tile, per-face flux map g x (dH - alpha n x dE), LIFT with
K = 4 faces per step, push-forward with J / det), then compared with the oracle's Mult.  This pins the plan — including
non-trivial node numberings and face-step orders — without a GPU; the GPU parity tests pin the kernel that consumes it."""
import numpy as np
import pytest

import dgtd_b200 as dg
from conftest import load_golden, product_mesh_and_kwargs, rel_l2
from oracle.dgtd_oracle import HesthavenOracle


def _plan(mesh, kw, **extra):
    q = lambda name, dt: dg.setup_query(mesh, name, dt, **kw, **extra)
    ngroups, NEpad, NT, KSV, nfv, nfl, ntab, GEO = q("wg_dims", np.int32)
    return dict(NT=NT, KSV=KSV, nfv=nfv, nfl=nfl, NEpad=NEpad,
                bfrag=q("wg_bfrag", np.float64).reshape(-1, 32), geo=q("wg_geo", np.float64).reshape(NEpad, GEO),
                desc=q("wg_desc", np.int32).reshape(NEpad, 4, 2), tab=q("wg_tab", np.uint8).reshape(-1, 16),
                d2r=q("wg_dev2ref", np.int32), gid=q("elem_gid", np.int32), dims=q("dims", np.int32),
                hpush=q("wg_hpush", np.int32).reshape(-1, 2), peers=q("peers5", np.int32).reshape(-1, 5))


def _frag_matrix(f):
    """32 lane values of a DMMA B fragment -> B[k][n], lane l holds B[k = l & 3][n = l >> 2]."""
    B = np.zeros((4, 8))
    for l in range(32):
        B[l & 3, l >> 2] = f[l]
    return B


def device_state(P, x_ref):
    """u[e][device node][c] of this rank's elements from a global reference-layout vector."""
    Np, NE = P["dims"][2], P["dims"][5]
    xr = x_ref.reshape(6, -1)
    U = np.zeros((NE, Np, 6))
    for le in range(NE):
        U[le] = xr[:, P["gid"][le] * Np + P["d2r"]].T
    return U


def pushed_traces(P, U):
    """What the stage kernel stores into its neighbours' halo buffers: {peer rank: {slot on the peer: [Nfp][6]}} —
    face behind my halo slot s, my nodes in the RECEIVER's face-node order (row hpush[s] >> 8 of the tables)."""
    Nfp = P["dims"][3]
    slot_face = {-2 - P["desc"][e, f, 0]: (e, f) for e in range(P["dims"][5]) for f in range(4) if P["desc"][e, f, 0] < -1}
    out = {}
    for s, (e, f) in slot_face.items():
        pi, row = P["hpush"][s, 0] & 0xff, P["hpush"][s, 0] >> 8
        out.setdefault(int(P["peers"][pi, 0]), {})[int(P["hpush"][s, 1])] = U[e, P["tab"][row, :Nfp]]
    return out


def replay_mult(P, x_ref, alpha, halo=None):
    """Mult(x) (no TF/SF source) from the plan tables; halo = {my halo slot: [Nfp][6]} on a multi-rank plan."""
    dim, p, Np, Nfp, nf, NE = P["dims"][:6]
    NT, KSV, VT = P["NT"], P["KSV"], (P["NT"] - 1) * 3 + 3
    NL = Np - 8 * (NT - 1)
    fragV = [_frag_matrix(f) for f in P["bfrag"][:P["nfv"]]]
    fragL = [_frag_matrix(f) for f in P["bfrag"][P["nfv"]:P["nfv"] + P["nfl"]]]
    d2r, gid, tab = P["d2r"], P["gid"], P["tab"]
    N = x_ref.size // 6
    U = device_state(P, x_ref)
    K = np.zeros_like(U)
    for e in range(NE):
        g = P["geo"][e]
        Jd, Ji, fs = g[0:9], g[9:18], g[18:22]          # Jd[3d+a] = J[d][a]/det ; Ji[3a+d] = dxi_a/dx_d
        de, dm, se = g[23], g[24], g[25]
        # ---- volume: u~ = (J/det)^T u (E negated); acc[c][tile] columns = DMMA N index
        acc = np.zeros((6, NT, 8))
        ut_all = np.zeros((4 * KSV, 6))
        for n in range(Np):
            u = U[e, n]
            for a in range(3):
                ut_all[n, a] = -(Jd[a] * u[0] + Jd[3 + a] * u[1] + Jd[6 + a] * u[2])
                ut_all[n, 3 + a] = Jd[a] * u[3] + Jd[3 + a] * u[4] + Jd[6 + a] * u[5]
        for ks in range(KSV):
            A = ut_all[4 * ks:4 * ks + 4]               # [k][comp]
            for nt in range(NT - 1):
                for d in range(3):
                    B = fragV[ks * VT + nt * 3 + d]
                    cp, cm = (d + 2) % 3, (d + 1) % 3
                    acc[cp, nt] += A[:, 3 + (d + 1) % 3] @ B
                    acc[cm, nt] -= A[:, 3 + (d + 2) % 3] @ B
                    acc[3 + cp, nt] += A[:, (d + 1) % 3] @ B
                    acc[3 + cm, nt] -= A[:, (d + 2) % 3] @ B
            for x in range(3):
                B = fragV[ks * VT + (NT - 1) * 3 + x]
                acc[x, NT - 1] += A[:, 3 + x] @ B
                acc[3 + x, NT - 1] += A[:, x] @ B
        # ---- flux -> LIFT: K index of the DMMA = face j
        ft = np.zeros((Nfp, 4, 6))
        for j in range(4):
            nb, code = P["desc"][e, j]
            ce = ch = 0.0
            al = alpha
            trace = None
            if nb >= 0:
                nrow, ne_ = tab[(code >> 4) & 0xff], nb
            elif nb < -1:                                # partition face: the neighbour rank's trace, canonical node order
                trace, nrow = halo[-2 - nb], tab[4 + j]
            else:
                bc = code & 3
                ce = -2.0 if bc == 1 else -1.0 if bc == 3 else 0.0
                ch = -2.0 if bc == 2 else -1.0 if bc == 3 else 0.0
                if bc == 3:
                    al = 1.0
                nrow, ne_ = tab[j], e
            Jim = Ji.reshape(3, 3)                       # [a][d]
            gn = (Jim[0] + Jim[1] + Jim[2]) if j == 0 else -Jim[j - 1]
            cross = np.array([[0, -gn[2], gn[1]], [gn[2], 0, -gn[0]], [-gn[1], gn[0], 0]])
            Ah = Jim @ cross                             # J^-1 [g x], g = fs n
            na = al * g[26 + j] * gn                     # alpha n  (1 / fs comes with the geometry record)
            assert abs(g[26 + j] * fs[j] - 1.0) < 1e-14
            for s in range(Nfp):
                uM = U[e, tab[j][s]]
                uP = trace[nrow[s]] if trace is not None else U[ne_, nrow[s]]
                dE = uP[:3] - (1.0 - ce) * uM[:3]
                dH = uP[3:] - (1.0 - ch) * uM[3:]
                ft[s, j, :3] = Ah @ (dH - np.cross(na, dE))      # g x (dH - alpha n x dE)
                ft[s, j, 3:] = -(Ah @ (dE + np.cross(na, dH)))
        for s in range(Nfp):
            for nt in range(NT - 1):
                B = fragL[s * NT + nt]
                for c in range(6):
                    acc[c, nt] += ft[s, :, c] @ B
            B = fragL[s * NT + NT - 1]
            for c in range(6):
                acc[3 * (c // 3) + (c % 3 + 2) % 3, NT - 1] += ft[s, :, c] @ B
        # ---- epilogue: column q of a full tile is node 8 nt + (q >> 1) + 4 (q & 1); mixed tile: even/odd columns
        Jdm = Jd.reshape(3, 3)
        for n in range(Np):
            nt = n // 8
            kr = np.zeros(6)
            if nt < NT - 1:
                r = n - 8 * nt
                q = 2 * (r % 4) + r // 4
                kr = acc[:, nt, q]
            else:
                r = n - 8 * (NT - 1)
                assert r < NL
                for c in range(6):
                    f3, cc = 3 * (c // 3), c % 3
                    kr[c] = acc[f3 + (cc + 2) % 3, NT - 1, 2 * r] + acc[f3 + (cc + 1) % 3, NT - 1, 2 * r + 1]
            K[e, n, :3] = de * (Jdm @ kr[:3]) - se * U[e, n, :3]
            K[e, n, 3:] = dm * (Jdm @ kr[3:])
    out = np.zeros((6, N))
    for le in range(NE):
        out[:, gid[le] * Np + d2r] = K[le].T
    return out.ravel()


@pytest.mark.parametrize("name", ["box3d_p1_gauss", "box3d_p2_mixed_centered", "box3d_p2_materials", "box3d_p3_pec_upwind", "box3d_p4_sma_partial"])
@pytest.mark.parametrize("face_order", ["file", "bank"])
def test_plan_replay_matches_the_oracle(name, face_order, monkeypatch):
    """face_order = "bank": DGTD_B200_FACE_ORDER=bank renumbers the device nodes and orders the face steps so that the four
    faces of an element read nodes with distinct residues mod 4 at every step (conflict-free LDS.128 gathers)."""
    monkeypatch.setenv("DGTD_B200_FACE_ORDER", face_order)
    pb, dat = load_golden(name)
    mesh, kw = product_mesh_and_kwargs(pb)
    kw = {k: v for k, v in kw.items() if k in ("order", "alpha", "bdr", "materials")}
    P = _plan(mesh, kw)
    O = HesthavenOracle(pb)
    x = np.random.default_rng(9).standard_normal(6 * O.N)
    assert rel_l2(replay_mult(P, x, pb.alpha), O.mult(0.0, x)) < 1e-12
    if face_order == "bank":
        Nfp = P["dims"][3]
        own = P["tab"][:4, :Nfp].astype(int)
        assert sorted(P["d2r"].tolist()) == list(range(P["dims"][2]))
        assert all(len(set(own[:, s] % 4)) == 4 for s in range(Nfp)), "a face step with a bank conflict"


@pytest.mark.parametrize("name,world,method", [("box3d_p3_pec_upwind", 2, "rcb"), ("box3d_p3_pec_upwind", 4, "metis"), ("box3d_p2_mixed_centered", 3, "rcb")])
def test_partitioned_plan_replay_with_pushed_traces(name, world, method):
    """The peer-memory halo path on the CPU: every rank 'stores' the traces of its partition faces into the slots of its
    neighbours' halo buffers as the stage kernel would (hpush tables), then evaluates its share from its own plan; the
    assembled result must be the oracle's Mult on the undivided mesh."""
    pb, _ = load_golden(name)
    mesh, kw = product_mesh_and_kwargs(pb)
    kw = {k: v for k, v in kw.items() if k in ("order", "alpha", "bdr", "materials")}
    O = HesthavenOracle(pb)
    x = np.random.default_rng(21).standard_normal(6 * O.N)
    part = mesh.partition(world, method)
    plans = [_plan(mesh, kw, rank=r, nranks=world, partitioning=part) for r in range(world)]
    halos = [dict() for _ in range(world)]
    for r, P in enumerate(plans):
        for peer, slots in pushed_traces(P, device_state(P, x)).items():
            halos[peer].update(slots)
    out = np.zeros(6 * O.N)
    owned = np.zeros(O.N, int)
    for r, P in enumerate(plans):
        n_halo = sum(1 for e in range(P["dims"][5]) for f in range(4) if P["desc"][e, f, 0] < -1)
        assert sorted(halos[r]) == list(range(n_halo)), "every halo slot of a rank is written exactly by its neighbours"
        k = replay_mult(P, x, pb.alpha, halos[r]).reshape(6, -1)
        Np = P["dims"][2]
        idx = (P["gid"][:, None] * Np + np.arange(Np)[None, :]).ravel()
        out.reshape(6, -1)[:, idx] = k[:, idx]
        owned[idx] += 1
    assert owned.min() == 1 and owned.max() == 1
    assert rel_l2(out, O.mult(0.0, x)) < 1e-12
