"""CPU oracle (numpy) for the DG evolution hot path of OpenSEMBA/dgtd.

TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module; the product
(dgtd_b200/) never does and fails loudly when its CUDA library is missing.

What it restates (reference file:line, /root/reference = OpenSEMBA/dgtd @ 47baddc):

* the matrix-free `hesthaven` right-hand side
    src/evolution/HesthavenEvolution.cpp:450-542   (Mult: jumps -> BCs -> TF/SF -> flux -> D + LIFT)
    :275-313  applyBoundaryConditionsToNodes        :97-124  evaluateTFSF
    :126-132  applyLIFT                             :150-205 D_d, normals, fscale
  with the SIGN/COEFFICIENT CONVENTIONS OF THE DEFAULT `global` OPERATOR where the two
  flavours differ: SMA uses full upwinding whatever alpha is
  (src/components/DGOperatorFactory.h:483-496, 555-568), materials enter through
  M^-1 only (:413-424) and conductivity as -sigma/eps E (:1349-1361), and the TF/SF
  injection is skipped for a Mult whose masked source has ||s||_2 < 1e-8
  (src/evolution/GlobalEvolution.cpp:584-598, GlobalEvolution.h:100);
* node set and numbering of DG_FECollection(p, dim, GaussLobatto) on simplices
    external/mfem-geg/fem/fe/fe_l2.cpp:23-37 (segment), :569-592 (triangle), :716-722 (tet)
  dof = e*Np + local, fields stored [Ex,Ey,Ez,Hx,Hy,Hz] blocks of N (src/evolution/Fields.h:45-65);
* TF/SF side rule in 3-D  src/components/SubMesher.cpp:677-771 and the +-1/2 mask
    src/solver/SourcesManager.cpp:158-188;
* plane wave  src/math/Function.h:71-136, 361-409;
* classical RK4 with the reference's SetTime sequence  external/mfem-geg/linalg/ode.cpp:109-136.

Pinning: tests/test_oracle.py checks this module against state vectors produced by
oracle/_ref/dgtd_ref (the reference's own integrators + MFEM compiled from
/root/reference, itself pinned to the reference's known-answer matrices), committed
under tests/golden/.  Agreement is ~1e-13 relative.

Everything is derived from scratch (orthonormalised polynomial basis, exact
quadrature, barycentric face matching); nothing is imported from the product.
"""
from __future__ import annotations

import itertools
import math
from dataclasses import dataclass, field

import numpy as np

PEC, PMC, SMA = 1, 2, 3           # boundary codes (0 = interior / untagged boundary)
BC_NAMES = {"pec": PEC, "pmc": PMC, "sma": SMA}


# ----------------------------------------------------------------------------------------------
# reference element
# ----------------------------------------------------------------------------------------------
def gll01(p: int) -> np.ndarray:
    """Gauss-Lobatto points on [0,1] (poly1d.OpenPoints(p, GaussLobatto))."""
    if p == 0:
        return np.array([0.5])
    if p == 1:
        return np.array([0.0, 1.0])
    # interior points: roots of P'_p on [-1,1]
    c = np.zeros(p + 1)
    c[p] = 1.0
    dr = np.polynomial.legendre.legroots(np.polynomial.legendre.legder(c))
    x = np.concatenate(([-1.0], np.sort(dr), [1.0]))
    for _ in range(3):  # Newton polish on (1-x^2) P'_p(x)
        xi = x[1:-1]
        d1 = np.polynomial.legendre.legval(xi, np.polynomial.legendre.legder(c))
        d2 = np.polynomial.legendre.legval(xi, np.polynomial.legendre.legder(c, 2))
        x[1:-1] = xi - d1 / d2
    x = 0.5 * (x + 1.0)
    return 0.5 * (x + (1.0 - x[::-1]))  # symmetrise


def simplex_index_tuples(dim: int, p: int) -> np.ndarray:
    """Integer index tuples (i[,j[,k]]) in MFEM's L2 ordering (last index outermost)."""
    if dim == 1:
        return np.array([[i] for i in range(p + 1)])
    if dim == 2:
        return np.array([[i, j] for j in range(p + 1) for i in range(p + 1 - j)])
    return np.array([[i, j, k] for k in range(p + 1) for j in range(p + 1 - k) for i in range(p + 1 - j - k)])


def simplex_nodes(dim: int, p: int):
    """Reference coordinates of the nodes on the unit simplex and their barycentric integers.

    bary[:, 0] belongs to vertex 0 (= p - sum of the others), bary[:, 1+d] to vertex 1+d.
    """
    g = gll01(p)
    idx = simplex_index_tuples(dim, p)
    last = p - idx.sum(axis=1)
    bary = np.concatenate([last[:, None], idx], axis=1)
    if dim == 1:
        # segment is a tensor element: nodes are the 1-D GLL points themselves
        return g[idx[:, 0]][:, None].copy(), bary
    w = g[bary].sum(axis=1)
    return g[idx] / w[:, None], bary


def _legendre01(n: int, x: np.ndarray):
    """P_0..P_n of (2x-1) and d/dx, shape (n+1, len(x))."""
    t = 2.0 * x - 1.0
    P = np.zeros((n + 1, x.size))
    dP = np.zeros((n + 1, x.size))
    P[0] = 1.0
    if n >= 1:
        P[1] = t
        dP[1] = 1.0
    for k in range(1, n):
        P[k + 1] = ((2 * k + 1) * t * P[k] - k * P[k - 1]) / (k + 1)
        dP[k + 1] = dP[k - 1] + (2 * k + 1) * P[k]
    return P, 2.0 * dP


def _raw_basis(dim: int, p: int, x: np.ndarray):
    """Legendre-product basis of P_p and its gradient at points x (npts, dim)."""
    exps = [a for a in itertools.product(range(p + 1), repeat=dim) if sum(a) <= p]
    Ps = [_legendre01(p, x[:, d]) for d in range(dim)]
    phi = np.ones((x.shape[0], len(exps)))
    dphi = np.ones((dim, x.shape[0], len(exps)))
    for m, a in enumerate(exps):
        for d in range(dim):
            phi[:, m] *= Ps[d][0][a[d]]
            for dd in range(dim):
                dphi[dd, :, m] *= Ps[d][1][a[d]] if d == dd else Ps[d][0][a[d]]
    return phi, dphi


def simplex_quadrature(dim: int, deg: int):
    """Duffy-collapsed Gauss-Legendre rule on the unit simplex, exact for total degree `deg`."""
    if dim == 0:
        return np.zeros((1, 0)), np.ones(1)
    n = deg // 2 + dim + 1
    t, w = np.polynomial.legendre.leggauss(n)
    t = 0.5 * (t + 1.0)
    w = 0.5 * w
    if dim == 1:
        return t[:, None], w
    if dim == 2:
        U, V = np.meshgrid(t, t, indexing="ij")
        W = np.outer(w, w) * (1.0 - U)
        return np.stack([U.ravel(), (V * (1.0 - U)).ravel()], axis=1), W.ravel()
    U, V, Z = np.meshgrid(t, t, t, indexing="ij")
    W = np.einsum("i,j,k->ijk", w, w, w) * (1.0 - U) ** 2 * (1.0 - V)
    pts = np.stack([U.ravel(), (V * (1.0 - U)).ravel(), (Z * (1.0 - U) * (1.0 - V)).ravel()], axis=1)
    return pts, W.ravel()


@dataclass
class RefElement:
    dim: int
    p: int
    nodes: np.ndarray          # (Np, dim)
    bary: np.ndarray           # (Np, dim+1) integer barycentric indices
    D: np.ndarray              # (dim, Np, Np) collocation derivatives d/dxi
    Minv: np.ndarray           # (Np, Np) inverse mass on the unit simplex
    fnodes: np.ndarray         # (dim+1, Nfp) local node ids on face f (opposite vertex f), ascending
    lift: np.ndarray           # (dim+1, Np, Nfp)  Minv @ face mass (unit (dim-1)-simplex measure)

    @property
    def Np(self):
        return self.nodes.shape[0]

    @property
    def Nfp(self):
        return self.fnodes.shape[1]


def build_ref_element(dim: int, p: int) -> RefElement:
    nodes, bary = simplex_nodes(dim, p)
    # orthonormalise the raw basis numerically -> well conditioned Vandermonde
    qx, qw = simplex_quadrature(dim, 2 * p)
    phi_q, _ = _raw_basis(dim, p, qx)
    # QR of the weighted samples instead of Cholesky of the Gram matrix: cond(A) = sqrt(cond(G)), which keeps
    # LIFT accurate to ~1e-13 in double precision (Cholesky of G loses ~1e-9 at order-3 tetrahedra)
    _, Rq = np.linalg.qr(np.sqrt(qw)[:, None] * phi_q)
    Linv_T = np.linalg.inv(Rq)              # psi = phi @ Rq^-1 is orthonormal
    phi_n, dphi_n = _raw_basis(dim, p, nodes)
    V = phi_n @ Linv_T                      # orthonormal Vandermonde
    Vinv = np.linalg.inv(V)
    D = np.stack([(dphi_n[d] @ Linv_T) @ Vinv for d in range(dim)])
    Minv = V @ V.T
    nf = dim + 1
    fnodes = np.stack([np.nonzero(bary[:, f] == 0)[0] for f in range(nf)])
    # face mass: integrate the element's Lagrange functions over face f
    verts = np.concatenate([np.zeros((1, dim)), np.eye(dim)])
    fq, fw = simplex_quadrature(dim - 1, 2 * p)
    lift = np.zeros((nf, nodes.shape[0], fnodes.shape[1]))
    for f in range(nf):
        fv = [v for v in range(nf) if v != f]
        base = verts[fv[0]]
        x = base[None, :] + fq @ (verts[fv[1:]] - base) if dim > 1 else base[None, :] + np.zeros((1, dim))
        ph, _ = _raw_basis(dim, p, x)
        ell = (ph @ Linv_T) @ Vinv           # (nq, Np) Lagrange functions at face quadrature points
        Mf = ell.T @ (fw[:, None] * ell)     # (Np, Np), zero outside face nodes
        lift[f] = Minv @ Mf[:, fnodes[f]]
    return RefElement(dim, p, nodes, bary, D, Minv, fnodes, lift)


# ----------------------------------------------------------------------------------------------
# problem description and setup
# ----------------------------------------------------------------------------------------------
@dataclass
class PlaneWave:
    spread: float
    mean1d: float
    pol: np.ndarray
    dir: np.ndarray
    freq: float = 0.0
    fieldtype: int = 0           # 0: polarisation is E, 1: polarisation is H

    def __post_init__(self):
        self.pol = np.asarray(self.pol, float) / np.linalg.norm(self.pol)
        self.dir = np.asarray(self.dir, float) / np.linalg.norm(self.dir)

    def eval6(self, xyz: np.ndarray, t: float) -> np.ndarray:
        """(6, n) incident field at points xyz (n,3) — Function.h:361-409."""
        if self.fieldtype == 0:
            pe, ph = self.pol, np.cross(self.dir, self.pol)
        else:
            ph, pe = self.pol, np.cross(self.pol, self.dir)
        u = xyz @ self.dir - t
        arg = u - self.mean1d
        if self.freq == 0.0:
            g = np.exp(-arg ** 2 / (2.0 * self.spread ** 2))
        else:
            g = np.exp(-arg * arg / (2.0 * self.spread * self.spread)) * np.cos(2.0 * math.pi * self.freq * arg)
        return np.concatenate([pe[:, None] * g[None, :], ph[:, None] * g[None, :]], axis=0)


@dataclass
class Problem:
    verts: np.ndarray            # (nv, 3)
    elems: np.ndarray            # (NE, dim+1) vertex ids, MFEM element-local order
    elem_attr: np.ndarray        # (NE,)
    bdr: np.ndarray              # (NBE, dim) vertex ids
    bdr_attr: np.ndarray         # (NBE,)
    order: int
    alpha: float = 1.0
    bdr_cond: dict = field(default_factory=dict)      # bdr attribute -> PEC/PMC/SMA
    tfsf_tags: tuple = ()
    materials: dict = field(default_factory=dict)     # element attribute -> (eps, mu, sigma)
    planewave: PlaneWave | None = None
    # "global": coefficients of GlobalEvolution / DGOperatorFactory (what the product computes and the goldens hold);
    # "hesthaven": the boundary encodings of HesthavenEvolution.cpp:275-313 (SMA jump -u/alpha with the face's alpha kept, so
    # its centred part is scaled 1/alpha; interior PEC/PMC/SMA jumps (-1,0)/(0,-1)/(-1/2,-1/2) on both sides).  The two agree
    # for alpha = 1 on meshes without interior boundaries (tests/test_oracle.py); driver.cpp:1596-1598 forbids alpha = 0 there.
    flavour: str = "global"


class HesthavenOracle:
    """Matrix-free restatement; `mult(t, x)` returns f(t, x) with x, f of shape (6N,)."""

    TFSF_SKIP = 1e-8   # GlobalEvolution.h:100

    def __init__(self, pb: Problem):
        self.pb = pb
        dim = pb.elems.shape[1] - 1
        self.dim = dim
        ref = self.ref = build_ref_element(dim, pb.order)
        NE, Np, Nfp, nf = pb.elems.shape[0], ref.Np, ref.Nfp, dim + 1
        self.NE, self.Np, self.Nfp, self.nf, self.N = NE, Np, Nfp, nf, NE * Np
        X = pb.verts[pb.elems][:, :, :3]                     # (NE, dim+1, 3)
        # affine map x = v0 + J xi ; J columns v_k - v_0 ; embed in 3x3
        J = np.zeros((NE, 3, 3))
        J[:, :, :dim] = np.transpose(X[:, 1:, :] - X[:, :1, :], (0, 2, 1))
        Jd = J[:, :dim, :dim]
        self.detJ = np.linalg.det(Jd)
        if np.any(self.detJ <= 0):
            raise ValueError("negatively oriented element")
        Jinv = np.zeros((NE, 3, 3))                           # Jinv[e, xi, d] = d xi / d x_d
        Jinv[:, :dim, :dim] = np.linalg.inv(Jd)
        self.Jinv = Jinv
        # grad lambda_f: vertex 0 -> -sum rows, vertex k -> row k-1; outward normal * fscale = -grad lambda_f
        gl = np.concatenate([-Jinv[:, :dim, :].sum(axis=1, keepdims=True), Jinv[:, :dim, :]], axis=1)  # (NE, nf, 3)
        self.fscale = np.linalg.norm(gl, axis=2)              # = |J_f| / |J_e|
        self.normal = -gl / self.fscale[:, :, None]
        # node coordinates
        self.xyz = X[:, 0, None, :] + np.einsum("edk,nk->end", J[:, :, :dim], ref.nodes)   # (NE, Np, 3)
        # materials
        eps = np.ones(NE)
        mu = np.ones(NE)
        sig = np.zeros(NE)
        for a, (e_, m_, s_) in pb.materials.items():
            sel = pb.elem_attr == a
            eps[sel], mu[sel], sig[sel] = e_, m_, s_
        self.inv_eps, self.inv_mu, self.sig_eps = 1.0 / eps, 1.0 / mu, sig / eps
        self._connect()

    # -- connectivity by barycentric matching ------------------------------------
    def _connect(self):
        pb, ref = self.pb, self.ref
        NE, Np, Nfp, nf, dim, p = self.NE, self.Np, self.Nfp, self.nf, self.dim, pb.order
        node_of = {tuple(b): n for n, b in enumerate(ref.bary)}
        faces = {}
        for e in range(NE):
            for f in range(nf):
                key = tuple(sorted(int(v) for k, v in enumerate(pb.elems[e]) if k != f))
                faces.setdefault(key, []).append((e, f))
        battr = {tuple(sorted(int(v) for v in b)): int(a) for b, a in zip(pb.bdr, pb.bdr_attr)}
        vmapM = np.zeros((NE, nf, Nfp), np.int64)
        vmapP = np.zeros((NE, nf, Nfp), np.int64)
        bc = np.zeros((NE, nf), np.int32)
        self.bc_interior = np.zeros((NE, nf), bool)           # boundary condition declared on a face between two elements
        nbr = -np.ones((NE, nf), np.int64)
        face_tag = np.zeros((NE, nf), np.int32)
        for key, sides in faces.items():
            tag = battr.get(key, 0)
            for s, (e, f) in enumerate(sides):
                vmapM[e, f] = e * Np + ref.fnodes[f]
                face_tag[e, f] = tag
                # a PEC/PMC/SMA tag on an INTERIOR face: the `global` operator skips the regular interior flux there
                # (ignore marker, DGOperatorFactory.h:373-389, bilinearform.cpp:634-690 of the fork) and assembles only the
                # two self blocks with the true-boundary coefficients (MaxwellDGInteriorJumpIntegrator, :575-675;
                # BilinearIntegrators.cpp:356-412): each side sees a boundary face
                if len(sides) == 1 or pb.bdr_cond.get(tag, 0):
                    vmapP[e, f] = vmapM[e, f]
                    bc[e, f] = pb.bdr_cond.get(tag, 0)
                    self.bc_interior[e, f] = len(sides) == 2
                    continue
                e2, f2 = sides[1 - s]
                nbr[e, f] = e2
                pos2 = {int(v): k for k, v in enumerate(pb.elems[e2])}
                for j, n in enumerate(ref.fnodes[f]):
                    b2 = [0] * (dim + 1)
                    for k, v in enumerate(pb.elems[e]):
                        if k != f:
                            b2[pos2[int(v)]] = int(ref.bary[n, k])
                    vmapP[e, f, j] = e2 * Np + node_of[tuple(b2)]
        self.vmapM, self.vmapP, self.bc, self.nbr, self.face_tag = vmapM, vmapP, bc, nbr, face_tag
        # TF/SF: side per element (1 TF, 2 SF), 3-D centroid rule
        self.tfsf_side = np.zeros(NE, np.int32)
        self.tfsf_face = np.zeros((NE, nf), np.int32)   # +1: this side is TF, -1: this side is SF
        tags = set(pb.tfsf_tags)
        if tags:
            vs = sorted({int(v) for b, a in zip(pb.bdr, pb.bdr_attr) if int(a) in tags for v in b})
            ctr = pb.verts[vs, :3].sum(axis=0) / len(vs)
            bary_e = pb.verts[pb.elems][:, :, :3].sum(axis=1) / (dim + 1)
            d2 = ((bary_e - ctr) ** 2).sum(axis=1)
            seen = 0
            if dim == 1 and sum(1 for a in pb.bdr_attr if int(a) in tags) > 2:
                raise ValueError("only one or two TF/SF points can be declared on a 1-D mesh")     # SubMesher.cpp:563
            for b, a in zip(pb.bdr, pb.bdr_attr):      # boundary-element order, as the reference loops
                if int(a) not in tags:
                    continue
                (e1, f1), (e2, f2) = sorted(faces[tuple(sorted(int(v) for v in b))])   # Elem1 = the lower element id (MFEM)
                if dim == 3:                                   # centroid rule, SubMesher.cpp:677-771
                    e1_tf = d2[e1] < d2[e2]
                elif dim == 2:                                 # orientation rule, SubMesher.cpp:568-660, 239-250, 281-290
                    ev = ((1, 2), (2, 0), (0, 1))[f1]          # Elem1's local edge opposite vertex f1, in its orientation
                    t = pb.verts[pb.elems[e1, ev[1]], :2] - pb.verts[pb.elems[e1, ev[0]], :2]
                    bb = bary_e[e2, :2] - bary_e[e1, :2]
                    e1_tf = not (bb[0] * t[1] - bb[1] * t[0] >= 0.0)
                else:                                          # SubMesher.cpp:476-547: SF|TF at the first point, TF|SF at the second
                    e1_tf = seen == 1
                seen += 1
                for (e, f, tf) in ((e1, f1, e1_tf), (e2, f2, not e1_tf)):
                    self.tfsf_face[e, f] = 1 if tf else -1
                    if not tf:
                        self.tfsf_side[e] = 2
                    elif self.tfsf_side[e] == 0:
                        self.tfsf_side[e] = 1
            # the `global` flavour masks per ELEMENT (SourcesManager.cpp:158-188): a face sees the
            # masks of its two elements, which is what tfsf_face must reproduce
            for e, f in zip(*np.nonzero(self.tfsf_face)):
                self.tfsf_face[e, f] = 1 if self.tfsf_side[e] == 1 else -1

    # -- the right-hand side ---------------------------------------------------------------------
    def tfsf_gate(self, t: float) -> bool:
        """True when the `global` operator injects at time t (norm test over the masked sub-mesh DOFs)."""
        pw = self.pb.planewave
        if pw is None or not self.tfsf_side.any():
            return False
        sel = self.tfsf_side > 0
        s = 0.5 * pw.eval6(self.xyz[sel].reshape(-1, 3), t)
        return float((s ** 2).sum()) >= self.TFSF_SKIP ** 2

    def mult(self, t: float, x: np.ndarray) -> np.ndarray:
        pb, ref = self.pb, self.ref
        NE, Np, Nfp, nf = self.NE, self.Np, self.Nfp, self.nf
        u = x.reshape(6, NE, Np)
        uf = x.reshape(6, NE * Np)
        # volume: curl via reference derivatives and geometric factors
        dxi = np.einsum("xij,cej->xcei", ref.D, u)                       # (dim, 6, NE, Np)
        grad = np.einsum("exd,xcei->dcei", self.Jinv[:, : self.dim, :], dxi)   # (3, 6, NE, Np)  d u_c / d x_d
        E, H = slice(0, 3), slice(3, 6)

        def curl(g):   # g: (3 d, 3 comp, NE, Np)
            return np.stack([g[1, 2] - g[2, 1], g[2, 0] - g[0, 2], g[0, 1] - g[1, 0]])

        rhsE = curl(grad[:, H])
        rhsH = -curl(grad[:, E])
        # jumps (neighbour - self), HesthavenEvolution.cpp:474-479
        uM = uf[:, self.vmapM]                                           # (6, NE, nf, Nfp)
        uP = uf[:, self.vmapP]
        dU = uP - uM
        alpha = np.full((NE, nf), pb.alpha)
        hest = pb.flavour == "hesthaven"
        if hest and pb.alpha == 0.0 and (self.bc == SMA).any():
            raise ValueError("alpha = 0 with SMA boundaries is not defined in the hesthaven flavour (driver.cpp:1596-1598)")
        sma = -1.0 / pb.alpha if hest and pb.alpha != 0.0 else -1.0   # HesthavenEvolution.cpp:292 vs DGOperatorFactory.h:483-496
        for code, (ce, ch) in {PEC: (-2.0, 0.0), PMC: (0.0, -2.0), SMA: (sma, sma)}.items():
            m = self.bc == code
            if m.any():
                dU[:3, m] = ce * uM[:3, m]
                dU[3:, m] = ch * uM[3:, m]
        if hest:                                              # interior boundaries, HesthavenEvolution.cpp:308-310
            for code, (ce, ch) in {PEC: (-1.0, 0.0), PMC: (0.0, -1.0), SMA: (-0.5, -0.5)}.items():
                m = (self.bc == code) & self.bc_interior
                if m.any():
                    dU[:3, m] = ce * uM[:3, m]
                    dU[3:, m] = ch * uM[3:, m]
        else:
            alpha[self.bc == SMA] = 1.0
        if pb.planewave is not None and self.tfsf_face.any() and self.tfsf_gate(t):
            e_i, f_i = np.nonzero(self.tfsf_face)
            pts = self.xyz.reshape(-1, 3)[self.vmapM[e_i, f_i]]          # (nfaces, Nfp, 3)
            inc = pb.planewave.eval6(pts.reshape(-1, 3), t).reshape(6, -1, Nfp)
            sgn = self.tfsf_face[e_i, f_i].astype(float)                # TF side: nbr = u_SF + inc ; SF side: nbr = u_TF - inc
            dU[:, e_i, f_i] += sgn[None, :, None] * inc
        n = np.transpose(self.normal, (2, 0, 1))[:, :, :, None]          # (3, NE, nf, 1)
        dE, dH = dU[:3], dU[3:]
        ndE = (n * dE).sum(axis=0)
        ndH = (n * dH).sum(axis=0)

        def cross(a, b):
            return np.stack([a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]])

        a = alpha[None, :, :, None]
        fluxE = cross(n, dH) + a * (dE - ndE[None] * n)                  # :518
        fluxH = -cross(n, dE) + a * (dH - ndH[None] * n)                 # :517
        sc = 0.5 * self.fscale[None, :, :, None]
        rhsE = rhsE + np.einsum("fij,cefj->cei", ref.lift, fluxE * sc)
        rhsH = rhsH + np.einsum("fij,cefj->cei", ref.lift, fluxH * sc)
        rhsE = rhsE * self.inv_eps[None, :, None] - self.sig_eps[None, :, None] * u[E]
        rhsH = rhsH * self.inv_mu[None, :, None]
        return np.concatenate([rhsE, rhsH]).reshape(-1)

    # -- mfem::RK4Solver::Step (ode.cpp:109-136) -------------------------------------------------
    def rk4_step(self, x: np.ndarray, t: float, dt: float) -> np.ndarray:
        k = self.mult(t, x)
        y = x + (dt / 2) * k
        z = x + (dt / 6) * k
        k = self.mult(t + dt / 2, y)
        y = x + (dt / 2) * k
        z = z + (dt / 3) * k
        k = self.mult(t + dt / 2, y)        # time is NOT advanced between k2 and k3
        y = x + dt * k
        z = z + (dt / 3) * k
        k = self.mult(t + dt, y)
        return z + (dt / 6) * k


def load_case(path: str, **over) -> tuple[Problem, dict]:
    """Read a fixture directory written by `dgtd_ref gen --out` (or tests/golden/*.npz)."""
    import json
    import os

    if path.endswith(".npz"):
        z = np.load(path, allow_pickle=False)
        meta = json.loads(str(z["meta"]))
        get = lambda k: z[k]
    else:
        meta = json.load(open(os.path.join(path, "meta.json")))
        dt = {"f64": np.float64, "i32": np.int32}
        get = lambda k: np.fromfile(os.path.join(path, k.replace("_f64", ".f64").replace("_i32", ".i32")),
                                    dt[k[-3:]])
    dim = meta["dim"]
    pb = Problem(
        verts=get("verts_f64").reshape(-1, 3),
        elems=get("elems_i32").reshape(-1, dim + 1).astype(np.int64),
        elem_attr=get("elem_attr_i32"),
        bdr=get("bdr_i32").reshape(-1, dim).astype(np.int64),
        bdr_attr=get("bdr_attr_i32"),
        order=meta["order"], alpha=meta["alpha"], **over)
    if meta.get("pw", {}).get("on"):
        w = meta["pw"]
        pb.planewave = PlaneWave(w["spread"], w["mean1d"], w["pol"], w["dir"], w["freq"])
    data = {k: get(k) for k in ("x0_f64", "k0_f64", "x_final_f64", "nodes_f64")}
    return pb, {"meta": meta, **data}
