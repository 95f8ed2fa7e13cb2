// Flat operator data for one rank: geometry factors, face connectivity, boundary / TF-SF codes, halo plan.
// Replaces (reference file:line)
//   Connectivities (vmapM/vmapP, mapB/vmapB, TF/SF maps)  src/evolution/HesthavenEvolutionMethods.cpp:501-534, 721-791
//   normals / fscale / per-element D                       src/evolution/HesthavenEvolution.cpp:150-205
//   TF/SF side classification (3-D centroid rule)          src/components/SubMesher.cpp:677-771
//   +-1/2 element mask of the `global` source vector       src/solver/SourcesManager.cpp:158-188
//   ParMesh ghost layer / send_face_nbr_ldof               external/mfem-geg/fem/pfespace.cpp:1258-1332 (whole elements
//                                                           there; face traces only here)
// Face nodes are matched combinatorially through integer barycentric indices, never by assembling
// two-element flux matrices as the reference does (HesthavenEvolutionMethods.cpp:59-75, 342-370).
#include "host.hpp"
#include "../../include/dgtd_b200.h"

#include <algorithm>
#include <cmath>
#include <map>
#include <numeric>
#include <set>

namespace dgtd {
namespace {

struct FaceKey {
    int v[3];
    int e, f;
    bool operator<(const FaceKey &o) const { return std::lexicographical_compare(v, v + 3, o.v, o.v + 3); }
    bool same(const FaceKey &o) const { return v[0] == o.v[0] && v[1] == o.v[1] && v[2] == o.v[2]; }
};

void elem_geometry(const Mesh &m, int e, double J[3][3], double Jinv[3][3], double &det)
{
    const int d = m.dim;
    const int *v = &m.elems[(size_t)e * (d + 1)];
    for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) { J[r][c] = 0; Jinv[r][c] = 0; }
    for (int k = 0; k < d; k++) for (int c = 0; c < 3; c++) J[c][k] = m.verts[3 * (size_t)v[k + 1] + c] - m.verts[3 * (size_t)v[0] + c];
    if (d == 1) { det = J[0][0]; Jinv[0][0] = 1.0 / det; }
    else if (d == 2) {
        det = J[0][0] * J[1][1] - J[0][1] * J[1][0];
        Jinv[0][0] = J[1][1] / det; Jinv[0][1] = -J[0][1] / det;
        Jinv[1][0] = -J[1][0] / det; Jinv[1][1] = J[0][0] / det;
    } else {
        double c00 = J[1][1] * J[2][2] - J[1][2] * J[2][1], c01 = J[1][2] * J[2][0] - J[1][0] * J[2][2], c02 = J[1][0] * J[2][1] - J[1][1] * J[2][0];
        det = J[0][0] * c00 + J[0][1] * c01 + J[0][2] * c02;
        Jinv[0][0] = c00 / det; Jinv[1][0] = c01 / det; Jinv[2][0] = c02 / det;
        Jinv[0][1] = (J[0][2] * J[2][1] - J[0][1] * J[2][2]) / det;
        Jinv[1][1] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) / det;
        Jinv[2][1] = (J[0][1] * J[2][0] - J[0][0] * J[2][1]) / det;
        Jinv[0][2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) / det;
        Jinv[1][2] = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) / det;
        Jinv[2][2] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) / det;
    }
    // Jinv[xi][d] = d xi / d x_d
}

// Sort the rank's elements along a Morton curve of their barycentres, quantised to about one element diameter
// (ties keep the mesh order, so the six tetrahedra of a Kuhn cube stay together).
void morton_order(const Mesh &m, std::vector<int> &ids)
{
    const int nf = m.dim + 1, n = (int)ids.size();
    if (n < 2) return;
    std::vector<double> bc(3 * (size_t)n, 0.0);
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    for (int i = 0; i < n; i++)
        for (int c = 0; c < 3; c++) {
            double s = 0;
            for (int k = 0; k < nf; k++) s += m.verts[3 * (size_t)m.elems[(size_t)ids[i] * nf + k] + c];
            s /= nf; bc[3 * (size_t)i + c] = s;
            lo[c] = std::min(lo[c], s); hi[c] = std::max(hi[c], s);
        }
    double vol = 1.0; int nd = 0;
    for (int c = 0; c < 3; c++) if (hi[c] - lo[c] > 0) { vol *= hi[c] - lo[c]; nd++; }
    if (nd == 0) return;
    const double cell = 1.817 * std::pow(vol / n, 1.0 / nd);
    auto spread = [](uint64_t v) {   // 21 bits -> every third bit
        v &= 0x1fffff;
        v = (v | v << 32) & 0x1f00000000ffffULL; v = (v | v << 16) & 0x1f0000ff0000ffULL; v = (v | v << 8) & 0x100f00f00f00f00fULL;
        v = (v | v << 4) & 0x10c30c30c30c30c3ULL; v = (v | v << 2) & 0x1249249249249249ULL;
        return v;
    };
    std::vector<std::pair<uint64_t, int>> key(n);
    for (int i = 0; i < n; i++) {
        uint64_t q[3];
        for (int c = 0; c < 3; c++) q[c] = (uint64_t)std::min(2097151.0, std::floor((bc[3 * (size_t)i + c] - lo[c]) / cell + 1e-9));
        key[i] = {spread(q[0]) | spread(q[1]) << 1 | spread(q[2]) << 2, ids[i]};
    }
    std::stable_sort(key.begin(), key.end(), [](const auto &a, const auto &b) { return a.first < b.first; });
    for (int i = 0; i < n; i++) ids[i] = key[i].second;
}

// Groups of BLK_E consecutive elements are what one warp carries through a stage (kernels_wg.cuh): a face whose neighbour
// sits in the same group is served from the warp's own shared-memory buffer instead of L2.  Morton order alone puts 41 % of
// the faces of a Kuhn box in-group (30 % on a rank's part of a partitioned box); growing each group from a Morton-ordered
// seed by always adding the unassigned element that shares most faces with the group reaches 50 %.
// `ids` (Morton-sorted, all with mark[e] == tag) is rewritten: whole groups first, in seed order; elements of groups that
// could not be filled (islands) are returned in `left`.  `init` pre-seeds groups (partition-face leftovers, see below).
void grow_groups(const std::vector<int> &nbrE, int nf, std::vector<int> &mark, int tag, const std::vector<int> &pos,
                 std::vector<int> &ids, std::vector<int> &left, const std::vector<std::vector<int>> &init = {})
{
    std::vector<int> out; out.reserve(ids.size());
    std::vector<std::pair<int, int>> cand;      // (element, faces shared with the group so far)
    std::vector<int> grp;
    auto add = [&](int e) {
        grp.push_back(e); mark[(size_t)e] = -1 - tag;          // assigned
        for (int f = 0; f < nf; f++) {
            const int e2 = nbrE[(size_t)e * nf + f];
            if (e2 < 0 || mark[(size_t)e2] != tag) continue;
            bool found = false;
            for (auto &c : cand) if (c.first == e2) { c.second++; found = true; break; }
            if (!found) cand.push_back({e2, 1});
        }
    };
    auto grow = [&]() {
        while ((int)grp.size() < BLK_E && !cand.empty()) {
            size_t best = 0;
            for (size_t i = 1; i < cand.size(); i++)
                if (cand[i].second > cand[best].second || (cand[i].second == cand[best].second && pos[(size_t)cand[i].first] < pos[(size_t)cand[best].first])) best = i;
            const int e = cand[best].first;
            cand.erase(cand.begin() + (long)best);
            add(e);
        }
    };
    size_t scan = 0;
    for (const auto &g0 : init) {               // pre-seeded groups are always emitted, filled from `ids` in order when adjacency runs out
        grp.clear(); cand.clear();
        for (int e : g0) { const int keep = mark[(size_t)e]; add(e); mark[(size_t)e] = keep; }
        grow();
        while ((int)grp.size() < BLK_E) {
            while (scan < ids.size() && mark[(size_t)ids[scan]] != tag) scan++;
            if (scan == ids.size()) break;
            add(ids[scan]); grow();
        }
        out.insert(out.end(), grp.begin(), grp.end());
    }
    for (int seed : ids) {
        if (mark[(size_t)seed] != tag) continue;
        grp.clear(); cand.clear();
        add(seed); grow();
        if ((int)grp.size() == BLK_E) out.insert(out.end(), grp.begin(), grp.end());
        else for (int e : grp) left.push_back(e);
    }
    ids.swap(out);
}

}  // namespace

void node_coords(const Mesh &m, const RefElem &ref, std::vector<double> &xyz)
{
    const int d = m.dim, Np = ref.Np;
    xyz.assign((size_t)m.ne() * Np * 3, 0.0);
    for (int e = 0; e < m.ne(); e++) {
        const int *v = &m.elems[(size_t)e * (d + 1)];
        for (int n = 0; n < Np; n++) for (int c = 0; c < 3; c++) {
            double x0 = m.verts[3 * (size_t)v[0] + c], s = x0;
            for (int k = 0; k < d; k++) s += (m.verts[3 * (size_t)v[k + 1] + c] - x0) * ref.nodes[(size_t)n * d + k];
            xyz[((size_t)e * Np + n) * 3 + c] = s;
        }
    }
}

// (element, local face) pairs lying on boundary elements with one of the given attributes; for interior surfaces both
// sides are listed, lower element id first.  Selection step of NearToFarFieldSubMesher (SubMesher.cpp:832-905).
std::vector<int> boundary_element_faces(const Mesh &m, const std::vector<int> &attrs)
{
    const int dim = m.dim, nf = dim + 1, NE = m.ne();
    std::set<int> want(attrs.begin(), attrs.end());
    std::vector<FaceKey> sorted((size_t)NE * nf);
    for (int e = 0; e < NE; e++) for (int f = 0; f < nf; f++) {
        FaceKey &k = sorted[(size_t)e * nf + f];
        k.v[0] = k.v[1] = k.v[2] = -1; int c = 0;
        for (int q = 0; q < nf; q++) if (q != f) k.v[c++] = m.elems[(size_t)e * nf + q];
        std::sort(k.v, k.v + dim);
        k.e = e; k.f = f;
    }
    std::sort(sorted.begin(), sorted.end(), [](const FaceKey &a, const FaceKey &b) { return a < b || (!(b < a) && a.e < b.e); });
    std::vector<int> out;
    for (int b = 0; b < m.nbe(); b++) {
        if (!want.count(m.bdr_attr[b])) continue;
        FaceKey k; k.v[0] = k.v[1] = k.v[2] = -1; k.e = k.f = 0;
        for (int c = 0; c < dim; c++) k.v[c] = m.bdr[(size_t)b * dim + c];
        std::sort(k.v, k.v + dim);
        auto it = std::lower_bound(sorted.begin(), sorted.end(), k);
        if (it == sorted.end() || !it->same(k)) throw Error(DGTD_ERR_MESH, "boundary element " + std::to_string(b) + " is not a face of any element");
        for (; it != sorted.end() && it->same(k); ++it) { out.push_back(it->e); out.push_back(it->f); }
    }
    return out;
}

HostOp build_host_op(const Mesh &m, const Options &o)
{
    HostOp H;
    const int dim = m.dim, nf = dim + 1, NE = m.ne();
    if (NE == 0) throw Error(DGTD_ERR_MESH, "empty mesh");
    if (o.order < 1 || o.order > (dim == 3 ? 5 : 6)) throw Error(DGTD_ERR_UNSUPPORTED, "order out of range for this dimension");
    if (!(o.alpha >= 0.0 && o.alpha <= 1.0)) throw Error(DGTD_ERR_ARG, "upwind_alpha must lie in [0,1]");
    if (o.nranks < 1 || o.rank < 0 || o.rank >= o.nranks) throw Error(DGTD_ERR_ARG, "bad rank/nranks");
    H.ref = build_ref_element(dim, o.order);
    const RefElem &R = H.ref;
    const int Np = R.Np, Nfp = R.Nfp;
    H.dim = dim; H.p = o.order; H.Np = Np; H.Nfp = Nfp; H.nf = nf; H.NEglob = NE;
    H.alpha = o.alpha; H.pw = o.pw; H.tfsf_gate = o.tfsf_gate;
    if ((long long)NE * Np > 2000000000LL) throw Error(DGTD_ERR_UNSUPPORTED, "more than 2^31 dofs per field on one rank");

    // ---- global face pairing -------------------------------------------------------------------------------
    std::vector<FaceKey> keys((size_t)NE * nf);
    for (int e = 0; e < NE; e++) for (int f = 0; f < nf; f++) {
        FaceKey &k = keys[(size_t)e * nf + f];
        k.v[0] = k.v[1] = k.v[2] = -1; int c = 0;
        for (int q = 0; q < nf; q++) if (q != f) k.v[c++] = m.elems[(size_t)e * nf + q];
        std::sort(k.v, k.v + dim);
        k.e = e; k.f = f;
    }
    std::vector<FaceKey> sorted = keys;
    std::sort(sorted.begin(), sorted.end());
    std::vector<int> nbrE((size_t)NE * nf, -1), nbrF((size_t)NE * nf, -1);
    for (size_t i = 0; i < sorted.size();) {
        size_t j = i + 1;
        while (j < sorted.size() && sorted[j].same(sorted[i])) j++;
        if (j - i > 2) throw Error(DGTD_ERR_MESH, "non-manifold mesh: a face is shared by more than two elements");
        if (j - i == 2) {
            const FaceKey &a = sorted[i], &b = sorted[i + 1];
            nbrE[(size_t)a.e * nf + a.f] = b.e; nbrF[(size_t)a.e * nf + a.f] = b.f;
            nbrE[(size_t)b.e * nf + b.f] = a.e; nbrF[(size_t)b.e * nf + b.f] = a.f;
        }
        i = j;
    }
    // ---- boundary elements -> faces ------------------------------------------------------------------------
    std::map<int, int> bcOf, matIdx;
    for (auto &kv : o.bdr) {
        if (kv.second < DGTD_BC_NONE || kv.second > DGTD_BC_SMA) throw Error(DGTD_ERR_ARG, "unknown boundary condition code");
        bcOf[kv.first] = kv.second;
    }
    std::set<int> tfsfTags(o.tfsf.begin(), o.tfsf.end());
    std::vector<int> faceBC((size_t)NE * nf, 0);
    std::vector<int> bdrElemFace(m.nbe(), -1);       // index into sorted[] of the first side
    for (int b = 0; b < m.nbe(); b++) {
        FaceKey k; k.v[0] = k.v[1] = k.v[2] = -1;
        for (int c = 0; c < dim; c++) k.v[c] = m.bdr[(size_t)b * dim + c];
        std::sort(k.v, k.v + dim);
        auto it = std::lower_bound(sorted.begin(), sorted.end(), k);
        if (it == sorted.end() || !it->same(k)) throw Error(DGTD_ERR_MESH, "boundary element " + std::to_string(b) + " is not a face of any element");
        bdrElemFace[b] = (int)(it - sorted.begin());
        const int attr = m.bdr_attr[b];
        const bool interior = nbrE[(size_t)it->e * nf + it->f] >= 0;
        if (tfsfTags.count(attr)) {
            if (!interior) throw Error(DGTD_ERR_UNSUPPORTED, "TF/SF tag on a true boundary face");
            continue;
        }
        auto bc = bcOf.find(attr);
        if (bc == bcOf.end()) continue;
        faceBC[(size_t)it->e * nf + it->f] = bc->second;
        if (interior) {
            // PEC/PMC/SMA sheet inside the mesh, `global` semantics: the regular interior flux skips the face (ignore marker,
            // DGOperatorFactory.h:373-389) and each side gets the self block of a true boundary face with the same
            // coefficients (MaxwellDGInteriorJumpIntegrator, DGOperatorFactory.h:575-675, BilinearIntegrators.cpp:356-412):
            // the two elements are disconnected here and both faces become boundary faces with this condition.
            // (`hesthaven` halves the coefficients instead, HesthavenEvolution.cpp:308-310; the default operator is followed.)
            const int e2 = nbrE[(size_t)it->e * nf + it->f], f2 = nbrF[(size_t)it->e * nf + it->f];
            faceBC[(size_t)e2 * nf + f2] = bc->second;
            nbrE[(size_t)it->e * nf + it->f] = nbrF[(size_t)it->e * nf + it->f] = -1;
            nbrE[(size_t)e2 * nf + f2] = nbrF[(size_t)e2 * nf + f2] = -1;
        }
    }
    // ---- TF/SF sides (SubMesher.cpp:677-771) ---------------------------------------------------------------
    std::vector<int> side(NE, 0), faceTF((size_t)NE * nf, 0);
    if (!tfsfTags.empty() && o.pw.enabled) {
        std::set<int> counted; double ctr[3] = {0, 0, 0};
        for (int b = 0; b < m.nbe(); b++) if (tfsfTags.count(m.bdr_attr[b]))
            for (int c = 0; c < dim; c++) { int v = m.bdr[(size_t)b * dim + c]; if (counted.insert(v).second) for (int q = 0; q < 3; q++) ctr[q] += m.verts[3 * (size_t)v + q]; }
        if (!counted.empty()) for (int q = 0; q < 3; q++) ctr[q] /= (double)counted.size();
        auto bary = [&](int e, double *bc3) {
            for (int q = 0; q < 3; q++) bc3[q] = 0;
            for (int k = 0; k < nf; k++) for (int q = 0; q < 3; q++) bc3[q] += m.verts[3 * (size_t)m.elems[(size_t)e * nf + k] + q];
            for (int q = 0; q < 3; q++) bc3[q] /= nf;
        };
        auto dist2 = [&](int e) {
            double bc3[3]; bary(e, bc3);
            double s = 0; for (int q = 0; q < 3; q++) s += (bc3[q] - ctr[q]) * (bc3[q] - ctr[q]);
            return s;
        };
        int nTfsfBdr = 0;
        for (int b = 0; b < m.nbe(); b++) nTfsfBdr += tfsfTags.count(m.bdr_attr[b]) ? 1 : 0;
        if (dim == 1 && nTfsfBdr > 2) throw Error(DGTD_ERR_ARG, "only one or two TF/SF points can be declared on a 1-D mesh");   // SubMesher.cpp:563
        std::vector<std::pair<int, int>> tfFaces;
        int seen = 0;
        for (int b = 0; b < m.nbe(); b++) {
            if (!tfsfTags.count(m.bdr_attr[b])) continue;
            // Elem1 = the face's first element in MFEM = lower element id
            const FaceKey &s0 = sorted[bdrElemFace[b]], &s1 = sorted[bdrElemFace[b] + 1];
            const FaceKey &a = s0.e < s1.e ? s0 : s1, &c = s0.e < s1.e ? s1 : s0;
            bool e1tf;
            if (dim == 3) e1tf = dist2(a.e) < dist2(c.e);          // centroid rule (SubMesher.cpp:677-771)
            else if (dim == 2) {
                // orientation rule (SubMesher.cpp:568-660): cross(barycentre(Elem1) -> barycentre(Elem2), face tangent)_z >= 0
                // makes Elem1 the scattered-field side; the face's vertices are Elem1's local edge (MFEM GenerateFaces)
                static const int ev[3][2] = {{1, 2}, {2, 0}, {0, 1}};      // edge opposite vertex f, in the element's orientation
                const int v0 = m.elems[(size_t)a.e * nf + ev[a.f][0]], v1 = m.elems[(size_t)a.e * nf + ev[a.f][1]];
                double b1[3], b2[3]; bary(a.e, b1); bary(c.e, b2);
                const double tx = m.verts[3 * (size_t)v1] - m.verts[3 * (size_t)v0], ty = m.verts[3 * (size_t)v1 + 1] - m.verts[3 * (size_t)v0 + 1];
                const double ori = (b2[0] - b1[0]) * ty - (b2[1] - b1[1]) * tx;
                e1tf = !(ori >= 0.0);
            } else {
                // one point: Elem1 scattered, Elem2 total; two points, in boundary-element order: the first like that, the
                // second the other way round, so that the total field lies between them (SubMesher.cpp:476-547)
                e1tf = seen == 1;
            }
            seen++;
            auto mark = [&](int e, bool tf) { if (!tf) side[e] = 2; else if (side[e] == 0) side[e] = 1; };
            mark(a.e, e1tf); mark(c.e, !e1tf);
            tfFaces.push_back({a.e, a.f}); tfFaces.push_back({c.e, c.f});
        }
        for (auto &ef : tfFaces) faceTF[(size_t)ef.first * nf + ef.second] = side[ef.first];   // the element's mask decides
    }
    // ---- neighbour node tables -----------------------------------------------------------------------------
    std::map<std::vector<uint8_t>, int> tabIdx;
    auto tabOf = [&](const std::vector<uint8_t> &row) {
        auto it = tabIdx.find(row);
        if (it != tabIdx.end()) return it->second;
        int id = (int)tabIdx.size(); tabIdx[row] = id;
        H.ftab.insert(H.ftab.end(), row.begin(), row.end());
        return id;
    };
    for (int f = 0; f < nf; f++) {   // rows 0..nf-1: identity (boundary faces read their own trace)
        std::vector<uint8_t> row(Nfp); for (int j = 0; j < Nfp; j++) row[j] = (uint8_t)R.fnodes[(size_t)f * Nfp + j];
        tabOf(row);
    }
    auto nbrRow = [&](int e, int f, int e2) {
        // node of e2 coinciding with face node j of (e,f): move the barycentric integers onto e2's vertex order
        std::vector<uint8_t> row(Nfp);
        int pos2[4];
        for (int k = 0; k < nf; k++) {
            pos2[k] = -1;
            if (k == f) continue;
            for (int q = 0; q < nf; q++) if (m.elems[(size_t)e2 * nf + q] == m.elems[(size_t)e * nf + k]) pos2[k] = q;
            if (pos2[k] < 0) throw Error(DGTD_ERR_MESH, "face vertex mismatch");
        }
        for (int j = 0; j < Nfp; j++) {
            int n = R.fnodes[(size_t)f * Nfp + j], b2[4] = {0, 0, 0, 0};
            for (int k = 0; k < nf; k++) if (k != f) b2[pos2[k]] = R.bary[(size_t)n * nf + k];
            int n2 = R.lookup(b2);
            if (n2 < 0) throw Error(DGTD_ERR_MESH, "face node match failed");
            row[j] = (uint8_t)n2;
        }
        return row;
    };
    // ---- partition -----------------------------------------------------------------------------------------
    std::vector<int> part;
    if (o.nranks > 1) {
        part = o.partitioning.empty() ? partition_rcb(m, o.nranks) : o.partitioning;
        if ((int)part.size() != NE) throw Error(DGTD_ERR_ARG, "partitioning must have one entry per element");
        for (int r : part) if (r < 0 || r >= o.nranks) throw Error(DGTD_ERR_ARG, "partitioning entry out of range");
    } else part.assign(NE, 0);
    std::vector<int> g2l(NE, -1);
    for (int e = 0; e < NE; e++) if (part[e] == o.rank) H.elem_gid.push_back(e);
    const int NEloc = H.NEloc = (int)H.elem_gid.size();
    if (NEloc == 0) throw Error(DGTD_ERR_ARG, "rank owns no elements");
    if (dim == 3) {
        // locality: groups of BLK_E consecutive local elements share as many faces as possible (grow_groups), groups follow
        // a Morton curve.  Elements that own a partition face come FIRST, grouped among themselves: they fill whole groups
        // instead of being scattered over three times as many, so the per-group cost of the halo hand-shake is paid in fewer
        // groups and at the start of a launch, and their traces are at the neighbours long before the next launch asks.
        std::vector<int> front, rest;
        std::vector<int> mark((size_t)NE, 0), pos((size_t)NE, 0);          // mark: 1 partition-face element, 2 other owned element
        for (int e : H.elem_gid) {
            bool cut = false;
            for (int f = 0; f < nf; f++) { const int e2 = nbrE[(size_t)e * nf + f]; cut |= e2 >= 0 && part[e2] != o.rank; }
            (cut ? front : rest).push_back(e);
            mark[(size_t)e] = cut ? 1 : 2;
        }
        morton_order(m, front); morton_order(m, rest);
        // Measured (profiles/r2_group_order.txt): on a rank's part of a partitioned mesh Morton order alone leaves only 30 % of the
        // faces in-group and the grown groups are 0.7 % faster; on a whole mesh they are neutral (23.6 M DOFs) to 2-4 % slower
        // (L2-resident configs 3 and 4: the in-group records are read through shared-memory banks that alias every second
        // element), so a single rank keeps plain Morton order.  DGTD_B200_ORDER=morton|grow overrides.
        const char *oenv = std::getenv("DGTD_B200_ORDER");
        const bool grow = oenv ? std::string(oenv) == "grow" : o.nranks > 1;
        if (grow) {
            std::vector<int> all(front); all.insert(all.end(), rest.begin(), rest.end());
            { std::vector<int> t(all); morton_order(m, t); for (size_t i = 0; i < t.size(); i++) pos[(size_t)t[i]] = (int)i; }
            auto by_pos = [&](int a, int b) { return pos[(size_t)a] < pos[(size_t)b]; };
            std::vector<int> leftF, leftR, left2;
            grow_groups(nbrE, nf, mark, 1, pos, front, leftF);
            std::sort(leftF.begin(), leftF.end(), by_pos);
            std::vector<std::vector<int>> init;                              // partition-face leftovers: filled up with other elements
            for (size_t i = 0; i < leftF.size(); i += BLK_E) init.emplace_back(leftF.begin() + (long)i, leftF.begin() + (long)std::min(leftF.size(), i + BLK_E));
            grow_groups(nbrE, nf, mark, 2, pos, rest, leftR, init);
            std::sort(leftR.begin(), leftR.end(), by_pos);
            for (int e : leftR) mark[(size_t)e] = 3;                         // islands: one more pass among themselves
            grow_groups(nbrE, nf, mark, 3, pos, leftR, left2);
            std::sort(left2.begin(), left2.end(), by_pos);
            rest.insert(rest.end(), leftR.begin(), leftR.end());
            rest.insert(rest.end(), left2.begin(), left2.end());
        }
        H.elem_gid = front;
        H.elem_gid.insert(H.elem_gid.end(), rest.begin(), rest.end());
    }
    for (int le = 0; le < NEloc; le++) g2l[H.elem_gid[le]] = le;
    // shared faces, ordered per peer by (owner-of-lower-rank element id, its face): both sides enumerate identically
    struct Shared { int peer, keyE, keyF, le, f, ge2, f2; };
    std::vector<Shared> shared;
    for (int le = 0; le < NEloc; le++) {
        int e = H.elem_gid[le];
        for (int f = 0; f < nf; f++) {
            int e2 = nbrE[(size_t)e * nf + f];
            if (e2 < 0 || part[e2] == o.rank) continue;
            int f2 = nbrF[(size_t)e * nf + f], pr = part[e2];
            bool mineLow = o.rank < pr;
            shared.push_back({pr, mineLow ? e : e2, mineLow ? f : f2, le, f, e2, f2});
        }
    }
    std::sort(shared.begin(), shared.end(), [](const Shared &a, const Shared &b) {
        if (a.peer != b.peer) return a.peer < b.peer;
        if (a.keyE != b.keyE) return a.keyE < b.keyE;
        return a.keyF < b.keyF;
    });
    H.n_halo_faces = (int)shared.size();
    // every rank holds the whole mesh and the partitioning, so the peer's own enumeration is known without a message:
    // shared[r][q] = faces rank r shares with rank q; a peer lists its blocks by ascending peer rank (as above)
    std::map<std::pair<int, int>, int> sharedCnt;
    if (o.nranks > 1)
        for (int e = 0; e < NE; e++)
            for (int f = 0; f < nf; f++) {
                const int e2 = nbrE[(size_t)e * nf + f];
                if (e2 >= 0 && part[e2] != part[e]) sharedCnt[{part[e], part[e2]}]++;
            }
    auto remote_block = [&](int peer, int &off, int &idx) {   // where MY block starts in the peer's halo slots / peer list
        off = idx = 0;
        for (auto it = sharedCnt.lower_bound({peer, 0}); it != sharedCnt.end() && it->first.first == peer && it->first.second < o.rank; ++it) { off += it->second; idx++; }
    };
    std::map<std::pair<int, int>, int> haloSlot;   // (le, f) -> slot
    for (size_t s = 0; s < shared.size(); s++) {
        const Shared &sh = shared[s];
        if (H.peers.empty() || H.peers.back().rank != sh.peer) { PeerPlan pp; pp.rank = sh.peer; pp.send_off = pp.recv_off = (int)s; remote_block(sh.peer, pp.remote_off, pp.remote_idx); H.peers.push_back(pp); }
        H.peers.back().nfaces++;
        haloSlot[{sh.le, sh.f}] = (int)s;
        // what I send for this face: my nodes in the RECEIVER's face-node order
        int e = H.elem_gid[sh.le];
        auto row = nbrRow(sh.ge2, sh.f2, e);          // for receiver's face node j -> my local node
        for (int j = 0; j < Nfp; j++) H.send_node.push_back(sh.le * Np + row[j]);
    }
    // ---- per-element records -------------------------------------------------------------------------------
    std::map<int, std::array<double, 3>> mats;
    for (auto &kv : o.mat) {
        if (!(kv.second[0] > 0 && kv.second[1] > 0) || kv.second[2] < 0) throw Error(DGTD_ERR_ARG, "material needs eps > 0, mu > 0, sigma >= 0");
        mats[kv.first] = kv.second;
    }
    H.geo.assign((size_t)NEloc * GEO_STRIDE, 0.0);
    H.jac.assign((size_t)NEloc * 10, 0.0);
    H.finfo.assign((size_t)NEloc * 4 * 2, 0);
    H.tfsf_side.assign(NEloc, 0);
    std::vector<double> xyzTF;
    for (int le = 0; le < NEloc; le++) {
        const int e = H.elem_gid[le];
        double J[3][3], Ji[3][3], det;
        elem_geometry(m, e, J, Ji, det);
        if (!(det > 0)) throw Error(DGTD_ERR_MESH, "element with non-positive Jacobian");
        double *g = &H.geo[(size_t)le * GEO_STRIDE];
        for (int x = 0; x < 3; x++) for (int d = 0; d < 3; d++) g[3 * x + d] = Ji[x][d];
        for (int d = 0; d < 3; d++) for (int a = 0; a < 3; a++) H.jac[(size_t)le * 10 + 3 * d + a] = J[d][a];
        H.jac[(size_t)le * 10 + 9] = det;
        for (int f = 0; f < nf; f++) {   // |grad lambda_f| = |J_f| / |J_e|
            double gl[3];
            for (int d = 0; d < 3; d++) {
                if (f == 0) { gl[d] = 0; for (int x = 0; x < dim; x++) gl[d] -= Ji[x][d]; }
                else gl[d] = Ji[f - 1][d];
            }
            g[9 + f] = std::sqrt(gl[0] * gl[0] + gl[1] * gl[1] + gl[2] * gl[2]);
        }
        std::array<double, 3> mt{1.0, 1.0, 0.0};
        auto mi = mats.find(m.elem_attr[e]); if (mi != mats.end()) mt = mi->second;
        g[13] = 1.0 / mt[0]; g[14] = 1.0 / mt[1]; g[15] = mt[2] / mt[0];
        H.tfsf_side[le] = side[e];
        for (int f = 0; f < nf; f++) {
            int *fi = &H.finfo[((size_t)le * 4 + f) * 2];
            int e2 = nbrE[(size_t)e * nf + f], code = 0, tab = f, nb = -1;
            if (e2 >= 0) {
                tab = tabOf(nbrRow(e, f, e2));
                if (part[e2] == o.rank) nb = g2l[e2];
                else nb = -2 - haloSlot[{le, f}];
            } else code |= (faceBC[(size_t)e * nf + f] & FI_BC_MASK) << FI_BC_SHIFT;
            if (tab > FI_TAB_MASK) throw Error(DGTD_ERR_UNSUPPORTED, "too many distinct face orientations");
            code |= tab << FI_TAB_SHIFT;
            int tf = faceTF[(size_t)e * nf + f];
            if (tf) {
                code |= (tf & FI_TFSF_MASK) << FI_TFSF_SHIFT;
                code |= H.n_tfsf_faces << FI_TIDX_SHIFT;
                if (H.n_tfsf_faces >= (1 << (31 - FI_TIDX_SHIFT))) throw Error(DGTD_ERR_UNSUPPORTED, "too many TF/SF faces");
                for (int j = 0; j < Nfp; j++) {
                    int n = R.fnodes[(size_t)f * Nfp + j];
                    for (int c = 0; c < 3; c++) {
                        double x0 = m.verts[3 * (size_t)m.elems[(size_t)e * nf] + c], s = x0;
                        for (int k = 0; k < dim; k++) s += J[c][k] * R.nodes[(size_t)n * dim + k];
                        H.tfsf_xyz.push_back(s);
                    }
                }
                H.n_tfsf_faces++;
            }
            fi[0] = nb; fi[1] = code;
        }
    }
    // gate: every node of EVERY TF/SF-adjacent element of the global mesh (SourcesManager.cpp:136-156); the norm test
    // is a function of time only, so each rank evaluates the global sum itself instead of all-reducing it
    for (int e = 0; e < NE; e++) {
        if (!side[e]) continue;
        double J[3][3], Ji[3][3], det;
        elem_geometry(m, e, J, Ji, det);
        for (int n = 0; n < Np; n++) for (int c = 0; c < 3; c++) {
            double x0 = m.verts[3 * (size_t)m.elems[(size_t)e * nf] + c], s = x0;
            for (int k = 0; k < dim; k++) s += J[c][k] * R.nodes[(size_t)n * dim + k];
            H.gate_xyz.push_back(s);
        }
    }
    H.ntab = (int)tabIdx.size();
    return H;
}

}  // namespace dgtd
