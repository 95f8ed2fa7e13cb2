// Simplex meshes on the host: readers for the two formats the reference's cases use
// (Gmsh 2.2 ASCII as read by mfem::Mesh::LoadFromFile, src/driver/driver.cpp:1176-1183, and "MFEM mesh v1.0"),
// a Cartesian tetrahedral box generator (config 5: Mesh::MakeCartesian3D(n,n,n,TETRAHEDRON)) and the two partitioners
// behind the reference's int[NE] contract (Mesh::GeneratePartitioning, driver.cpp:1269): recursive coordinate bisection
// and METIS k-way on the element dual graph (the static METIS shipped inside the CUDA toolkit).
#include "host.hpp"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <fstream>
#include <map>
#include <numeric>
#include <sstream>

namespace dgtd {

static double det_of(const Mesh &m, int e)
{
    const int d = m.dim;
    const int *v = &m.elems[(size_t)e * (d + 1)];
    const double *x0 = &m.verts[3 * (size_t)v[0]];
    double J[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    for (int k = 0; k < d; k++) for (int c = 0; c < d; c++) J[c][k] = m.verts[3 * (size_t)v[k + 1] + c] - x0[c];
    if (d == 1) return J[0][0];
    if (d == 2) return J[0][0] * J[1][1] - J[0][1] * J[1][0];
    return J[0][0] * (J[1][1] * J[2][2] - J[1][2] * J[2][1]) - J[0][1] * (J[1][0] * J[2][2] - J[1][2] * J[2][0]) +
           J[0][2] * (J[1][0] * J[2][1] - J[1][1] * J[2][0]);
}

void Mesh::validate_and_orient(bool reorient)
{
    if (dim < 1 || dim > 3) throw Error(-3, "mesh dimension must be 1, 2 or 3");
    if (verts.size() % 3) throw Error(-3, "vertex array must hold 3 doubles per vertex");
    if (elems.size() != (size_t)ne() * (dim + 1) || bdr.size() != (size_t)nbe() * dim) throw Error(-3, "element/boundary array sizes");
    const int n = nv();
    for (int v : elems) if (v < 0 || v >= n) throw Error(-3, "element vertex id out of range");
    for (int v : bdr) if (v < 0 || v >= n) throw Error(-3, "boundary vertex id out of range");
    for (int e = 0; e < ne(); e++) {
        double d = det_of(*this, e);
        if (d == 0.0 || !std::isfinite(d)) throw Error(-3, "degenerate element " + std::to_string(e));
        if (d < 0.0) {
            if (dim == 1) throw Error(-3, "inverted segment " + std::to_string(e));
            // caller-owned numbering (dgtd_mesh_from_arrays): a silent swap would permute this element's dofs relative to the
            // caller's FE space, so it is an error there; the file loaders and generators own their numbering
            if (!reorient) throw Error(-3, "inverted element " + std::to_string(e) + " (negative Jacobian): orient the mesh first, as mfem::Mesh does on load");
            std::swap(elems[(size_t)e * (dim + 1)], elems[(size_t)e * (dim + 1) + 1]);
        }
    }
}

static Mesh load_gmsh22(std::istream &in)
{
    std::string tok;
    std::map<long long, int> node_id;          // file node number -> index in xyz
    std::vector<double> xyz;
    struct El { int type, phys; std::vector<long long> nodes; };
    std::vector<El> els;
    while (in >> tok) {
        if (tok == "$MeshFormat") {
            double ver; int ft, ds; in >> ver >> ft >> ds;
            if (ver < 2.0 || ver >= 3.0 || ft != 0) throw Error(-3, "only Gmsh 2.x ASCII meshes are supported");
        } else if (tok == "$Nodes") {
            long long n; in >> n;
            for (long long i = 0; i < n; i++) {
                long long id; double x, y, z; in >> id >> x >> y >> z;
                node_id[id] = (int)(xyz.size() / 3);
                xyz.push_back(x); xyz.push_back(y); xyz.push_back(z);
            }
        } else if (tok == "$Elements") {
            long long n; in >> n;
            static const int nn[16] = {0, 2, 3, 4, 4, 8, 6, 5, 3, 6, 9, 10, 27, 18, 14, 1};
            for (long long i = 0; i < n; i++) {
                long long id; int type, ntags; in >> id >> type >> ntags;
                if (type < 1 || type > 15) throw Error(-3, "unsupported Gmsh element type " + std::to_string(type));
                El e; e.type = type; e.phys = 1;
                for (int t = 0; t < ntags; t++) { int tag; in >> tag; if (t == 0) e.phys = tag; }
                e.nodes.resize(nn[type]);
                for (auto &v : e.nodes) in >> v;
                els.push_back(std::move(e));
            }
        }
    }
    if (!in.eof() && in.fail()) throw Error(-3, "malformed Gmsh file");
    int dim = 0;
    for (auto &e : els) dim = std::max(dim, e.type == 4 ? 3 : e.type == 2 ? 2 : e.type == 1 ? 1 : 0);
    if (dim == 0) throw Error(-3, "no simplex elements in Gmsh file");
    const int etype[4] = {15, 1, 2, 4};
    Mesh m; m.dim = dim;
    // vertices: only those used by top-dimensional elements, in order of first use (MFEM renumbers the same way)
    std::map<long long, int> used;
    auto vid = [&](long long fileid) {
        auto it = used.find(fileid);
        if (it != used.end()) return it->second;
        auto nt = node_id.find(fileid);
        if (nt == node_id.end()) throw Error(-3, "element references unknown node");
        int k = (int)used.size(); used[fileid] = k;
        for (int c = 0; c < 3; c++) m.verts.push_back(xyz[3 * (size_t)nt->second + c]);
        return k;
    };
    for (auto &e : els) if (e.type == etype[dim]) { for (auto v : e.nodes) m.elems.push_back(vid(v)); m.elem_attr.push_back(e.phys); }
    for (auto &e : els) if (e.type == etype[dim - 1]) {
        bool ok = true; for (auto v : e.nodes) ok &= used.count(v) > 0;
        if (!ok) continue;
        for (auto v : e.nodes) m.bdr.push_back(used[v]);
        m.bdr_attr.push_back(e.phys);
    }
    return m;
}

static Mesh load_mfem_v10(std::istream &in)
{
    Mesh m; std::string line, tok;
    auto next = [&](std::string &t) {
        while (in >> t) { if (t[0] == '#') { std::getline(in, line); continue; } return true; }
        return false;
    };
    while (next(tok)) {
        if (tok == "dimension") { in >> m.dim; }
        else if (tok == "elements") {
            int n; in >> n;
            for (int i = 0; i < n; i++) {
                int attr, geom; in >> attr >> geom;
                int nvx = geom == 1 ? 2 : geom == 2 ? 3 : geom == 4 ? 4 : -1;
                if (nvx != m.dim + 1) throw Error(-3, "only simplex elements are supported");
                for (int k = 0; k < nvx; k++) { int v; in >> v; m.elems.push_back(v); }
                m.elem_attr.push_back(attr);
            }
        } else if (tok == "boundary") {
            int n; in >> n;
            for (int i = 0; i < n; i++) {
                int attr, geom; in >> attr >> geom;
                int nvx = geom == 0 ? 1 : geom == 1 ? 2 : geom == 2 ? 3 : -1;
                if (nvx != m.dim) throw Error(-3, "boundary element geometry does not match the mesh dimension");
                for (int k = 0; k < nvx; k++) { int v; in >> v; m.bdr.push_back(v); }
                m.bdr_attr.push_back(attr);
            }
        } else if (tok == "vertices") {
            int n, sd; in >> n; std::string s; in >> s;
            if (s == "nodes") throw Error(-4, "curved (nodal) MFEM meshes are not supported");
            sd = std::stoi(s);
            for (int i = 0; i < n; i++) { double c[3] = {0, 0, 0}; for (int k = 0; k < sd; k++) in >> c[k]; for (int k = 0; k < 3; k++) m.verts.push_back(c[k]); }
        }
    }
    return m;
}

Mesh load_mesh(const std::string &path)
{
    std::ifstream in(path);
    if (!in) throw Error(-3, "cannot open mesh file " + path);
    std::string first; std::getline(in, first);
    in.seekg(0);
    Mesh m;
    if (first.rfind("$MeshFormat", 0) == 0) m = load_gmsh22(in);
    else if (first.rfind("MFEM mesh v1.0", 0) == 0) { std::getline(in, first); m = load_mfem_v10(in); }
    else throw Error(-3, "unrecognised mesh format in " + path);
    m.validate_and_orient();
    return m;
}

Mesh cartesian3d(int nx, int ny, int nz, double sx, double sy, double sz)
{
    if (nx < 1 || ny < 1 || nz < 1) throw Error(-1, "cartesian3d: cell counts must be positive");
    if ((long long)nx * ny * nz * 6 > 2000000000LL / 4) throw Error(-1, "cartesian3d: too many elements for 32-bit ids");
    Mesh m; m.dim = 3;
    auto vid = [&](int i, int j, int k) { return (k * (ny + 1) + j) * (nx + 1) + i; };
    for (int k = 0; k <= nz; k++) for (int j = 0; j <= ny; j++) for (int i = 0; i <= nx; i++) {
        m.verts.push_back(sx * i / nx); m.verts.push_back(sy * j / ny); m.verts.push_back(sz * k / nz);
    }
    // Kuhn subdivision: 6 tets around the diagonal (0,0,0)-(1,1,1); conforming across cells
    static const int perm[6][3] = {{0, 1, 2}, {0, 2, 1}, {1, 0, 2}, {1, 2, 0}, {2, 0, 1}, {2, 1, 0}};
    for (int k = 0; k < nz; k++) for (int j = 0; j < ny; j++) for (int i = 0; i < nx; i++)
        for (int t = 0; t < 6; t++) {
            int c[3] = {0, 0, 0}, v[4];
            v[0] = vid(i, j, k);
            for (int s = 0; s < 3; s++) { c[perm[t][s]] = 1; v[s + 1] = vid(i + c[0], j + c[1], k + c[2]); }
            for (int s = 0; s < 4; s++) m.elems.push_back(v[s]);
            m.elem_attr.push_back(1);
        }
    // boundary triangles: each boundary square is split along the same diagonal as the Kuhn tets
    auto quad = [&](int a, int b, int c, int d, int attr) {   // a-b-c-d around, diagonal a-c
        m.bdr.insert(m.bdr.end(), {a, b, c}); m.bdr_attr.push_back(attr);
        m.bdr.insert(m.bdr.end(), {a, c, d}); m.bdr_attr.push_back(attr);
    };
    for (int j = 0; j < ny; j++) for (int i = 0; i < nx; i++) {
        quad(vid(i, j, 0), vid(i + 1, j, 0), vid(i + 1, j + 1, 0), vid(i, j + 1, 0), 1);
        quad(vid(i, j, nz), vid(i + 1, j, nz), vid(i + 1, j + 1, nz), vid(i, j + 1, nz), 6);
    }
    for (int k = 0; k < nz; k++) for (int i = 0; i < nx; i++) {
        quad(vid(i, 0, k), vid(i + 1, 0, k), vid(i + 1, 0, k + 1), vid(i, 0, k + 1), 2);
        quad(vid(i, ny, k), vid(i + 1, ny, k), vid(i + 1, ny, k + 1), vid(i, ny, k + 1), 4);
    }
    for (int k = 0; k < nz; k++) for (int j = 0; j < ny; j++) {
        quad(vid(nx, j, k), vid(nx, j + 1, k), vid(nx, j + 1, k + 1), vid(nx, j, k + 1), 3);
        quad(vid(0, j, k), vid(0, j + 1, k), vid(0, j + 1, k + 1), vid(0, j, k + 1), 5);
    }
    m.validate_and_orient();
    return m;
}

static void rcb(const std::vector<double> &bc, std::vector<int> &ids, int lo, int hi, int r0, int nr, std::vector<int> &part)
{
    if (nr == 1) { for (int i = lo; i < hi; i++) part[ids[i]] = r0; return; }
    double mn[3] = {1e300, 1e300, 1e300}, mx[3] = {-1e300, -1e300, -1e300};
    for (int i = lo; i < hi; i++) for (int c = 0; c < 3; c++) { double v = bc[3 * (size_t)ids[i] + c]; mn[c] = std::min(mn[c], v); mx[c] = std::max(mx[c], v); }
    int ax = 0; for (int c = 1; c < 3; c++) if (mx[c] - mn[c] > mx[ax] - mn[ax]) ax = c;
    int nl = nr / 2;
    int mid = lo + (int)((long long)(hi - lo) * nl / nr);
    std::nth_element(ids.begin() + lo, ids.begin() + mid, ids.begin() + hi, [&](int a, int b) {
        double va = bc[3 * (size_t)a + ax], vb = bc[3 * (size_t)b + ax];
        return va < vb || (va == vb && a < b);
    });
    rcb(bc, ids, lo, mid, r0, nl, part);
    rcb(bc, ids, mid, hi, r0 + nl, nr - nl, part);
}

std::vector<int> partition_rcb(const Mesh &m, int nranks)
{
    if (nranks < 1) throw Error(-1, "nranks must be positive");
    const int ne = m.ne(), d = m.dim;
    std::vector<double> bc(3 * (size_t)ne, 0.0);
    for (int e = 0; e < ne; e++) for (int k = 0; k <= d; k++) for (int c = 0; c < 3; c++)
        bc[3 * (size_t)e + c] += m.verts[3 * (size_t)m.elems[(size_t)e * (d + 1) + k] + c] / (d + 1);
    std::vector<int> ids(ne), part(ne, 0);
    std::iota(ids.begin(), ids.end(), 0);
    rcb(bc, ids, 0, ne, 0, nranks, part);
    return part;
}

// ---- METIS k-way partition of the element dual graph — what the reference does through
// Mesh::GeneratePartitioning (src/driver/driver.cpp:1269; external/mfem-geg/mesh/mesh.cpp:8378 -> METIS_PartGraphKway).
// The METIS here is the static library that ships in the CUDA toolkit (for cuSOLVER; 64-bit idx_t, no header installed),
// so its two entry points are declared by hand; real_t arguments are passed as NULL (defaults).
extern "C" {
int METIS_SetDefaultOptions(long long *options);
int METIS_PartGraphKway(long long *nvtxs, long long *ncon, long long *xadj, long long *adjncy, long long *vwgt, long long *vsize,
                        long long *adjwgt, long long *nparts, void *tpwgts, void *ubvec, long long *options, long long *edgecut, long long *part);
}

std::vector<int> partition_metis(const Mesh &m, int nranks)
{
    if (nranks < 1) throw Error(-1, "nranks must be positive");
    const int ne = m.ne(), nf = m.dim + 1;
    std::vector<int> part(ne, 0);
    if (nranks == 1) return part;
    // dual graph: elements sharing a face
    struct Key { int v[3]; int e; };
    std::vector<Key> keys((size_t)ne * nf);
    for (int e = 0; e < ne; e++) for (int f = 0; f < nf; f++) {
        Key &k = keys[(size_t)e * nf + f];
        k.v[0] = k.v[1] = k.v[2] = -1; k.e = e; int c = 0;
        for (int q = 0; q < nf; q++) if (q != f) k.v[c++] = m.elems[(size_t)e * nf + q];
        std::sort(k.v, k.v + m.dim);
    }
    std::sort(keys.begin(), keys.end(), [](const Key &a, const Key &b) { return std::lexicographical_compare(a.v, a.v + 3, b.v, b.v + 3); });
    std::vector<std::vector<long long>> adj(ne);
    for (size_t i = 0; i + 1 < keys.size(); i++) {
        const Key &a = keys[i], &b = keys[i + 1];
        if (a.v[0] == b.v[0] && a.v[1] == b.v[1] && a.v[2] == b.v[2]) { adj[a.e].push_back(b.e); adj[b.e].push_back(a.e); i++; }
    }
    std::vector<long long> xadj(ne + 1, 0), adjncy;
    for (int e = 0; e < ne; e++) { adjncy.insert(adjncy.end(), adj[e].begin(), adj[e].end()); xadj[e + 1] = (long long)adjncy.size(); }
    long long nv = ne, ncon = 1, np = nranks, cut = 0, options[40];
    METIS_SetDefaultOptions(options);
    // load imbalance is lost time on ranks that run in lock step: ask for 0.2 % (METIS' default tolerance is 3 %, which
    // the reference inherits, mesh.cpp:8378); DGTD_B200_METIS_UFACTOR overrides (30 = the reference's behaviour)
    {
        const char *uf = std::getenv("DGTD_B200_METIS_UFACTOR");
        options[16 /* METIS_OPTION_UFACTOR */] = uf ? std::atoll(uf) : 2;
        options[11 /* METIS_OPTION_CONTIG */] = 1;
    }
    std::vector<long long> p64(ne, 0);
    const int rc = METIS_PartGraphKway(&nv, &ncon, xadj.data(), adjncy.data(), nullptr, nullptr, nullptr, &np, nullptr, nullptr, options, &cut, p64.data());
    if (rc != 1) throw Error(-1, "METIS_PartGraphKway failed (" + std::to_string(rc) + ")");
    std::vector<int> count(nranks, 0);
    for (int e = 0; e < ne; e++) { part[e] = (int)p64[e]; count[part[e]]++; }
    for (int r = 0; r < nranks; r++) if (count[r] == 0) return partition_rcb(m, nranks);   // tiny graphs: METIS may leave a part empty
    return part;
}

}  // namespace dgtd
