// Host-side data model of the B200 DG-Maxwell operator (C++17, no MFEM, no CUDA).
// Everything here replaces the SETUP the reference does in
//   HesthavenEvolution::HesthavenEvolution      src/evolution/HesthavenEvolution.cpp:315-437
//   Connectivities                               src/evolution/HesthavenEvolutionMethods.cpp:501-534, 721-791
// (one ParSubMesh + operator assembly PER ELEMENT there; flat arrays built in O(NE) here).
#pragma once
#include <array>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

namespace dgtd {

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string &m) : std::runtime_error(m), code(c) {}
};

struct Mesh {
    int dim = 0;
    std::vector<double> verts;                 // nv*3
    std::vector<int> elems, elem_attr;         // ne*(dim+1), ne
    std::vector<int> bdr, bdr_attr;            // nbe*dim, nbe
    int nv() const { return (int)(verts.size() / 3); }
    int ne() const { return (int)elem_attr.size(); }
    int nbe() const { return (int)bdr_attr.size(); }
    void validate_and_orient(bool reorient = true);   // simplex check; swap v0/v1 of inverted elements (mfem/mesh/mesh.cpp:6437-6493) or reject them
};
Mesh load_mesh(const std::string &path);
Mesh cartesian3d(int nx, int ny, int nz, double sx, double sy, double sz);
std::vector<int> partition_rcb(const Mesh &m, int nranks);
std::vector<int> partition_metis(const Mesh &m, int nranks);      // k-way on the element dual graph (the reference's partitioner)

// Reference element on the unit simplex with MFEM's L2 Gauss-Lobatto nodes.
struct RefElem {
    int dim = 0, p = 0, Np = 0, Nfp = 0, nf = 0;
    std::vector<double> nodes;     // Np*dim
    std::vector<int> bary;         // Np*(dim+1): integer barycentric index w.r.t. vertex k
    std::vector<double> D;         // dim*Np*Np : D[x][i][j] = d l_j / d xi_x (r_i)
    std::vector<double> Minv;      // Np*Np
    std::vector<int> fnodes;       // nf*Nfp local node ids on face f (opposite vertex f), ascending
    std::vector<double> lift;      // nf*Np*Nfp : Minv * (face mass, unit (dim-1)-simplex measure)
    std::vector<int> node_of;      // lookup: sum_k bary[k+1]*(p+1)^k -> node id (or -1)
    int lookup(const int *bary_full) const;
};
RefElem build_ref_element(int dim, int p);
std::vector<double> gll01(int p);

struct PlaneWave {
    bool enabled = false;
    double spread = 1, mean1d = 0, freq = 0;
    double pe[3] = {0, 0, 0}, ph[3] = {0, 0, 0}, dir[3] = {0, 0, 1};   // E / H polarisation vectors, unit propagation
};

struct Options {
    int order = 3;
    double alpha = 1.0;
    std::vector<std::pair<int, int>> bdr;                          // attribute -> DGTD_BC_*
    std::vector<int> tfsf;
    std::vector<std::pair<int, std::array<double, 3>>> mat;        // attribute -> eps, mu, sigma
    PlaneWave pw;
    bool tfsf_gate = true;
    int rank = 0, nranks = 1;
    std::vector<int> partitioning;
};

// face-info code bits (finfo[e][f].y)
constexpr int FI_BC_SHIFT = 0, FI_BC_MASK = 3;          // DGTD_BC_*
constexpr int FI_TFSF_SHIFT = 2, FI_TFSF_MASK = 3;      // 0 none, 1 this side is TF, 2 this side is SF
constexpr int FI_TAB_SHIFT = 4, FI_TAB_MASK = 0xff;     // row of ftab: neighbour-local node per face node
constexpr int FI_TIDX_SHIFT = 12;                       // TF/SF face slot (node coordinates)
constexpr int GEO_STRIDE = 16;                          // doubles per element: Jinv[9], fscale[4], 1/eps, 1/mu, sigma/eps

struct PeerPlan {
    int rank = -1;
    int nfaces = 0;
    int send_off = 0, recv_off = 0;      // in faces, into the packed send list / the halo slots
    int remote_off = 0;                  // first halo slot of MY block on the peer (its recv_off for me)
    int remote_idx = 0;                  // my position in the peer's peer list
};

// Everything a rank uploads to its GPU.
struct HostOp {
    int dim = 0, p = 0, Np = 0, Nfp = 0, nf = 0;
    long long NEglob = 0;
    int NEloc = 0;
    RefElem ref;
    std::vector<int> elem_gid;            // NEloc
    std::vector<double> geo;              // NEloc*GEO_STRIDE
    std::vector<double> jac;              // NEloc*10 : J[d][a] = dx_d/dxi_a (3d+a), det J
    std::vector<int> finfo;               // NEloc*4*2 : {nbr local element | -1 boundary | -2-haloFace, code}
    std::vector<uint8_t> ftab;            // ntab*Nfp
    int ntab = 0;
    // TF/SF
    std::vector<double> tfsf_xyz;         // nTfsfFaces*Nfp*3 face-node coordinates
    std::vector<double> gate_xyz;         // V*3 : all nodes of TF/SF-adjacent elements (norm test of `global`)
    int n_tfsf_faces = 0;
    std::vector<int> tfsf_side;           // NEloc: 0 none, 1 TF, 2 SF
    // halo
    std::vector<PeerPlan> peers;
    std::vector<int> send_node;           // nSendFaces*Nfp local dof ids, receiver's face-node order
    int n_halo_faces = 0;
    // node coordinates of the GLOBAL mesh are produced on demand (node_coords)
    double alpha = 1.0;
    PlaneWave pw;
    bool tfsf_gate = true;
};
HostOp build_host_op(const Mesh &m, const Options &o);
std::vector<int> boundary_element_faces(const Mesh &m, const std::vector<int> &attrs);

constexpr int BLK_E = 8;               // elements per group (= DMMA m dimension of the transposed contraction)
// ---- "wg" plan of the warp-per-group DMMA stage kernel (3-D only) --------------------------------------------------------
// Device state layout "aos": element-major node records, offset(e, n, c) = ((e * Np) + n) * 6 + c with n the DEVICE node
// id (dev2ref maps it to the reference node); 8 consecutive local elements form a group = one warp's unit of work, one
// contiguous chunk of Np*8*6 doubles.  The contraction runs "transposed" (DMMA A = data [8 elements x 4 nodes],
// B = operator [4 nodes x 8 output nodes]) so that a lane owns one element through all phases.
// geometry record of the wg / wh kernels, 30 doubles: J / det J (dx_d/dxi_a at 3d+a), Jinv[9] (dxi_a/dx_d at 9+3a+d),
// fscale[4] at 18, 1/det J at 22, det/eps 23, det/mu 24, sigma/eps 25, 1/fscale[4] at 26.  A stride of 30 doubles puts the
// records of the 8 elements of a group 28 banks apart in shared memory (distinct multiples of 4 banks: conflict-free
// LDS.128); a stride of 32 would put them all on the same banks, every geometry LDS.128 costing 8 wavefronts instead of 1
#ifndef DGTD_WG_GEO
#define DGTD_WG_GEO 30
#endif
constexpr int WG_GEO = DGTD_WG_GEO;
// rows of the node tables kept in shared memory: 8 own / canonical + (my face, neighbour's face, rotation) <= 4 x 4 x 6
// + push rows (my face in the receiver's order) <= 4 x 6: 128 covers every conforming tetrahedral mesh and partition
// (72 was too few for the METIS parts of BASELINE config 4's sphere mesh: 76-80 rows, which then ran the generic kernel)
#ifndef DGTD_WG_TABROWS
#define DGTD_WG_TABROWS 128
#endif
constexpr int WG_TABROWS = DGTD_WG_TABROWS;
struct WgPlan {
    int ngroups = 0, NEpad = 0;
    int NT = 0, KSV = 0;               // output n-tiles (the last one is "mixed"), k-steps of the volume contraction
    int nfrag_vol = 0, nfrag_lift = 0; // B fragments of 32 doubles: volume [KSV][(NT-1)*3 + 3], LIFT [Nfp][NT]
    std::vector<int> dev2ref, ref2dev; // Np
    std::vector<int> forder;           // 4*Nfp : step s of face f handles canonical face node forder[f*Nfp+s]
    std::vector<double> geo;           // NEpad * WG_GEO
    std::vector<int> desc;             // NEpad*4*2 : {nbr local element | -1 boundary | -2-haloFace, code (FI_TAB = row of tab)}
    std::vector<uint8_t> tab;          // ntab*16 : rows 0..3 own device node per step, 4..7 canonical index per step, 8.. neighbour device node per step
    int ntab = 0;
    std::vector<double> bfrag;         // (nfrag_vol + nfrag_lift) * 32
    std::vector<long long> send_off;   // nSendFaces*Nfp : offset (doubles) of the node record in the aos state
    std::vector<int> hpush;            // nHaloFaces*2 : {peer index | tab row << 8 (own device node per RECEIVER face node), slot on the peer}
};
WgPlan build_wg_plan(const HostOp &H);
void node_coords(const Mesh &m, const RefElem &ref, std::vector<double> &xyz);   // [NE*Np][3], global numbering

}  // namespace dgtd
