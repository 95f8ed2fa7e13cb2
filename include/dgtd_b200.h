/* dgtd_b200 — C ABI of the B200-native DG-Maxwell evolution hot path.
 *
 * Drop-in boundary for OpenSEMBA/dgtd's evolution operators.  The reference has
 * no FFI: its seam is the C++ virtual interface
 *     mfem::TimeDependentOperator::Mult(const Vector&, Vector&) const   (external/mfem-geg/linalg/operator.hpp:89, 391-394)
 * implemented by
 *     maxwell::GlobalEvolution      src/evolution/GlobalEvolution.h:19-22, GlobalEvolution.cpp:628-1088
 *     maxwell::HesthavenEvolution   src/evolution/HesthavenEvolution.h:20-21, HesthavenEvolution.cpp:450-542
 * and driven by mfem::RK4Solver::Step (external/mfem-geg/linalg/ode.cpp:109-136) from
 * maxwell::Solver::step (src/solver/Solver.cpp:535-551).  The two thin C++ shells in
 * dgtd_b200/mfem_shell/ (B200Evolution, B200RK4Solver) are the only callers a
 * maintainer adds on the reference side; they bind exactly the entry points below
 * (see INTEGRATION.md).
 *
 * Conventions: every function returns DGTD_OK (0) or a negative error code and never
 * throws; dgtd_last_error() gives the message of the last failure on the calling
 * thread (the shells turn it into std::runtime_error like the reference does,
 * src/solver/Solver.cpp:37,74).  All pointers are caller-owned unless returned by a
 * *_create / *_load function.  One host thread per context; calls on a context are
 * serialised (the reference's Mult is not re-entrant either, GlobalEvolution.h:82-90).
 * State layout is the reference's (src/evolution/Fields.h:45-65): 6 blocks
 * [Ex,Ey,Ez,Hx,Hy,Hz] of N doubles, dof = element*Np + local node, MFEM L2
 * Gauss-Lobatto simplex node order (external/mfem-geg/fem/fe/fe_l2.cpp:716-722).
 * There is NO CPU fallback: without a CUDA device every compute entry point fails.
 *
 * Environment variables read by the library (A/B measurements and diagnostics; none is needed in normal use):
 *   DGTD_B200_KERNEL=wg|wh|generic   force a stage-kernel family (default: wh at order 4, wg below, generic otherwise)
 *   DGTD_B200_HALO=nccl              force the pack + ncclSend/ncclRecv halo path instead of peer-memory stores
 *   DGTD_B200_ORDER=morton|grow      local element order (default: Morton on one rank, face-sharing groups on several)
 *   DGTD_B200_DYNAMIC=1              counter-based group scheduling on a single rank too
 *   DGTD_B200_METIS_UFACTOR=n        METIS load-imbalance tolerance in 1/1000 (default 2; the reference's is 30)
 *   DGTD_B200_FACE_ORDER=bank        alternative order of the face steps inside the kernel plan
 */
#ifndef DGTD_B200_H
#define DGTD_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define DGTD_OK               0
#define DGTD_ERR_ARG         -1   /* invalid argument / null pointer / size mismatch            */
#define DGTD_ERR_CUDA        -2   /* CUDA runtime failure or no usable device                   */
#define DGTD_ERR_MESH        -3   /* unreadable / inconsistent / non-simplex / inverted mesh    */
#define DGTD_ERR_UNSUPPORTED -4   /* order/dimension/feature outside what the kernels cover     */
#define DGTD_ERR_COMM        -5   /* NCCL / halo exchange failure                               */

/* boundary conditions — BdrCond in src/components/Types.h:48-56 */
#define DGTD_BC_NONE 0
#define DGTD_BC_PEC  1
#define DGTD_BC_PMC  2
#define DGTD_BC_SMA  3

typedef struct dgtd_mesh dgtd_mesh;   /* host-side simplex mesh (segments / triangles / tetrahedra) */
typedef struct dgtd_ctx  dgtd_ctx;    /* one per GPU (rank): owns all device memory, streams, comms  */

/* Gaussian / modulated-Gaussian plane wave — maxwell::Planewave, src/math/Function.h:328-409,
 * built by driver.cpp:483-519.  g(u) = exp(-(u-mean1d)^2/(2 spread^2)) [* cos(2 pi freq (u-mean1d))],
 * u = dir.x - t (c = 1).  pol/dir are normalised by the library as the reference ctor does. */
typedef struct dgtd_planewave {
    int    enabled;
    double spread, mean1d, freq;   /* freq == 0: plain Gaussian */
    double pol[3], dir[3];
    int    fieldtype;              /* 0: pol is the E polarisation, 1: pol is H */
} dgtd_planewave;

/* What maxwell::Model + EvolutionOptions + Sources give the reference ctor
 * (src/evolution/GlobalEvolution.cpp:32-364; EvolutionOptions.h:13-20). */
typedef struct dgtd_options {
    int    order;                       /* 1..6 (tets: 1..5)                                       */
    double alpha;                       /* upwind_alpha in [0,1]                                   */
    int    n_bdr;                       /* boundary attribute -> condition                         */
    const int *bdr_attr, *bdr_cond;
    int    n_tfsf;                      /* attributes of interior TF/SF faces (3-D)                */
    const int *tfsf_attr;
    int    n_mat;                       /* element attribute -> (eps, mu, sigma); default vacuum   */
    const int *mat_attr;
    const double *mat_eps_mu_sigma;     /* 3 doubles per entry                                     */
    dgtd_planewave pw;
    int    tfsf_gate;                   /* 1: skip injection when ||s|| < 1e-8 like `global`
                                           (GlobalEvolution.cpp:584-598); 0: never skip (`hesthaven`) */
    int    device;                      /* CUDA device ordinal                                     */
    int    rank, nranks;                /* partition of the mesh this context owns                 */
    const int *partitioning;            /* [NE] element -> rank, or NULL = built-in partitioner
                                           (same contract as Mesh::GeneratePartitioning, driver.cpp:1269) */
} dgtd_options;

/* ---- mesh (host) ---------------------------------------------------------------------------- */
/* verts: nv*3 doubles (unused coordinates 0); elems: ne*(dim+1) vertex ids in the host code's
 * element-local order (mfem::Mesh::GetElementVertices); bdr: nbe*dim vertex ids.  Elements must be positively oriented
 * (mfem::Mesh orients on load): an inverted element fails with DGTD_ERR_MESH instead of being silently renumbered. */
int  dgtd_mesh_from_arrays(int dim, int nv, const double *verts, int ne, const int *elems,
                           const int *elem_attr, int nbe, const int *bdr, const int *bdr_attr,
                           dgtd_mesh **out);
/* Gmsh 2.2 ASCII (.msh, attribute = physical tag as MFEM reads it) or "MFEM mesh v1.0" (.mesh).  Elements keep the file's
 * order; for "MFEM mesh v1.0" the arrays equal what mfem::Mesh holds after loading, for Gmsh files MFEM renumbers the
 * vertices, so state vectors are interchangeable with the reference's only when the mesh comes from the reference's
 * mfem::Mesh through dgtd_mesh_from_arrays (what B200Evolution does).                                               */
int  dgtd_mesh_load(const char *path, dgtd_mesh **out);
/* nx*ny*nz cubes of 6 tetrahedra on [0,sx]x[0,sy]x[0,sz]; boundary attributes 1..6 =
 * bottom(z=0), front(y=0), right(x=sx), back(y=sy), left(x=0), top(z=sz) (MFEM's MakeCartesian3D). */
int  dgtd_mesh_cartesian3d(int nx, int ny, int nz, double sx, double sy, double sz, dgtd_mesh **out);
int  dgtd_mesh_info(const dgtd_mesh *, int *dim, int *nv, int *ne, int *nbe);
/* copies out what dgtd_mesh_from_arrays takes in (any pointer may be NULL)                        */
int  dgtd_mesh_get_arrays(const dgtd_mesh *, double *verts, int *elems, int *elem_attr, int *bdr, int *bdr_attr);
/* element -> rank by recursive coordinate bisection of element barycentres (slabs / bricks on Cartesian boxes; the default
 * when dgtd_options.partitioning is NULL)                                                                         */
int  dgtd_mesh_partition(const dgtd_mesh *, int nranks, int *partitioning);
/* element -> rank by METIS k-way on the element dual graph, the reference's partitioner (Mesh::GeneratePartitioning,
 * src/driver/driver.cpp:1269); hand the result to dgtd_options.partitioning                                          */
int  dgtd_mesh_partition_metis(const dgtd_mesh *, int nranks, int *partitioning);
void dgtd_mesh_destroy(dgtd_mesh *);

/* ---- context = the evolution operator ----------------------------------------------------------- */
/* Replaces GlobalEvolution::GlobalEvolution / HesthavenEvolution::HesthavenEvolution.             */
int  dgtd_create(const dgtd_mesh *, const dgtd_options *, dgtd_ctx **out);
void dgtd_destroy(dgtd_ctx *);
/* sizes: global scalar dofs N (state = 6N), dofs per element, owned elements/dofs of this rank   */
int  dgtd_sizes(const dgtd_ctx *, long long *n_global, int *np, long long *ne_local, long long *n_local);
/* global element id of every local element, in local order [ne_local]                            */
int  dgtd_local_elements(const dgtd_ctx *, int *elem_ids);
/* physical node coordinates, global numbering, [N][3] (what GridFunction::ProjectCoefficient
 * evaluates initial fields at, SourcesManager.cpp:28-32)                                         */
int  dgtd_node_coords(const dgtd_ctx *, double *xyz);
/* use a caller-provided cudaStream_t for all work (NULL: the context's own stream)                */
int  dgtd_set_stream(dgtd_ctx *, void *cuda_stream);

/* state: host vectors are GLOBAL [6N] in the reference layout; a rank reads/writes only the
 * entries of the elements it owns (other entries of `host_6N` are left untouched on get).         */
int  dgtd_set_state(dgtd_ctx *, const double *host_6N);
int  dgtd_get_state(dgtd_ctx *, double *host_6N);
/* the same with LOCAL host vectors [6][n_local] in this rank's element order (dgtd_local_elements)  */
int  dgtd_set_state_local(dgtd_ctx *, const double *host_6nlocal);
int  dgtd_get_state_local(dgtd_ctx *, double *host_6nlocal);
/* the same with LOCAL host vectors [6][n_local] in the element order of the rank's mfem::ParMesh — the owned elements by
 * ascending global id, which is what a ParFiniteElementSpace-sized mfem::Vector holds in the reference's MPI build
 * (Model.cpp:59 builds the ParMesh from the serial mesh and the partitioning).  dgtd_mult_parlocal is
 * TimeDependentOperator::Mult on such vectors (src/evolution/GlobalEvolution.cpp:628 works on fes_.GetNDofs() local dofs). */
int  dgtd_set_state_parlocal(dgtd_ctx *, const double *host_6nlocal);
int  dgtd_get_state_parlocal(dgtd_ctx *, double *host_6nlocal);
int  dgtd_mult_parlocal(dgtd_ctx *, double t, const double *in_6nlocal, double *out_6nlocal);
/* device-resident state of this rank in the kernels' NATIVE layout: tetrahedra of order <= 4 use the "aos" layout
 * offset(e, n, c) = (e * Np + n) * 6 + c over local elements in Morton order, n = device node id ("wg_dev2ref" of
 * dgtd_setup_query maps it to the reference node), padded to whole groups of 8 elements; everything else [6][n_local].
 * Use dgtd_get_state_local / dgtd_mult(on_device=1) for layout-independent access.                                */
int  dgtd_state_device_ptr(dgtd_ctx *, double **dev);

/* TimeDependentOperator::Mult at time t (SetTime + Mult).  Host pointers: global [6N] vectors,
 * copied in/out (each rank fills its owned entries of out).  Device pointers (on_device=1):
 * local [6][n_local] vectors, no copies.                                                          */
int  dgtd_mult(dgtd_ctx *, double t, const double *in, double *out, int on_device);
/* mfem::RK4Solver::Step on the resident state, fused: 4 launches, k never stored
 * (ode.cpp:109-136 semantics incl. the SetTime sequence t, t+dt/2, t+dt/2, t+dt).                 */
int  dgtd_rk4_step(dgtd_ctx *, double t, double dt);
/* nsteps of the above from t0 (Solver::run loop body, Solver.cpp:483-533, without probes)         */
int  dgtd_rk4_run(dgtd_ctx *, double t0, double dt, int nsteps);
/* Solver::run (Solver.cpp:497-533) with Solver::step's final short step (Solver.cpp:535-537): advances *t from its value
 * to t_final with steps of min(dt, t_final - t).  check_every > 0 evaluates the reference's per-step stability test
 * (!isfinite(norm) || norm > 1e20, Solver.cpp:500-516) on this rank's dofs every that many steps and stops with
 * *unstable = 1 when it fires; on multi-rank contexts the flag is max-reduced over the ranks (the reference's
 * MPI_Allreduce, Solver.cpp:503) so that all of them stop at the same step.                                        */
int  dgtd_run_until(dgtd_ctx *, double *t, double dt, double t_final, int check_every, long long *nsteps, int *unstable);
/* ||state||_2 over all ranks' owned dofs of THIS rank (caller reduces) — Fields::getNorml2        */
int  dgtd_norm2_local(dgtd_ctx *, double *sumsq);
/* point probes: field values at npts (local element, Np shape weights) -> out[npts][6]            */
int  dgtd_sample(dgtd_ctx *, int npts, const int *local_elem, const double *shape, double *out6);
/* ---- probes, field / RCS-surface export: periodic asynchronous device-to-host copies ------------------------------
 * A gather is a fixed list of scalar dofs in the reference numbering (element * Np + node, Fields.h:45-65).  Every launch
 * snapshots the six field values of the locally owned ones, in list order, into host_out[6][n_local] WITHOUT stopping
 * the time loop: gather kernel on the compute stream, device-to-host copy on a side stream (pin host_out for a truly
 * asynchronous copy), dgtd_gather_wait before reading.  Steps issued after the launch do not alter the snapshot.
 * Replaces RCSSurfaceExporter::transferFields (RCSSurfaceExporter.cpp:71-79: six TransferMaps + host write per export
 * step), the FieldProbe / PointProbe reads of ProbesManager (Solver.cpp:550) and the per-step full-state D2H they imply. */
typedef struct dgtd_gather dgtd_gather;
int  dgtd_gather_create(dgtd_ctx *, long long n, const long long *dofs, dgtd_gather **out, long long *n_local);
/* the owned dofs of the gather (global numbering), in output order: dofs_local[n_local]           */
int  dgtd_gather_dofs(const dgtd_gather *, long long *dofs_local);
int  dgtd_gather_launch(dgtd_ctx *, dgtd_gather *, double *host_out);
int  dgtd_gather_wait(dgtd_ctx *, dgtd_gather *);
/* either order of dgtd_gather_destroy and dgtd_destroy is fine: a gather whose context is gone only frees its host shell */
void dgtd_gather_destroy(dgtd_gather *);
/* Host-only helper for surface exports: the (global element, local face) pairs lying on boundary elements with one of
 * the given attributes — NearToFarFieldSubMesher's selection (SubMesher.cpp:832-905) keeps one element per face; here
 * both sides are listed, lower element id first (MFEM's Elem1), and the caller picks.  pairs[2*k] = element, [2*k+1] =
 * face; returns the number of pairs through n_pairs (call with pairs = NULL to size the buffer).                    */
int  dgtd_mesh_boundary_elements(const dgtd_mesh *, int n_attr, const int *bdr_attr, long long cap_pairs, int *pairs, long long *n_pairs);
int  dgtd_synchronize(dgtd_ctx *);
/* number of kernel launches issued by this context so far (bench.py's gpu_launches)               */
long long dgtd_launch_count(const dgtd_ctx *);
/* which stage kernel this context runs (name, tiling, launch shape) — for benchmark reports                */
int  dgtd_kernel_info(const dgtd_ctx *, char *buf, int cap);

/* ---- multi-GPU halo exchange (one context per rank/GPU over NVLink) ------------------------------
 * Replaces the six blocking MPI neighbour exchanges of GlobalEvolution::Mult (GlobalEvolution.cpp:763-774).
 * dgtd_comm_init is collective: NCCL bootstrap, then (tetrahedra, order <= 4) every rank maps its neighbours' halo
 * buffers with CUDA IPC; from then on the stage kernel itself stores the traces of its partition faces into the
 * neighbour's buffer and only the warps owning such a face wait for the neighbour's epoch flag (no pack kernel, no
 * collective on the path).  If IPC is unavailable on any rank, all ranks use ncclSend/ncclRecv instead.
 * dgtd_destroy of a multi-rank context is collective too (neighbours must stop storing before a buffer is freed).
 * Like the reference's MPI calls, every entry point that evaluates the operator or replaces the state (dgtd_mult,
 * dgtd_rk4_step/run/run_until, dgtd_set_state*) must be called by ALL ranks in the same order: each one is one numbered
 * exchange of the neighbours' traces.                                                                                */
int  dgtd_comm_unique_id(void *id128);                    /* rank 0: ncclGetUniqueId              */
int  dgtd_comm_init(dgtd_ctx *, const void *id128);       /* all ranks                            */
#define DGTD_HALO_NONE 0   /* single rank                                         */
#define DGTD_HALO_NCCL 1   /* pack kernel + ncclSend/ncclRecv per RHS evaluation  */
#define DGTD_HALO_P2P  2   /* peer-memory stores fused into the stage kernel      */
int  dgtd_halo_mode(const dgtd_ctx *);
/* bytes this rank sends per RHS evaluation (6 * Nfp * shared faces * 8)                          */
int  dgtd_halo_bytes(const dgtd_ctx *, long long *bytes);

/* Host-only diagnostic (no CUDA, no compute): copies one of the flat tables a rank would upload — "dims", "D", "lift",
 * "nodes", "fnodes", "geo", "finfo", "ftab", "elem_gid", "tfsf_xyz", "gate_xyz", "tfsf_side", "send_node", "peers", "peers5",
 * "node_coords", the plan of the warp-per-group kernel "wg_dims", "wg_bfrag", "wg_geo", "wg_forder", "wg_hpush", "wg_tab", "wg_desc", "wg_send_off", "wg_dev2ref"
 * — so that tests can check the setup against the oracle without a GPU.                            */
int  dgtd_setup_query(const dgtd_mesh *, const dgtd_options *, const char *name, void *buf, long long cap_bytes, long long *size_bytes);

const char *dgtd_last_error(void);
const char *dgtd_version(void);

#ifdef __cplusplus
}
#endif
#endif /* DGTD_B200_H */
