// C ABI of libdgtd_b200.so (include/dgtd_b200.h): contexts, device memory, launches, halo exchange.
// Host logic mirrors what the reference does around its operator:
//   Mult        src/evolution/GlobalEvolution.cpp:628-823   (halo exchange -> operator -> TF/SF injection)
//   RK4 step    external/mfem-geg/linalg/ode.cpp:109-136     (fused here: 4 launches, no k vector, no AXPY passes)
//   time loop   src/solver/Solver.cpp:483-551
// There is no CPU path in this file: every compute entry point needs a CUDA device.
#include "../../include/dgtd_b200.h"
#include "host.hpp"
#include "kernels.cuh"
#include "kernels_wg.cuh"
#include "kernels_wh.cuh"

#include <dlfcn.h>
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

using namespace dgtd;

static thread_local std::string g_err;
static int fail(int code, const std::string &msg) { g_err = msg; return code; }

#define CU(call)                                                                                                   \
    do {                                                                                                           \
        cudaError_t e_ = (call);                                                                                   \
        if (e_ != cudaSuccess) throw Error(DGTD_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));    \
    } while (0)

#define GUARD_BEGIN try {
#define GUARD_END                                                                       \
    }                                                                                   \
    catch (const Error &e) { return fail(e.code, e.what()); }                           \
    catch (const std::bad_alloc &) { return fail(DGTD_ERR_ARG, "out of host memory"); } \
    catch (const std::exception &e) { return fail(DGTD_ERR_ARG, e.what()); }            \
    return DGTD_OK;

struct dgtd_mesh { Mesh m; };

// ---- minimal NCCL binding, resolved at run time from the libnccl.so.2 already in the process (torch's) -------------
namespace {
typedef struct ncclComm *ncclComm_t;
struct NcclId { char internal[128]; };
struct Nccl {
    void *h = nullptr;
    int (*GetUniqueId)(NcclId *) = nullptr;
    int (*CommInitRank)(ncclComm_t *, int, NcclId, int) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    int (*Send)(const void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*Recv)(void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*AllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    void load()
    {
        if (h) return;
        h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) throw Error(DGTD_ERR_COMM, std::string("cannot load libnccl.so.2: ") + dlerror());
        auto sym = [&](const char *n) { void *p = dlsym(h, n); if (!p) throw Error(DGTD_ERR_COMM, std::string("missing NCCL symbol ") + n); return p; };
        GetUniqueId = (decltype(GetUniqueId))sym("ncclGetUniqueId");
        CommInitRank = (decltype(CommInitRank))sym("ncclCommInitRank");
        CommDestroy = (decltype(CommDestroy))sym("ncclCommDestroy");
        GroupStart = (decltype(GroupStart))sym("ncclGroupStart");
        GroupEnd = (decltype(GroupEnd))sym("ncclGroupEnd");
        Send = (decltype(Send))sym("ncclSend");
        Recv = (decltype(Recv))sym("ncclRecv");
        AllGather = (decltype(AllGather))sym("ncclAllGather");
        AllReduce = (decltype(AllReduce))sym("ncclAllReduce");
        GetErrorString = (decltype(GetErrorString))sym("ncclGetErrorString");
    }
    void check(int r, const char *what) { if (r != 0) throw Error(DGTD_ERR_COMM, std::string(what) + ": " + (GetErrorString ? GetErrorString(r) : "nccl error")); }
};
Nccl g_nccl;
constexpr int NCCL_FLOAT64 = 8;   // ncclDouble
constexpr int NCCL_UINT8 = 1, NCCL_INT32 = 2, NCCL_MAX = 2, NCCL_MIN = 3;
}  // namespace

template <class T> struct DevBuf {
    T *p = nullptr; size_t n = 0;
    void alloc(size_t count) { release(); n = count; if (count) CU(cudaMalloc(&p, count * sizeof(T))); }
    void upload(const std::vector<T> &h, size_t padTo = 0)
    {
        alloc(std::max(h.size(), padTo));
        if (n) CU(cudaMemset(p, 0, n * sizeof(T)));
        if (!h.empty()) CU(cudaMemcpy(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
    }
    void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
    ~DevBuf() { release(); }
};

typedef void (*StageFn)(const StageArgs);
struct KernelSet { StageFn fn[4]; int threads; size_t smem; };

template <int DIM, int P> static KernelSet kset()
{
    using E = Elem<DIM, P>;
    return {{stage_kernel<DIM, P, 0>, stage_kernel<DIM, P, 1>, stage_kernel<DIM, P, 2>, stage_kernel<DIM, P, 3>}, E::T, E::smem_bytes};
}
static KernelSet select_kernels(int dim, int p)
{
    switch (dim * 10 + p) {
        case 11: return kset<1, 1>(); case 12: return kset<1, 2>(); case 13: return kset<1, 3>();
        case 14: return kset<1, 4>(); case 15: return kset<1, 5>(); case 16: return kset<1, 6>();
        case 21: return kset<2, 1>(); case 22: return kset<2, 2>(); case 23: return kset<2, 3>();
        case 24: return kset<2, 4>(); case 25: return kset<2, 5>(); case 26: return kset<2, 6>();
        case 31: return kset<3, 1>(); case 32: return kset<3, 2>(); case 33: return kset<3, 3>();
        case 34: return kset<3, 4>(); case 35: return kset<3, 5>();
    }
    throw Error(DGTD_ERR_UNSUPPORTED, "no kernel for this dimension/order");
}

// the warp-per-group kernel: tetrahedra of order 1..4, element-major "aos" layout, one warp per group of 8 elements
typedef void (*WgFn)(const WgArgs);
struct WgSet { WgFn fn[4]; int threads; size_t smem; };
template <int P, bool TF> static WgSet wgset()
{
    using B = Wg<P>;
    return {{stage_wg_kernel<P, 0, TF>, stage_wg_kernel<P, 1, TF>, stage_wg_kernel<P, 2, TF>, stage_wg_kernel<P, 3, TF>}, B::T, B::smem_bytes};
}
// tf = the context injects a TF/SF plane wave
static bool select_wg(int dim, int p, bool tf, WgSet &ws)
{
    if (dim != 3) return false;
    switch (p) {
        case 1: ws = tf ? wgset<1, true>() : wgset<1, false>(); return true; case 2: ws = tf ? wgset<2, true>() : wgset<2, false>(); return true;
        case 3: ws = tf ? wgset<3, true>() : wgset<3, false>(); return true; case 4: ws = tf ? wgset<4, true>() : wgset<4, false>(); return true;
    }
    return false;
}
// the half-row kernel (kernels_wh.cuh): same plan and layout, one warp per group of FOUR elements, rows = (element, field)
template <int P, bool TF> static WgSet whset()
{
    using B = Wh<P>;
    return {{stage_wh_kernel<P, 0, TF>, stage_wh_kernel<P, 1, TF>, stage_wh_kernel<P, 2, TF>, stage_wh_kernel<P, 3, TF>}, B::T, B::smem_bytes};
}
static bool select_wh(int dim, int p, bool tf, WgSet &ws, int &groups_per_cta)
{
    if (dim != 3) return false;
    switch (p) {
        case 1: ws = tf ? whset<1, true>() : whset<1, false>(); groups_per_cta = Wh<1>::NW; return true;
        case 2: ws = tf ? whset<2, true>() : whset<2, false>(); groups_per_cta = Wh<2>::NW; return true;
        case 3: ws = tf ? whset<3, true>() : whset<3, false>(); groups_per_cta = Wh<3>::NW; return true;
        case 4: ws = tf ? whset<4, true>() : whset<4, false>(); groups_per_cta = Wh<4>::NW; return true;
    }
    return false;
}

struct dgtd_ctx {
    HostOp H;
    WgPlan WP;
    bool has_sigma = false;
    bool wh = false;                 // ... or its half-row form (kernels_wh.cuh, groups of 4 elements); wg stays set: same plan and layout
    int wg_groups_per_cta = 0;
    bool wg = false;                 // aos layout + warp-per-group kernel (the state needs a layout conversion at the ABI)
    WgSet wgs{};
    DevBuf<uint8_t> wtab;
    DevBuf<unsigned int> wwork;       // two group counters of the warp-per-group kernels: a launch draws from one and zeroes the other
    unsigned wlaunch = 0;
    DevBuf<int> dev2ref;
    bool identity = true;            // local element order == global order (single rank, no reordering)
    long long Nalloc = 0;            // scalar dofs allocated per component (padded to whole groups of 8 elements in the aos layout)
    DevBuf<double> bgeo, bafrag, stage_ref;
    DevBuf<int> bdesc;
    DevBuf<long long> bsend_off;
    Mesh mesh;                       // kept for node_coords
    int device = 0;
    int rank = 0, nranks = 1;
    long long Nloc = 0, Nglob = 0;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    KernelSet ks{};
    int grid = 0;
    DevBuf<double> D, LIFT, geo, tfsf_xyz, gate_xyz, gate, x, ya, yb, z, halo, sendbuf, scratch, tmp_in, tmp_out;
    DevBuf<int> finfo, send_node;
    DevBuf<uint8_t> ftab;
    DevPlaneWave pw{};
    int pw_on = 0;
    ncclComm_t comm = nullptr;
    // direct halo exchange over peer memory (kernels_wg.cuh: WgP2P).  p2p_mem = [face flags 0][face flags 1][halo buffer 0][halo buffer 1]
    bool p2p = false;
    DevBuf<unsigned char> p2p_mem;
    size_t p2p_hb = 0, p2p_fb = 0;                   // bytes of one halo buffer / of one per-face flag array (128-byte multiples)
    void *p2p_peer_base[P2P_MAXPEERS] = {};          // peers' p2p_mem, mapped with cudaIpcOpenMemHandle
    size_t p2p_peer_hb[P2P_MAXPEERS] = {}, p2p_peer_fb[P2P_MAXPEERS] = {};
    unsigned long long epoch = 0;                    // exchanges produced so far (all ranks run the same sequence)
    const double *pushed = nullptr;                  // vector whose traces exchange `epoch` carries (nullptr: none valid)
    DevBuf<int> p2p_err;
    DevBuf<int> hpush, dgid, dlidx;                  // dgid: local element -> caller's element index (H.elem_gid); dlidx: -> its rank among the owned elements by global id
    long long launches = 0;
    std::vector<double> hostbuf;     // staging of the by-element gather/scatter of multi-rank contexts
    DevBuf<int> flagbuf;             // one int: collective verdicts (run_until's stability flag, teardown barriers)
    DevBuf<int> smp_elem;            // dgtd_sample staging, grown on demand (point probes are called every step)
    DevBuf<double> smp_shape, smp_out;
    std::vector<struct dgtd_gather *> gathers;       // live gathers: their device resources go with the context
    ~dgtd_ctx();
};

static void fill_args(dgtd_ctx *c, StageArgs &A)
{
    A.D = c->D.p; A.LIFT = c->LIFT.p; A.geo = c->geo.p; A.finfo = reinterpret_cast<const int2 *>(c->finfo.p);
    A.ftab = c->ftab.p; A.tfsf_xyz = c->tfsf_xyz.p; A.gate = nullptr; A.halo = c->halo.p;
    A.NE = c->H.NEloc; A.stride = c->Nloc; A.hstride = (long long)c->H.n_halo_faces * c->H.Nfp;
    A.alpha = c->H.alpha; A.pw = c->pw; A.pw_on = c->pw_on;
}

// halo exchange of the face traces of `y` (GlobalEvolution.cpp:763-774 ships whole neighbour elements, 6 blocking
// MPI exchanges; here one packed NCCL group per RHS evaluation)
static void exchange(dgtd_ctx *c, const double *y)
{
    if (c->nranks == 1 || c->H.n_halo_faces == 0) return;
    if (!c->comm) throw Error(DGTD_ERR_COMM, "multi-rank context used before dgtd_comm_init");
    const int Nfp = c->H.Nfp;
    const int ns = c->H.n_halo_faces * Nfp;
    if (c->wg) {   // node records [face][node][6], one contiguous message per peer
        pack_records_kernel<<<std::min(1024, (3 * ns + 255) / 256), 256, 0, c->stream>>>(y, c->bsend_off.p, ns, c->sendbuf.p);
        c->launches++;
        g_nccl.check(g_nccl.GroupStart(), "ncclGroupStart");
        for (auto &pp : c->H.peers) {
            const size_t off = (size_t)pp.send_off * Nfp * 6, cnt = (size_t)pp.nfaces * Nfp * 6;
            g_nccl.check(g_nccl.Send(c->sendbuf.p + off, cnt, NCCL_FLOAT64, pp.rank, c->comm, c->stream), "ncclSend");
            g_nccl.check(g_nccl.Recv(c->halo.p + off, cnt, NCCL_FLOAT64, pp.rank, c->comm, c->stream), "ncclRecv");
        }
        g_nccl.check(g_nccl.GroupEnd(), "ncclGroupEnd");
        return;
    }
    pack_kernel<<<std::min(1024, (ns + 255) / 256), 256, 0, c->stream>>>(y, c->Nloc, c->send_node.p, ns, c->sendbuf.p, ns);
    c->launches++;
    g_nccl.check(g_nccl.GroupStart(), "ncclGroupStart");
    for (auto &pp : c->H.peers)
        for (int comp = 0; comp < 6; comp++) {
            const size_t off = (size_t)comp * ns + (size_t)pp.send_off * Nfp, cnt = (size_t)pp.nfaces * Nfp;
            g_nccl.check(g_nccl.Send(c->sendbuf.p + off, cnt, NCCL_FLOAT64, pp.rank, c->comm, c->stream), "ncclSend");
            g_nccl.check(g_nccl.Recv(c->halo.p + off, cnt, NCCL_FLOAT64, pp.rank, c->comm, c->stream), "ncclRecv");
        }
    g_nccl.check(g_nccl.GroupEnd(), "ncclGroupEnd");
}

// Peer-memory halo path of the warp-per-group kernel: every rank exports [flags | halo 0 | halo 1] with CUDA IPC, the
// handles travel through one ncclAllGather, every rank maps its neighbours' buffers.  All ranks must agree (ncclAllReduce
// min) or all stay on the NCCL send/recv path.  DGTD_B200_HALO=nccl forces the latter.
static void p2p_setup(dgtd_ctx *c)
{
    struct Card { cudaIpcMemHandle_t h; unsigned long long hb; unsigned long long ok; unsigned long long fb; char pad[128 - sizeof(cudaIpcMemHandle_t) - 24]; };
    static_assert(sizeof(Card) == 128, "card size");
    const char *env = std::getenv("DGTD_B200_HALO");
    int want = c->wg && c->nranks <= 512 && (int)c->H.peers.size() <= P2P_MAXPEERS && !(env && std::string(env) == "nccl");
    Card mine{};
    if (want) {
        c->p2p_hb = (((size_t)c->H.n_halo_faces * c->H.Nfp * 6 * sizeof(double)) + 127) / 128 * 128;
        c->p2p_fb = (((size_t)c->H.n_halo_faces * sizeof(unsigned long long)) + 127) / 128 * 128;
        c->p2p_mem.alloc(2 * c->p2p_fb + 2 * c->p2p_hb + 128);
        CU(cudaMemset(c->p2p_mem.p, 0, c->p2p_mem.n));
        c->p2p_err.alloc(1);
        CU(cudaMemset(c->p2p_err.p, 0, sizeof(int)));
        if (cudaIpcGetMemHandle(&mine.h, c->p2p_mem.p) != cudaSuccess) { cudaGetLastError(); want = 0; }
        mine.hb = c->p2p_hb; mine.fb = c->p2p_fb;
    }
    mine.ok = (unsigned long long)want;
    DevBuf<unsigned char> dsend, drecv;
    dsend.alloc(sizeof(Card)); drecv.alloc(sizeof(Card) * (size_t)c->nranks);
    CU(cudaMemcpyAsync(dsend.p, &mine, sizeof(Card), cudaMemcpyHostToDevice, c->stream));
    g_nccl.check(g_nccl.AllGather(dsend.p, drecv.p, sizeof(Card), NCCL_UINT8, c->comm, c->stream), "ncclAllGather");
    std::vector<Card> all((size_t)c->nranks);
    CU(cudaMemcpyAsync(all.data(), drecv.p, sizeof(Card) * (size_t)c->nranks, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    int ok = want;
    for (auto &cd : all) ok &= (int)cd.ok;
    if (ok)
        for (size_t p = 0; p < c->H.peers.size(); p++) {
            const Card &cd = all[(size_t)c->H.peers[p].rank];
            if (cudaIpcOpenMemHandle(&c->p2p_peer_base[p], cd.h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); c->p2p_peer_base[p] = nullptr; ok = 0; break; }
            c->p2p_peer_hb[p] = (size_t)cd.hb; c->p2p_peer_fb[p] = (size_t)cd.fb;
        }
    // collective verdict
    DevBuf<int> dflag; dflag.alloc(1);
    CU(cudaMemcpyAsync(dflag.p, &ok, sizeof(int), cudaMemcpyHostToDevice, c->stream));
    g_nccl.check(g_nccl.AllReduce(dflag.p, dflag.p, 1, NCCL_INT32, NCCL_MIN, c->comm, c->stream), "ncclAllReduce");
    CU(cudaMemcpyAsync(&ok, dflag.p, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    if (!ok) {
        for (void *&b : c->p2p_peer_base) if (b) { cudaIpcCloseMemHandle(b); b = nullptr; }
        c->p2p = false;
        return;
    }
    c->p2p = true; c->epoch = 0; c->pushed = nullptr;
}
// collective barrier on the context's stream (teardown: all ranks' kernels must have drained before any rank unmaps or
// frees a halo buffer its neighbours store into).  No allocation, no exception: it runs inside dgtd_destroy.
static void comm_barrier(dgtd_ctx *c)
{
    if (!c->comm || !c->flagbuf.p) return;
    cudaStreamSynchronize(c->stream);
    cudaMemsetAsync(c->flagbuf.p, 0, sizeof(int), c->stream);
    g_nccl.AllReduce(c->flagbuf.p, c->flagbuf.p, 1, NCCL_INT32, NCCL_MIN, c->comm, c->stream);
    cudaStreamSynchronize(c->stream);
}
static void p2p_check(dgtd_ctx *c)
{
    if (!c->p2p) return;
    int e = 0;
    CU(cudaMemcpy(&e, c->p2p_err.p, sizeof(int), cudaMemcpyDeviceToHost));
    if (e) throw Error(DGTD_ERR_COMM, "halo exchange: a neighbour rank did not signal within 20 s");
}

static void launch_gate(dgtd_ctx *c, const double *ts, int nt)
{
    CU(cudaMemsetAsync(c->gate.p, 0, 4 * sizeof(double), c->stream));
    const int V = (int)(c->H.gate_xyz.size() / 3);
    gate_kernel<<<std::min(296, (V + 255) / 256), 256, 0, c->stream>>>(c->gate_xyz.p, V, c->pw, ts[0], ts[1], ts[2], ts[3], nt, c->gate.p);
    c->launches++;
}

// WgP2P of one launch: consume exchange `wait` (my flags and halo buffer of parity wait & 1), produce exchange `signal` into
// the peers' buffers of parity signal & 1
static WgP2P p2p_args(dgtd_ctx *c, unsigned long long wait, unsigned long long signal)
{
    WgP2P q{};
    if (!c->p2p) return q;
    q.hpush = reinterpret_cast<const int2 *>(c->hpush.p);
    q.flags = reinterpret_cast<const unsigned long long *>(c->p2p_mem.p + (wait & 1) * c->p2p_fb);
    for (size_t p = 0; p < c->H.peers.size(); p++) {
        unsigned char *base = static_cast<unsigned char *>(c->p2p_peer_base[p]);
        q.peer_flag[p] = reinterpret_cast<unsigned long long *>(base + (signal & 1) * c->p2p_peer_fb[p]);
        q.peer_out[p] = reinterpret_cast<double *>(base + 2 * c->p2p_peer_fb[p] + (signal & 1) * c->p2p_peer_hb[p]);
    }
    q.wait_epoch = wait; q.signal_epoch = signal;
    q.err = c->p2p_err.p;
    return q;
}
static const double *p2p_halo_in(dgtd_ctx *c, unsigned long long k)
{
    return reinterpret_cast<const double *>(c->p2p_mem.p + 2 * c->p2p_fb + (k & 1) * c->p2p_hb);
}

static void launch_stage(dgtd_ctx *c, int mode, StageArgs &A)
{
    const bool p2p = c->p2p && c->H.n_halo_faces > 0;
    if (p2p) {
        if (c->pushed != A.yin) {   // the traces of y_in are not at the peers yet: stand-alone producer of the next exchange
            const int nf = c->H.n_halo_faces;
            c->epoch++;
            halo_push_kernel<<<std::min(592, (nf + 63) / 64), 64, 0, c->stream>>>(A.yin, c->bsend_off.p, nf, c->H.Nfp, p2p_args(c, c->epoch - 1, c->epoch));
            c->launches++;
            c->pushed = A.yin;
        }
    } else exchange(c, A.yin);
    if (c->wg) {
        WgArgs W;
        W.pp = p2p_args(c, 0, 0);
        if (p2p) {
            A.halo = p2p_halo_in(c, c->epoch);
            if (mode != MODE_MULT) { W.pp = p2p_args(c, c->epoch, c->epoch + 1); c->epoch++; c->pushed = A.yout; }
            else W.pp = p2p_args(c, c->epoch, 0);
        }
        W.bfrag = c->bafrag.p; W.geo = c->bgeo.p; W.desc = c->bdesc.p; W.tab = c->wtab.p; W.ntab = c->WP.ntab;
        W.tfsf_xyz = A.tfsf_xyz; W.gate = A.gate; W.halo = A.halo; W.ngroups = c->WP.ngroups; W.has_sigma = c->has_sigma ? 1 : 0;
        W.alpha = A.alpha; W.pw = A.pw; W.pw_on = A.pw_on;
        W.yin = A.yin; W.x = A.x; W.z = A.z; W.yout = A.yout; W.a = A.a; W.b = A.b; W.t = A.t;
        // dynamic group scheduling on multi-rank contexts (partition-face groups cost more) and for the half-row kernel
        // (order 4: +3 %); a single rank at order <= 3 keeps the static split (118.6 vs 119.8 G)
        const bool dyn = c->nranks > 1 || c->wh || std::getenv("DGTD_B200_DYNAMIC") != nullptr;
        W.work = dyn ? c->wwork.p + (c->wlaunch & 1) : nullptr; W.work_next = c->wwork.p + ((c->wlaunch + 1) & 1); c->wlaunch++;
        c->wgs.fn[mode]<<<c->grid, c->wgs.threads, c->wgs.smem, c->stream>>>(W);
    } else {
        c->ks.fn[mode]<<<c->grid, c->ks.threads, c->ks.smem, c->stream>>>(A);
    }
    c->launches++;
    CU(cudaGetLastError());
}

static void rk4_step(dgtd_ctx *c, double t, double dt)
{
    StageArgs A; fill_args(c, A);
    const bool gated = c->pw_on && c->H.tfsf_gate;
    if (gated) { double ts[4] = {t, t + dt / 2, t + dt, 0}; launch_gate(c, ts, 3); }
    // k1 = f(t, x); y = x + dt/2 k1; z = x + dt/6 k1
    A.yin = c->x.p; A.x = c->x.p; A.z = c->z.p; A.yout = c->ya.p; A.a = dt / 2; A.b = dt / 6; A.t = t; A.gate = gated ? c->gate.p + 0 : nullptr;
    launch_stage(c, MODE_STAGE1, A);
    // k2 = f(t + dt/2, y); y = x + dt/2 k2; z += dt/3 k2
    A.yin = c->ya.p; A.yout = c->yb.p; A.a = dt / 2; A.b = dt / 3; A.t = t + dt / 2; A.gate = gated ? c->gate.p + 1 : nullptr;
    launch_stage(c, MODE_STAGE23, A);
    // k3 = f(t + dt/2, y)  (time not advanced, ode.cpp:127); y = x + dt k3; z += dt/3 k3
    A.yin = c->yb.p; A.yout = c->ya.p; A.a = dt; A.b = dt / 3;
    launch_stage(c, MODE_STAGE23, A);
    // k4 = f(t + dt, y); x = z + dt/6 k4
    A.yin = c->ya.p; A.yout = c->x.p; A.b = dt / 6; A.t = t + dt; A.gate = gated ? c->gate.p + 2 : nullptr;
    launch_stage(c, MODE_STAGE4, A);
}

static void mult_device(dgtd_ctx *c, double t, const double *in, double *out)
{
    if (in == out) throw Error(DGTD_ERR_ARG, "Mult: in and out must not alias");
    StageArgs A; fill_args(c, A);
    const bool gated = c->pw_on && c->H.tfsf_gate;
    if (gated) { double ts[4] = {t, 0, 0, 0}; launch_gate(c, ts, 1); }
    A.yin = in; A.x = in; A.z = nullptr; A.yout = out; A.a = 0; A.b = 0; A.t = t; A.gate = gated ? c->gate.p : nullptr;
    launch_stage(c, MODE_MULT, A);
}

// reference-layout device vector [6][Nloc] <-> the aos state layout of the warp-per-group kernels;
// gid: the vector is in the caller's element order (single rank) and the Morton permutation is applied on the fly
static void to_device_layout(dgtd_ctx *c, const double *ref, double *dev, const int *gid = nullptr)
{
    to_aos_kernel<<<1184, 256, 0, c->stream>>>(ref, c->Nloc, c->H.Np, c->H.NEloc, c->WP.NEpad, c->dev2ref.p, gid, dev);
    c->launches++;
}
static void from_device_layout(dgtd_ctx *c, const double *dev, double *ref, const int *gid = nullptr)
{
    from_aos_kernel<<<1184, 256, 0, c->stream>>>(dev, c->Nloc, c->H.Np, c->H.NEloc, c->dev2ref.p, gid, ref);
    c->launches++;
}

// host vector [6][Nloc] <-> device state (reference layout, or aos through a staging buffer).  perm = nullptr: the host
// vector is in this rank's own (Morton) element order; otherwise local element le is element perm[le] of the host vector
// (dgid: caller's global order of a single-rank context; dlidx: the owned elements by ascending global id = the element
// order of the rank's mfem::ParMesh)
static void upload_local(dgtd_ctx *c, const double *hloc, double *dev, const int *perm = nullptr)
{
    const long long Nl = c->Nloc;
    if (c->pushed == dev) c->pushed = nullptr;
    if (!c->wg && !perm) {
        CU(cudaMemcpyAsync(dev, hloc, sizeof(double) * 6 * Nl, cudaMemcpyHostToDevice, c->stream));
    } else {
        if (c->stage_ref.n != (size_t)6 * Nl) c->stage_ref.alloc((size_t)6 * Nl);
        CU(cudaMemcpyAsync(c->stage_ref.p, hloc, sizeof(double) * 6 * Nl, cudaMemcpyHostToDevice, c->stream));
        if (c->wg) to_device_layout(c, c->stage_ref.p, dev, perm);
        else { permute_elements_kernel<<<1184, 256, 0, c->stream>>>(c->stage_ref.p, perm, c->H.Np, c->H.NEloc, Nl, 1, dev); c->launches++; }
    }
    CU(cudaStreamSynchronize(c->stream));
}
static void download_local(dgtd_ctx *c, const double *dev, double *hloc, const int *perm = nullptr)
{
    const long long Nl = c->Nloc;
    p2p_check(c);
    if (!c->wg && !perm) {
        CU(cudaMemcpyAsync(hloc, dev, sizeof(double) * 6 * Nl, cudaMemcpyDeviceToHost, c->stream));
    } else {
        if (c->stage_ref.n != (size_t)6 * Nl) c->stage_ref.alloc((size_t)6 * Nl);
        if (c->wg) from_device_layout(c, dev, c->stage_ref.p, perm);
        else { permute_elements_kernel<<<1184, 256, 0, c->stream>>>(dev, perm, c->H.Np, c->H.NEloc, Nl, 0, c->stage_ref.p); c->launches++; }
        CU(cudaMemcpyAsync(hloc, c->stage_ref.p, sizeof(double) * 6 * Nl, cudaMemcpyDeviceToHost, c->stream));
    }
    CU(cudaStreamSynchronize(c->stream));
}
// global [6N] host vector <-> device state of this rank
static void scatter_to_device(dgtd_ctx *c, const double *host, double *dev)
{
    const int Np = c->H.Np; const long long Ng = c->Nglob, Nl = c->Nloc;
    if (c->identity) { upload_local(c, host, dev); return; }
    if (c->nranks == 1) { upload_local(c, host, dev, c->dgid.p); return; }   // the permutation runs on the device
    c->hostbuf.resize((size_t)6 * Nl);
    for (int comp = 0; comp < 6; comp++)
        for (int le = 0; le < c->H.NEloc; le++)
            std::memcpy(&c->hostbuf[(size_t)comp * Nl + (size_t)le * Np], host + (size_t)comp * Ng + (size_t)c->H.elem_gid[le] * Np, sizeof(double) * Np);
    upload_local(c, c->hostbuf.data(), dev);
}
static void gather_from_device(dgtd_ctx *c, const double *dev, double *host)
{
    const int Np = c->H.Np; const long long Ng = c->Nglob, Nl = c->Nloc;
    if (c->identity) { download_local(c, dev, host); return; }
    if (c->nranks == 1) { download_local(c, dev, host, c->dgid.p); return; }
    c->hostbuf.resize((size_t)6 * Nl);
    download_local(c, dev, c->hostbuf.data());
    for (int comp = 0; comp < 6; comp++)
        for (int le = 0; le < c->H.NEloc; le++)
            std::memcpy(host + (size_t)comp * Ng + (size_t)c->H.elem_gid[le] * Np, &c->hostbuf[(size_t)comp * Nl + (size_t)le * Np], sizeof(double) * Np);
}

static Options options_from_c(const dgtd_mesh *mesh, const dgtd_options *o)
{
    Options op;
    op.order = o->order; op.alpha = o->alpha; op.rank = o->rank; op.nranks = o->nranks < 1 ? 1 : o->nranks; op.tfsf_gate = o->tfsf_gate != 0;
    if ((o->n_bdr && (!o->bdr_attr || !o->bdr_cond)) || (o->n_tfsf && !o->tfsf_attr) || (o->n_mat && (!o->mat_attr || !o->mat_eps_mu_sigma)))
        throw Error(DGTD_ERR_ARG, "dgtd_create: option arrays missing");
    for (int i = 0; i < o->n_bdr; i++) op.bdr.push_back({o->bdr_attr[i], o->bdr_cond[i]});
    for (int i = 0; i < o->n_tfsf; i++) op.tfsf.push_back(o->tfsf_attr[i]);
    for (int i = 0; i < o->n_mat; i++) op.mat.push_back({o->mat_attr[i], {o->mat_eps_mu_sigma[3 * i], o->mat_eps_mu_sigma[3 * i + 1], o->mat_eps_mu_sigma[3 * i + 2]}});
    if (o->partitioning) op.partitioning.assign(o->partitioning, o->partitioning + mesh->m.ne());
    if (o->pw.enabled) {
        // Planewave ctor normalises pol and dir (Function.h:330-339); H polarisation = dir x pol (E-type) / pol x dir gives E (H-type)
        const dgtd_planewave &w = o->pw;
        double pn = std::sqrt(w.pol[0] * w.pol[0] + w.pol[1] * w.pol[1] + w.pol[2] * w.pol[2]);
        double kn = std::sqrt(w.dir[0] * w.dir[0] + w.dir[1] * w.dir[1] + w.dir[2] * w.dir[2]);
        if (!(pn > 0) || !(kn > 0) || !(w.spread > 0)) throw Error(DGTD_ERR_ARG, "plane wave needs non-zero polarisation, propagation and spread");
        double p[3], k[3];
        for (int d = 0; d < 3; d++) { p[d] = w.pol[d] / pn; k[d] = w.dir[d] / kn; }
        double kxp[3] = {k[1] * p[2] - k[2] * p[1], k[2] * p[0] - k[0] * p[2], k[0] * p[1] - k[1] * p[0]};
        op.pw.enabled = true; op.pw.spread = w.spread; op.pw.mean1d = w.mean1d; op.pw.freq = w.freq;
        for (int d = 0; d < 3; d++) {
            op.pw.dir[d] = k[d];
            if (w.fieldtype == 0) { op.pw.pe[d] = p[d]; op.pw.ph[d] = kxp[d]; }
            else { op.pw.ph[d] = p[d]; op.pw.pe[d] = -kxp[d]; }
        }
    }
    return op;
}

// =====================================================================================================================
extern "C" {

const char *dgtd_last_error(void) { return g_err.c_str(); }
const char *dgtd_version(void) { return "dgtd_b200 0.1 (sm_100a, fp64)"; }

int dgtd_mesh_from_arrays(int dim, int nv, const double *verts, int ne, const int *elems, const int *elem_attr, int nbe,
                          const int *bdr, const int *bdr_attr, dgtd_mesh **out)
{
    GUARD_BEGIN
    if (!out || !verts || !elems || nv <= 0 || ne <= 0 || nbe < 0 || (nbe > 0 && (!bdr || !bdr_attr))) throw Error(DGTD_ERR_ARG, "dgtd_mesh_from_arrays: bad arguments");
    if (dim < 1 || dim > 3) throw Error(DGTD_ERR_MESH, "mesh dimension must be 1, 2 or 3");
    auto m = std::make_unique<dgtd_mesh>();
    m->m.dim = dim;
    m->m.verts.assign(verts, verts + 3 * (size_t)nv);
    m->m.elems.assign(elems, elems + (size_t)ne * (dim + 1));
    if (elem_attr) m->m.elem_attr.assign(elem_attr, elem_attr + ne); else m->m.elem_attr.assign(ne, 1);
    if (nbe) { m->m.bdr.assign(bdr, bdr + (size_t)nbe * dim); m->m.bdr_attr.assign(bdr_attr, bdr_attr + nbe); }
    m->m.validate_and_orient(false);     // the caller's element-local order is a contract: inverted elements are rejected, not swapped
    *out = m.release();
    GUARD_END
}
int dgtd_mesh_load(const char *path, dgtd_mesh **out)
{
    GUARD_BEGIN
    if (!path || !out) throw Error(DGTD_ERR_ARG, "dgtd_mesh_load: null argument");
    auto m = std::make_unique<dgtd_mesh>();
    m->m = load_mesh(path);
    *out = m.release();
    GUARD_END
}
int dgtd_mesh_cartesian3d(int nx, int ny, int nz, double sx, double sy, double sz, dgtd_mesh **out)
{
    GUARD_BEGIN
    if (!out) throw Error(DGTD_ERR_ARG, "null out");
    auto m = std::make_unique<dgtd_mesh>();
    m->m = cartesian3d(nx, ny, nz, sx, sy, sz);
    *out = m.release();
    GUARD_END
}
int dgtd_mesh_info(const dgtd_mesh *m, int *dim, int *nv, int *ne, int *nbe)
{
    if (!m) return fail(DGTD_ERR_ARG, "null mesh");
    if (dim) *dim = m->m.dim; if (nv) *nv = m->m.nv(); if (ne) *ne = m->m.ne(); if (nbe) *nbe = m->m.nbe();
    return DGTD_OK;
}
int dgtd_mesh_get_arrays(const dgtd_mesh *m, double *verts, int *elems, int *elem_attr, int *bdr, int *bdr_attr)
{
    if (!m) return fail(DGTD_ERR_ARG, "null mesh");
    auto cp = [](auto *dst, const auto &v) { if (dst && !v.empty()) std::memcpy(dst, v.data(), v.size() * sizeof(v[0])); };
    cp(verts, m->m.verts); cp(elems, m->m.elems); cp(elem_attr, m->m.elem_attr); cp(bdr, m->m.bdr); cp(bdr_attr, m->m.bdr_attr);
    return DGTD_OK;
}
int dgtd_mesh_partition(const dgtd_mesh *m, int nranks, int *partitioning)
{
    GUARD_BEGIN
    if (!m || !partitioning) throw Error(DGTD_ERR_ARG, "null argument");
    auto p = partition_rcb(m->m, nranks);
    std::memcpy(partitioning, p.data(), p.size() * sizeof(int));
    GUARD_END
}
int dgtd_mesh_partition_metis(const dgtd_mesh *m, int nranks, int *partitioning)
{
    GUARD_BEGIN
    if (!m || !partitioning) throw Error(DGTD_ERR_ARG, "null argument");
    auto p = partition_metis(m->m, nranks);
    std::memcpy(partitioning, p.data(), p.size() * sizeof(int));
    GUARD_END
}
void dgtd_mesh_destroy(dgtd_mesh *m) { delete m; }

int dgtd_create(const dgtd_mesh *mesh, const dgtd_options *o, dgtd_ctx **out)
{
    GUARD_BEGIN
    if (!mesh || !o || !out) throw Error(DGTD_ERR_ARG, "dgtd_create: null argument");
    Options op = options_from_c(mesh, o);
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) throw Error(DGTD_ERR_CUDA, "no CUDA device: dgtd_b200 has no CPU fallback");
    if (o->device < 0 || o->device >= ndev) throw Error(DGTD_ERR_CUDA, "CUDA device ordinal out of range");
    CU(cudaSetDevice(o->device));
    cudaDeviceProp prop; CU(cudaGetDeviceProperties(&prop, o->device));
    if (prop.major < 10) throw Error(DGTD_ERR_CUDA, std::string("dgtd_b200 is built for sm_100a only; device is ") + prop.name);

    auto c = std::make_unique<dgtd_ctx>();
    c->device = o->device; c->rank = op.rank; c->nranks = op.nranks;
    c->mesh = mesh->m;
    c->H = build_host_op(mesh->m, op);
    HostOp &H = c->H;
    c->Nloc = (long long)H.NEloc * H.Np; c->Nglob = H.NEglob * H.Np;
    c->identity = c->nranks == 1;
    for (int le = 0; le < H.NEloc && c->identity; le++) c->identity = H.elem_gid[le] == le;
    // kernel choice: DGTD_B200_KERNEL = wg | wh | generic overrides the default
    const char *kenv = std::getenv("DGTD_B200_KERNEL");
    const std::string ksel = kenv ? kenv : "";
    bool has_sigma = false;
    for (int le = 0; le < H.NEloc; le++) has_sigma |= H.geo[(size_t)le * GEO_STRIDE + 15] != 0.0;
    c->has_sigma = has_sigma;
    const bool has_tf = H.pw.enabled && H.n_tfsf_faces > 0;
    // default for tetrahedra: the half-row kernel at order 4 (8 instead of 4 warps per SM), the one-warp-per-8-elements
    // kernel below it; everything else (segments, triangles, order-5 tetrahedra): the generic kernel
    if ((ksel == "wh" || (ksel.empty() && H.p == 4)) && select_wh(H.dim, H.p, has_tf, c->wgs, c->wg_groups_per_cta) && c->wgs.smem <= (size_t)prop.sharedMemPerBlockOptin) {
        c->WP = build_wg_plan(H);
        if (c->WP.ntab <= Wg<3>::TABROWS) c->wg = c->wh = true;
    }
    if (!c->wg && (ksel == "wg" || ksel == "wh" || ksel.empty()) && select_wg(H.dim, H.p, has_tf, c->wgs) && c->wgs.smem <= (size_t)prop.sharedMemPerBlockOptin) {
        c->WP = build_wg_plan(H);
        c->wg_groups_per_cta = c->wgs.threads / 32;
        if (c->WP.ntab <= Wg<3>::TABROWS) c->wg = true;
    }
    if (c->wg) {
        c->Nalloc = (long long)c->WP.NEpad * H.Np;
        for (int m = 0; m < 4; m++) CU(cudaFuncSetAttribute((const void *)c->wgs.fn[m], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->wgs.smem));
        const int nw = c->wg_groups_per_cta;
        const long long units = (long long)c->WP.ngroups * (c->wh ? BLK_E / WH_E : 1);      // groups of 8, or of 4 elements
        c->grid = (int)std::min<long long>((units + nw - 1) / nw, (long long)prop.multiProcessorCount);
        c->bgeo.upload(c->WP.geo); c->bafrag.upload(c->WP.bfrag); c->bdesc.upload(c->WP.desc); c->bsend_off.upload(c->WP.send_off, 1);
        c->wtab.upload(c->WP.tab, 16); c->dev2ref.upload(c->WP.dev2ref); c->hpush.upload(c->WP.hpush, 2);
        c->wwork.alloc(2); CU(cudaMemset(c->wwork.p, 0, 2 * sizeof(unsigned int)));
    } else {
        if (H.ntab > 256) throw Error(DGTD_ERR_UNSUPPORTED, "too many distinct face orientations for the generic kernel");
        c->Nalloc = c->Nloc;
        c->ks = select_kernels(H.dim, H.p);
        if (c->ks.smem > (size_t)prop.sharedMemPerBlockOptin) throw Error(DGTD_ERR_UNSUPPORTED, "order too high for the shared-memory tiling");
        for (int m = 0; m < 4; m++) CU(cudaFuncSetAttribute((const void *)c->ks.fn[m], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->ks.smem));
        int occ = 0; CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, (const void *)c->ks.fn[2], c->ks.threads, c->ks.smem));
        if (occ < 1) throw Error(DGTD_ERR_UNSUPPORTED, "stage kernel does not fit on an SM");
        int EB = c->ks.threads / H.Np;
        long long nbatch = (H.NEloc + EB - 1) / EB;
        c->grid = (int)std::min<long long>(nbatch, (long long)prop.multiProcessorCount * occ);
    }
    CU(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
    c->stream = c->own_stream;
    // operator tables, transposed for the kernels
    {
        const int Np = H.Np, Nfp = H.Nfp, nf = H.nf, dim = H.dim;
        std::vector<double> Dt((size_t)dim * Np * Np), Lt((size_t)nf * Nfp * Np);
        for (int x = 0; x < dim; x++) for (int i = 0; i < Np; i++) for (int j = 0; j < Np; j++) Dt[((size_t)x * Np + j) * Np + i] = H.ref.D[((size_t)x * Np + i) * Np + j];
        for (int f = 0; f < nf; f++) for (int i = 0; i < Np; i++) for (int j = 0; j < Nfp; j++) Lt[((size_t)f * Nfp + j) * Np + i] = H.ref.lift[((size_t)f * Np + i) * Nfp + j];
        c->D.upload(Dt); c->LIFT.upload(Lt);
    }
    c->geo.upload(H.geo); c->finfo.upload(H.finfo); c->ftab.upload(H.ftab, (size_t)256 * H.Nfp);
    c->tfsf_xyz.upload(H.tfsf_xyz, 3); c->gate_xyz.upload(H.gate_xyz, 3); c->gate.alloc(4);
    CU(cudaMemset(c->gate.p, 0, 4 * sizeof(double)));
    c->send_node.upload(H.send_node, 1);
    c->dgid.upload(H.elem_gid, 1);
    {   // the rank's mfem::ParMesh numbers its elements by ascending global id (ParMesh(comm, mesh, partitioning))
        std::vector<int> sorted(H.elem_gid), lidx((size_t)H.NEloc);
        std::sort(sorted.begin(), sorted.end());
        for (int le = 0; le < H.NEloc; le++) lidx[(size_t)le] = (int)(std::lower_bound(sorted.begin(), sorted.end(), H.elem_gid[(size_t)le]) - sorted.begin());
        c->dlidx.upload(lidx, 1);
    }
    const size_t hn = std::max<size_t>(1, (size_t)6 * H.n_halo_faces * H.Nfp);
    c->halo.alloc(hn); c->sendbuf.alloc(hn); c->scratch.alloc(4); c->flagbuf.alloc(1);
    CU(cudaMemset(c->halo.p, 0, hn * sizeof(double)));
    const size_t n6 = (size_t)6 * c->Nalloc;
    c->x.alloc(n6); c->ya.alloc(n6); c->yb.alloc(n6); c->z.alloc(n6);
    CU(cudaMemset(c->x.p, 0, n6 * sizeof(double))); CU(cudaMemset(c->ya.p, 0, n6 * sizeof(double)));
    CU(cudaMemset(c->yb.p, 0, n6 * sizeof(double))); CU(cudaMemset(c->z.p, 0, n6 * sizeof(double)));
    c->pw_on = H.pw.enabled && H.n_tfsf_faces > 0;
    if (H.pw.enabled) {
        c->pw.inv2s2 = 1.0 / (2.0 * H.pw.spread * H.pw.spread); c->pw.mean1d = H.pw.mean1d;
        c->pw.twopif = 2.0 * M_PI * H.pw.freq;
        for (int d = 0; d < 3; d++) { c->pw.pe[d] = H.pw.pe[d]; c->pw.ph[d] = H.pw.ph[d]; c->pw.dir[d] = H.pw.dir[d]; }
    }
    CU(cudaDeviceSynchronize());
    *out = c.release();
    GUARD_END
}
void dgtd_destroy(dgtd_ctx *c)
{
    if (!c) return;
    try {
        cudaSetDevice(c->device);
        delete c;               // ~dgtd_ctx: collective teardown of the halo mappings (two barriers), then the device memory
    } catch (...) {             // nothing may cross the C ABI
    }
}
int dgtd_sizes(const dgtd_ctx *c, long long *n_global, int *np, long long *ne_local, long long *n_local)
{
    if (!c) return fail(DGTD_ERR_ARG, "null context");
    if (n_global) *n_global = c->Nglob; if (np) *np = c->H.Np; if (ne_local) *ne_local = c->H.NEloc; if (n_local) *n_local = c->Nloc;
    return DGTD_OK;
}
int dgtd_local_elements(const dgtd_ctx *c, int *ids)
{
    if (!c || !ids) return fail(DGTD_ERR_ARG, "null argument");
    std::memcpy(ids, c->H.elem_gid.data(), sizeof(int) * c->H.elem_gid.size());
    return DGTD_OK;
}
int dgtd_node_coords(const dgtd_ctx *c, double *xyz)
{
    GUARD_BEGIN
    if (!c || !xyz) throw Error(DGTD_ERR_ARG, "null argument");
    std::vector<double> v; node_coords(c->mesh, c->H.ref, v);
    std::memcpy(xyz, v.data(), v.size() * sizeof(double));
    GUARD_END
}
int dgtd_set_stream(dgtd_ctx *c, void *s)
{
    if (!c) return fail(DGTD_ERR_ARG, "null context");
    c->stream = s ? (cudaStream_t)s : c->own_stream;
    return DGTD_OK;
}
int dgtd_set_state(dgtd_ctx *c, const double *h)
{
    GUARD_BEGIN
    if (!c || !h) throw Error(DGTD_ERR_ARG, "null argument");
    CU(cudaSetDevice(c->device));
    scatter_to_device(c, h, c->x.p);
    GUARD_END
}
int dgtd_get_state(dgtd_ctx *c, double *h)
{
    GUARD_BEGIN
    if (!c || !h) throw Error(DGTD_ERR_ARG, "null argument");
    CU(cudaSetDevice(c->device));
    gather_from_device(c, c->x.p, h);
    GUARD_END
}
int dgtd_set_state_local(dgtd_ctx *c, const double *h)
{
    GUARD_BEGIN
    if (!c || !h) throw Error(DGTD_ERR_ARG, "null argument");
    CU(cudaSetDevice(c->device));
    upload_local(c, h, c->x.p);
    GUARD_END
}
int dgtd_get_state_local(dgtd_ctx *c, double *h)
{
    GUARD_BEGIN
    if (!c || !h) throw Error(DGTD_ERR_ARG, "null argument");
    CU(cudaSetDevice(c->device));
    download_local(c, c->x.p, h);
    GUARD_END
}
int dgtd_set_state_parlocal(dgtd_ctx *c, const double *h)
{
    GUARD_BEGIN
    if (!c || !h) throw Error(DGTD_ERR_ARG, "null argument");
    CU(cudaSetDevice(c->device));
    upload_local(c, h, c->x.p, c->dlidx.p);
    GUARD_END
}
int dgtd_get_state_parlocal(dgtd_ctx *c, double *h)
{
    GUARD_BEGIN
    if (!c || !h) throw Error(DGTD_ERR_ARG, "null argument");
    CU(cudaSetDevice(c->device));
    download_local(c, c->x.p, h, c->dlidx.p);
    GUARD_END
}
int dgtd_mult_parlocal(dgtd_ctx *c, double t, const double *in, double *out)
{
    GUARD_BEGIN
    if (!c || !in || !out) throw Error(DGTD_ERR_ARG, "null argument");
    CU(cudaSetDevice(c->device));
    const size_t n6 = (size_t)6 * c->Nalloc;
    c->pushed = nullptr;
    if (c->tmp_in.n != n6) { c->tmp_in.alloc(n6); c->tmp_out.alloc(n6); }
    upload_local(c, in, c->tmp_in.p, c->dlidx.p);
    mult_device(c, t, c->tmp_in.p, c->tmp_out.p);
    download_local(c, c->tmp_out.p, out, c->dlidx.p);
    GUARD_END
}
int dgtd_state_device_ptr(dgtd_ctx *c, double **dev)
{
    if (!c || !dev) return fail(DGTD_ERR_ARG, "null argument");
    *dev = c->x.p;
    c->pushed = nullptr;      // the caller may write the state behind our back
    return DGTD_OK;
}
int dgtd_mult(dgtd_ctx *c, double t, const double *in, double *out, int on_device)
{
    GUARD_BEGIN
    if (!c || !in || !out) throw Error(DGTD_ERR_ARG, "null argument");
    CU(cudaSetDevice(c->device));
    const size_t n6 = (size_t)6 * c->Nalloc;
    c->pushed = nullptr;   // tmp_in / a caller-owned vector is about to change under the same address
    if (on_device && !c->wg) { mult_device(c, t, in, out); }
    else {
        if (c->tmp_in.n != n6) { c->tmp_in.alloc(n6); c->tmp_out.alloc(n6); }
        if (on_device) {   // reference-layout device vectors [6][n_local] <-> blocked
            to_device_layout(c, in, c->tmp_in.p);
            mult_device(c, t, c->tmp_in.p, c->tmp_out.p);
            from_device_layout(c, c->tmp_out.p, out);
        } else {
            scatter_to_device(c, in, c->tmp_in.p);
            mult_device(c, t, c->tmp_in.p, c->tmp_out.p);
            gather_from_device(c, c->tmp_out.p, out);
        }
    }
    GUARD_END
}
int dgtd_rk4_step(dgtd_ctx *c, double t, double dt)
{
    GUARD_BEGIN
    if (!c) throw Error(DGTD_ERR_ARG, "null context");
    CU(cudaSetDevice(c->device));
    rk4_step(c, t, dt);
    GUARD_END
}
int dgtd_rk4_run(dgtd_ctx *c, double t0, double dt, int nsteps)
{
    GUARD_BEGIN
    if (!c || nsteps < 0) throw Error(DGTD_ERR_ARG, "bad argument");
    CU(cudaSetDevice(c->device));
    double t = t0;
    for (int s = 0; s < nsteps; s++) { rk4_step(c, t, dt); t += dt; }
    GUARD_END
}
int dgtd_run_until(dgtd_ctx *c, double *t, double dt, double t_final, int check_every, long long *nsteps, int *unstable)
{
    GUARD_BEGIN
    if (!c || !t || !(dt > 0) || check_every < 0) throw Error(DGTD_ERR_ARG, "dgtd_run_until: bad argument");
    CU(cudaSetDevice(c->device));
    long long n = 0;
    int bad = 0;
    // Solver::run / Solver::step (Solver.cpp:497-551): while (time <= final - 1e-8 dt) { truedt = min(dt, final - time); Step; norm check }
    while (*t <= t_final - 1e-8 * dt) {
        const double truedt = std::min(dt, t_final - *t);
        rk4_step(c, *t, truedt);
        *t += truedt;
        n++;
        if (check_every && n % check_every == 0) {
            double ss = 0;
            CU(cudaMemsetAsync(c->scratch.p, 0, sizeof(double), c->stream));
            sumsq_kernel<<<296, 256, 0, c->stream>>>(c->x.p, 6 * c->Nalloc, c->scratch.p);
            c->launches++;
            CU(cudaMemcpyAsync(&ss, c->scratch.p, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
            CU(cudaStreamSynchronize(c->stream));
            const double nrm = std::sqrt(ss);
            int flag = (!std::isfinite(nrm) || nrm > 1e20) ? 1 : 0;
            if (c->nranks > 1 && c->comm) {      // all ranks must leave the loop together (Solver.cpp:503: MPI_Allreduce MAX)
                int *d = c->flagbuf.p;
                CU(cudaMemcpyAsync(d, &flag, sizeof(int), cudaMemcpyHostToDevice, c->stream));
                g_nccl.check(g_nccl.AllReduce(d, d, 1, NCCL_INT32, NCCL_MAX, c->comm, c->stream), "ncclAllReduce");
                CU(cudaMemcpyAsync(&flag, d, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
                CU(cudaStreamSynchronize(c->stream));
            }
            if (flag) { bad = 1; break; }        // the reference warns and goes on; here the caller decides
        }
    }
    if (nsteps) *nsteps = n;
    if (unstable) *unstable = bad;
    GUARD_END
}
int dgtd_norm2_local(dgtd_ctx *c, double *sumsq)
{
    GUARD_BEGIN
    if (!c || !sumsq) throw Error(DGTD_ERR_ARG, "null argument");
    CU(cudaSetDevice(c->device));
    CU(cudaMemsetAsync(c->scratch.p, 0, sizeof(double), c->stream));
    sumsq_kernel<<<296, 256, 0, c->stream>>>(c->x.p, 6 * c->Nalloc, c->scratch.p);
    c->launches++;
    CU(cudaMemcpyAsync(sumsq, c->scratch.p, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    GUARD_END
}
int dgtd_sample(dgtd_ctx *c, int npts, const int *elem, const double *shape, double *out6)
{
    GUARD_BEGIN
    if (!c || npts < 0 || (npts && (!elem || !shape || !out6))) throw Error(DGTD_ERR_ARG, "bad argument");
    if (npts == 0) return DGTD_OK;
    for (int p = 0; p < npts; p++) if (elem[p] < 0 || elem[p] >= c->H.NEloc) throw Error(DGTD_ERR_ARG, "probe element out of range");
    CU(cudaSetDevice(c->device));
    DevBuf<int> &de = c->smp_elem; DevBuf<double> &ds = c->smp_shape, &dout = c->smp_out;
    if (de.n < (size_t)npts) { de.alloc(npts); ds.alloc((size_t)npts * c->H.Np); dout.alloc((size_t)npts * 6); }
    CU(cudaMemcpyAsync(de.p, elem, sizeof(int) * npts, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(ds.p, shape, sizeof(double) * npts * c->H.Np, cudaMemcpyHostToDevice, c->stream));
    if (c->wg) sample_aos_kernel<<<(npts + 127) / 128, 128, 0, c->stream>>>(c->x.p, c->H.Np, npts, de.p, ds.p, c->dev2ref.p, dout.p);
    else sample_kernel<<<(npts + 127) / 128, 128, 0, c->stream>>>(c->x.p, c->Nloc, c->H.Np, npts, de.p, ds.p, dout.p);
    c->launches++;
    CU(cudaMemcpyAsync(out6, dout.p, sizeof(double) * npts * 6, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    GUARD_END
}
struct dgtd_gather {
    dgtd_ctx *c = nullptr;                // nullptr once the context is gone (dgtd_destroy before dgtd_gather_destroy)
    int device = 0;
    std::vector<long long> dofs;          // owned dofs, global numbering, output order
    DevBuf<long long> off;                // offset of component 0 in the device state
    long long cstride = 1;
    DevBuf<double> stage;                 // [6][n]
    cudaStream_t side = nullptr;
    cudaEvent_t ready = nullptr, done = nullptr;
    bool pending = false;
    void release_device()                 // streams, events and buffers live on the context's device
    {
        if (done) { if (pending) cudaEventSynchronize(done); cudaEventDestroy(done); done = nullptr; }
        if (ready) { cudaEventDestroy(ready); ready = nullptr; }
        if (side) { cudaStreamDestroy(side); side = nullptr; }
        off.release(); stage.release(); pending = false;
    }
    ~dgtd_gather() { release_device(); }
};
dgtd_ctx::~dgtd_ctx()
{
    for (dgtd_gather *g : gathers) { g->release_device(); g->c = nullptr; }     // a later dgtd_gather_destroy only frees the shell
    if (p2p && comm) {
        // neighbours store into my halo buffer and I into theirs: (1) everybody's kernels have drained, (2) everybody has
        // closed its mappings of the others' buffers, only then is an exported buffer freed (cudaFree of memory a peer
        // still maps is undefined)
        comm_barrier(this);
        for (void *&b : p2p_peer_base) if (b) { cudaIpcCloseMemHandle(b); b = nullptr; }
        comm_barrier(this);
        p2p_mem.release();
    }
    for (void *b : p2p_peer_base) if (b) cudaIpcCloseMemHandle(b);
    if (comm && g_nccl.CommDestroy) g_nccl.CommDestroy(comm);
    if (own_stream) cudaStreamDestroy(own_stream);
}

int dgtd_gather_create(dgtd_ctx *c, long long n, const long long *dofs, dgtd_gather **out, long long *n_local)
{
    GUARD_BEGIN
    if (!c || !out || n < 0 || (n && !dofs)) throw Error(DGTD_ERR_ARG, "dgtd_gather_create: bad argument");
    CU(cudaSetDevice(c->device));
    const int Np = c->H.Np;
    std::vector<int> g2l((size_t)c->H.NEglob, -1);
    for (int le = 0; le < c->H.NEloc; le++) g2l[(size_t)c->H.elem_gid[le]] = le;
    auto g = std::make_unique<dgtd_gather>();
    g->c = c; g->device = c->device;
    std::vector<long long> off;
    for (long long i = 0; i < n; i++) {
        const long long d = dofs[i];
        if (d < 0 || d >= c->Nglob) throw Error(DGTD_ERR_ARG, "dgtd_gather_create: dof out of range");
        const int le = g2l[(size_t)(d / Np)], node = (int)(d % Np);
        if (le < 0) continue;   // another rank's
        g->dofs.push_back(d);
        if (c->wg) off.push_back(((long long)le * Np + c->WP.ref2dev[node]) * 6);
        else off.push_back((long long)le * Np + node);
    }
    g->cstride = c->wg ? 1 : c->Nloc;
    g->off.upload(off, 1);
    g->stage.alloc(std::max<size_t>(1, 6 * off.size()));
    CU(cudaStreamCreateWithFlags(&g->side, cudaStreamNonBlocking));
    CU(cudaEventCreateWithFlags(&g->ready, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&g->done, cudaEventDisableTiming));
    if (n_local) *n_local = (long long)g->dofs.size();
    c->gathers.push_back(g.get());
    *out = g.release();
    GUARD_END
}
int dgtd_gather_dofs(const dgtd_gather *g, long long *dofs_local)
{
    if (!g || !dofs_local) return fail(DGTD_ERR_ARG, "null argument");
    std::memcpy(dofs_local, g->dofs.data(), g->dofs.size() * sizeof(long long));
    return DGTD_OK;
}
int dgtd_gather_launch(dgtd_ctx *c, dgtd_gather *g, double *host_out)
{
    GUARD_BEGIN
    if (!c || !g || g->c != c || !host_out) throw Error(DGTD_ERR_ARG, "dgtd_gather_launch: bad argument");
    CU(cudaSetDevice(c->device));
    const long long n = (long long)g->dofs.size();
    if (n == 0) return DGTD_OK;
    if (g->pending) CU(cudaStreamWaitEvent(c->stream, g->done, 0));          // the previous snapshot must have left the staging buffer
    gather_kernel<<<(int)std::min<long long>(592, (n + 255) / 256), 256, 0, c->stream>>>(c->x.p, g->off.p, g->cstride, n, g->stage.p);
    c->launches++;
    CU(cudaGetLastError());
    CU(cudaEventRecord(g->ready, c->stream));
    CU(cudaStreamWaitEvent(g->side, g->ready, 0));
    CU(cudaMemcpyAsync(host_out, g->stage.p, sizeof(double) * 6 * (size_t)n, cudaMemcpyDeviceToHost, g->side));
    CU(cudaEventRecord(g->done, g->side));
    g->pending = true;
    GUARD_END
}
int dgtd_gather_wait(dgtd_ctx *c, dgtd_gather *g)
{
    GUARD_BEGIN
    if (!c || !g || g->c != c) throw Error(DGTD_ERR_ARG, "dgtd_gather_wait: bad argument");
    if (g->pending) { CU(cudaEventSynchronize(g->done)); g->pending = false; }
    GUARD_END
}
void dgtd_gather_destroy(dgtd_gather *g)
{
    if (!g) return;
    if (g->c) {                            // context still alive: free the device side now and leave its list
        cudaSetDevice(g->device);
        auto &v = g->c->gathers;
        for (size_t i = 0; i < v.size(); i++) if (v[i] == g) { v.erase(v.begin() + (long)i); break; }
    }
    delete g;
}
int dgtd_mesh_boundary_elements(const dgtd_mesh *m, int n_attr, const int *bdr_attr, long long cap_pairs, int *pairs, long long *n_pairs)
{
    GUARD_BEGIN
    if (!m || !n_pairs || n_attr < 0 || (n_attr && !bdr_attr)) throw Error(DGTD_ERR_ARG, "dgtd_mesh_boundary_elements: bad argument");
    std::vector<int> attrs(bdr_attr, bdr_attr + n_attr);
    std::vector<int> pr = boundary_element_faces(m->m, attrs);
    *n_pairs = (long long)pr.size() / 2;
    if (pairs) {
        if ((long long)pr.size() / 2 > cap_pairs) throw Error(DGTD_ERR_ARG, "buffer too small");
        std::memcpy(pairs, pr.data(), pr.size() * sizeof(int));
    }
    GUARD_END
}
int dgtd_synchronize(dgtd_ctx *c)
{
    GUARD_BEGIN
    if (!c) throw Error(DGTD_ERR_ARG, "null context");
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->stream));
    p2p_check(c);
    GUARD_END
}
long long dgtd_launch_count(const dgtd_ctx *c) { return c ? c->launches : 0; }
int dgtd_kernel_info(const dgtd_ctx *c, char *buf, int cap)
{
    if (!c || !buf || cap < 1) return fail(DGTD_ERR_ARG, "bad argument");
    char tmp[320];
    if (c->wh)
        std::snprintf(tmp, sizeof tmp, "stage_wh_kernel<P=%d,MODE> DMMA m8n8k4 transposed, warp per group of 4 elements (rows = element x field), aos layout, %d threads, %zu B smem, grid %d%s",
                      c->H.p, c->wgs.threads, c->wgs.smem, c->grid, c->nranks == 1 ? "" : c->p2p ? ", halo: fused peer-memory stores" : ", halo: NCCL send/recv");
    else if (c->wg)
        std::snprintf(tmp, sizeof tmp, "stage_wg_kernel<P=%d,MODE> DMMA m8n8k4 transposed, warp per group of 8 elements, aos layout, %d threads, %zu B smem, grid %d%s",
                      c->H.p, c->wgs.threads, c->wgs.smem, c->grid, c->nranks == 1 ? "" : c->p2p ? ", halo: fused peer-memory stores" : ", halo: NCCL send/recv");
    else
        std::snprintf(tmp, sizeof tmp, "stage_kernel<DIM=%d,P=%d,MODE> generic, %d threads, %zu B smem, grid %d", c->H.dim, c->H.p, c->ks.threads, c->ks.smem, c->grid);
    std::snprintf(buf, (size_t)cap, "%s", tmp);
    return DGTD_OK;
}

// Host-only diagnostic: the flat operator tables a rank would upload (no CUDA involved, no compute).
int dgtd_setup_query(const dgtd_mesh *mesh, const dgtd_options *o, const char *name, void *buf, long long cap_bytes, long long *size_bytes)
{
    GUARD_BEGIN
    if (!mesh || !o || !name || !size_bytes) throw Error(DGTD_ERR_ARG, "null argument");
    Options op = options_from_c(mesh, o);
    HostOp H = build_host_op(mesh->m, op);
    const std::string n = name;
    const void *src = nullptr; size_t bytes = 0;
    std::vector<int> dims;
    std::vector<double> xyz;
    if (n == "D") { src = H.ref.D.data(); bytes = H.ref.D.size() * 8; }
    else if (n == "lift") { src = H.ref.lift.data(); bytes = H.ref.lift.size() * 8; }
    else if (n == "nodes") { src = H.ref.nodes.data(); bytes = H.ref.nodes.size() * 8; }
    else if (n == "fnodes") { src = H.ref.fnodes.data(); bytes = H.ref.fnodes.size() * 4; }
    else if (n == "geo") { src = H.geo.data(); bytes = H.geo.size() * 8; }
    else if (n == "finfo") { src = H.finfo.data(); bytes = H.finfo.size() * 4; }
    else if (n == "ftab") { src = H.ftab.data(); bytes = H.ftab.size(); }
    else if (n == "elem_gid") { src = H.elem_gid.data(); bytes = H.elem_gid.size() * 4; }
    else if (n == "tfsf_xyz") { src = H.tfsf_xyz.data(); bytes = H.tfsf_xyz.size() * 8; }
    else if (n == "gate_xyz") { src = H.gate_xyz.data(); bytes = H.gate_xyz.size() * 8; }
    else if (n == "tfsf_side") { src = H.tfsf_side.data(); bytes = H.tfsf_side.size() * 4; }
    else if (n == "send_node") { src = H.send_node.data(); bytes = H.send_node.size() * 4; }
    else if (n == "peers") { for (auto &p : H.peers) { dims.push_back(p.rank); dims.push_back(p.nfaces); dims.push_back(p.send_off); } src = dims.data(); bytes = dims.size() * 4; }
    else if (n == "peers5") { for (auto &p : H.peers) { dims.push_back(p.rank); dims.push_back(p.nfaces); dims.push_back(p.send_off); dims.push_back(p.remote_off); dims.push_back(p.remote_idx); } src = dims.data(); bytes = dims.size() * 4; }
    else if (n.rfind("wg_", 0) == 0) {   // tables of the warp-per-group kernel's plan (tetrahedra)
        static thread_local WgPlan WP;
        WP = build_wg_plan(H);
        if (n == "wg_hpush") { src = WP.hpush.data(); bytes = WP.hpush.size() * 4; }
        else if (n == "wg_tab") { src = WP.tab.data(); bytes = WP.tab.size(); }
        else if (n == "wg_dims") { dims = {WP.ngroups, WP.NEpad, WP.NT, WP.KSV, WP.nfrag_vol, WP.nfrag_lift, WP.ntab, WG_GEO}; src = dims.data(); bytes = dims.size() * 4; }
        else if (n == "wg_bfrag") { src = WP.bfrag.data(); bytes = WP.bfrag.size() * 8; }
        else if (n == "wg_geo") { src = WP.geo.data(); bytes = WP.geo.size() * 8; }
        else if (n == "wg_forder") { src = WP.forder.data(); bytes = WP.forder.size() * 4; }
        else if (n == "wg_desc") { src = WP.desc.data(); bytes = WP.desc.size() * 4; }
        else if (n == "wg_send_off") { src = WP.send_off.data(); bytes = WP.send_off.size() * 8; }
        else if (n == "wg_dev2ref") { src = WP.dev2ref.data(); bytes = WP.dev2ref.size() * 4; }
        else throw Error(DGTD_ERR_ARG, "unknown setup table " + n);
    }
    else if (n == "dims") { dims = {H.dim, H.p, H.Np, H.Nfp, H.nf, H.NEloc, H.ntab, H.n_tfsf_faces, H.n_halo_faces}; src = dims.data(); bytes = dims.size() * 4; }
    else if (n == "node_coords") { node_coords(mesh->m, H.ref, xyz); src = xyz.data(); bytes = xyz.size() * 8; }
    else throw Error(DGTD_ERR_ARG, "unknown setup table " + n);
    *size_bytes = (long long)bytes;
    if (buf) { if ((long long)bytes > cap_bytes) throw Error(DGTD_ERR_ARG, "buffer too small"); if (bytes) std::memcpy(buf, src, bytes); }
    GUARD_END
}

int dgtd_comm_unique_id(void *id128)
{
    GUARD_BEGIN
    if (!id128) throw Error(DGTD_ERR_ARG, "null argument");
    g_nccl.load();
    NcclId id; g_nccl.check(g_nccl.GetUniqueId(&id), "ncclGetUniqueId");
    std::memcpy(id128, &id, 128);
    GUARD_END
}
int dgtd_comm_init(dgtd_ctx *c, const void *id128)
{
    GUARD_BEGIN
    if (!c || !id128) throw Error(DGTD_ERR_ARG, "null argument");
    if (c->nranks == 1) return DGTD_OK;
    if (c->comm) throw Error(DGTD_ERR_COMM, "dgtd_comm_init: this context already has a communicator");
    CU(cudaSetDevice(c->device));
    g_nccl.load();
    NcclId id; std::memcpy(&id, id128, 128);
    g_nccl.check(g_nccl.CommInitRank(&c->comm, c->nranks, id, c->rank), "ncclCommInitRank");
    p2p_setup(c);
    GUARD_END
}
int dgtd_halo_mode(const dgtd_ctx *c)
{
    if (!c) return -1;
    if (c->nranks == 1) return DGTD_HALO_NONE;
    return c->p2p ? DGTD_HALO_P2P : DGTD_HALO_NCCL;
}
int dgtd_halo_bytes(const dgtd_ctx *c, long long *bytes)
{
    if (!c || !bytes) return fail(DGTD_ERR_ARG, "null argument");
    *bytes = (long long)6 * c->H.n_halo_faces * c->H.Nfp * 8;
    return DGTD_OK;
}

}  // extern "C"
