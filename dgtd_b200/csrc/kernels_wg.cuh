// sm_100a warp-per-group DMMA stage kernel (tetrahedra): one warp carries a group of 8 elements through the whole fused
// stage — covariant transform, volume contraction, face flux, LIFT, push-forward, Runge-Kutta update — in registers.
//
// Why: kernels that hand intermediates (covariant field, flux, partial results) from warp to warp through
// shared memory put the LSU pipe at 54 % next to an FP64 pipe at 61 % and left both waiting on hand-over barriers
// (round-1 measurements, 70-81 G DOF-updates/s).  Here the contraction is TRANSPOSED: the DMMA A operand is the data
// [8 elements x 4 nodes] and the B operand the operator [4 nodes x 8 output nodes], so lane (e = l>>2, j = l&3) owns
// element e as A-row, as accumulator row and in the element-wise phases:
//   volume   lane reads the node record (e, 4ks + j) of y_in, forms u~ = J^T u / det J in registers, feeds it as A;
//   flux     lane = (element e, face j): per step s it forms the contravariant flux at face node s of its face and feeds
//            it as A of the LIFT contraction (K = 4 faces x Nfp nodes, k-lane j = face j: geometry set up once per lane);
//   epilogue lane holds k~ of (e, nodes 8nt + j, 8nt + j + 4) for all six components: push-forward with J, material,
//            RK update on the x / z records, in place in shared memory.
// Warps never synchronise with each other: a warp has its own three group buffers (y_in, x -> y_out, z -> z_new), filled
// and drained by bulk-TMA copies on per-warp mbarriers, and its own geometry/descriptor slice.  The operator B fragments
// (19 KB at order 3) sit in shared memory once per CTA.  Neighbour traces of faces leaving the group are read straight from
// y_in in global memory (L2) through the generic address path, prefetched one step ahead.
// State layout "aos" (host.hpp, WgPlan): offset(e, n, c) = (e * Np + n) * 6 + c.
// Reference semantics: src/evolution/HesthavenEvolution.cpp:450-542 with the `global` operator's coefficients
// (src/components/DGOperatorFactory.h:469-573, 1268-1361), external/mfem-geg/linalg/ode.cpp:109-136.
#pragma once
#include "kernels.cuh"
#include "host.hpp"
#include "ptx.cuh"

namespace dgtd {

// Direct halo exchange over NVLink peer memory (one process per GPU, buffers mapped with CUDA IPC): the stage kernel that
// PRODUCES y_out stores the traces of its partition faces straight into the neighbour rank's halo buffer, and the kernel
// that CONSUMES them waits, only in the lanes that own a partition face, for that face's flag.
// Per-face handshake: exchange number k uses halo buffer k & 1 and flag array k & 1 on every rank.  The warp that owns a
// face stores its Nfp records (lane p the p-th 16-byte piece), synchronises, and the owning lane writes that face's
// flag = k with a release store AFTER them (data and flag are ordered inside one warp by bar.warp.sync + the lane's
// release; no ordering between different SMs is relied upon); the lane of the neighbour that owns the same face polls
// the flag (relaxed polls, then an acquire load) before it reads the records.  Reuse of a buffer is safe pairwise:
// a lane pushes exchange k+1 of a face only after it has read the neighbour's exchange k of that face, and the neighbour
// wrote exchange k after reading my exchange k-1 — the data k+1 overwrites.  No grid-wide election, no counters.
// Round 2 history: one flag per peer raised by the last CTA of the launch (round 1) and per-peer arrival
// counters both gave N-rank results that differed from the single-GPU run in a handful of elements per step at 786 K
// elements per rank (bench.py's parity check), although small fixtures and 400-step stress runs passed.
// Replaces GlobalEvolution.cpp:763-774 (six blocking MPI exchanges of whole neighbour elements per Mult).
constexpr int P2P_MAXPEERS = 8;
struct WgP2P {
    const int2 *hpush;                                   // WgPlan::hpush
    double *peer_out[P2P_MAXPEERS];                      // peer's halo buffer of the exchange being produced
    unsigned long long *peer_flag[P2P_MAXPEERS];         // peer's per-face flags of that exchange (indexed by the slot on the peer)
    const unsigned long long *flags;                     // my per-face flags of the exchange being consumed (indexed by my halo slot)
    unsigned long long wait_epoch, signal_epoch;         // 0: nothing to wait for / nothing to produce
    int *err;                                            // set when a wait timed out
};
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long ld_relaxed_sys(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long globaltimer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
// wait until the neighbour has delivered exchange wait_epoch of my halo face `slot`: relaxed polls (no L1 invalidation per
// poll), then one acquire load; 20 s without progress sets *err and gives up
__device__ __forceinline__ void p2p_wait_face(const WgP2P &pp, int slot)
{
    const unsigned long long *f = pp.flags + slot;
    if (ld_relaxed_sys(f) < pp.wait_epoch) {
        const unsigned long long t0 = globaltimer_ns();
        while (ld_relaxed_sys(f) < pp.wait_epoch) {
            if (globaltimer_ns() - t0 > 20000000000ull) { atomicExch(pp.err, 1); break; }
        }
    }
    (void)ld_acquire_sys(f);
}

struct WgArgs {
    const double *bfrag;      // WgPlan::bfrag
    const double *geo;        // [NEpad][32]
    const int *desc;          // [NEpad][4] int2
    const uint8_t *tab;       // [ntab][16]
    int ntab;
    const double *tfsf_xyz;
    const double *gate;
    const double *halo;       // [haloFace][Nfp][6]
    unsigned int *work, *work_next;   // group counter of this launch (groups beyond the first wave), and the next launch's (zeroed here)
    int ngroups;
    int has_sigma;
    double alpha;
    DevPlaneWave pw;
    int pw_on;
    const double *yin, *x;    // aos layout
    double *z, *yout;
    double a, b, t;
    WgP2P pp;
};

// Warps per SM: the three group buffers of a warp (y_in, x -> y_out, z -> z_new: 23 KB at order 3) and 222 registers put
// 8 warps on an SM at order <= 3 and 4 at order 4.  Two ways to a third warp per scheduler were built and measured in
// round 2 (both parity-green, both slower: profiles/r2_warp_count_experiments.txt): x / z through per-lane
// global accesses instead of staging (12 warps, 91 G vs 120 G) and two buffers with x loaded over y_in after the flux
// (12 warps at 168 registers, 99-105 G).
template <int P> struct Wg {
    static constexpr int Np = (P + 1) * (P + 2) * (P + 3) / 6, Nfp = (P + 1) * (P + 2) / 2;
    static constexpr int NT = (Np + 7) / 8, KSV = (Np + 3) / 4, VT = (NT - 1) * 3 + 3;
#ifndef DGTD_WG_NW
#define DGTD_WG_NW 8
#endif
    static constexpr int NW = P <= 3 ? DGTD_WG_NW : 4, T = 32 * NW;   // warps per CTA = groups in flight per SM
    static constexpr int GS = Np * BLK_E * 6;                    // doubles per group of one state vector
    static constexpr int NFV = KSV * VT, NFL = Nfp * NT, NFR = NFV + NFL;
    static constexpr int WGEO = BLK_E * WG_GEO, WDESC = BLK_E * 4 * 2;   // doubles / ints per group
    static constexpr int WDBL = 3 * GS + WGEO + WDESC / 2;       // doubles per warp: Y, X, Z, geometry, descriptors
    static constexpr int TABROWS = WG_TABROWS;
    static constexpr int oWarp = NFR * 32;
    static constexpr size_t bTab = (size_t)(oWarp + NW * WDBL) * 8;
    static constexpr size_t bBar = bTab + (size_t)TABROWS * 16;
    static constexpr size_t smem_bytes = bBar + (size_t)(NW * 3 + 1) * 8;     // per-warp barriers (y, x, z) + one for the operator fragments
    static_assert(Np - 8 * (NT - 1) <= 4, "mixed last tile");
    static_assert((GS % 2) == 0 && (bTab % 16) == 0, "alignment");
};

__device__ __forceinline__ int tab_byte(const uint4 &r, int s)
{
    const uint32_t w = (s >> 2) == 0 ? r.x : (s >> 2) == 1 ? r.y : (s >> 2) == 2 ? r.z : r.w;
    return (int)((w >> (8 * (s & 3))) & 0xffu);
}
// 48-byte node record through the generic address path (shared or global)
__device__ __forceinline__ void load_rec(const double *p, double *u)
{
    const double2 *q = reinterpret_cast<const double2 *>(p);
    const double2 a = q[0], b = q[1], c = q[2];
    u[0] = a.x; u[1] = a.y; u[2] = b.x; u[3] = b.y; u[4] = c.x; u[5] = c.y;
}
// neighbour record whose address space is known per lane: a generic LD whose lanes fall partly into shared memory costs
// ~8.5 shared-memory wavefronts per instruction against 2 for the same lanes through LDS (profiles/r1_final_stage_wg_ncu_summary.txt)
__device__ __forceinline__ void load_rec_split(const double *p, bool in_smem, double *u)
{
    if (in_smem) {
        const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
        asm volatile("ld.shared.v2.f64 {%0,%1}, [%6];\n\tld.shared.v2.f64 {%2,%3}, [%6+16];\n\tld.shared.v2.f64 {%4,%5}, [%6+32];"
                     : "=d"(u[0]), "=d"(u[1]), "=d"(u[2]), "=d"(u[3]), "=d"(u[4]), "=d"(u[5]) : "r"(a));
    } else {
        asm volatile("ld.global.v2.f64 {%0,%1}, [%6];\n\tld.global.v2.f64 {%2,%3}, [%6+16];\n\tld.global.v2.f64 {%4,%5}, [%6+32];"
                     : "=d"(u[0]), "=d"(u[1]), "=d"(u[2]), "=d"(u[3]), "=d"(u[4]), "=d"(u[5]) : "l"(p));
    }
}
__device__ __forceinline__ void store_rec(double *p, const double *u)
{
    double2 *q = reinterpret_cast<double2 *>(p);
    q[0] = make_double2(u[0], u[1]); q[1] = make_double2(u[2], u[3]); q[2] = make_double2(u[4], u[5]);
}

// The first PF neighbour records of a face are requested BEFORE the volume contraction (their L2 latency hides behind it)
// and the prefetch then runs PF face steps ahead; the flux is one 6x6 map per (element, face):
//   F~_E = Ah (dH - na x dE),  F~_H = -Ah (dE + na x dH),  Ah = J^-1 [fs n x],  na = alpha n.
// Measured alternatives (profiles/): one-step prefetch issued inside the face loop with the flux in physical
// components, 97.9 G vs 110 G; asm-volatile (pinned) prefetch loads, +-0; prefetch.global.L1 of the records, -9 %.
// TF: the context has a TF/SF plane-wave source (the injection code sits in the face loop only then).
template <int P, int MODE, bool TF>
__global__ void __launch_bounds__(Wg<P>::T, 1) stage_wg_kernel(const WgArgs A)
{
    using B = Wg<P>;
    constexpr int Np = B::Np, Nfp = B::Nfp, NT = B::NT, KSV = B::KSV, VT = B::VT, GS = B::GS;
    constexpr int NL = Np - 8 * (NT - 1);                                          // nodes of the mixed tile
#ifndef DGTD_WG_PF
#define DGTD_WG_PF 3
#endif
    constexpr int PF = P >= 4 ? 1 : DGTD_WG_PF;                                    // neighbour-record prefetch distance (face steps)
    constexpr bool LOAD_X = MODE == MODE_STAGE1 || MODE == MODE_STAGE23;           // stage 1: x == y_in, fetched again (L2 hit)
    constexpr bool LOAD_Z = MODE == MODE_STAGE23 || MODE == MODE_STAGE4;
    constexpr bool STORE_X = MODE != MODE_STAGE4, STORE_Z = MODE != MODE_MULT;      // stage 4 forms the new x in the z buffer
    extern __shared__ __align__(128) unsigned char smem_wg[];
    double *sm = reinterpret_cast<double *>(smem_wg);
    const double *sFragV = sm, *sFragL = sm + B::NFV * 32;
    const uint4 *sTab = reinterpret_cast<const uint4 *>(smem_wg + B::bTab);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, e = lane >> 2, j = lane & 3;
    double *wY = sm + B::oWarp + warp * B::WDBL, *wX = wY + GS, *wZ = wX + GS, *wGeo = wZ + GS;
    const int2 *wDesc = reinterpret_cast<const int2 *>(wGeo + B::WGEO);
    uint64_t *barY = reinterpret_cast<uint64_t *>(smem_wg + B::bBar) + 3 * warp, *barX = barY + 1, *barZ = barY + 2;

    uint64_t *barF = reinterpret_cast<uint64_t *>(smem_wg + B::bBar) + 3 * B::NW;   // operator fragments: one bulk copy per CTA
    {
        uint4 *dst = reinterpret_cast<uint4 *>(smem_wg + B::bTab);
        const uint4 *src = reinterpret_cast<const uint4 *>(A.tab);
        for (int i = tid; i < min(A.ntab, B::TABROWS); i += B::T) dst[i] = src[i];
    }
    if (tid == 0) mbar_init(barF, 1);
    if (lane == 0) { mbar_init(barY, 1); mbar_init(barX, 1); mbar_init(barZ, 1); }
    if (tid == 0) asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    fence_async_smem();
    __syncthreads();

    const int gstride = gridDim.x * B::NW;
    int g = blockIdx.x * B::NW + warp;
    const bool has_work = g < A.ngroups;

    const uint4 ownrow = sTab[j];
    const bool inject = A.pw_on && (A.gate == nullptr || *A.gate >= 1e-16);
    // flag of the partition face this lane pushed in an earlier group, not raised yet: the release store waits for the
    // lane's peer stores to be performed (an NVLink round trip), so it is issued one group later, when they long are
    unsigned long long *pend_flag = nullptr;

    auto issue_y = [&](int gg) {
        mbar_expect_tx(barY, (uint32_t)(GS * 8 + B::WGEO * 8 + B::WDESC * 4));
        bulk_load(wY, A.yin + (size_t)gg * GS, GS * 8, barY);
        bulk_load(wGeo, A.geo + (size_t)gg * B::WGEO, B::WGEO * 8, barY);
        bulk_load(wGeo + B::WGEO, A.desc + (size_t)gg * B::WDESC, B::WDESC * 4, barY);
    };
    auto issue_x = [&](int gg) {
        mbar_expect_tx(barX, (uint32_t)(GS * 8));
        bulk_load(wX, (MODE == MODE_STAGE1 ? A.yin : A.x) + (size_t)gg * GS, GS * 8, barX);
    };
    auto issue_z = [&](int gg) {
        mbar_expect_tx(barZ, (uint32_t)(GS * 8));
        bulk_load(wZ, A.z + (size_t)gg * GS, GS * 8, barZ);
    };
    if (tid == 0) { mbar_expect_tx(barF, (uint32_t)(B::NFR * 32 * 8)); bulk_load(sm, A.bfrag, B::NFR * 32 * 8, barF); }
    if (lane == 0 && has_work) { issue_y(g); if (LOAD_X) issue_x(g); if (LOAD_Z) issue_z(g); }
    mbar_wait(barF, 0);

    // Groups beyond the first wave are handed out by a counter, not by a fixed stride: a warp that loses time (partition
    // faces: waits, NVLink stores; boundary groups) then simply takes fewer groups, instead of every launch ending with
    // the slowest warp's fixed share; the partial last wave of a static split (20.76 groups per warp at the bench size)
    // disappears too.  The next group is drawn at the top of an iteration, a whole group ahead of its use.
    if (A.work && blockIdx.x == 0 && tid == 0) *A.work_next = 0;
    for (int it = 0; g < A.ngroups; it++) {
        const uint32_t par = it & 1;
        int gnext = 0;
        if (A.work) {
            if (lane == 0) gnext = gstride + (int)atomicAdd(A.work, 1u);
            gnext = __shfl_sync(0xffffffffu, gnext, 0);
        } else gnext = g + gstride;                     // static split (single-rank contexts at order <= 3: measured 1 % faster there)
        if (gnext >= A.ngroups) gnext = -1;
        const double *ge = wGeo + e * WG_GEO;
        const double *yrec = wY + e * Np * 6;
        mbar_wait(barY, par);

        double acc[6][NT][2];
#pragma unroll
        for (int c = 0; c < 6; c++)
#pragma unroll
            for (int nt = 0; nt < NT; nt++) acc[c][nt][0] = acc[c][nt][1] = 0.0;
        // The mixed tile (nt = NT-1) of field fo keeps, for input component x, the accumulator acc[3 fo + x][NT-1]:
        // element 0 -> k~_{x+1}, element 1 -> k~_{x+2} (volume); the LIFT of component c adds to acc[3 fo + (c+2)%3][NT-1][0].

        // ---------------- face (e, j): where the exterior trace comes from; first neighbour records requested now ----------
        const int2 info = wDesc[e * 4 + j];
        const int code = info.y;
        const double *nbase = yrec;
        uint4 nrow = ownrow;
        double ce = 0.0, ch = 0.0, al = A.alpha;
        bool nb_smem = true;                                  // the exterior trace lies in this warp's own y_in buffer
        if (info.x >= 0) {
            nrow = sTab[(code >> FI_TAB_SHIFT) & FI_TAB_MASK];
            nb_smem = (info.x >> 3) == g;
            nbase = nb_smem ? wY + (info.x & 7) * Np * 6 : A.yin + (size_t)info.x * Np * 6;
        } else if (info.x == -1) {
            const int bc = code & FI_BC_MASK;
            ce = bc == 1 ? -2.0 : bc == 3 ? -1.0 : 0.0;
            ch = bc == 2 ? -2.0 : bc == 3 ? -1.0 : 0.0;
            if (bc == 3) al = 1.0;
        } else {
            nrow = sTab[4 + j];
            nb_smem = false;
            nbase = A.halo + (size_t)(-2 - info.x) * Nfp * 6;
        }
        if (A.pp.wait_epoch != 0 && __any_sync(0xffffffffu, info.x < -1)) {      // the neighbour's traces of my partition faces
            if (info.x < -1) p2p_wait_face(A.pp, -2 - info.x);
            __syncwarp();
        }
        double uQ[PF + 1][6];
#pragma unroll
        for (int q = 0; q < PF; q++) load_rec_split(nbase + tab_byte(nrow, q) * 6, nb_smem, uQ[q]);

        // ---------------- volume: k~E_c = D_{c+1} u~H_{c+2} - D_{c+2} u~H_{c+1},  k~H likewise from u~E = -J^T E / det ------
        {
            double jm[9];
#pragma unroll
            for (int i = 0; i < 9; i++) jm[i] = ge[i];
#pragma unroll
            for (int ks = 0; ks < KSV; ks++) {
                const int node = 4 * ks + j;
                double u[6] = {0, 0, 0, 0, 0, 0};
                if (4 * ks + 3 < Np || node < Np) load_rec(yrec + node * 6, u);
                double ut[6];     // [0..2] u~E = -(J/det)^T E, [3..5] u~H = (J/det)^T H   (jm = J / det J, WgPlan::geo)
#pragma unroll
                for (int a = 0; a < 3; a++) {
                    ut[a] = -fma(jm[a], u[0], fma(jm[3 + a], u[1], jm[6 + a] * u[2]));
                    ut[3 + a] = fma(jm[a], u[3], fma(jm[3 + a], u[4], jm[6 + a] * u[5]));
                }
                const double *fr = sFragV + (ks * VT) * 32 + lane;
#pragma unroll
                for (int nt = 0; nt < NT - 1; nt++)
#pragma unroll
                    for (int d = 0; d < 3; d++) {
                        const double bv = fr[(nt * 3 + d) * 32];
                        const int cp = (d + 2) % 3, cm = (d + 1) % 3;      // acc[cp] += D_d u~[d+1] ; acc[cm] -= D_d u~[d+2]
                        dmma884(acc[cp][nt][0], acc[cp][nt][1], ut[3 + (d + 1) % 3], bv);
                        dmma884(acc[cm][nt][0], acc[cm][nt][1], -ut[3 + (d + 2) % 3], bv);
                        dmma884(acc[3 + cp][nt][0], acc[3 + cp][nt][1], ut[(d + 1) % 3], bv);
                        dmma884(acc[3 + cm][nt][0], acc[3 + cm][nt][1], -ut[(d + 2) % 3], bv);
                    }
#pragma unroll
                for (int x = 0; x < 3; x++) {
                    const double bv = fr[((NT - 1) * 3 + x) * 32];
                    dmma884(acc[x][NT - 1][0], acc[x][NT - 1][1], ut[3 + x], bv);
                    dmma884(acc[3 + x][NT - 1][0], acc[3 + x][NT - 1][1], ut[x], bv);
                }
            }
        }
        // x / z of this group: the previous group's stores have long drained the buffers
        if ((LOAD_X || LOAD_Z) && it > 0 && lane == 0) {
            bulk_wait_read();
            if (LOAD_X) issue_x(g);
            if (LOAD_Z) issue_z(g);
        }

        // ---------------- face flux of (element e, face j) -> LIFT --------------------------------------------------------
        {
            const int tf = (code >> FI_TFSF_SHIFT) & FI_TFSF_MASK;
            // dU = u+ - u- + c u-  =  u+ - (1 - c) u-   (boundary faces read u+ = u-; 1 - c is 1, 2 or 3: exact)
            const double se1 = 1.0 - ce, sh1 = 1.0 - ch;
            double ji[9];
#pragma unroll
            for (int i = 0; i < 9; i++) ji[i] = ge[9 + i];
            double gn[3];
#pragma unroll
            for (int d = 0; d < 3; d++) gn[d] = j == 0 ? (ji[d] + ji[3 + d]) + ji[6 + d] : -ji[3 * (j - 1) + d];   // -grad lambda_j: outward
            // F_E = g x dH + alpha fs (dE - n (n.dE)) = g x (dH - na x dE),  F_H = -g x (dE + na x dH)   with g = fs n, na = alpha n:
            // one cross product with na and one 3x3 map Ah = J^-1 [g x] per field (30 FMA per face node instead of the 36 of
            // two 3x3 maps per field; 12 instead of 18 doubles of per-face set-up; 1 / fs comes with the geometry record)
            const double ans = al * ge[26 + j];                      // alpha / fs
            const double na0 = ans * gn[0], na1 = ans * gn[1], na2 = ans * gn[2];
            double Ah[9];
#pragma unroll
            for (int a = 0; a < 3; a++) {
                const double j0 = ji[3 * a], j1 = ji[3 * a + 1], j2 = ji[3 * a + 2];
                Ah[3 * a + 0] = j1 * gn[2] - j2 * gn[1];
                Ah[3 * a + 1] = j2 * gn[0] - j0 * gn[2];
                Ah[3 * a + 2] = j0 * gn[1] - j1 * gn[0];
            }
#pragma unroll
            for (int s = 0; s < Nfp; s++) {
                double uM[6], dU[6];
                const double *uP = uQ[s % (PF + 1)];
                load_rec(yrec + tab_byte(ownrow, s) * 6, uM);
                if (s + PF < Nfp) load_rec_split(nbase + tab_byte(nrow, s + PF) * 6, nb_smem, uQ[(s + PF) % (PF + 1)]);
#pragma unroll
                for (int c = 0; c < 3; c++) { dU[c] = fma(-se1, uM[c], uP[c]); dU[3 + c] = fma(-sh1, uM[3 + c], uP[3 + c]); }   // u+ - u- (+ c u-)
                if (TF && tf && inject) {
                    double inc[6];
                    const int m = tab_byte(sTab[4 + j], s);
                    planewave6(A.pw, A.tfsf_xyz + ((long long)(code >> FI_TIDX_SHIFT) * Nfp + m) * 3, A.t, inc);
                    const double sg = tf == 1 ? 1.0 : -1.0;
#pragma unroll
                    for (int c = 0; c < 6; c++) dU[c] += sg * inc[c];
                }
                double wE[3], wH[3];
                wE[0] = fma(na2, dU[1], fma(-na1, dU[2], dU[3]));    // dH - na x dE
                wE[1] = fma(na0, dU[2], fma(-na2, dU[0], dU[4]));
                wE[2] = fma(na1, dU[0], fma(-na0, dU[1], dU[5]));
                wH[0] = fma(-na2, dU[4], fma(na1, dU[5], dU[0]));    // dE + na x dH
                wH[1] = fma(-na0, dU[5], fma(na2, dU[3], dU[1]));
                wH[2] = fma(-na1, dU[3], fma(na0, dU[4], dU[2]));
                double ft[6];
#pragma unroll
                for (int a = 0; a < 3; a++) {
                    ft[a] = fma(Ah[3 * a + 2], wE[2], fma(Ah[3 * a + 1], wE[1], Ah[3 * a] * wE[0]));
                    ft[3 + a] = -fma(Ah[3 * a + 2], wH[2], fma(Ah[3 * a + 1], wH[1], Ah[3 * a] * wH[0]));
                }
                const double *fr = sFragL + (s * NT) * 32 + lane;
#pragma unroll
                for (int nt = 0; nt < NT - 1; nt++) {
                    const double bv = fr[nt * 32];
#pragma unroll
                    for (int c = 0; c < 6; c++) dmma884(acc[c][nt][0], acc[c][nt][1], ft[c], bv);
                }
                {
                    const double bv = fr[(NT - 1) * 32];
#pragma unroll
                    for (int c = 0; c < 6; c++) {
                        const int xa = 3 * (c / 3) + (c % 3 + 2) % 3;
                        dmma884(acc[xa][NT - 1][0], acc[xa][NT - 1][1], ft[c], bv);
                    }
                }
            }
        }

        // ---------------- push forward, material, Runge-Kutta stage ---------------------------------------------------------
        double jm[9];
#pragma unroll
        for (int i = 0; i < 9; i++) jm[i] = ge[i];
        const double de = ge[23], dm = ge[24], se = ge[25];          // det/eps, det/mu, sigma/eps  (jm = J / det)
        const double ae = A.a * de, am = A.a * dm, be = A.b * de, bm = A.b * dm;
        const bool keep_y = A.has_sigma != 0;        // the conductivity term reads E of y_in in the epilogue
        if (!keep_y) {
            __syncwarp();                            // every lane has read y_in for the last time
            if (lane == 0 && gnext >= 0) issue_y(gnext);
        }
        if (LOAD_X) mbar_wait(barX, par);
        if (LOAD_Z) mbar_wait(barZ, par);
        // node slots of lane (e, j): nodes 8 nt + j and 8 nt + j + 4 of the full tiles, then node 8 (NT-1) + j of the mixed one
        constexpr int NS = 2 * (NT - 1) + 1;
        auto slot_node = [&](int sl) { return sl < NS - 1 ? 8 * (sl >> 1) + j + 4 * (sl & 1) : 8 * (NT - 1) + j; };
        auto slot_valid = [&](int sl) { return sl < NS - 1 || NL == 4 || j < NL; };
#pragma unroll
        for (int sl = 0; sl < NS; sl++) {
            const int nt = sl < NS - 1 ? sl >> 1 : NT - 1, h = sl < NS - 1 ? sl & 1 : 0;
            const int node = slot_node(sl);
            if (!slot_valid(sl)) continue;
            double kr[6];
            if (nt < NT - 1) {
#pragma unroll
                for (int c = 0; c < 6; c++) kr[c] = acc[c][nt][h];
            } else {
#pragma unroll
                for (int c = 0; c < 6; c++) {
                    const int f3 = 3 * (c / 3), cc = c % 3;
                    kr[c] = acc[f3 + (cc + 2) % 3][NT - 1][0] + acc[f3 + (cc + 1) % 3][NT - 1][1];
                }
            }
            double k[6];     // (J/det) k~ : the material factor det/eps, det/mu joins the Runge-Kutta coefficient below
#pragma unroll
            for (int d = 0; d < 3; d++) {
                k[d] = fma(jm[3 * d], kr[0], fma(jm[3 * d + 1], kr[1], jm[3 * d + 2] * kr[2]));
                k[3 + d] = fma(jm[3 * d], kr[3], fma(jm[3 * d + 1], kr[4], jm[3 * d + 2] * kr[5]));
            }
            const int off = (e * Np + node) * 6;
            double ca[6] = {ae, ae, ae, am, am, am}, cb[6] = {be, be, be, bm, bm, bm};
            if (keep_y || MODE == MODE_MULT) {       // the plain k is needed: conductivity term, or Mult's output
                double uo[6] = {0, 0, 0, 0, 0, 0};
                if (keep_y) load_rec(wY + off, uo);
#pragma unroll
                for (int d = 0; d < 3; d++) { k[d] = fma(de, k[d], -(se * uo[d])); k[3 + d] *= dm; }
#pragma unroll
                for (int c = 0; c < 6; c++) { ca[c] = A.a; cb[c] = A.b; }
            }
            double xv[6], zv[6], o[6], zn[6];
            if (LOAD_X) load_rec(wX + off, xv);
            if (LOAD_Z) load_rec(wZ + off, zv);
#pragma unroll
            for (int c = 0; c < 6; c++) {
                if (MODE == MODE_MULT) o[c] = k[c];
                else if (MODE == MODE_STAGE1) { o[c] = fma(ca[c], k[c], xv[c]); zn[c] = fma(cb[c], k[c], xv[c]); }
                else if (MODE == MODE_STAGE23) { o[c] = fma(ca[c], k[c], xv[c]); zn[c] = fma(cb[c], k[c], zv[c]); }
                else zn[c] = fma(cb[c], k[c], zv[c]);          // stage 4: new x, formed in the z buffer
            }
            if (STORE_X) store_rec(wX + off, o);
            if (STORE_Z) store_rec(wZ + off, zn);
        }
        fence_async_smem();
        __syncwarp();
        if (MODE != MODE_MULT && A.pp.signal_epoch != 0 && __any_sync(0xffffffffu, pend_flag != nullptr)) {
            if (pend_flag) { st_release_sys(pend_flag, A.pp.signal_epoch); pend_flag = nullptr; }
        }
        if (MODE != MODE_MULT && A.pp.signal_epoch != 0) {   // my traces of the new stage vector -> the peers' halo buffers
            // The warp stores one face at a time: its Nfp records are 3 Nfp consecutive 16-byte pieces in the receiver's
            // node order, piece p by lane p -> one store instruction covers the face's 48 Nfp contiguous bytes (full NVLink
            // write packets; lane = face with 3 Nfp stores of 16 scattered bytes each costs ten times the requests).
            unsigned pm = __ballot_sync(0xffffffffu, info.x < -1);
            int2 hp = make_int2(0, 0);
            if (info.x < -1) hp = A.pp.hpush[-2 - info.x];
            const double *srcb = MODE == MODE_STAGE4 ? wZ : wX;
            while (pm) {
                const int owner = __ffs(pm) - 1;
                pm &= pm - 1;
                const int hx = __shfl_sync(0xffffffffu, hp.x, owner), hy = __shfl_sync(0xffffffffu, hp.y, owner);
                const uint4 prow = sTab[hx >> 8];
                double2 *dst = reinterpret_cast<double2 *>(A.pp.peer_out[hx & 0xff] + (size_t)hy * Nfp * 6);
                const double *src = srcb + (owner >> 2) * Np * 6;
                for (int p = lane; p < 3 * Nfp; p += 32) {
                    const int m = p / 3;
                    dst[p] = *reinterpret_cast<const double2 *>(src + tab_byte(prow, m) * 6 + (p - 3 * m) * 2);
                }
            }
            __syncwarp();                                                                   // every lane's pieces before the owner's flag
            if (info.x < -1) pend_flag = A.pp.peer_flag[hp.x & 0xff] + hp.y;               // raised in the next epilogue, or at the end
        }
        if (lane == 0) {
            const size_t goff = (size_t)g * GS;
            if (STORE_X) bulk_store(A.yout + goff, wX, GS * 8);
            if (MODE == MODE_STAGE4) bulk_store(A.yout + goff, wZ, GS * 8);
            else if (STORE_Z) bulk_store(A.z + goff, wZ, GS * 8);
            bulk_commit();
            if (keep_y && gnext >= 0) issue_y(gnext);
            if (!(LOAD_X || LOAD_Z)) bulk_wait_read();            // the next epilogue writes these buffers again
        }
        __syncwarp();
        g = gnext < 0 ? A.ngroups : gnext;
    }
    if (lane == 0) bulk_wait_all();
    if (MODE != MODE_MULT && pend_flag) st_release_sys(pend_flag, A.pp.signal_epoch);
}

// Stand-alone producer of an exchange (the state came from the host, or Mult was called on a foreign vector): one thread
// per partition face copies the traces of `y` (aos layout) into the neighbour's halo buffer and raises that face's flag.
// Flow control: exchange e overwrites the records the neighbour read for exchange e-2.  The neighbour produces exchange e-1
// of a face only after it has consumed e-2 of it (stream order, or the read-then-push order inside a stage kernel), so the
// thread first waits for ITS OWN flag of that face to reach wait_epoch = e-1 (pp.flags = my flags of parity (e-1) & 1): a
// rank that runs ahead of a slow neighbour cannot clobber traces that neighbour is still consuming.
__global__ void halo_push_kernel(const double *y, const long long *send_off, int nfaces, int Nfp, const WgP2P pp)
{
    for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < nfaces; s += gridDim.x * blockDim.x) {
        if (pp.wait_epoch != 0) p2p_wait_face(pp, s);
        const int2 hp = pp.hpush[s];
        double *dst = pp.peer_out[hp.x & 0xff] + (size_t)hp.y * Nfp * 6;
        for (int m = 0; m < Nfp; m++) {
            double r[6];
            load_rec(y + send_off[(size_t)s * Nfp + m], r);
            store_rec(dst + m * 6, r);
        }
        st_release_sys(pp.peer_flag[hp.x & 0xff] + hp.y, pp.signal_epoch);
    }
}

// halo pack of the NCCL send/recv path: send[s][c] = y[send_off[s] + c]  (48-byte node records, receiver's face-node order)
__global__ void pack_records_kernel(const double *y, const long long *send_off, int ns, double *send)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < ns * 3; i += gridDim.x * blockDim.x) {
        const int s = i / 3, ch = i - 3 * s;
        reinterpret_cast<double2 *>(send)[i] = *reinterpret_cast<const double2 *>(y + send_off[s] + 2 * ch);
    }
}

// ---- layout conversion between the reference layout [6][Nloc] (Fields.h:45-65) and the aos device layout ---------------
// gid != nullptr: `ref` is in the caller's element order and local element e is element gid[e] there (single rank: the
// Morton permutation is applied here instead of in a host loop)
__global__ void to_aos_kernel(const double *ref, long long stride, int Np, long long NEloc, long long NEpad, const int *dev2ref, const int *gid, double *aos)
{
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < NEpad * Np; idx += (long long)gridDim.x * blockDim.x) {
        const long long e = idx / Np; const int n = (int)(idx - e * Np);
        double *o = aos + idx * 6;
        const long long src = (gid && e < NEloc ? (long long)gid[e] : e) * Np + dev2ref[n];
#pragma unroll
        for (int c = 0; c < 6; c++) o[c] = e < NEloc ? ref[c * stride + src] : 0.0;
    }
}
__global__ void from_aos_kernel(const double *aos, long long stride, int Np, long long NEloc, const int *dev2ref, const int *gid, double *ref)
{
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < NEloc * Np; idx += (long long)gridDim.x * blockDim.x) {
        const long long e = idx / Np; const int n = (int)(idx - e * Np);
        const double *o = aos + idx * 6;
        const long long dst = (gid ? (long long)gid[e] : e) * Np + dev2ref[n];
#pragma unroll
        for (int c = 0; c < 6; c++) ref[c * stride + dst] = o[c];
    }
}
__global__ void sample_aos_kernel(const double *y, int Np, int npts, const int *elem, const double *shape, const int *dev2ref, double *out)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npts) return;
    const long long e = elem[p];
    for (int c = 0; c < 6; c++) {
        double s = 0;
        for (int n = 0; n < Np; n++) s = fma(shape[(long long)p * Np + dev2ref[n]], y[(e * Np + n) * 6 + c], s);
        out[p * 6 + c] = s;
    }
}

}  // namespace dgtd
