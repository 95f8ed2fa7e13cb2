// sm_100a DMMA stage kernel of the DG-Maxwell hot path on tetrahedra (FP64 tensor-core contraction).
//
// Same fused stage as kernels.cuh (volume curl + face flux + LIFT + RK update in ONE pass over HBM), re-tiled for the
// machine balance measured on B200 (tools/fp64_peaks.cu: DFMA 34.1 TF/s, DMMA 37.1 TF/s on the same pipe, while a DFMA
// fed from shared memory is LSU-bound at ~1/3 of that, profiles/r1_v1_stage_kernel_ncu_summary.txt):
//   * state layout "blocked": groups of 8 elements, [node][element-in-group][6 components]; a batch of G groups is one
//     contiguous chunk, moved by ONE bulk-TMA copy per vector (UBLKCP) into / out of shared memory;
//   * neighbour traces of faces that leave the batch are gathered one batch ahead with cp.async (48-byte node records);
//   * volume term as a reference-space curl of the covariant field (12 instead of 18 mat-vecs), LIFT on the
//     contravariant flux, both as m8n8k4 DMMA tiles  [nodes x 8 elements]  accumulated in the same registers;
//   * the contravariant result is pushed forward with J, scaled by the material and folded into the RK stage.
// Reference semantics: src/evolution/HesthavenEvolution.cpp:450-542 with the `global` operator's coefficients
// (src/components/DGOperatorFactory.h:469-573, 1268-1361), external/mfem-geg/linalg/ode.cpp:109-136.
#pragma once
#include "kernels.cuh"
#include "host.hpp"

namespace dgtd {

struct MmaArgs {
    const double *afrag;      // DMMA A fragments (BlockedPlan::afrag)
    const double *geo;        // [NEpad][32]
    const int *desc;          // [nbatch][DS] per-batch descriptors: finfo int2[EB*4], tdesc int2[SL], tcount, 3 pad
    const uint8_t *ftab;      // [ntab][Nfp]
    int ntab;
    const double *tfsf_xyz;
    const double *gate;
    const double *halo;       // [haloFace][Nfp][6]
    int nbatch;
    double alpha;
    DevPlaneWave pw;
    int pw_on;
    const double *yin, *x;    // blocked layout
    double *z, *yout;
    double a, b, t;
};

// ---- PTX helpers -------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) { while (!mbar_try_wait(bar, parity)) {} }
// TMA 1-D bulk copies (SASS: UBLKCP)
__device__ __forceinline__ void bulk_load(void *sdst, const void *gsrc, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(sdst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_store(void *gdst, const void *ssrc, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void cp_async16(void *sdst, const void *gsrc)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(sdst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
// D(8x8) += A(8x4) * B(4x8); lane l holds A[l>>2][l&3], B[l&3][l>>2], C[l>>2][2(l&3) + {0,1}]
__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// bank swizzle of the [row][8 elements] operand tiles: a DMMA B-fragment load touches rows 4ks..4ks+3 x elements
// (lane>>2); flipping bit 2 of the element index on rows with bit 1 set spreads a half-warp over all 16 bank pairs
__device__ __forceinline__ int swz8(int row, int e) { return row * BLK_E + (e ^ ((row & 2) << 1)); }

template <int P, int G> struct Blk {
    static constexpr int Np = (P + 1) * (P + 2) * (P + 3) / 6, Nfp = (P + 1) * (P + 2) / 2, NFN = 4 * Nfp;
    static constexpr int MT = (Np + 7) / 8, KSV = (Np + 3) / 4, KH = 4 * KSV, KSL = NFN / 4;
    static constexpr int EB = BLK_E * G, NW = 6 * G, T = 32 * NW;
    static constexpr int GS = Np * BLK_E * 6;                 // doubles per group of one state vector
    static constexpr int SL = 4 * EB;                         // trace slots per batch
    static constexpr int DS = EB * 8 + SL * 2 + 4;            // ints per batch descriptor block (multiple of 4)
    static constexpr int TABROWS = 128;
    static constexpr int GST = BLK_GEO + 2;                   // geometry record stride in shared memory (bank spread over the 8 elements of a group)
    static constexpr bool AREG = 2 * MT * KSV <= 32;          // volume A fragments live in registers (else in shared memory)
    static constexpr int NAV = 3 * MT * KSV * 32, NAL = MT * KSL * 32;
    static constexpr int NA = (AREG ? 0 : NAV) + NAL;
    // shared-memory carve-up (doubles)
    static constexpr int oRaw = 0;                            // [2][G*GS]   stage input as it lies in HBM
    static constexpr int oU = oRaw + 2 * G * GS;              // [G][6][KH][8]  covariant field (E negated) ; later: yout staging
    static constexpr int oF = oU + G * 6 * KH * 8;            // [G][6][NFN][8] contravariant flux ; its head rows: k (in place)
    static constexpr int oX = oF + G * 6 * NFN * 8;           // [G*GS]
    static constexpr int oZ = oX + G * GS;                    // [G*GS]
    static constexpr int oTr = oZ + G * GS;                   // [2][SL][Nfp][6]
    static constexpr int oGeo = oTr + 2 * SL * Nfp * 6;       // [3][EB][GST]
    static constexpr int oA = oGeo + 3 * EB * GST;            // [NA]
    static constexpr int nDoubles = oA + NA;
    static constexpr size_t bDesc = (size_t)nDoubles * 8;     // int [3][DS]
    static constexpr size_t bTab = bDesc + (size_t)3 * DS * 4;
    static constexpr size_t bBar = bTab + (size_t)(TABROWS + 8) * Nfp;   // (128+8)*Nfp is a multiple of 8; rows 128.. : identity row
    static constexpr size_t smem_bytes = bBar + 4 * 8;
    static_assert(6 * KH * 8 >= GS, "yout staging must fit in the U buffer");
    static_assert(NFN >= Np, "k must fit in the head rows of its flux slice");
    static_assert((oA % 2) == 0 && (bDesc % 16) == 0 && (bBar % 8) == 0, "alignment");
};

template <int P, int G, int MODE>
__global__ void __launch_bounds__(Blk<P, G>::T, G == 1 ? 2 : 1) stage_mma_kernel(const MmaArgs A)
{
    using B = Blk<P, G>;
    constexpr int Np = B::Np, Nfp = B::Nfp, NFN = B::NFN, MT = B::MT, KSV = B::KSV, KH = B::KH, KSL = B::KSL;
    constexpr int EB = B::EB, NW = B::NW, T = B::T, GS = B::GS, SL = B::SL, GST = B::GST, DS = B::DS, NQ = (Nfp + 1) / 2;
    constexpr uint32_t BATCH_BYTES = (uint32_t)G * GS * 8;
    constexpr bool LOAD_X = MODE == MODE_STAGE23, LOAD_Z = MODE == MODE_STAGE23 || MODE == MODE_STAGE4;
    constexpr bool STORE_Z = MODE == MODE_STAGE1 || MODE == MODE_STAGE23;
    extern __shared__ __align__(128) unsigned char smem_mma[];
    unsigned char *smem_raw = smem_mma;
    double *sm = reinterpret_cast<double *>(smem_raw);
    double *sU = sm + B::oU, *sF = sm + B::oF, *sX = sm + B::oX, *sZ = sm + B::oZ, *sA = sm + B::oA;
    int *sDesc = reinterpret_cast<int *>(smem_raw + B::bDesc);
    uint8_t *sTab = smem_raw + B::bTab;
    uint8_t *sIdent = sTab + B::TABROWS * Nfp;                           // 0..Nfp-1: trace slots are already in face-node order
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw + B::bBar);   // [0],[1]: stage input buffers ; [2]: x/z

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    {
        const double *asrc = A.afrag + (B::AREG ? B::NAV : 0);
        for (int i = tid; i < B::NA; i += T) sA[i] = asrc[i];
        const int nb = min(A.ntab, B::TABROWS) * Nfp;
        for (int i = tid; i < nb; i += T) sTab[i] = A.ftab[i];
        if (tid < Nfp) sIdent[tid] = (uint8_t)tid;
    }
    if (tid == 0) {
        mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); mbar_init(&bars[2], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        fence_async_smem();
    }
    // this warp's output tile type is fixed: field f (0: dE/dt rows, 1: dH/dt rows), reference component cp
    const int wr = warp % 6, wf = wr / 3, cp = wr - 3 * wf;
    const int c1 = (cp + 2) % 3, x1 = (cp + 1) % 3;
    double aV[B::AREG ? 2 * MT * KSV : 1];
    if (B::AREG) {
#pragma unroll
        for (int mt = 0; mt < MT; mt++)
#pragma unroll
            for (int ks = 0; ks < KSV; ks++) {
                aV[mt * KSV + ks] = __ldg(A.afrag + ((x1 * MT + mt) * KSV + ks) * 32 + lane);
                aV[(MT + mt) * KSV + ks] = __ldg(A.afrag + ((c1 * MT + mt) * KSV + ks) * 32 + lane);
            }
    }
    const bool inject = A.pw_on && (A.gate == nullptr || *A.gate >= 1e-16);

    // stage 1 of the prefetch (two batches ahead): geometry records and the batch descriptor block, plain copies
    auto prefetch_desc = [&](int bb, int ring) {
        const double *gsrc = A.geo + (size_t)bb * EB * BLK_GEO;
        double *gdst = sm + B::oGeo + ring * EB * GST;
        for (int i = tid; i < EB * BLK_GEO / 2; i += T) cp_async16(gdst + (i / (BLK_GEO / 2)) * GST + 2 * (i % (BLK_GEO / 2)), gsrc + 2 * i);
        const int *dsrc = A.desc + (size_t)bb * DS;
        int *ddst = sDesc + ring * DS;
        for (int i = tid; i < DS / 4; i += T) cp_async16(ddst + 4 * i, dsrc + 4 * i);
    };
    // stage 2 (one batch ahead): the stage input by bulk TMA, the traces of faces leaving the batch by cp.async
    auto prefetch_data = [&](int bb, int bf, int ring) {
        if (tid == 0) {
            mbar_expect_tx(&bars[bf], BATCH_BYTES);
            bulk_load(sm + B::oRaw + bf * G * GS, A.yin + (size_t)bb * G * GS, BATCH_BYTES, &bars[bf]);
        }
        const int *dsc = sDesc + ring * DS;
        const int nt = dsc[EB * 8 + SL * 2];
        const int2 *td = reinterpret_cast<const int2 *>(dsc + EB * 8);
        double *tdst = sm + B::oTr + bf * SL * Nfp * 6;
        for (int it = tid; it < nt * Nfp * 3; it += T) {
            const int slot = it / (Nfp * 3), r = it - slot * (Nfp * 3), m = r / 3, ch = r - 3 * m;
            const int2 d = td[slot];
            const double *src;
            if (d.x >= 0) {
                const int nn = sTab[d.y * Nfp + m];
                src = A.yin + (((size_t)(d.x >> 3) * Np + nn) * BLK_E + (d.x & 7)) * 6;
            } else src = A.halo + ((size_t)(-1 - d.x) * Nfp + m) * 6;
            cp_async16(tdst + (slot * Nfp + m) * 6 + 2 * ch, src + 2 * ch);
        }
    };

    int b = blockIdx.x;
    const int gstep = gridDim.x;
    if (b < A.nbatch) prefetch_desc(b, 0);
    cp_async_commit();
    cp_async_wait_all();
    __syncthreads();
    if (b < A.nbatch) prefetch_data(b, 0, 0);
    if (b + gstep < A.nbatch) prefetch_desc(b + gstep, 1);
    cp_async_commit();
    int ring = 0;
    for (int it = 0; b < A.nbatch; b += gstep, it++) {
        const int cur = it & 1;
        const int ring1 = ring == 2 ? 0 : ring + 1, ring2 = ring1 == 2 ? 0 : ring1 + 1;
        const double *raw = sm + B::oRaw + cur * G * GS;
        const double *tr = sm + B::oTr + cur * SL * Nfp * 6;
        const double *geo = sm + B::oGeo + ring * EB * GST;
        const int2 *fi = reinterpret_cast<const int2 *>(sDesc + ring * DS);
        cp_async_wait_all();
        mbar_wait(&bars[cur], (it >> 1) & 1);
        if (tid == 0) bulk_wait_read();          // the previous batch's stores have drained their staging buffers
        __syncthreads();
        if (tid == 0 && (LOAD_X || LOAD_Z)) {
            mbar_expect_tx(&bars[2], BATCH_BYTES * ((LOAD_X ? 1 : 0) + (LOAD_Z ? 1 : 0)));
            if (LOAD_X) bulk_load(sX, A.x + (size_t)b * G * GS, BATCH_BYTES, &bars[2]);
            if (LOAD_Z) bulk_load(sZ, A.z + (size_t)b * G * GS, BATCH_BYTES, &bars[2]);
        }
        if (b + gstep < A.nbatch) prefetch_data(b + gstep, cur ^ 1, ring1);
        if (b + 2 * gstep < A.nbatch) prefetch_desc(b + 2 * gstep, ring2);
        cp_async_commit();

        // ---- face flux: jumps, boundary ghost states, TF/SF, upwind flux, pulled back to reference components -------
        // one thread = (element, face, pair of face nodes): the per-face geometry is set up once, the two nodes give ILP
        for (int item = tid; item < EB * 4 * NQ; item += T) {
            const int e8 = item & 7;
            int r = item >> 3;
            const int q = r % NQ; r /= NQ;
            const int f = r & 3, g = r >> 2, el = g * BLK_E + e8;
            const int2 info = fi[el * 4 + f];
            const int code = info.y;
            const uint8_t *srow = sTab + f * Nfp;
            const double *ubase = raw + (g * Np * BLK_E + e8) * 6;       // + node * 48
            // where the exterior trace comes from: an element of this batch, a prefetched trace slot, or (boundary) the element itself
            const uint8_t *nrow; const double *nbase; int nstride;
            double ce = 0.0, ch = 0.0, al = A.alpha;
            if (info.x >= 0) {
                nrow = sTab + ((code >> FI_TAB_SHIFT) & FI_TAB_MASK) * Nfp;
                nbase = raw + ((info.x >> 3) * Np * BLK_E + (info.x & 7)) * 6; nstride = BLK_E * 6;
            } else if (info.x == -1) {
                // boundary ghost states (HesthavenEvolution.cpp:275-313 with the global operator's SMA)
                const int bc = code & FI_BC_MASK;
                ce = bc == 1 ? -2.0 : bc == 3 ? -1.0 : 0.0;
                ch = bc == 2 ? -2.0 : bc == 3 ? -1.0 : 0.0;
                if (bc == 3) al = 1.0;
                nrow = srow; nbase = ubase; nstride = BLK_E * 6;
            } else {
                nrow = sIdent; nbase = tr + (-2 - info.x) * Nfp * 6; nstride = 6;
            }
            const int tf = (code >> FI_TFSF_SHIFT) & FI_TFSF_MASK;
            const double *ge = geo + el * GST;
            double ji[9];                           // Jinv[a][d] at 3a+d
#pragma unroll
            for (int i = 0; i < 9; i++) ji[i] = ge[9 + i];
            double gn[3];                           // outward normal * fscale = -grad lambda_f
#pragma unroll
            for (int d = 0; d < 3; d++) gn[d] = f == 0 ? (ji[d] + ji[3 + d]) + ji[6 + d] : -ge[9 + 3 * (f - 1) + d];
            const double fs = ge[18 + f];
            const double ifs = 1.0 / fs, ifs2 = ifs * ifs;
            const double af = al * fs;
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const int m = 2 * q + h;
                if (m >= Nfp) break;
                double uM[6], dU[6];
                {
                    const double2 *pu = reinterpret_cast<const double2 *>(ubase + srow[m] * (BLK_E * 6));
                    const double2 *pn = reinterpret_cast<const double2 *>(nbase + nrow[m] * nstride);
                    const double2 v0 = pu[0], v1 = pu[1], v2 = pu[2], w0 = pn[0], w1 = pn[1], w2 = pn[2];
                    uM[0] = v0.x; uM[1] = v0.y; uM[2] = v1.x; uM[3] = v1.y; uM[4] = v2.x; uM[5] = v2.y;
                    dU[0] = fma(ce, uM[0], w0.x - uM[0]); dU[1] = fma(ce, uM[1], w0.y - uM[1]); dU[2] = fma(ce, uM[2], w1.x - uM[2]);
                    dU[3] = fma(ch, uM[3], w1.y - uM[3]); dU[4] = fma(ch, uM[4], w2.x - uM[4]); dU[5] = fma(ch, uM[5], w2.y - uM[5]);
                }
                if (tf && inject) {
                    double inc[6];
                    planewave6(A.pw, A.tfsf_xyz + ((long long)(code >> FI_TIDX_SHIFT) * Nfp + m) * 3, A.t, inc);
                    const double sg = tf == 1 ? 1.0 : -1.0;
#pragma unroll
                    for (int c = 0; c < 6; c++) dU[c] += sg * inc[c];
                }
                const double gdE = (gn[0] * dU[0] + gn[1] * dU[1] + gn[2] * dU[2]) * ifs2;
                const double gdH = (gn[0] * dU[3] + gn[1] * dU[4] + gn[2] * dU[5]) * ifs2;
                double fl[6];   // fscale * ( n x dH + alpha (dE - n (n.dE)) ),  fscale * ( -n x dE + alpha (dH - n (n.dH)) )
                fl[0] = (gn[1] * dU[5] - gn[2] * dU[4]) + af * (dU[0] - gdE * gn[0]);
                fl[1] = (gn[2] * dU[3] - gn[0] * dU[5]) + af * (dU[1] - gdE * gn[1]);
                fl[2] = (gn[0] * dU[4] - gn[1] * dU[3]) + af * (dU[2] - gdE * gn[2]);
                fl[3] = -(gn[1] * dU[2] - gn[2] * dU[1]) + af * (dU[3] - gdH * gn[0]);
                fl[4] = -(gn[2] * dU[0] - gn[0] * dU[2]) + af * (dU[4] - gdH * gn[1]);
                fl[5] = -(gn[0] * dU[1] - gn[1] * dU[0]) + af * (dU[5] - gdH * gn[2]);
                double *pf = sF + (g * 6) * NFN * BLK_E + swz8(f * Nfp + m, e8);
#pragma unroll
                for (int a = 0; a < 3; a++) {
                    pf[a * NFN * BLK_E] = fma(ji[3 * a], fl[0], fma(ji[3 * a + 1], fl[1], ji[3 * a + 2] * fl[2]));
                    pf[(3 + a) * NFN * BLK_E] = fma(ji[3 * a], fl[3], fma(ji[3 * a + 1], fl[4], ji[3 * a + 2] * fl[5]));
                }
            }
        }
        // ---- covariant field  u~ = J^T u / det J  (E stored negated: it only feeds dH/dt = -curl E) ----------------
        for (int item = T - 1 - tid; item < EB * KH; item += T) {
            const int e8 = item & 7, r = item >> 3, j = r % KH, g = r / KH, el = g * BLK_E + e8;
            double *po = sU + (g * 6) * KH * BLK_E + swz8(j, e8);
            if (KH > Np && j >= Np) {   // k-padding rows of the DMMA (the yout staging of the previous batch lay here)
#pragma unroll
                for (int c = 0; c < 6; c++) po[c * KH * BLK_E] = 0.0;
                continue;
            }
            const double2 *pu = reinterpret_cast<const double2 *>(raw + ((g * Np + j) * BLK_E + e8) * 6);
            const double2 v0 = pu[0], v1 = pu[1], v2 = pu[2];
            const double *ge = geo + el * GST;   // J[d][a] at 3d+a
            const double idet = ge[22], nidet = -idet;
#pragma unroll
            for (int a = 0; a < 3; a++) {
                po[a * KH * BLK_E] = fma(ge[a], v0.x, fma(ge[3 + a], v0.y, ge[6 + a] * v1.x)) * nidet;
                po[(3 + a) * KH * BLK_E] = fma(ge[a], v1.y, fma(ge[3 + a], v2.x, ge[6 + a] * v2.y)) * idet;
            }
        }
        __syncthreads();

        // ---- contraction: k~[i][e] = sum_j D_{c+1}[i][j] u~_{c+2}[j][e] - D_{c+2}[i][j] u~_{c+1}[j][e] + LIFT/2 [i][m] F~_c[m][e] ----
        for (int tile = warp; tile < 6 * G; tile += NW) {
            const int g = tile / 6;
            const double *Ub = sU + ((g * 6 + (1 - wf) * 3) * KH) * BLK_E;
            const int boff = (lane & 3) * BLK_E + ((lane >> 2) ^ ((lane & 2) << 1));   // swz8(4ks + (lane&3), lane>>2) - 32 ks
            double acc[MT][2];
#pragma unroll
            for (int mt = 0; mt < MT; mt++) acc[mt][0] = acc[mt][1] = 0.0;
#pragma unroll
            for (int ks = 0; ks < KSV; ks++) {
                const double bv = Ub[(c1 * KH + 4 * ks) * BLK_E + boff];
#pragma unroll
                for (int mt = 0; mt < MT; mt++)
                    dmma884(acc[mt][0], acc[mt][1], B::AREG ? aV[mt * KSV + ks] : sA[((x1 * MT + mt) * KSV + ks) * 32 + lane], bv);
            }
#pragma unroll
            for (int ks = 0; ks < KSV; ks++) {
                const double bv = -Ub[(x1 * KH + 4 * ks) * BLK_E + boff];
#pragma unroll
                for (int mt = 0; mt < MT; mt++)
                    dmma884(acc[mt][0], acc[mt][1], B::AREG ? aV[(MT + mt) * KSV + ks] : sA[((c1 * MT + mt) * KSV + ks) * 32 + lane], bv);
            }
            double *Fb = sF + ((g * 6 + 3 * wf + cp) * NFN) * BLK_E;
            const double *sL = sA + (B::AREG ? 0 : B::NAV);
#pragma unroll
            for (int ks = 0; ks < KSL; ks++) {
                const double bv = Fb[4 * ks * BLK_E + boff];
#pragma unroll
                for (int mt = 0; mt < MT; mt++) dmma884(acc[mt][0], acc[mt][1], sL[(mt * KSL + ks) * 32 + lane], bv);
            }
            __syncwarp();
#pragma unroll
            for (int mt = 0; mt < MT; mt++) {
                const int i = mt * 8 + (lane >> 2);
                if (i < Np) *reinterpret_cast<double2 *>(Fb + swz8(i, 2 * (lane & 3))) = make_double2(acc[mt][0], acc[mt][1]);
            }
        }
        __syncthreads();

        // ---- push forward, material, Runge-Kutta stage ---------------------------------------------------------
        if (LOAD_X || LOAD_Z) mbar_wait(&bars[2], it & 1);
        for (int item = tid; item < EB * Np; item += T) {
            const int e8 = item & 7, r = item >> 3, i = r % Np, g = r / Np, el = g * BLK_E + e8;
            const double *pk = sF + (g * 6) * NFN * BLK_E + swz8(i, e8);
            double kr[6];
#pragma unroll
            for (int c = 0; c < 6; c++) kr[c] = pk[c * NFN * BLK_E];
            const double *ge = geo + el * GST;
            const double ie = ge[23], im = ge[24], se = ge[25];
            const int off = ((g * Np + i) * BLK_E + e8) * 6;
            const double2 *pu = reinterpret_cast<const double2 *>(raw + off);
            const double2 u0 = pu[0], u1 = pu[1], u2 = pu[2];
            const double uo[6] = {u0.x, u0.y, u1.x, u1.y, u2.x, u2.y};
            double k[6];
#pragma unroll
            for (int d = 0; d < 3; d++) {
                k[d] = fma(ge[3 * d], kr[0], fma(ge[3 * d + 1], kr[1], ge[3 * d + 2] * kr[2])) * ie - se * uo[d];
                k[3 + d] = fma(ge[3 * d], kr[3], fma(ge[3 * d + 1], kr[4], ge[3 * d + 2] * kr[5])) * im;
            }
            double o[6], zn[6];
            if (MODE == MODE_MULT) {
#pragma unroll
                for (int c = 0; c < 6; c++) o[c] = k[c];
            } else if (MODE == MODE_STAGE1) {
#pragma unroll
                for (int c = 0; c < 6; c++) { o[c] = fma(A.a, k[c], uo[c]); zn[c] = fma(A.b, k[c], uo[c]); }
            } else if (MODE == MODE_STAGE23) {
#pragma unroll
                for (int c = 0; c < 6; c++) { o[c] = fma(A.a, k[c], sX[off + c]); zn[c] = fma(A.b, k[c], sZ[off + c]); }
            } else {
#pragma unroll
                for (int c = 0; c < 6; c++) o[c] = fma(A.b, k[c], sZ[off + c]);
            }
            double2 *po = reinterpret_cast<double2 *>(sU + off);
            po[0] = make_double2(o[0], o[1]); po[1] = make_double2(o[2], o[3]); po[2] = make_double2(o[4], o[5]);
            if (STORE_Z) {
                double2 *pz = reinterpret_cast<double2 *>(sZ + off);
                pz[0] = make_double2(zn[0], zn[1]); pz[1] = make_double2(zn[2], zn[3]); pz[2] = make_double2(zn[4], zn[5]);
            }
        }
        fence_async_smem();
        __syncthreads();
        if (tid == 0) {
            bulk_store(A.yout + (size_t)b * G * GS, sU, BATCH_BYTES);
            if (STORE_Z) bulk_store(A.z + (size_t)b * G * GS, sZ, BATCH_BYTES);
            bulk_commit();
        }
        ring = ring1;
    }
    if (tid == 0) bulk_wait_all();
}

// ---- layout conversion between the reference layout [6][Nloc] (Fields.h:45-65) and the blocked device layout ----------
// gid != nullptr: `ref` is in the caller's element order, local element e is element gid[e] there
__global__ void to_blocked_kernel(const double *ref, long long stride, int Np, long long NEloc, long long NEpad, const int *gid, double *blk)
{
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < NEpad * Np; idx += (long long)gridDim.x * blockDim.x) {
        const long long e = idx / Np; const int j = (int)(idx - e * Np);
        double *o = blk + (((e >> 3) * Np + j) * BLK_E + (e & 7)) * 6;
        const long long src = (gid && e < NEloc ? (long long)gid[e] : e) * Np + j;
#pragma unroll
        for (int c = 0; c < 6; c++) o[c] = e < NEloc ? ref[c * stride + src] : 0.0;
    }
}
__global__ void from_blocked_kernel(const double *blk, long long stride, int Np, long long NEloc, const int *gid, double *ref)
{
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < NEloc * Np; idx += (long long)gridDim.x * blockDim.x) {
        const long long e = idx / Np; const int j = (int)(idx - e * Np);
        const double *o = blk + (((e >> 3) * Np + j) * BLK_E + (e & 7)) * 6;
        const long long dst = (gid ? (long long)gid[e] : e) * Np + j;
#pragma unroll
        for (int c = 0; c < 6; c++) ref[c * stride + dst] = o[c];
    }
}
// halo pack: send[s][c] = y[send_off[s] + c]  (48-byte node records, receiver's face-node order)
__global__ void pack_blocked_kernel(const double *y, const long long *send_off, int ns, double *send)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < ns * 3; i += gridDim.x * blockDim.x) {
        const int s = i / 3, ch = i - 3 * s;
        reinterpret_cast<double2 *>(send)[i] = *reinterpret_cast<const double2 *>(y + send_off[s] + 2 * ch);
    }
}
__global__ void sample_blocked_kernel(const double *y, int Np, int npts, const int *elem, const double *shape, double *out)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npts) return;
    const long long e = elem[p];
    for (int c = 0; c < 6; c++) {
        double s = 0;
        for (int i = 0; i < Np; i++) s = fma(shape[(long long)p * Np + i], y[(((e >> 3) * Np + i) * BLK_E + (e & 7)) * 6 + c], s);
        out[p * 6 + c] = s;
    }
}

}  // namespace dgtd
