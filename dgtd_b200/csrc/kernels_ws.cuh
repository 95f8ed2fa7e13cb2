// sm_100a warp-specialised DMMA stage kernel (tetrahedra, sigma = 0): one persistent 20-warp CTA per SM.
//
// Same arithmetic and data layout as stage_mma_kernel (kernels_mma.cuh), re-organised after its ncu profile
// (profiles/r1_v2d_stage_mma_ncu_summary.txt: 12 warps/SM, phases serialised by block barriers, LSU pipe 62 %, DMMA 36 %):
//   warps  0..11  contraction: 6 output tile types (field, reference component) x {volume, LIFT}; each warp keeps ITS 30
//                 DMMA A-fragments in registers for the whole kernel (no operator traffic at all) and the 12 warps
//                 load the four SM sub-partitions evenly (3 each);
//   warps 12..17  flux: thread = (element, face, third of the face nodes) — geometry set up once per 3-4 nodes; the traces
//                 of faces leaving the batch are gathered (cp.async -> mbarrier) by the LIFT contraction warps;
//   warps 18..19  covariant transform of batch b+1 and push-forward + Runge-Kutta stage of batch b-1,
//                 thread = (element, quarter of the nodes).
// The roles run concurrently on different batches (16 elements each) and hand buffers over through mbarriers; the only
// block-wide barrier is at start-up.  HBM traffic is unchanged: bulk-TMA loads of y_in / x / z, bulk-TMA stores of
// y_out / z, cp.async gathers of the traces of faces leaving the batch.
#pragma once
#include "kernels_mma.cuh"

namespace dgtd {

__device__ __forceinline__ void named_bar(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory"); }
__device__ __forceinline__ void cp_async_wait_1() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }

template <int P> struct Ws {
    static constexpr int G = 2;
    static constexpr int Np = (P + 1) * (P + 2) * (P + 3) / 6, Nfp = (P + 1) * (P + 2) / 2, NFN = 4 * Nfp;
    static constexpr int MT = (Np + 7) / 8, KSV = (Np + 3) / 4, KH = 4 * KSV, KSL = NFN / 4;
    static constexpr int EB = BLK_E * G;
    static constexpr int GS = Np * BLK_E * 6, BS = G * GS;     // doubles per group / per batch of one state vector
    static constexpr int SL = 4 * EB;                          // trace descriptors per batch (worst case)
    static constexpr int CAP = 3 * EB;                         // trace slots staged in shared memory; the rest is read from L2 directly
    static constexpr int DS = EB * 8 + SL * 2 + 4;
    static constexpr int TABROWS = 128, GST = BLK_GEO + 2;
    static constexpr int NWG = 12, NWF = 6, NWR = 2, T = 32 * (NWG + NWF + NWR);
    static constexpr int NF = 32 * NWF, NR = 32 * NWR, NL = 32 * (NWG / 2);   // flux / transform+RK / LIFT-contraction threads
    static constexpr int NH = 3, HN = (Nfp + NH - 1) / NH;     // parts of a face (nodes h, h+NH, ...), nodes per part
    static constexpr int NQ = NR / EB, QN = (Np + NQ - 1) / NQ;   // node quarters per element, nodes per quarter
    static constexpr int UB = G * 6 * KH * BLK_E, FB = G * 6 * NFN * BLK_E;
    static constexpr int oRaw = 0;                             // [2][BS]
    static constexpr int oU = oRaw + 2 * BS;                   // [2][G][6][KH][8]
    static constexpr int oF = oU + 2 * UB;                     // [2][G][6][NFN][8] ; after the contraction rows 0..Np-1: LIFT part of k~, rows Np..2Np-1: volume part
    static constexpr int oX = oF + 2 * FB;                     // [BS] x (or y_in) -> y_out staging
    static constexpr int oZ = oX + BS;                         // [BS] z -> z staging
    static constexpr int oGeo = oZ + BS;                       // [3][EB][GST]
    static constexpr int oTr = oGeo + 3 * EB * GST;            // [2][CAP][Nfp][6]
    static constexpr int nDoubles = oTr + 2 * CAP * Nfp * 6;
    static constexpr size_t bDesc = (size_t)nDoubles * 8;      // int [3][DS]
    static constexpr size_t bTab = bDesc + (size_t)3 * DS * 4;
    static constexpr size_t bBar = bTab + (size_t)(TABROWS + 8) * Nfp;
    static constexpr int NBAR = 16;
    static constexpr size_t smem_bytes = bBar + NBAR * 8;
    static_assert(NFN >= 2 * Np, "both partial results must fit in a flux slice");
    static_assert(2 * MT * KSV <= 32 && MT * KSL <= 32, "A fragments must fit in registers");
    static_assert(NF == EB * 4 * NH && NR == EB * NQ && NWG == 12, "role thread counts are tied to the batch shape");
    static_assert((bDesc % 16) == 0 && (bBar % 8) == 0 && (DS % 4) == 0, "alignment");
};

// barrier slots
enum { WB_RAW = 0, WB_DG = 2, WB_XZ = 5, WB_FULLU = 6, WB_FULLF = 8, WB_DONEK = 10, WB_FREE = 12, WB_TR = 14 };

template <int P, int MODE>
__global__ void __launch_bounds__(Ws<P>::T, 1) stage_ws_kernel(const MmaArgs A)
{
    using B = Ws<P>;
    constexpr int G = B::G, Np = B::Np, Nfp = B::Nfp, NFN = B::NFN, MT = B::MT, KSV = B::KSV, KH = B::KH, KSL = B::KSL;
    constexpr int EB = B::EB, GS = B::GS, BS = B::BS, SL = B::SL, CAP = B::CAP, DS = B::DS, GST = B::GST, HN = B::HN, QN = B::QN;
    constexpr uint32_t BATCH_BYTES = (uint32_t)BS * 8;
    constexpr bool LOAD_X = MODE == MODE_STAGE1 || MODE == MODE_STAGE23;          // stage 1: x == y_in, fetched again (L2 hit)
    constexpr bool LOAD_Z = MODE == MODE_STAGE23 || MODE == MODE_STAGE4;
    constexpr bool STORE_X = MODE != MODE_STAGE4, STORE_Z = MODE != MODE_MULT;     // stage 4 stages y_out in the z buffer
    extern __shared__ __align__(128) unsigned char smem_ws[];
    double *sm = reinterpret_cast<double *>(smem_ws);
    int *sDesc = reinterpret_cast<int *>(smem_ws + B::bDesc);
    uint8_t *sTab = smem_ws + B::bTab;
    uint8_t *sIdent = sTab + B::TABROWS * Nfp;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_ws + B::bBar);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b0 = blockIdx.x, gstep = gridDim.x;
    const int nmine = b0 < A.nbatch ? (A.nbatch - b0 + gstep - 1) / gstep : 0;     // batches of this CTA: b0 + it*gstep
    {
        const int nb = min(A.ntab, B::TABROWS) * Nfp;
        for (int i = tid; i < nb; i += B::T) sTab[i] = A.ftab[i];
        if (tid < Nfp) sIdent[tid] = (uint8_t)tid;
        if (KH > Np) for (int i = tid; i < 2 * B::UB; i += B::T) sm[B::oU + i] = 0.0;   // k-padding rows stay zero
    }
    if (tid == 0) {
        mbar_init(&bars[WB_RAW], 1); mbar_init(&bars[WB_RAW + 1], 1);
        mbar_init(&bars[WB_DG], 1); mbar_init(&bars[WB_DG + 1], 1); mbar_init(&bars[WB_DG + 2], 1);
        mbar_init(&bars[WB_XZ], 1);
        mbar_init(&bars[WB_FULLU], B::NR); mbar_init(&bars[WB_FULLU + 1], B::NR);
        mbar_init(&bars[WB_FULLF], B::NF); mbar_init(&bars[WB_FULLF + 1], B::NF);
        mbar_init(&bars[WB_DONEK], B::NWG); mbar_init(&bars[WB_DONEK + 1], B::NWG);
        mbar_init(&bars[WB_FREE], 1); mbar_init(&bars[WB_FREE + 1], 1);
        mbar_init(&bars[WB_TR], B::NL); mbar_init(&bars[WB_TR + 1], B::NL);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        fence_async_smem();
    }
    __syncthreads();
    if (nmine == 0) return;

    // descriptor block + geometry records of this CTA's batch number `n` into ring slot n % 3 (bulk TMA, one thread)
    auto load_desc_geo = [&](int n) {
        const int slot = n % 3, bb = b0 + n * gstep;
        mbar_expect_tx(&bars[WB_DG + slot], (uint32_t)(DS * 4 + EB * BLK_GEO * 8));
        bulk_load(sDesc + slot * DS, A.desc + (size_t)bb * DS, DS * 4, &bars[WB_DG + slot]);
        for (int e = 0; e < EB; e++)
            bulk_load(sm + B::oGeo + (slot * EB + e) * GST, A.geo + ((size_t)bb * EB + e) * BLK_GEO, BLK_GEO * 8, &bars[WB_DG + slot]);
    };
    auto load_raw = [&](int n) {
        mbar_expect_tx(&bars[WB_RAW + (n & 1)], BATCH_BYTES);
        bulk_load(sm + B::oRaw + (n & 1) * BS, A.yin + (size_t)(b0 + n * gstep) * BS, BATCH_BYTES, &bars[WB_RAW + (n & 1)]);
    };

    // traces of faces leaving batch n -> buffer n & 1, issued by the NL threads of the LIFT contraction warps (they idle
    // while the producers work); completion is signalled on an mbarrier by the copies themselves
    auto issue_traces = [&](int n, int tl) {
        mbar_wait(&bars[WB_DG + n % 3], (n / 3) & 1);
        const int *dsc = sDesc + (n % 3) * DS;
        const int nt = min(dsc[EB * 8 + SL * 2], CAP);
        const int2 *td = reinterpret_cast<const int2 *>(dsc + EB * 8);
        double *tdst = sm + B::oTr + (n & 1) * CAP * Nfp * 6;
        constexpr int PER = Nfp * 3, LANES = B::NL / PER;      // (node, 16-byte chunk) pairs ; slots handled concurrently
        const int mc = tl % PER, m = mc / 3, ch = mc - 3 * m;
        if (tl < LANES * PER)
            for (int slot = tl / PER; slot < nt; slot += LANES) {
                const int2 d = td[slot];
                const double *src;
                if (d.x >= 0) src = A.yin + (((size_t)(d.x >> 3) * Np + sTab[d.y * Nfp + m]) * BLK_E + (d.x & 7)) * 6;
                else src = A.halo + ((size_t)(-1 - d.x) * Nfp + m) * 6;
                cp_async16(tdst + (slot * Nfp + m) * 6 + 2 * ch, src + 2 * ch);
            }
        asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(&bars[WB_TR + (n & 1)])) : "memory");
    };

    if (warp < B::NWG) {
        // =============================== contraction warps =====================================================
        const int tt = warp % 6, part = warp / 6, wf = tt / 3, cp = tt - 3 * wf;
        const int c1 = (cp + 2) % 3, x1 = (cp + 1) % 3;
        constexpr int NAF = MT * (2 * KSV > KSL ? 2 * KSV : KSL);
        double aF[NAF];                           // volume: [mt][2*KSV] = D_{x1} | D_{c1} ; LIFT: [mt][KSL]
        if (part == 0) {
#pragma unroll
            for (int mt = 0; mt < MT; mt++)
#pragma unroll
                for (int ks = 0; ks < KSV; ks++) {
                    aF[mt * 2 * KSV + ks] = __ldg(A.afrag + ((x1 * MT + mt) * KSV + ks) * 32 + lane);
                    aF[mt * 2 * KSV + KSV + ks] = __ldg(A.afrag + ((c1 * MT + mt) * KSV + ks) * 32 + lane);
                }
        } else {
#pragma unroll
            for (int mt = 0; mt < MT; mt++)
#pragma unroll
                for (int ks = 0; ks < KSL; ks++) aF[mt * KSL + ks] = __ldg(A.afrag + 3 * MT * KSV * 32 + (mt * KSL + ks) * 32 + lane);
        }
        const int boff = (lane & 3) * BLK_E + ((lane >> 2) ^ ((lane & 2) << 1));
        const int tl = tid - 32 * (B::NWG / 2);               // LIFT warps: index among the trace-issuing threads
        if (part == 1) issue_traces(0, tl);
        for (int it = 0; it < nmine; it++) {
            const int par = it & 1, ph = (it >> 1) & 1;
            // buffer (it+1)&1 was last read by the flux of batch it-1, which this warp's previous contraction waited for
            if (part == 1 && it + 1 < nmine) issue_traces(it + 1, tl);
            const double *sU = sm + B::oU + par * B::UB;
            double *sF = sm + B::oF + par * B::FB;
            mbar_wait(&bars[(part == 0 ? WB_FULLU : WB_FULLF) + par], ph);
#pragma unroll
            for (int g = 0; g < G; g++) {
                double acc[MT][2];
#pragma unroll
                for (int mt = 0; mt < MT; mt++) acc[mt][0] = acc[mt][1] = 0.0;
                double *Fb = sF + ((g * 6 + 3 * wf + cp) * NFN) * BLK_E;
                if (part == 0) {
                    const double *Ub = sU + ((g * 6 + (1 - wf) * 3) * KH) * BLK_E;
#pragma unroll
                    for (int ks = 0; ks < KSV; ks++) {
                        const double bv = Ub[(c1 * KH + 4 * ks) * BLK_E + boff];
#pragma unroll
                        for (int mt = 0; mt < MT; mt++) dmma884(acc[mt][0], acc[mt][1], aF[mt * 2 * KSV + ks], bv);
                    }
#pragma unroll
                    for (int ks = 0; ks < KSV; ks++) {
                        const double bv = -Ub[(x1 * KH + 4 * ks) * BLK_E + boff];
#pragma unroll
                        for (int mt = 0; mt < MT; mt++) dmma884(acc[mt][0], acc[mt][1], aF[mt * 2 * KSV + KSV + ks], bv);
                    }
                } else {
#pragma unroll
                    for (int ks = 0; ks < KSL; ks++) {
                        const double bv = Fb[4 * ks * BLK_E + boff];
#pragma unroll
                        for (int mt = 0; mt < MT; mt++) dmma884(acc[mt][0], acc[mt][1], aF[mt * KSL + ks], bv);
                    }
                }
                // the volume warp overwrites rows Np.. of the slice its LIFT partner is reading: meet first
                named_bar(3 + tt, 64);
#pragma unroll
                for (int mt = 0; mt < MT; mt++) {
                    const int i = mt * 8 + (lane >> 2);
                    if (i < Np) *reinterpret_cast<double2 *>(Fb + swz8((part == 0 ? Np : 0) + i, 2 * (lane & 3))) = make_double2(acc[mt][0], acc[mt][1]);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars[WB_DONEK + par]);
        }
    } else if (warp < B::NWG + B::NWF) {
        // =============================== flux warps ===========================================================
        const int t = tid - 32 * B::NWG;                      // 0..NF-1
        const int e8 = t & 7, h = (t >> 3) % B::NH, f = ((t >> 3) / B::NH) & 3, g = (t >> 3) / (4 * B::NH), el = g * BLK_E + e8;
        const bool inject = A.pw_on && (A.gate == nullptr || *A.gate >= 1e-16);
        if (t == 0) {   // start-up loads of the whole CTA
            for (int n = 0; n < 3 && n < nmine; n++) load_desc_geo(n);
            for (int n = 0; n < 2 && n < nmine; n++) load_raw(n);
        }
        for (int it = 0; it < nmine; it++) {
            const int par = it & 1, ph = (it >> 1) & 1, slot = it % 3;
            mbar_wait(&bars[WB_TR + par], ph);                 // the traces of this batch have landed (issued by the LIFT warps)
            mbar_wait(&bars[WB_DG + slot], (it / 3) & 1);
            mbar_wait(&bars[WB_RAW + par], ph);
            if (it >= 2) mbar_wait(&bars[WB_FREE + par], ((it - 2) >> 1) & 1);
            const double *raw = sm + B::oRaw + par * BS;
            const double *tr = sm + B::oTr + par * CAP * Nfp * 6;
            const double *ge = sm + B::oGeo + (slot * EB + el) * GST;
            const int *dsc = sDesc + slot * DS;
            double *sF = sm + B::oF + par * B::FB;
            const int2 info = reinterpret_cast<const int2 *>(dsc)[el * 4 + f];
            const int code = info.y;
            const uint8_t *srow = sTab + f * Nfp;
            const double *ubase = raw + (g * Np * BLK_E + e8) * 6;
            const uint8_t *nrow; const double *nbase = ubase; const double *gbase = nullptr; int nstride;
            double ce = 0.0, ch = 0.0, al = A.alpha;
            if (info.x >= 0) {
                nrow = sTab + ((code >> FI_TAB_SHIFT) & FI_TAB_MASK) * Nfp;
                nbase = raw + ((info.x >> 3) * Np * BLK_E + (info.x & 7)) * 6; nstride = BLK_E * 6;
            } else if (info.x == -1) {
                const int bc = code & FI_BC_MASK;
                ce = bc == 1 ? -2.0 : bc == 3 ? -1.0 : 0.0;
                ch = bc == 2 ? -2.0 : bc == 3 ? -1.0 : 0.0;
                if (bc == 3) al = 1.0;
                nrow = srow; nstride = BLK_E * 6;
            } else {
                const int s = -2 - info.x;
                if (s < CAP) { nrow = sIdent; nbase = tr + s * Nfp * 6; nstride = 6; }
                else {   // more faces leave this batch than the trace buffer stages: read the trace from L2
                    const int2 d = reinterpret_cast<const int2 *>(dsc + EB * 8)[s];
                    if (d.x >= 0) { nrow = sTab + d.y * Nfp; gbase = A.yin + ((size_t)(d.x >> 3) * Np * BLK_E + (d.x & 7)) * 6; nstride = BLK_E * 6; }
                    else { nrow = sIdent; gbase = A.halo + (size_t)(-1 - d.x) * Nfp * 6; nstride = 6; }
                }
            }
            const int tf = (code >> FI_TFSF_SHIFT) & FI_TFSF_MASK;
            double ji[9];
#pragma unroll
            for (int i = 0; i < 9; i++) ji[i] = ge[9 + i];
            double gn[3];
#pragma unroll
            for (int d = 0; d < 3; d++) gn[d] = f == 0 ? (ji[d] + ji[3 + d]) + ji[6 + d] : -ge[9 + 3 * (f - 1) + d];
            const double fs = ge[18 + f];
            const double ifs = 1.0 / fs, ifs2 = ifs * ifs;
            const double af = al * fs;
#pragma unroll
            for (int mm = 0; mm < HN; mm++) {
                const int m = h + B::NH * mm;
                if (m >= Nfp) break;
                double uM[6], dU[6];
                {
                    const double2 *pu = reinterpret_cast<const double2 *>(ubase + srow[m] * (BLK_E * 6));
                    const double2 v0 = pu[0], v1 = pu[1], v2 = pu[2];
                    double2 w0, w1, w2;
                    if (gbase == nullptr) {
                        const double2 *pn = reinterpret_cast<const double2 *>(nbase + nrow[m] * nstride);
                        w0 = pn[0]; w1 = pn[1]; w2 = pn[2];
                    } else {
                        const double2 *pn = reinterpret_cast<const double2 *>(gbase + (size_t)nrow[m] * nstride);
                        w0 = __ldg(pn); w1 = __ldg(pn + 1); w2 = __ldg(pn + 2);
                    }
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
                double fl[6];
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
            mbar_arrive(&bars[WB_FULLF + par]);
            named_bar(1, B::NF);                               // all flux threads are done with raw[par] and the trace buffer
            if (t == 0 && it + 2 < nmine) {
                mbar_wait(&bars[WB_FULLU + par], ph);          // ... and so is the transform
                load_raw(it + 2);
            }
        }
    } else {
        // =============================== transform + Runge-Kutta warps =============================================
        const int r = tid - 32 * (B::NWG + B::NWF);            // 0..NR-1
        const int e8 = r & 7, q = (r >> 3) % B::NQ, g = r / (BLK_E * B::NQ), el = g * BLK_E + e8;
        auto rk_stage = [&](int n) {                           // push forward + RK update of this CTA's batch n
            const int par = n & 1, slot = n % 3;
            mbar_wait(&bars[WB_DONEK + par], (n >> 1) & 1);
            if (LOAD_X || LOAD_Z) mbar_wait(&bars[WB_XZ], n & 1);
            const double *sF = sm + B::oF + par * B::FB;
            const double *ge = sm + B::oGeo + (slot * EB + el) * GST;
            double *sX = sm + B::oX, *sZ = sm + B::oZ;
            double jm[9];
#pragma unroll
            for (int i = 0; i < 9; i++) jm[i] = ge[i];
            const double ie = ge[23], im = ge[24];
#pragma unroll
            for (int k5 = 0; k5 < QN; k5++) {
                const int i = q * QN + k5;
                if (i >= Np) break;
                const double *pl = sF + (g * 6) * NFN * BLK_E + swz8(i, e8);
                const double *pv = sF + (g * 6) * NFN * BLK_E + swz8(Np + i, e8);
                double kr[6];
#pragma unroll
                for (int c = 0; c < 6; c++) kr[c] = pl[c * NFN * BLK_E] + pv[c * NFN * BLK_E];
                double k[6];
#pragma unroll
                for (int d = 0; d < 3; d++) {
                    k[d] = fma(jm[3 * d], kr[0], fma(jm[3 * d + 1], kr[1], jm[3 * d + 2] * kr[2])) * ie;
                    k[3 + d] = fma(jm[3 * d], kr[3], fma(jm[3 * d + 1], kr[4], jm[3 * d + 2] * kr[5])) * im;
                }
                const int off = ((g * Np + i) * BLK_E + e8) * 6;
                double2 *px = reinterpret_cast<double2 *>(sX + off), *pz = reinterpret_cast<double2 *>(sZ + off);
                double xv[6], zv[6], o[6], zn[6];
                if (LOAD_X) { const double2 a0 = px[0], a1 = px[1], a2 = px[2]; xv[0] = a0.x; xv[1] = a0.y; xv[2] = a1.x; xv[3] = a1.y; xv[4] = a2.x; xv[5] = a2.y; }
                if (LOAD_Z) { const double2 a0 = pz[0], a1 = pz[1], a2 = pz[2]; zv[0] = a0.x; zv[1] = a0.y; zv[2] = a1.x; zv[3] = a1.y; zv[4] = a2.x; zv[5] = a2.y; }
#pragma unroll
                for (int c = 0; c < 6; c++) {
                    if (MODE == MODE_MULT) o[c] = k[c];
                    else if (MODE == MODE_STAGE1) { o[c] = fma(A.a, k[c], xv[c]); zn[c] = fma(A.b, k[c], xv[c]); }
                    else if (MODE == MODE_STAGE23) { o[c] = fma(A.a, k[c], xv[c]); zn[c] = fma(A.b, k[c], zv[c]); }
                    else zn[c] = fma(A.b, k[c], zv[c]);            // stage 4: new x, staged in the z buffer
                }
                if (STORE_X) { px[0] = make_double2(o[0], o[1]); px[1] = make_double2(o[2], o[3]); px[2] = make_double2(o[4], o[5]); }
                if (STORE_Z) { pz[0] = make_double2(zn[0], zn[1]); pz[1] = make_double2(zn[2], zn[3]); pz[2] = make_double2(zn[4], zn[5]); }
            }
            fence_async_smem();
            named_bar(2, B::NR);
            if (r == 0) {
                const size_t goff = (size_t)(b0 + n * gstep) * BS;
                if (STORE_X) bulk_store(A.yout + goff, sX, BATCH_BYTES);
                if (MODE == MODE_STAGE4) bulk_store(A.yout + goff, sZ, BATCH_BYTES);
                else if (STORE_Z) bulk_store(A.z + goff, sZ, BATCH_BYTES);
                bulk_commit();
                mbar_arrive(&bars[WB_FREE + par]);             // U/F[par] (holding k~ of batch n) may be refilled
                if (n + 3 < nmine) load_desc_geo(n + 3);       // ring slot n % 3 is free again
            }
        };
        for (int it = 0; it < nmine; it++) {
            const int par = it & 1, ph = (it >> 1) & 1, slot = it % 3;
            if (r == 0 && it >= 1 && (LOAD_X || LOAD_Z)) {    // x / z of the batch whose RK stage follows this transform
                bulk_wait_read();                              // the previous stage's stores have drained the buffers
                const size_t goff = (size_t)(b0 + (it - 1) * gstep) * BS;
                mbar_expect_tx(&bars[WB_XZ], BATCH_BYTES * ((LOAD_X ? 1 : 0) + (LOAD_Z ? 1 : 0)));
                if (LOAD_X) bulk_load(sm + B::oX, (MODE == MODE_STAGE1 ? A.yin : A.x) + goff, BATCH_BYTES, &bars[WB_XZ]);
                if (LOAD_Z) bulk_load(sm + B::oZ, A.z + goff, BATCH_BYTES, &bars[WB_XZ]);
            } else if (r == 0 && it >= 1) bulk_wait_read();
            // ---- covariant field of batch `it` -------------------------------------------------------------------
            mbar_wait(&bars[WB_DG + slot], (it / 3) & 1);
            mbar_wait(&bars[WB_RAW + par], ph);
            {
                const double *raw = sm + B::oRaw + par * BS;
                const double *ge = sm + B::oGeo + (slot * EB + el) * GST;
                double *sU = sm + B::oU + par * B::UB;
                double jm[9];
#pragma unroll
                for (int i = 0; i < 9; i++) jm[i] = ge[i];
                const double idet = ge[22], nidet = -idet;
#pragma unroll
                for (int k5 = 0; k5 < QN; k5++) {
                    const int j = q * QN + k5;
                    if (j >= Np) break;
                    const double2 *pu = reinterpret_cast<const double2 *>(raw + ((g * Np + j) * BLK_E + e8) * 6);
                    const double2 v0 = pu[0], v1 = pu[1], v2 = pu[2];
                    double *po = sU + (g * 6) * KH * BLK_E + swz8(j, e8);
#pragma unroll
                    for (int a = 0; a < 3; a++) {
                        po[a * KH * BLK_E] = fma(jm[a], v0.x, fma(jm[3 + a], v0.y, jm[6 + a] * v1.x)) * nidet;
                        po[(3 + a) * KH * BLK_E] = fma(jm[a], v1.y, fma(jm[3 + a], v2.x, jm[6 + a] * v2.y)) * idet;
                    }
                }
            }
            mbar_arrive(&bars[WB_FULLU + par]);
            if (it >= 1) rk_stage(it - 1);
        }
        // the last batch: its x / z, then its stage
        if (r == 0) {
            bulk_wait_read();
            if (LOAD_X || LOAD_Z) {
                const size_t goff = (size_t)(b0 + (nmine - 1) * gstep) * BS;
                mbar_expect_tx(&bars[WB_XZ], BATCH_BYTES * ((LOAD_X ? 1 : 0) + (LOAD_Z ? 1 : 0)));
                if (LOAD_X) bulk_load(sm + B::oX, (MODE == MODE_STAGE1 ? A.yin : A.x) + goff, BATCH_BYTES, &bars[WB_XZ]);
                if (LOAD_Z) bulk_load(sm + B::oZ, A.z + goff, BATCH_BYTES, &bars[WB_XZ]);
            }
        }
        rk_stage(nmine - 1);
        if (r == 0) bulk_wait_all();
    }
}

}  // namespace dgtd
