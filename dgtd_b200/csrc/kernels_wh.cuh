// sm_100a "half-row" DMMA stage kernel (tetrahedra): one warp carries a group of FOUR elements; the eight DMMA rows are
// (element, field): row r = 2 e + h feeds / accumulates the E rows (h = 0) or the H rows (h = 1) of element e.
//
// Why: the one-warp-per-8-elements kernel (kernels_wg.cuh) needs 72 accumulator registers and 25 KB of buffers per warp,
// so 8 warps fill an SM at order 3 and only 4 at order 4.  Splitting a group between two WARPS by field (tried: 12 warps at
// order 3, 8 at order 4) halves the accumulators but doubles the operator-fragment and trace loads and needs two named
// barriers per group: 82 G at order 3 (shared-memory bound), 69 G at order 4.  Here the same field split happens between
// the ROWS of one DMMA: a fragment load still serves eight rows, the two lanes that need the same node record read the same
// address (one broadcast), the accumulator set is 36 registers, a warp's buffers are 12.7 KB — 12 warps per SM at order
// <= 3 (3 per scheduler), 8 at order 4, no barrier between warps.  Cost: both rows of an element form the jumps
// (+25 % of the flux's FP64 ALU work, +3 % of the pipe time).
//   lane l: DMMA row r = l >> 2 -> element e = r >> 1 of the group, field h = r & 1; j = l & 3 is the k index (node of a
//   k-step in the volume contraction, face in the flux) exactly as in kernels_wg.cuh; own = 3h / other = 3 - 3h are the
//   offsets of "my" and "the other" field inside a 6-double node record.
// State layout, plan (WgPlan: per-element geometry and descriptors, fragments, tables, halo push) and the peer-memory
// halo are those of kernels_wg.cuh; a group is simply half of one of its groups.
// Reference semantics: src/evolution/HesthavenEvolution.cpp:450-542 with the `global` operator's coefficients
// (src/components/DGOperatorFactory.h:469-573, 1268-1361), external/mfem-geg/linalg/ode.cpp:109-136.
#pragma once
#include "kernels_wg.cuh"

namespace dgtd {

constexpr int WH_E = 4;                // elements per group
__device__ __forceinline__ void load3(const double *p, double *u) { u[0] = p[0]; u[1] = p[1]; u[2] = p[2]; }
__device__ __forceinline__ void store3(double *p, const double *u) { p[0] = u[0]; p[1] = u[1]; p[2] = u[2]; }
#ifndef DGTD_WH_NW
#define DGTD_WH_NW 12
#endif
template <int P> struct Wh {
    static constexpr int Np = (P + 1) * (P + 2) * (P + 3) / 6, Nfp = (P + 1) * (P + 2) / 2;
    static constexpr int NT = (Np + 7) / 8, KSV = (Np + 3) / 4, VT = (NT - 1) * 3 + 3;
    static constexpr int NW = P <= 3 ? DGTD_WH_NW : 8, T = 32 * NW;             // warps per CTA = groups in flight per SM
    static constexpr int GS = Np * WH_E * 6;                                    // doubles per group of one state vector
    static constexpr int NFV = KSV * VT, NFL = Nfp * NT, NFR = NFV + NFL;
    static constexpr int WGEO = WH_E * WG_GEO, WDESC = WH_E * 4 * 2;
    static constexpr int WDBL = 3 * GS + WGEO + WDESC / 2;
    static constexpr int TABROWS = WG_TABROWS;
    static constexpr int oWarp = NFR * 32;
    static constexpr size_t bTab = (size_t)(oWarp + NW * WDBL) * 8;
    static constexpr size_t bBar = bTab + (size_t)TABROWS * 16;
    static constexpr size_t smem_bytes = bBar + (size_t)(NW * 2 + 1) * 8;     // per-warp barriers + one for the operator fragments
    static_assert(Np - 8 * (NT - 1) <= 4, "mixed last tile");
    static_assert((GS % 2) == 0 && (bTab % 16) == 0 && (WDBL % 2) == 0, "alignment");
};

template <int P, int MODE, bool TF>
__global__ void __launch_bounds__(Wh<P>::T, 1) stage_wh_kernel(const WgArgs A)
{
    using B = Wh<P>;
    constexpr int Np = B::Np, Nfp = B::Nfp, NT = B::NT, KSV = B::KSV, VT = B::VT, GS = B::GS;
    constexpr int NL = Np - 8 * (NT - 1);
#ifndef DGTD_WH_PF
#define DGTD_WH_PF 2
#endif
    constexpr int PF = DGTD_WH_PF;                      // neighbour-record prefetch distance (face steps)
    constexpr bool LOAD_X = MODE == MODE_STAGE1 || MODE == MODE_STAGE23;
    constexpr bool LOAD_Z = MODE == MODE_STAGE23 || MODE == MODE_STAGE4;
    constexpr bool STORE_X = MODE != MODE_STAGE4, STORE_Z = MODE != MODE_MULT;
    extern __shared__ __align__(128) unsigned char smem_wg[];
    double *sm = reinterpret_cast<double *>(smem_wg);
    const double *sFragV = sm, *sFragL = sm + B::NFV * 32;
    const uint4 *sTab = reinterpret_cast<const uint4 *>(smem_wg + B::bTab);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, j = lane & 3;
    const int e = lane >> 3, h = (lane >> 2) & 1;        // element of the group, field of my DMMA row (0: E rows, 1: H rows)
    const int own = 3 * h, oth = 3 - own;                // doubles into a node record: my output field / the other one
    const bool leader = lane == 0;
    double *wY = sm + B::oWarp + warp * B::WDBL, *wX = wY + GS, *wZ = wX + GS, *wGeo = wZ + GS;
    const int2 *wDesc = reinterpret_cast<const int2 *>(wGeo + B::WGEO);
    uint64_t *barY = reinterpret_cast<uint64_t *>(smem_wg + B::bBar) + 2 * warp, *barXZ = barY + 1;

    uint64_t *barF = reinterpret_cast<uint64_t *>(smem_wg + B::bBar) + 2 * B::NW;   // operator fragments: one bulk copy per CTA
    {
        uint4 *dst = reinterpret_cast<uint4 *>(smem_wg + B::bTab);
        const uint4 *src = reinterpret_cast<const uint4 *>(A.tab);
        for (int i = tid; i < min(A.ntab, B::TABROWS); i += B::T) dst[i] = src[i];
    }
    if (tid == 0) mbar_init(barF, 1);
    if (leader) { mbar_init(barY, 1); mbar_init(barXZ, 1); }
    if (tid == 0) asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    fence_async_smem();
    __syncthreads();

    const int gstride = gridDim.x * B::NW, ngroups = A.ngroups * (BLK_E / WH_E);   // WgArgs counts groups of 8
    int g = blockIdx.x * B::NW + warp;

    const uint4 ownrow = sTab[j];
    const bool inject = A.pw_on && (A.gate == nullptr || *A.gate >= 1e-16);
    unsigned long long *pend_flag = nullptr;             // see kernels_wg.cuh: the flag of a pushed face is raised one group later
    const double sgn = h ? -1.0 : 1.0;                   // u~_E = -(J/det)^T E feeds the H rows, u~_H = +(J/det)^T H the E rows

    auto issue_y = [&](int gg) {
        mbar_expect_tx(barY, (uint32_t)(GS * 8 + B::WGEO * 8 + B::WDESC * 4));
        bulk_load(wY, A.yin + (size_t)gg * GS, GS * 8, barY);
        bulk_load(wGeo, A.geo + (size_t)gg * B::WGEO, B::WGEO * 8, barY);
        bulk_load(wGeo + B::WGEO, A.desc + (size_t)gg * B::WDESC, B::WDESC * 4, barY);
    };
    auto issue_xz = [&](int gg) {
        mbar_expect_tx(barXZ, (uint32_t)(GS * 8) * ((LOAD_X ? 1 : 0) + (LOAD_Z ? 1 : 0)));
        if (LOAD_X) bulk_load(wX, (MODE == MODE_STAGE1 ? A.yin : A.x) + (size_t)gg * GS, GS * 8, barXZ);
        if (LOAD_Z) bulk_load(wZ, A.z + (size_t)gg * GS, GS * 8, barXZ);
    };
    if (tid == 0) { mbar_expect_tx(barF, (uint32_t)(B::NFR * 32 * 8)); bulk_load(sm, A.bfrag, B::NFR * 32 * 8, barF); }
    if (leader && g < ngroups) { issue_y(g); if (LOAD_X || LOAD_Z) issue_xz(g); }
    mbar_wait(barF, 0);

    if (A.work && blockIdx.x == 0 && tid == 0) *A.work_next = 0;      // groups beyond the first wave come from a counter (kernels_wg.cuh)
    for (int it = 0; g < ngroups; it++) {
        const uint32_t par = it & 1;
        int gnext = 0;
        if (A.work) {
            if (lane == 0) gnext = gstride + (int)atomicAdd(A.work, 1u);
            gnext = __shfl_sync(0xffffffffu, gnext, 0);
        } else gnext = g + gstride;                     // static split (single-rank contexts at order <= 3: measured 1 % faster there)
        if (gnext >= ngroups) gnext = -1;
        const double *ge = wGeo + e * WG_GEO;
        const double *yrec = wY + e * Np * 6;
        mbar_wait(barY, par);

        double acc[3][NT][2];
#pragma unroll
        for (int c = 0; c < 3; c++)
#pragma unroll
            for (int nt = 0; nt < NT; nt++) acc[c][nt][0] = acc[c][nt][1] = 0.0;

        // ---------------- face (e, j): where the exterior trace comes from; first neighbour record requested now -----------
        const int2 info = wDesc[e * 4 + j];
        const int code = info.y;
        const double *nbase = yrec;
        uint4 nrow = ownrow;
        double ce = 0.0, ch = 0.0, al = A.alpha;
        bool nb_smem = true;                                  // the exterior trace lies in this warp's own y_in buffer
        if (info.x >= 0) {
            nrow = sTab[(code >> FI_TAB_SHIFT) & FI_TAB_MASK];
            nb_smem = (info.x >> 2) == g;
            nbase = nb_smem ? wY + (info.x & 3) * Np * 6 : A.yin + (size_t)info.x * Np * 6;
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

        // ---------------- volume: my field's k~_c = D_{c+1} u~_{c+2} - D_{c+2} u~_{c+1} of the OTHER field ---------------------
        {
            double jm[9];
#pragma unroll
            for (int i = 0; i < 9; i++) jm[i] = sgn * ge[i];
#pragma unroll
            for (int ks = 0; ks < KSV; ks++) {
                const int node = 4 * ks + j;
                double u[3] = {0, 0, 0};
                if (4 * ks + 3 < Np || node < Np) load3(yrec + node * 6 + oth, u);
                double ut[3];
#pragma unroll
                for (int a = 0; a < 3; a++) ut[a] = fma(jm[a], u[0], fma(jm[3 + a], u[1], jm[6 + a] * u[2]));
                const double *fr = sFragV + (ks * VT) * 32 + lane;
#pragma unroll
                for (int nt = 0; nt < NT - 1; nt++)
#pragma unroll
                    for (int d = 0; d < 3; d++) {
                        const double bv = fr[(nt * 3 + d) * 32];
                        const int cp = (d + 2) % 3, cm = (d + 1) % 3;
                        dmma884(acc[cp][nt][0], acc[cp][nt][1], ut[(d + 1) % 3], bv);
                        dmma884(acc[cm][nt][0], acc[cm][nt][1], -ut[(d + 2) % 3], bv);
                    }
#pragma unroll
                for (int x = 0; x < 3; x++) {
                    const double bv = fr[((NT - 1) * 3 + x) * 32];
                    dmma884(acc[x][NT - 1][0], acc[x][NT - 1][1], ut[x], bv);
                }
            }
        }
        if ((LOAD_X || LOAD_Z) && it > 0 && leader) { bulk_wait_read(); issue_xz(g); }

        // ---------------- face flux of (element e, face j), my field's rows -> LIFT ---------------------------------------------
        {
            const int tf = (code >> FI_TFSF_SHIFT) & FI_TFSF_MASK;
            const double sOwn = 1.0 - (h ? ch : ce), sOth = 1.0 - (h ? ce : ch);   // dU = u+ - (1 - c) u-
            double ji[9];
#pragma unroll
            for (int i = 0; i < 9; i++) ji[i] = ge[9 + i];
            double gn[3];
#pragma unroll
            for (int d = 0; d < 3; d++) gn[d] = j == 0 ? (ji[d] + ji[3 + d]) + ji[6 + d] : -ji[3 * (j - 1) + d];
            // F~_own = Ax (dU_other - (+-na) x dU_own):  F_E = g x (dH - na x dE), F_H = -g x (dE + na x dH), g = fs n, na = alpha n,
            // Ax = +-J^-1 [g x]  (E rows +, H rows -); see kernels_wg.cuh
            const double ans = sgn * al * ge[26 + j];               // +- alpha / fs
            const double na0 = ans * gn[0], na1 = ans * gn[1], na2 = ans * gn[2];
            double Ax[9];
#pragma unroll
            for (int a = 0; a < 3; a++) {
                const double j0 = ji[3 * a], j1 = ji[3 * a + 1], j2 = ji[3 * a + 2];
                Ax[3 * a + 0] = sgn * (j1 * gn[2] - j2 * gn[1]);
                Ax[3 * a + 1] = sgn * (j2 * gn[0] - j0 * gn[2]);
                Ax[3 * a + 2] = sgn * (j0 * gn[1] - j1 * gn[0]);
            }
#pragma unroll
            for (int s = 0; s < Nfp; s++) {
                const double *uP = uQ[s % (PF + 1)];
                const double *mrec = yrec + tab_byte(ownrow, s) * 6;
                double mO[3], mX[3], dO[3], dX[3];
                load3(mrec + own, mO);
                load3(mrec + oth, mX);
                if (s + PF < Nfp) load_rec_split(nbase + tab_byte(nrow, s + PF) * 6, nb_smem, uQ[(s + PF) % (PF + 1)]);
#pragma unroll
                for (int c = 0; c < 3; c++) {
                    const double pO = h ? uP[3 + c] : uP[c], pX = h ? uP[c] : uP[3 + c];
                    dO[c] = fma(-sOwn, mO[c], pO);
                    dX[c] = fma(-sOth, mX[c], pX);
                }
                if (TF && tf && inject) {
                    double inc[6];
                    const int m = tab_byte(sTab[4 + j], s);
                    planewave6(A.pw, A.tfsf_xyz + ((long long)(code >> FI_TIDX_SHIFT) * Nfp + m) * 3, A.t, inc);
                    const double sg = tf == 1 ? 1.0 : -1.0;
#pragma unroll
                    for (int c = 0; c < 3; c++) { dO[c] += sg * (h ? inc[3 + c] : inc[c]); dX[c] += sg * (h ? inc[c] : inc[3 + c]); }
                }
                double w[3], ft[3];
                w[0] = fma(na2, dO[1], fma(-na1, dO[2], dX[0]));      // dU_other - (+-na) x dU_own
                w[1] = fma(na0, dO[2], fma(-na2, dO[0], dX[1]));
                w[2] = fma(na1, dO[0], fma(-na0, dO[1], dX[2]));
#pragma unroll
                for (int a = 0; a < 3; a++) ft[a] = fma(Ax[3 * a + 2], w[2], fma(Ax[3 * a + 1], w[1], Ax[3 * a] * w[0]));
                const double *fr = sFragL + (s * NT) * 32 + lane;
#pragma unroll
                for (int nt = 0; nt < NT - 1; nt++) {
                    const double bv = fr[nt * 32];
#pragma unroll
                    for (int c = 0; c < 3; c++) dmma884(acc[c][nt][0], acc[c][nt][1], ft[c], bv);
                }
                {
                    const double bv = fr[(NT - 1) * 32];
#pragma unroll
                    for (int c = 0; c < 3; c++) dmma884(acc[(c + 2) % 3][NT - 1][0], acc[(c + 2) % 3][NT - 1][1], ft[c], bv);
                }
            }
        }

        // ---------------- push forward, material, Runge-Kutta stage (my three components of every record) -------------------
        double jm[9];
#pragma unroll
        for (int i = 0; i < 9; i++) jm[i] = ge[i];
        const double dmat = h ? ge[24] : ge[23];            // det/mu : det/eps
        const double se = h ? 0.0 : ge[25];                 // sigma/eps acts on E only
        const bool keep_y = A.has_sigma != 0;
        if (!keep_y) {
            __syncwarp();                                   // every lane has read y_in for the last time
            if (leader && gnext >= 0) issue_y(gnext);
        }
        if (LOAD_X || LOAD_Z) mbar_wait(barXZ, par);
        const bool plain = keep_y || MODE == MODE_MULT;
        const double ca = plain ? A.a : A.a * dmat, cb = plain ? A.b : A.b * dmat;
#pragma unroll
        for (int nt = 0; nt < NT; nt++)
#pragma unroll
            for (int hh = 0; hh < 2; hh++) {
                const int node = nt < NT - 1 ? 8 * nt + j + 4 * hh : 8 * (NT - 1) + j;
                if (nt == NT - 1 && hh == 1) continue;
                if (nt == NT - 1 && NL < 4 && j >= NL) continue;
                double kr[3];
                if (nt < NT - 1) {
#pragma unroll
                    for (int c = 0; c < 3; c++) kr[c] = acc[c][nt][hh];
                } else {
#pragma unroll
                    for (int c = 0; c < 3; c++) kr[c] = acc[(c + 2) % 3][NT - 1][0] + acc[(c + 1) % 3][NT - 1][1];
                }
                double k[3];
#pragma unroll
                for (int d = 0; d < 3; d++) k[d] = fma(jm[3 * d], kr[0], fma(jm[3 * d + 1], kr[1], jm[3 * d + 2] * kr[2]));
                const int off = (e * Np + node) * 6 + own;
                if (plain) {
                    double uo[3] = {0, 0, 0};
                    if (keep_y) load3(wY + off, uo);
#pragma unroll
                    for (int d = 0; d < 3; d++) k[d] = fma(dmat, k[d], -(se * uo[d]));
                }
                double xv[3], zv[3], o[3], zn[3];
                if (LOAD_X) load3(wX + off, xv);
                if (LOAD_Z) load3(wZ + off, zv);
#pragma unroll
                for (int c = 0; c < 3; c++) {
                    if (MODE == MODE_MULT) o[c] = k[c];
                    else if (MODE == MODE_STAGE1) { o[c] = fma(ca, k[c], xv[c]); zn[c] = fma(cb, k[c], xv[c]); }
                    else if (MODE == MODE_STAGE23) { o[c] = fma(ca, k[c], xv[c]); zn[c] = fma(cb, k[c], zv[c]); }
                    else zn[c] = fma(cb, k[c], zv[c]);
                }
                if (STORE_X) store3(wX + off, o);
                if (STORE_Z) store3(wZ + off, zn);
            }
        fence_async_smem();
        __syncwarp();                                       // complete records in wX / wZ
        if (MODE != MODE_MULT && A.pp.signal_epoch != 0 && __any_sync(0xffffffffu, pend_flag != nullptr)) {
            if (pend_flag) { st_release_sys(pend_flag, A.pp.signal_epoch); pend_flag = nullptr; }
        }
        if (MODE != MODE_MULT && A.pp.signal_epoch != 0 && h == 0 && info.x < -1) {   // the E-row lane of (element, face) ships the whole records
            const int2 hp = A.pp.hpush[-2 - info.x];
            const uint4 prow = sTab[hp.x >> 8];
            double *dst = A.pp.peer_out[hp.x & 0xff] + (size_t)hp.y * Nfp * 6;
            const double *src = (MODE == MODE_STAGE4 ? wZ : wX) + e * Np * 6;
#pragma unroll
            for (int m = 0; m < Nfp; m++) {
                double r[6];
                load_rec(src + tab_byte(prow, m) * 6, r);
                store_rec(dst + m * 6, r);
            }
            pend_flag = A.pp.peer_flag[hp.x & 0xff] + hp.y;
        }
        if (lane == 0) {
            const size_t goff = (size_t)g * GS;
            if (STORE_X) bulk_store(A.yout + goff, wX, GS * 8);
            if (MODE == MODE_STAGE4) bulk_store(A.yout + goff, wZ, GS * 8);
            else if (STORE_Z) bulk_store(A.z + goff, wZ, GS * 8);
            bulk_commit();
            if (keep_y && gnext >= 0) issue_y(gnext);
            if (!(LOAD_X || LOAD_Z)) bulk_wait_read();
        }
        __syncwarp();
        g = gnext < 0 ? ngroups : gnext;
    }
    if (leader) bulk_wait_all();
    if (MODE != MODE_MULT && pend_flag) st_release_sys(pend_flag, A.pp.signal_epoch);
}

}  // namespace dgtd
