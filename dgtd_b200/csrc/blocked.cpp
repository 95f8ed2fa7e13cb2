// Host-side plan of the DMMA ("blocked") stage kernel: element groups, batches, trace-gather descriptors, operator
// fragments.  Pure data re-arrangement of HostOp (setup.cpp); the arithmetic it prepares is the reference's
//   per-element D_x/D_y/D_z, normals, fscale, LIFT     src/evolution/HesthavenEvolution.cpp:150-205, 56-81
//   vmapM/vmapP                                         src/evolution/HesthavenEvolutionMethods.cpp:501-534
// expressed for tensor-core tiles: the volume term is evaluated as a REFERENCE-space curl of the covariant field
// (curl_x u = J curl_xi(J^T u) / det J for an affine element), which needs 12 instead of 18 matrix-vector products.
#include "host.hpp"
#include "../../include/dgtd_b200.h"

#include <algorithm>

namespace dgtd {

BlockedPlan build_blocked_plan(const HostOp &H, int G)
{
    if (H.dim != 3) throw Error(DGTD_ERR_UNSUPPORTED, "the blocked kernel covers tetrahedra only");
    BlockedPlan B;
    const int Np = H.Np, Nfp = H.Nfp, NE = H.NEloc;
    B.G = G;
    const int EB = BLK_E * G;
    B.nbatch = (NE + EB - 1) / EB;
    B.ngroups = B.nbatch * G;
    B.NEpad = B.ngroups * BLK_E;
    B.slots = 4 * EB;
    B.MT = (Np + 7) / 8;
    B.KSV = (Np + 3) / 4;
    B.KSL = (4 * Nfp + 3) / 4;

    // ---- geometry records ------------------------------------------------------------------------------------
    B.geo.assign((size_t)B.NEpad * BLK_GEO, 0.0);
    for (int e = 0; e < B.NEpad; e++) {
        double *g = &B.geo[(size_t)e * BLK_GEO];
        if (e < NE) {
            const double *v1 = &H.geo[(size_t)e * GEO_STRIDE], *jc = &H.jac[(size_t)e * 10];
            for (int i = 0; i < 9; i++) { g[i] = jc[i]; g[9 + i] = v1[i]; }
            for (int f = 0; f < 4; f++) g[18 + f] = v1[9 + f];
            g[22] = 1.0 / jc[9]; g[23] = v1[13]; g[24] = v1[14]; g[25] = v1[15];
        } else {   // padding element: unit geometry, vacuum; its state stays zero
            g[0] = g[4] = g[8] = 1.0; g[9] = g[13] = g[17] = 1.0;
            g[18] = g[19] = g[20] = g[21] = 1.0; g[22] = g[23] = g[24] = 1.0;
        }
    }
    // ---- faces: in-batch neighbour, boundary, or a trace slot filled by the prefetch -------------------------------
    // per batch one descriptor block: finfo int2[EB*4] {>=0 in-batch element | -1 boundary | -2-slot, code},
    // tdesc int2[slots] {>=0 source local element | -1-haloFace, ftab row}, used-slot count, 3 pad ints
    B.desc_stride = EB * 8 + B.slots * 2 + 4;
    B.desc.assign((size_t)B.nbatch * B.desc_stride, 0);
    for (int e = 0; e < B.NEpad; e++) {
        const int b = e / EB;
        int *blk = &B.desc[(size_t)b * B.desc_stride];
        int *tcount = blk + EB * 8 + B.slots * 2;
        for (int f = 0; f < 4; f++) {
            int *fo = blk + ((size_t)(e - b * EB) * 4 + f) * 2;
            if (e >= NE) { fo[0] = -1; fo[1] = f << FI_TAB_SHIFT; continue; }   // boundary with BC none: zero jump
            const int nb = H.finfo[((size_t)e * 4 + f) * 2], code = H.finfo[((size_t)e * 4 + f) * 2 + 1];
            fo[1] = code;
            if (nb == -1) fo[0] = -1;
            else if (nb >= 0 && nb / EB == b) fo[0] = nb - b * EB;
            else {
                const int s = (*tcount)++;
                int *td = blk + EB * 8 + (size_t)s * 2;
                td[0] = nb >= 0 ? nb : -1 - (-2 - nb);      // local element, or -1-haloFace
                td[1] = (code >> FI_TAB_SHIFT) & FI_TAB_MASK;
                fo[0] = -2 - s;
            }
        }
    }
    // ---- DMMA A fragments (m8n8k4: lane l holds A[row l>>2][col l&3]) ------------------------------------------
    const int MT = B.MT, KSV = B.KSV, KSL = B.KSL;
    B.afrag.assign(((size_t)3 * MT * KSV + (size_t)MT * KSL) * 32, 0.0);
    for (int x = 0; x < 3; x++) for (int mt = 0; mt < MT; mt++) for (int ks = 0; ks < KSV; ks++) for (int l = 0; l < 32; l++) {
        const int i = mt * 8 + (l >> 2), j = ks * 4 + (l & 3);
        if (i < Np && j < Np) B.afrag[(((size_t)x * MT + mt) * KSV + ks) * 32 + l] = H.ref.D[((size_t)x * Np + i) * Np + j];
    }
    const size_t lbase = (size_t)3 * MT * KSV * 32;
    for (int mt = 0; mt < MT; mt++) for (int ks = 0; ks < KSL; ks++) for (int l = 0; l < 32; l++) {
        const int i = mt * 8 + (l >> 2), m = ks * 4 + (l & 3);
        if (i < Np && m < 4 * Nfp) {
            const int f = m / Nfp, j = m - f * Nfp;
            B.afrag[lbase + ((size_t)mt * KSL + ks) * 32 + l] = 0.5 * H.ref.lift[((size_t)f * Np + i) * Nfp + j];   // the 1/2 of applyLIFT (exact)
        }
    }
    // ---- halo pack list ----------------------------------------------------------------------------------------
    B.send_off.resize(H.send_node.size());
    for (size_t s = 0; s < H.send_node.size(); s++) B.send_off[s] = blocked_offset(Np, H.send_node[s] / Np, H.send_node[s] % Np);
    return B;
}

}  // namespace dgtd
