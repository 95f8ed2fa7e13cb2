// Host-side plan of the warp-per-group DMMA stage kernel (kernels_wg.cuh): element-major state layout, per-face
// descriptors, neighbour node tables and the operator B-fragments of the "transposed" contraction.
// Pure data re-arrangement of HostOp (setup.cpp); the arithmetic it prepares is the reference's
//   per-element D_x/D_y/D_z, normals, fscale, LIFT     src/evolution/HesthavenEvolution.cpp:150-205, 56-81
//   vmapM/vmapP                                         src/evolution/HesthavenEvolutionMethods.cpp:501-534
// with the volume term evaluated as a reference-space curl of the covariant field (see blocked.cpp).
#include "host.hpp"
#include "../../include/dgtd_b200.h"

#include <algorithm>
#include <cstdlib>
#include <functional>
#include <string>
#include <map>

namespace dgtd {

WgPlan build_wg_plan(const HostOp &H)
{
    if (H.dim != 3) throw Error(DGTD_ERR_UNSUPPORTED, "the warp-per-group kernel covers tetrahedra only");
    WgPlan W;
    const int Np = H.Np, Nfp = H.Nfp, NE = H.NEloc;
    if (Nfp > 16) throw Error(DGTD_ERR_UNSUPPORTED, "face node rows are 16 bytes");
    W.ngroups = (NE + BLK_E - 1) / BLK_E;
    W.NEpad = W.ngroups * BLK_E;
    const int NT = W.NT = (Np + 7) / 8, KSV = W.KSV = (Np + 3) / 4;
    if (Np - 8 * (NT - 1) > 4) throw Error(DGTD_ERR_UNSUPPORTED, "the mixed last tile holds at most 4 output nodes");

    // ---- device node numbering and per-face step order ------------------------------------------------------------------
    // default ("file"): the reference element's numbering, face nodes in ascending order.
    // DGTD_B200_FACE_ORDER=bank: at face step s the four lanes (e, j = 0..3) of an element gather the node records
    // tab[j][s]; with 48-byte records an LDS.128 of a quarter warp (two elements x four faces) is conflict-free exactly when
    // the four node ids are distinct mod 4.  So the nodes are renumbered into four residue classes whose face incidences sum
    // to Nfp each, and every step takes one node of each class (a perfect matching of faces and classes; it exists at every
    // step because the face x class incidence counts form a regular bipartite multigraph).
    W.dev2ref.resize(Np); W.ref2dev.resize(Np);
    for (int n = 0; n < Np; n++) W.dev2ref[n] = W.ref2dev[n] = n;
    W.forder.resize((size_t)4 * Nfp);
    for (int f = 0; f < 4; f++) for (int s = 0; s < Nfp; s++) W.forder[(size_t)f * Nfp + s] = s;
    const char *fo_env = std::getenv("DGTD_B200_FACE_ORDER");
    if (fo_env && std::string(fo_env) == "bank") {
        std::vector<int> mult(Np, 0);
        for (int f = 0; f < 4; f++) for (int s = 0; s < Nfp; s++) mult[H.ref.fnodes[(size_t)f * Nfp + s]]++;
        // how many nodes of each face multiplicity (3: vertices, 2: edge nodes, 1: face nodes, 0: interior) go into each
        // residue class: class r holds cap[r] nodes whose multiplicities sum to Nfp (a search over counts, not over nodes)
        int cap[4], have[4] = {0, 0, 0, 0}, x[4][4];
        for (int r = 0; r < 4; r++) cap[r] = (Np - r + 3) / 4;
        for (int n = 0; n < Np; n++) have[mult[n]]++;
        std::function<bool(int)> fill = [&](int r) {
            if (r == 4) return have[0] == 0 && have[1] == 0 && have[2] == 0 && have[3] == 0;
            for (int a = 0; a <= have[3]; a++)
                for (int b = 0; b <= have[2]; b++) {
                    const int c = Nfp - 3 * a - 2 * b, d = cap[r] - a - b - c;
                    if (c < 0 || c > have[1] || d < 0 || d > have[0]) continue;
                    x[r][3] = a; x[r][2] = b; x[r][1] = c; x[r][0] = d;
                    have[3] -= a; have[2] -= b; have[1] -= c; have[0] -= d;
                    if (fill(r + 1)) return true;
                    have[3] += a; have[2] += b; have[1] += c; have[0] += d;
                }
            return false;
        };
        std::vector<int> cls(Np, -1);
        auto place = [&](int) {
            if (!fill(0)) return false;
            for (int r = 0; r < 4; r++)
                for (int m = 0; m < 4; m++)
                    for (int n = 0; n < Np && x[r][m] > 0; n++)
                        if (cls[n] < 0 && mult[n] == m) { cls[n] = r; x[r][m]--; }
            return true;
        };
        if (place(0)) {
            int next[4] = {0, 1, 2, 3};
            for (int n = 0; n < Np; n++) { W.ref2dev[n] = next[cls[n]]; W.dev2ref[next[cls[n]]] = n; next[cls[n]] += 4; }
            std::vector<char> used((size_t)4 * Nfp, 0);
            int cnt[4][4] = {};
            for (int f = 0; f < 4; f++) for (int s = 0; s < Nfp; s++) cnt[f][cls[H.ref.fnodes[(size_t)f * Nfp + s]]]++;
            for (int s = 0; s < Nfp; s++) {
                int perm[4] = {0, 1, 2, 3}, pick[4] = {-1, -1, -1, -1};
                do {
                    bool ok = true;
                    for (int f = 0; f < 4; f++) ok &= cnt[f][perm[f]] > 0;
                    if (ok) { for (int f = 0; f < 4; f++) pick[f] = perm[f]; break; }
                } while (std::next_permutation(perm, perm + 4));
                if (pick[0] < 0) throw Error(DGTD_ERR_UNSUPPORTED, "face-step matching failed");
                for (int f = 0; f < 4; f++) {
                    cnt[f][pick[f]]--;
                    for (int m = 0; m < Nfp; m++)
                        if (!used[(size_t)f * Nfp + m] && cls[H.ref.fnodes[(size_t)f * Nfp + m]] == pick[f]) { used[(size_t)f * Nfp + m] = 1; W.forder[(size_t)f * Nfp + s] = m; break; }
                }
            }
        }
    }

    // ---- geometry records (same record as the blocked plan) --------------------------------------------------------------
    W.geo.assign((size_t)W.NEpad * WG_GEO, 0.0);
    for (int e = 0; e < W.NEpad; e++) {
        double *g = &W.geo[(size_t)e * WG_GEO];
        if (e < NE) {
            const double *v1 = &H.geo[(size_t)e * GEO_STRIDE], *jc = &H.jac[(size_t)e * 10];
            // J / det J (covariant transform and push-forward share it), J^-1, fscale, 1/det, det/eps, det/mu, sigma/eps, 1/fscale
            for (int i = 0; i < 9; i++) { g[i] = jc[i] / jc[9]; g[9 + i] = v1[i]; }
            for (int f = 0; f < 4; f++) g[18 + f] = v1[9 + f];
            g[22] = 1.0 / jc[9]; g[23] = jc[9] * v1[13]; g[24] = jc[9] * v1[14]; g[25] = v1[15];
            for (int f = 0; f < 4; f++) g[26 + f] = 1.0 / v1[9 + f];
        } else {   // padding element: unit geometry, vacuum; its state stays zero
            g[0] = g[4] = g[8] = 1.0; g[9] = g[13] = g[17] = 1.0;
            g[18] = g[19] = g[20] = g[21] = 1.0; g[22] = g[23] = g[24] = 1.0; g[26] = g[27] = g[28] = g[29] = 1.0;
        }
    }

    // ---- node tables: 16-byte rows indexed by the step s of a face ---------------------------------------------------------
    auto put_row = [&](const std::vector<uint8_t> &row) { W.tab.insert(W.tab.end(), row.begin(), row.end()); return W.ntab++; };
    for (int f = 0; f < 4; f++) {   // rows 0..3: own device node
        std::vector<uint8_t> row(16, 0);
        for (int s = 0; s < Nfp; s++) row[s] = (uint8_t)W.ref2dev[H.ref.fnodes[(size_t)f * Nfp + W.forder[(size_t)f * Nfp + s]]];
        put_row(row);
    }
    for (int f = 0; f < 4; f++) {   // rows 4..7: canonical face-node index (halo buffer, TF/SF coordinates)
        std::vector<uint8_t> row(16, 0);
        for (int s = 0; s < Nfp; s++) row[s] = (uint8_t)W.forder[(size_t)f * Nfp + s];
        put_row(row);
    }
    std::map<std::pair<int, int>, int> rowOf;   // (my face, HostOp::ftab row) -> row here
    auto nbr_row = [&](int f, int old) {
        auto it = rowOf.find({f, old});
        if (it != rowOf.end()) return it->second;
        std::vector<uint8_t> row(16, 0);
        for (int s = 0; s < Nfp; s++) row[s] = (uint8_t)W.ref2dev[H.ftab[(size_t)old * Nfp + W.forder[(size_t)f * Nfp + s]]];
        const int id = put_row(row);
        rowOf[{f, old}] = id;
        return id;
    };

    // ---- face descriptors ---------------------------------------------------------------------------------------------------
    W.desc.assign((size_t)W.NEpad * 8, 0);
    for (int e = 0; e < W.NEpad; e++)
        for (int f = 0; f < 4; f++) {
            int *fo = &W.desc[((size_t)e * 4 + f) * 2];
            if (e >= NE) { fo[0] = -1; fo[1] = 0; continue; }   // boundary with BC none: zero jump
            const int nb = H.finfo[((size_t)e * 4 + f) * 2];
            int code = H.finfo[((size_t)e * 4 + f) * 2 + 1];
            const int old = (code >> FI_TAB_SHIFT) & FI_TAB_MASK;
            int row = nb >= 0 ? nbr_row(f, old) : nb == -1 ? f : 4 + f;
            if (row > FI_TAB_MASK) throw Error(DGTD_ERR_UNSUPPORTED, "too many distinct face orientations");
            code = (code & ~(FI_TAB_MASK << FI_TAB_SHIFT)) | (row << FI_TAB_SHIFT);
            fo[0] = nb; fo[1] = code;
        }

    // ---- DMMA B fragments (m8n8k4: lane l holds B[k = l&3][n = l>>2]) ------------------------------------------------------
    // Output column q of tile nt is device node 8nt + (q>>1) + 4(q&1): the accumulator pair of lane (e, j) then holds the
    // nodes 8nt + j and 8nt + j + 4.  The last tile is "mixed": its even columns hold the node 8(NT-1) + (q>>1) of one
    // operator and its odd columns the same node of another one, so no DMMA column is spent on padding:
    //   volume, input component x:  even = +D_{x+2}, odd = -D_{x+1}   (k~_c = D_{c+1} u~_{c+2} - D_{c+2} u~_{c+1})
    //   LIFT:                       even = LIFT/2,   odd = 0
    const int VT = (NT - 1) * 3 + 3;
    W.nfrag_vol = KSV * VT; W.nfrag_lift = Nfp * NT;
    W.bfrag.assign((size_t)(W.nfrag_vol + W.nfrag_lift) * 32, 0.0);
    auto D = [&](int d, int outDev, int inDev) { return H.ref.D[((size_t)d * Np + W.dev2ref[outDev]) * Np + W.dev2ref[inDev]]; };
    for (int ks = 0; ks < KSV; ks++)
        for (int l = 0; l < 32; l++) {
            const int in = 4 * ks + (l & 3), q = l >> 2;
            if (in >= Np) continue;
            for (int nt = 0; nt < NT - 1; nt++)
                for (int d = 0; d < 3; d++)
                    W.bfrag[((size_t)ks * VT + nt * 3 + d) * 32 + l] = D(d, 8 * nt + (q >> 1) + 4 * (q & 1), in);
            const int out = 8 * (NT - 1) + (q >> 1);
            if (out < Np)
                for (int x = 0; x < 3; x++)
                    W.bfrag[((size_t)ks * VT + (NT - 1) * 3 + x) * 32 + l] = (q & 1) ? -D((x + 1) % 3, out, in) : D((x + 2) % 3, out, in);
        }
    const size_t lbase = (size_t)W.nfrag_vol * 32;
    for (int s = 0; s < Nfp; s++)
        for (int l = 0; l < 32; l++) {
            const int f = l & 3, q = l >> 2, m = W.forder[(size_t)f * Nfp + s];
            auto L = [&](int outDev) { return 0.5 * H.ref.lift[((size_t)f * Np + W.dev2ref[outDev]) * Nfp + m]; };   // the 1/2 of applyLIFT (exact)
            for (int nt = 0; nt < NT - 1; nt++) W.bfrag[lbase + ((size_t)s * NT + nt) * 32 + l] = L(8 * nt + (q >> 1) + 4 * (q & 1));
            const int out = 8 * (NT - 1) + (q >> 1);
            if (out < Np && !(q & 1)) W.bfrag[lbase + ((size_t)s * NT + NT - 1) * 32 + l] = L(out);
        }

    // ---- direct halo push: what the face behind halo slot s sends, in the receiver's face-node order, and where ------------
    W.hpush.assign((size_t)H.n_halo_faces * 2, 0);
    {
        std::map<std::vector<uint8_t>, int> pushRow;
        for (size_t pi = 0; pi < H.peers.size(); pi++) {
            const PeerPlan &pp = H.peers[pi];
            for (int s = pp.send_off; s < pp.send_off + pp.nfaces; s++) {
                std::vector<uint8_t> row(16, 0);
                for (int m = 0; m < Nfp; m++) row[m] = (uint8_t)W.ref2dev[H.send_node[(size_t)s * Nfp + m] % Np];
                auto it = pushRow.find(row);
                const int id = it != pushRow.end() ? it->second : (pushRow[row] = put_row(row));
                W.hpush[(size_t)s * 2] = (int)pi | (id << 8);
                W.hpush[(size_t)s * 2 + 1] = pp.remote_off + (s - pp.send_off);
            }
        }
    }
    // ---- halo pack list ---------------------------------------------------------------------------------------------------
    W.send_off.resize(H.send_node.size());
    for (size_t s = 0; s < H.send_node.size(); s++)
        W.send_off[s] = ((long long)(H.send_node[s] / Np) * Np + W.ref2dev[H.send_node[s] % Np]) * 6;
    return W;
}

}  // namespace dgtd
