// sm_100a kernels of the DG-Maxwell evolution hot path (FP64).
//
// One fused kernel per RK stage: volume curl (Dr/Ds/Dt + geometric factors), face flux (vmapM/vmapP jumps,
// centred/upwind, PEC/PMC/SMA, TF/SF injection), LIFT and the Runge-Kutta stage update, so that a stage makes ONE pass
// over HBM (reads y_in, x, z; writes y_out, z).  Reference semantics:
//   RHS      src/evolution/HesthavenEvolution.cpp:450-542 (matrix-free algorithm) with the coefficient conventions of
//            the default `global` operator, src/components/DGOperatorFactory.h:469-573, 1268-1361
//   RK4      external/mfem-geg/linalg/ode.cpp:109-136 (classical tableau; k is never stored)
//   TF/SF    src/evolution/GlobalEvolution.cpp:551-626 (incl. the ||s|| < 1e-8 skip), src/math/Function.h:361-409
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace dgtd {

enum StageMode { MODE_MULT = 0, MODE_STAGE1 = 1, MODE_STAGE23 = 2, MODE_STAGE4 = 3 };

struct DevPlaneWave {
    double inv2s2, mean1d, twopif;     // 1/(2 spread^2), mean, 2 pi freq (0 = plain Gaussian)
    double pe[3], ph[3], dir[3];
};

struct StageArgs {
    // operator
    const double *D;          // [dim][Np(j)][Np(i)]   transposed: D[x][j][i] = d l_j/d xi_x (r_i)
    const double *LIFT;       // [nf*Nfp (m)][Np (i)]  transposed
    const double *geo;        // [NE][16]
    const int2 *finfo;        // [NE][4]
    const uint8_t *ftab;      // [ntab][Nfp]
    const double *tfsf_xyz;   // [nTfsfFaces][Nfp][3]
    const double *gate;       // device scalar: sum of squares of the masked source at this stage time (or null)
    const double *halo;       // [6][hstride]
    int NE;                   // local elements
    long long stride;         // component stride of the state vectors (= NE*Np)
    long long hstride;        // component stride of the halo buffer
    double alpha;
    DevPlaneWave pw;
    int pw_on;
    // vectors (MODE decides which are used)
    const double *yin;        // stage input (neighbours are read from here)
    const double *x;          // step start state
    double *z;                // accumulator
    double *yout;             // MULT: k ; STAGE1/23: next stage input ; STAGE4: new x
    double a, b;              // yout = x + a k ; z (+)= b k ;  STAGE4: yout = z + b k
    double t;                 // stage time (TF/SF)
};

__device__ __forceinline__ void planewave6(const DevPlaneWave &pw, const double *p, double t, double *inc)
{
    // Planewave::eval (Function.h:361-409): g(dir.x - t) * polarisation
    const double u = (p[0] * pw.dir[0] + p[1] * pw.dir[1] + p[2] * pw.dir[2]) - t;
    const double arg = u - pw.mean1d;
    double g = exp(-(arg * arg) * pw.inv2s2);
    if (pw.twopif != 0.0) g *= cos(pw.twopif * arg);
    inc[0] = pw.pe[0] * g; inc[1] = pw.pe[1] * g; inc[2] = pw.pe[2] * g;
    inc[3] = pw.ph[0] * g; inc[4] = pw.ph[1] * g; inc[5] = pw.ph[2] * g;
}

template <int DIM, int P> struct Elem {
    static constexpr int Np = DIM == 1 ? P + 1 : DIM == 2 ? (P + 1) * (P + 2) / 2 : (P + 1) * (P + 2) * (P + 3) / 6;
    static constexpr int Nfp = DIM == 1 ? 1 : DIM == 2 ? P + 1 : (P + 1) * (P + 2) / 2;
    static constexpr int NF = DIM + 1;
    static constexpr int NFN = NF * Nfp;
    // elements per CTA pass: aim at ~192-256 threads
    static constexpr int EB = (256 / Np) < 1 ? 1 : ((256 / Np) > 32 ? 32 : (256 / Np));
    static constexpr int T = EB * Np;
    static constexpr size_t smem_bytes = sizeof(double) * ((size_t)DIM * Np * Np + (size_t)NFN * Np + 6 * (size_t)EB * Np + 6 * (size_t)EB * NFN + (size_t)EB * 16)
                                         + sizeof(int2) * EB * 4 + 256 * Nfp;
};

// ---------------------------------------------------------------------------------------------------------------------
// v1 kernel: thread = (element, node).  Persistent CTAs loop over batches of EB elements.
// ---------------------------------------------------------------------------------------------------------------------
template <int DIM, int P, int MODE>
__global__ void __launch_bounds__(Elem<DIM, P>::T) stage_kernel(const StageArgs A)
{
    using E = Elem<DIM, P>;
    constexpr int Np = E::Np, Nfp = E::Nfp, NF = E::NF, NFN = E::NFN, EB = E::EB, T = E::T;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *sD = reinterpret_cast<double *>(smem_raw);            // DIM*Np*Np
    double *sL = sD + DIM * Np * Np;                              // NFN*Np
    double *su = sL + NFN * Np;                                   // 6*EB*Np
    double *sfl = su + 6 * EB * Np;                               // 6*EB*NFN
    double *sgeo = sfl + 6 * EB * NFN;                            // EB*16
    int2 *sinfo = reinterpret_cast<int2 *>(sgeo + EB * 16);       // EB*4
    uint8_t *stab = reinterpret_cast<uint8_t *>(sinfo + EB * 4);  // ntab*Nfp (<= 256 rows)

    const int tid = threadIdx.x;
    for (int i = tid; i < DIM * Np * Np; i += T) sD[i] = A.D[i];
    for (int i = tid; i < NFN * Np; i += T) sL[i] = A.LIFT[i];
    // ftab rows actually used are few; copy a bounded prefix (host guarantees ntab <= 256)
    {
        const int ntabBytes = 256 * Nfp;
        for (int i = tid; i < ntabBytes; i += T) stab[i] = A.ftab[i];
    }
    const bool inject = A.pw_on && (A.gate == nullptr || *A.gate >= 1e-16);
    const int nbatch = (A.NE + EB - 1) / EB;

    for (int batch = blockIdx.x; batch < nbatch; batch += gridDim.x) {
        const int e0 = batch * EB;
        const int ne = min(EB, A.NE - e0);
        __syncthreads();   // previous pass done with smem (also covers the matrix loads on the first pass)
        if (tid < ne * Np) {
#pragma unroll
            for (int c = 0; c < 6; c++) su[c * EB * Np + tid] = A.yin[c * A.stride + (long long)e0 * Np + tid];
        }
        for (int i = tid; i < ne * 16; i += T) sgeo[i] = A.geo[(long long)e0 * 16 + i];
        for (int i = tid; i < ne * 4; i += T) sinfo[i] = A.finfo[(long long)e0 * 4 + i];
        __syncthreads();

        // ---- face flux -------------------------------------------------------------------------------------
        for (int m = tid; m < ne * NFN; m += T) {
            const int el = m / NFN, fm = m - el * NFN, f = fm / Nfp, j = fm - f * Nfp;
            const int2 info = sinfo[el * 4 + f];
            const int code = info.y;
            const int nself = stab[f * Nfp + j];
            double uM[6], dU[6];
#pragma unroll
            for (int c = 0; c < 6; c++) uM[c] = su[c * EB * Np + el * Np + nself];
            double al = A.alpha;
            if (info.x >= 0) {
                const int nn = stab[((code >> 4) & 0xff) * Nfp + j];
                const int le2 = info.x - e0;
                if (le2 >= 0 && le2 < ne) {
#pragma unroll
                    for (int c = 0; c < 6; c++) dU[c] = su[c * EB * Np + le2 * Np + nn] - uM[c];
                } else {
                    const long long off = (long long)info.x * Np + nn;
#pragma unroll
                    for (int c = 0; c < 6; c++) dU[c] = __ldg(A.yin + c * A.stride + off) - uM[c];
                }
            } else if (info.x == -1) {
                // boundary ghost states (HesthavenEvolution.cpp:275-313 with the global operator's SMA)
                const int bc = code & 3;
                const double ce = bc == 1 ? -2.0 : bc == 3 ? -1.0 : 0.0;
                const double ch = bc == 2 ? -2.0 : bc == 3 ? -1.0 : 0.0;
                if (bc == 3) al = 1.0;
#pragma unroll
                for (int c = 0; c < 3; c++) { dU[c] = ce * uM[c]; dU[3 + c] = ch * uM[3 + c]; }
            } else {
                const long long off = (long long)(-2 - info.x) * Nfp + j;
#pragma unroll
                for (int c = 0; c < 6; c++) dU[c] = A.halo[c * A.hstride + off] - uM[c];
            }
            const int tf = (code >> 2) & 3;
            if (tf && inject) {
                double inc[6];
                planewave6(A.pw, A.tfsf_xyz + ((long long)(code >> 12) * Nfp + j) * 3, A.t, inc);
                const double sg = tf == 1 ? 1.0 : -1.0;
#pragma unroll
                for (int c = 0; c < 6; c++) dU[c] += sg * inc[c];
            }
            // outward normal * fscale = -grad lambda_f
            const double *g = sgeo + el * 16;
            double gn[3];
#pragma unroll
            for (int d = 0; d < 3; d++) {
                if (f == 0) { double s = 0; for (int x = 0; x < DIM; x++) s += g[3 * x + d]; gn[d] = s; }
                else gn[d] = -g[3 * (f - 1) + d];
            }
            const double fs = g[9 + f];
            const double ifs = 1.0 / fs;
            const double gdE = (gn[0] * dU[0] + gn[1] * dU[1] + gn[2] * dU[2]) * ifs * ifs;
            const double gdH = (gn[0] * dU[3] + gn[1] * dU[4] + gn[2] * dU[5]) * ifs * ifs;
            // fscale/2 * ( n x dH + alpha (dE - n (n.dE)) ),  fscale/2 * ( -n x dE + alpha (dH - n (n.dH)) )
            double fl[6];
            const double af = al * fs;
            fl[0] = (gn[1] * dU[5] - gn[2] * dU[4]) + af * (dU[0] - gdE * gn[0]);
            fl[1] = (gn[2] * dU[3] - gn[0] * dU[5]) + af * (dU[1] - gdE * gn[1]);
            fl[2] = (gn[0] * dU[4] - gn[1] * dU[3]) + af * (dU[2] - gdE * gn[2]);
            fl[3] = -(gn[1] * dU[2] - gn[2] * dU[1]) + af * (dU[3] - gdH * gn[0]);
            fl[4] = -(gn[2] * dU[0] - gn[0] * dU[2]) + af * (dU[4] - gdH * gn[1]);
            fl[5] = -(gn[0] * dU[1] - gn[1] * dU[0]) + af * (dU[5] - gdH * gn[2]);
#pragma unroll
            for (int c = 0; c < 6; c++) sfl[c * EB * NFN + m] = 0.5 * fl[c];
        }
        __syncthreads();

        // ---- volume + LIFT + RK update ---------------------------------------------------------------------
        if (tid < ne * Np) {
            const int el = tid / Np, i = tid - el * Np;
            double dr[DIM][6];
#pragma unroll
            for (int x = 0; x < DIM; x++)
#pragma unroll
                for (int c = 0; c < 6; c++) dr[x][c] = 0.0;
            for (int j = 0; j < Np; j++) {
                double u[6];
#pragma unroll
                for (int c = 0; c < 6; c++) u[c] = su[c * EB * Np + el * Np + j];
#pragma unroll
                for (int x = 0; x < DIM; x++) {
                    const double d = sD[(x * Np + j) * Np + i];
#pragma unroll
                    for (int c = 0; c < 6; c++) dr[x][c] = fma(d, u[c], dr[x][c]);
                }
            }
            const double *g = sgeo + el * 16;
            // physical gradient: du_c/dx_d = sum_x Jinv[x][d] * dr[x][c]
            double gr[3][6];
#pragma unroll
            for (int d = 0; d < 3; d++)
#pragma unroll
                for (int c = 0; c < 6; c++) {
                    double s = 0.0;
#pragma unroll
                    for (int x = 0; x < DIM; x++) s = fma(g[3 * x + d], dr[x][c], s);
                    gr[d][c] = s;
                }
            double k[6];
            // dE/dt = curl H ; dH/dt = -curl E
            k[0] = gr[1][5] - gr[2][4];
            k[1] = gr[2][3] - gr[0][5];
            k[2] = gr[0][4] - gr[1][3];
            k[3] = -(gr[1][2] - gr[2][1]);
            k[4] = -(gr[2][0] - gr[0][2]);
            k[5] = -(gr[0][1] - gr[1][0]);
            for (int m = 0; m < NFN; m++) {
                const double l = sL[m * Np + i];
#pragma unroll
                for (int c = 0; c < 6; c++) k[c] = fma(l, sfl[c * EB * NFN + el * NFN + m], k[c]);
            }
            const double ie = g[13], im = g[14], se = g[15];
#pragma unroll
            for (int c = 0; c < 3; c++) {
                k[c] = k[c] * ie - se * su[c * EB * Np + tid];
                k[3 + c] *= im;
            }
            const long long base = (long long)e0 * Np + tid;
#pragma unroll
            for (int c = 0; c < 6; c++) {
                const long long idx = c * A.stride + base;
                if (MODE == MODE_MULT) {
                    A.yout[idx] = k[c];
                } else if (MODE == MODE_STAGE1) {
                    const double xv = su[c * EB * Np + tid];   // yin == x
                    A.yout[idx] = fma(A.a, k[c], xv);
                    A.z[idx] = fma(A.b, k[c], xv);
                } else if (MODE == MODE_STAGE23) {
                    A.yout[idx] = fma(A.a, k[c], A.x[idx]);
                    A.z[idx] = fma(A.b, k[c], A.z[idx]);
                } else {
                    A.yout[idx] = fma(A.b, k[c], A.z[idx]);
                }
            }
        }
    }
}

// masked TF/SF source norm (GlobalEvolution.cpp:584-598): sum over the nodes of TF/SF-adjacent elements of
// |0.5 * planewave|^2 over the six components, for up to 4 stage times at once.
__global__ void gate_kernel(const double *xyz, int V, DevPlaneWave pw, double t0, double t1, double t2, double t3, int nt, double *out)
{
    double s[4] = {0, 0, 0, 0};
    const double ts[4] = {t0, t1, t2, t3};
    const double pn = 0.25 * (pw.pe[0] * pw.pe[0] + pw.pe[1] * pw.pe[1] + pw.pe[2] * pw.pe[2] + pw.ph[0] * pw.ph[0] + pw.ph[1] * pw.ph[1] + pw.ph[2] * pw.ph[2]);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < V; i += gridDim.x * blockDim.x) {
        const double u0 = xyz[3 * i] * pw.dir[0] + xyz[3 * i + 1] * pw.dir[1] + xyz[3 * i + 2] * pw.dir[2];
        for (int q = 0; q < nt; q++) {
            const double arg = (u0 - ts[q]) - pw.mean1d;
            double g = exp(-(arg * arg) * pw.inv2s2);
            if (pw.twopif != 0.0) g *= cos(pw.twopif * arg);
            s[q] += g * g * pn;
        }
    }
    __shared__ double red[4][32];
    for (int q = 0; q < 4; q++) {
        double v = s[q];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0) red[q][threadIdx.x >> 5] = v;
    }
    __syncthreads();
    if (threadIdx.x < 4) {
        double v = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); w++) v += red[threadIdx.x][w];
        if (threadIdx.x < nt) atomicAdd(out + threadIdx.x, v);
    }
}

__global__ void sumsq_kernel(const double *x, long long n, double *out)
{
    double s = 0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) s = fma(x[i], x[i], s);
    __shared__ double red[32];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double v = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); w++) v += red[w];
        atomicAdd(out, v);
    }
}

// halo pack: send[c][s] = y[c][send_node[s]]
// Probe / surface-export gather: out[c][i] = component c of the scalar dof whose component-0 value sits at off[i]
// (component stride cstride: 1 in the record layouts, Nloc in the reference layout).
__global__ void gather_kernel(const double *x, const long long *off, long long cstride, long long n, double *out)
{
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const double *p = x + off[i];
#pragma unroll
        for (int c = 0; c < 6; c++) out[c * n + i] = p[c * cstride];
    }
}
// reference-layout vectors [6][N] in the caller's element order <-> in this rank's (Morton) element order; to_local = 1:
// out[c][le * Np + n] = in[c][gid[le] * Np + n], 0: the inverse
__global__ void permute_elements_kernel(const double *in, const int *gid, int Np, long long NEloc, long long stride, int to_local, double *out)
{
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < NEloc * Np; idx += (long long)gridDim.x * blockDim.x) {
        const long long le = idx / Np; const int n = (int)(idx - le * Np);
        const long long other = (long long)gid[le] * Np + n;
#pragma unroll
        for (int c = 0; c < 6; c++) {
            if (to_local) out[c * stride + idx] = in[c * stride + other];
            else out[c * stride + other] = in[c * stride + idx];
        }
    }
}
__global__ void pack_kernel(const double *y, long long stride, const int *send_node, int ns, double *send, long long sstride)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < ns; i += gridDim.x * blockDim.x) {
        const int n = send_node[i];
#pragma unroll
        for (int c = 0; c < 6; c++) send[c * sstride + i] = y[c * stride + n];
    }
}

// point probes: out[p][c] = sum_i shape[p][i] * y[c][elem[p]*Np + i]
__global__ void sample_kernel(const double *y, long long stride, int Np, int npts, const int *elem, const double *shape, double *out)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npts) return;
    for (int c = 0; c < 6; c++) {
        double s = 0;
        for (int i = 0; i < Np; i++) s = fma(shape[(long long)p * Np + i], y[c * stride + (long long)elem[p] * Np + i], s);
        out[p * 6 + c] = s;
    }
}

}  // namespace dgtd
