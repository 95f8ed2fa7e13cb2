// Reference-element operators on the unit simplex for MFEM's L2 Gauss-Lobatto node set.
//   node set / ordering : external/mfem-geg/fem/fe/fe_l2.cpp:23-37 (segment), :569-592 (triangle), :716-722 (tet)
//   what they replace   : refInvMass / refLIFT_ / per-element D_x,D_y,D_z of
//                         src/evolution/HesthavenEvolution.cpp:48-81, 150-170, 365 (assembled there through MFEM
//                         bilinear forms on one-element sub-meshes)
// Derived from scratch: Legendre-product basis of P_p, orthonormalised numerically (Gram + Cholesky with an exact
// Duffy/Gauss rule), collocation derivatives D = V_xi V^-1, M^-1 = V V^T, LIFT_f = M^-1 * (face mass).  All in long double.
#include "host.hpp"

#include <algorithm>
#include <cmath>

namespace dgtd {
namespace {
using real = long double;
using Mat = std::vector<real>;   // row-major

void legendre(int n, real t, real *P, real *dP)   // P_k(t), P_k'(t), k = 0..n
{
    P[0] = 1; dP[0] = 0;
    if (n >= 1) { P[1] = t; dP[1] = 1; }
    for (int k = 1; k < n; k++) {
        P[k + 1] = ((2 * k + 1) * t * P[k] - k * P[k - 1]) / (k + 1);
        dP[k + 1] = dP[k - 1] + (2 * k + 1) * P[k];
    }
}

void gauss_legendre01(int n, std::vector<real> &x, std::vector<real> &w)
{
    x.resize(n); w.resize(n);
    std::vector<real> P(n + 1), dP(n + 1);
    const real pi = acosl(-1.0L);
    for (int i = 0; i < n; i++) {
        real t = -cosl(pi * (i + 0.75L) / (n + 0.5L));
        for (int it = 0; it < 100; it++) {
            legendre(n, t, P.data(), dP.data());
            real dt = P[n] / dP[n];
            t -= dt;
            if (fabsl(dt) < 1e-19L) break;
        }
        legendre(n, t, P.data(), dP.data());
        x[i] = 0.5L * (t + 1);
        w[i] = 1.0L / ((1 - t * t) * dP[n] * dP[n]);   // = 0.5 * 2/((1-t^2) P'^2)
    }
}

std::vector<real> gll01_real(int p)
{
    if (p == 0) return {0.5L};
    std::vector<real> x(p + 1);
    x[0] = 0; x[p] = 1;
    std::vector<real> P(p + 1), dP(p + 1);
    const real pi = acosl(-1.0L);
    for (int i = 1; i < p; i++) {
        real t = -cosl(pi * i / p);                   // Chebyshev-Lobatto start, Newton on q(t) = P'_p(t)
        for (int it = 0; it < 100; it++) {
            legendre(p, t, P.data(), dP.data());
            // P''_p from the Legendre ODE: (1-t^2) P'' = 2 t P' - p(p+1) P
            real d2 = (2 * t * dP[p] - p * (p + 1) * P[p]) / (1 - t * t);
            real dt = dP[p] / d2;
            t -= dt;
            if (fabsl(dt) < 1e-19L) break;
        }
        x[i] = 0.5L * (t + 1);
    }
    for (int i = 0; i <= p / 2; i++) {                // symmetrise
        real a = 0.5L * (x[i] + (1 - x[p - i]));
        x[i] = a; x[p - i] = 1 - a;
    }
    return x;
}

// simplex quadrature (Duffy), exact for total degree `deg`
void simplex_quadrature(int dim, int deg, std::vector<real> &pts, std::vector<real> &w)
{
    pts.clear(); w.clear();
    if (dim == 0) { w.push_back(1); return; }
    int n = deg / 2 + dim + 1;
    std::vector<real> t, tw;
    gauss_legendre01(n, t, tw);
    if (dim == 1) { for (int i = 0; i < n; i++) { pts.push_back(t[i]); w.push_back(tw[i]); } return; }
    if (dim == 2) {
        for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) {
            pts.push_back(t[i]); pts.push_back(t[j] * (1 - t[i]));
            w.push_back(tw[i] * tw[j] * (1 - t[i]));
        }
        return;
    }
    for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) for (int k = 0; k < n; k++) {
        pts.push_back(t[i]); pts.push_back(t[j] * (1 - t[i])); pts.push_back(t[k] * (1 - t[i]) * (1 - t[j]));
        w.push_back(tw[i] * tw[j] * tw[k] * (1 - t[i]) * (1 - t[i]) * (1 - t[j]));
    }
}

struct RawBasis {
    int dim, p, nb;
    std::vector<std::array<int, 3>> exps;
    RawBasis(int d, int pp) : dim(d), p(pp) {
        for (int c = 0; c <= (d > 2 ? p : 0); c++)
            for (int b = 0; b <= (d > 1 ? p : 0); b++)
                for (int a = 0; a <= p; a++)
                    if (a + b + c <= p) exps.push_back({a, b, c});
        nb = (int)exps.size();
    }
    // phi[m], dphi[d][m] at point x
    void eval(const real *x, real *phi, real *dphi) const {
        real P[3][16], dP[3][16];
        for (int d = 0; d < dim; d++) {
            legendre(p, 2 * x[d] - 1, P[d], dP[d]);
            for (int k = 0; k <= p; k++) dP[d][k] *= 2;
        }
        for (int m = 0; m < nb; m++) {
            real v = 1;
            for (int d = 0; d < dim; d++) v *= P[d][exps[m][d]];
            phi[m] = v;
            for (int dd = 0; dd < dim; dd++) {
                real g = 1;
                for (int d = 0; d < dim; d++) g *= (d == dd) ? dP[d][exps[m][d]] : P[d][exps[m][d]];
                dphi[dd * nb + m] = g;
            }
        }
    }
};

Mat matmul(const Mat &A, const Mat &B, int n, int k, int m)   // (n x k) * (k x m)
{
    Mat C((size_t)n * m, 0);
    for (int i = 0; i < n; i++) for (int l = 0; l < k; l++) {
        real a = A[(size_t)i * k + l];
        for (int j = 0; j < m; j++) C[(size_t)i * m + j] += a * B[(size_t)l * m + j];
    }
    return C;
}
Mat transpose(const Mat &A, int n, int m)
{
    Mat T((size_t)n * m);
    for (int i = 0; i < n; i++) for (int j = 0; j < m; j++) T[(size_t)j * n + i] = A[(size_t)i * m + j];
    return T;
}
Mat inverse(Mat A, int n)   // Gauss-Jordan with partial pivoting
{
    Mat I((size_t)n * n, 0);
    for (int i = 0; i < n; i++) I[(size_t)i * n + i] = 1;
    for (int c = 0; c < n; c++) {
        int piv = c;
        for (int r = c + 1; r < n; r++) if (fabsl(A[(size_t)r * n + c]) > fabsl(A[(size_t)piv * n + c])) piv = r;
        if (A[(size_t)piv * n + c] == 0) throw Error(-4, "singular matrix in reference element setup");
        if (piv != c) for (int j = 0; j < n; j++) { std::swap(A[(size_t)c * n + j], A[(size_t)piv * n + j]); std::swap(I[(size_t)c * n + j], I[(size_t)piv * n + j]); }
        real d = 1 / A[(size_t)c * n + c];
        for (int j = 0; j < n; j++) { A[(size_t)c * n + j] *= d; I[(size_t)c * n + j] *= d; }
        for (int r = 0; r < n; r++) if (r != c) {
            real f = A[(size_t)r * n + c];
            if (f == 0) continue;
            for (int j = 0; j < n; j++) { A[(size_t)r * n + j] -= f * A[(size_t)c * n + j]; I[(size_t)r * n + j] -= f * I[(size_t)c * n + j]; }
        }
    }
    return I;
}
Mat cholesky(const Mat &G, int n)   // lower L with G = L L^T
{
    Mat L((size_t)n * n, 0);
    for (int i = 0; i < n; i++) for (int j = 0; j <= i; j++) {
        real s = G[(size_t)i * n + j];
        for (int k = 0; k < j; k++) s -= L[(size_t)i * n + k] * L[(size_t)j * n + k];
        if (i == j) { if (s <= 0) throw Error(-4, "Gram matrix not positive definite"); L[(size_t)i * n + i] = sqrtl(s); }
        else L[(size_t)i * n + j] = s / L[(size_t)j * n + j];
    }
    return L;
}
}  // namespace

std::vector<double> gll01(int p)
{
    auto g = gll01_real(p);
    return std::vector<double>(g.begin(), g.end());
}

int RefElem::lookup(const int *b) const
{
    int key = 0, mul = 1;
    for (int k = 1; k <= dim; k++) { key += b[k] * mul; mul *= (p + 1); }
    return node_of[key];
}

RefElem build_ref_element(int dim, int p)
{
    if (dim < 1 || dim > 3 || p < 1 || p > 8) throw Error(-4, "unsupported dimension/order");
    RefElem R;
    R.dim = dim; R.p = p; R.nf = dim + 1;
    // index tuples, MFEM ordering: last index outermost
    std::vector<std::array<int, 3>> idx;
    if (dim == 1) for (int i = 0; i <= p; i++) idx.push_back({i, 0, 0});
    if (dim == 2) for (int j = 0; j <= p; j++) for (int i = 0; i + j <= p; i++) idx.push_back({i, j, 0});
    if (dim == 3) for (int k = 0; k <= p; k++) for (int j = 0; j + k <= p; j++) for (int i = 0; i + j + k <= p; i++) idx.push_back({i, j, k});
    const int Np = R.Np = (int)idx.size();
    auto g = gll01_real(p);
    std::vector<real> nodes((size_t)Np * dim);
    R.bary.resize((size_t)Np * (dim + 1));
    int mul = 1; for (int k = 0; k < dim; k++) mul *= (p + 1);
    R.node_of.assign(mul, -1);
    for (int n = 0; n < Np; n++) {
        int s = 0; for (int d = 0; d < dim; d++) s += idx[n][d];
        R.bary[(size_t)n * (dim + 1)] = p - s;
        for (int d = 0; d < dim; d++) R.bary[(size_t)n * (dim + 1) + 1 + d] = idx[n][d];
        if (dim == 1) nodes[n] = g[idx[n][0]];
        else {
            real w = g[p - s]; for (int d = 0; d < dim; d++) w += g[idx[n][d]];
            for (int d = 0; d < dim; d++) nodes[(size_t)n * dim + d] = g[idx[n][d]] / w;
        }
        int key = 0, m2 = 1; for (int d = 0; d < dim; d++) { key += idx[n][d] * m2; m2 *= (p + 1); }
        R.node_of[key] = n;
    }
    RawBasis B(dim, p);
    if (B.nb != Np) throw Error(-4, "basis size mismatch");
    // Gram matrix and orthonormalisation
    std::vector<real> qx, qw;
    simplex_quadrature(dim, 2 * p, qx, qw);
    const int nq = (int)qw.size();
    Mat G((size_t)Np * Np, 0);
    std::vector<real> phi(Np), dphi((size_t)3 * Np);
    for (int q = 0; q < nq; q++) {
        B.eval(&qx[(size_t)q * dim], phi.data(), dphi.data());
        for (int a = 0; a < Np; a++) for (int b = 0; b < Np; b++) G[(size_t)a * Np + b] += qw[q] * phi[a] * phi[b];
    }
    Mat L = cholesky(G, Np);
    Mat LinvT = transpose(inverse(L, Np), Np, Np);
    // Vandermonde of the orthonormal basis at the nodes and of its derivatives
    Mat Vraw((size_t)Np * Np), dVraw[3];
    for (int d = 0; d < dim; d++) dVraw[d].resize((size_t)Np * Np);
    for (int n = 0; n < Np; n++) {
        B.eval(&nodes[(size_t)n * dim], phi.data(), dphi.data());
        for (int m = 0; m < Np; m++) {
            Vraw[(size_t)n * Np + m] = phi[m];
            for (int d = 0; d < dim; d++) dVraw[d][(size_t)n * Np + m] = dphi[(size_t)d * Np + m];
        }
    }
    Mat V = matmul(Vraw, LinvT, Np, Np, Np);
    Mat Vinv = inverse(V, Np);
    Mat Minv = matmul(V, transpose(V, Np, Np), Np, Np, Np);
    R.D.resize((size_t)dim * Np * Np);
    for (int d = 0; d < dim; d++) {
        Mat Dd = matmul(matmul(dVraw[d], LinvT, Np, Np, Np), Vinv, Np, Np, Np);
        for (size_t i = 0; i < Dd.size(); i++) R.D[(size_t)d * Np * Np + i] = (double)Dd[i];
    }
    R.Minv.assign(Minv.begin(), Minv.end());
    R.nodes.assign(nodes.begin(), nodes.end());
    // faces
    const int nf = R.nf;
    std::vector<std::vector<int>> fn(nf);
    for (int f = 0; f < nf; f++) for (int n = 0; n < Np; n++) if (R.bary[(size_t)n * (dim + 1) + f] == 0) fn[f].push_back(n);
    const int Nfp = R.Nfp = (int)fn[0].size();
    R.fnodes.resize((size_t)nf * Nfp);
    for (int f = 0; f < nf; f++) { if ((int)fn[f].size() != Nfp) throw Error(-4, "face node count"); for (int j = 0; j < Nfp; j++) R.fnodes[(size_t)f * Nfp + j] = fn[f][j]; }
    std::vector<real> fq, fw;
    simplex_quadrature(dim - 1, 2 * p, fq, fw);
    R.lift.assign((size_t)nf * Np * Nfp, 0.0);
    Mat LV = matmul(LinvT, Vinv, Np, Np, Np);   // raw basis values -> Lagrange values
    for (int f = 0; f < nf; f++) {
        real vx[4][3] = {{0, 0, 0}, {1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
        int fv[3], c = 0; for (int v = 0; v < nf; v++) if (v != f) fv[c++] = v;
        Mat Mf((size_t)Np * Np, 0);
        std::vector<real> ell(Np);
        for (size_t q = 0; q < fw.size(); q++) {
            real x[3] = {vx[fv[0]][0], vx[fv[0]][1], vx[fv[0]][2]};
            for (int k = 1; k < dim; k++) for (int d = 0; d < dim; d++) x[d] += fq[q * (dim - 1) + (k - 1)] * (vx[fv[k]][d] - vx[fv[0]][d]);
            B.eval(x, phi.data(), dphi.data());
            for (int i = 0; i < Np; i++) { real s = 0; for (int m = 0; m < Np; m++) s += phi[m] * LV[(size_t)m * Np + i]; ell[i] = s; }
            for (int i = 0; i < Np; i++) for (int j = 0; j < Np; j++) Mf[(size_t)i * Np + j] += fw[q] * ell[i] * ell[j];
        }
        for (int i = 0; i < Np; i++) for (int j = 0; j < Nfp; j++) {
            real s = 0;
            for (int k = 0; k < Np; k++) s += Minv[(size_t)i * Np + k] * Mf[(size_t)k * Np + fn[f][j]];
            R.lift[((size_t)f * Np + i) * Nfp + j] = (double)s;
        }
    }
    return R;
}

}  // namespace dgtd
