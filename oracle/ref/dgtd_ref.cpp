// dgtd_ref — CPU ORACLE for the DG evolution hot path of OpenSEMBA/dgtd.
//
// TEST INFRASTRUCTURE ONLY.  Nothing under dgtd_b200/ may include, link or run
// this file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs execute the binary built from it (oracle/_ref/dgtd_ref).
//
// What it is: a serial restatement of the reference's `global` evolution
// operator, built ON TOP OF THE REFERENCE'S OWN CODE compiled where it lies
// under /root/reference (see Makefile): the vendored MFEM fork
// (external/mfem-geg) and the reference's DG face integrators
// (src/mfemExtension/{BilinearIntegrators,IntegratorFunctions}.cpp).  The only
// restated part is the glue that the reference keeps in MPI-only translation
// units that cannot be compiled here (no MPI/HYPRE/Eigen/GSL in the image):
//   * DGOperatorFactory::buildGlobalOperator      src/components/DGOperatorFactory.h:1428-1602
//       - sub-operators                            :413-573
//       - block placement and signs                :1268-1361
//       - CSR merge + Threshold(1e-20)             :178-236, :1569-1570
//   * DGOperatorFactory::buildSourceFaceOperator  :1405-1425, :959-1045   (TF/SF operator)
//   * SourcesManager::initDirectPlanewaveEval / evalTimeVarFieldDirect
//                                                  src/solver/SourcesManager.cpp:129-231
//   * TF/SF side classification (3-D)              src/components/SubMesher.cpp:677-771
//   * Planewave / Gaussian / ModulatedGaussian     src/math/Function.h:71-136, 361-409
//   * GlobalEvolution::Mult (+ TF/SF skip test)    src/evolution/GlobalEvolution.cpp:551-626, 628-823
//   * time loop = mfem::RK4Solver, unmodified      external/mfem-geg/linalg/ode.cpp:109-136
//
// Pinning: `dgtd_ref known-answers <json>` rebuilds the nine M^-1*flux blocks
// that test/hesthavenComparison/Hesthaven2DTest.cpp:234-553 asserts (order 1,
// Maxwell2D_K2.mesh, PEC) and compares them with the literals extracted from
// that file (tests/golden/ref_known_answers.json, tolerance 1e-8 as in the
// reference test).
#include <mfem.hpp>
#include "mfemExtension/BilinearIntegrators.h"

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <map>
#include <memory>
#include <random>
#include <set>
#include <sstream>
#include <string>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

using namespace mfem;
namespace mx = maxwell::mfemExtension;
using maxwell::Direction;

enum FieldT { FE = 0, FH = 1 };
enum BC { BC_PEC = 0, BC_PMC = 1, BC_SMA = 2 };

// DGOperatorFactory.h:21-35 (centred and upwind tables are identical)
static double bdrCoeff(BC bc, int f)
{
   switch (bc)
   {
      case BC_PEC: return f == FE ? 2.0 : 0.0;
      case BC_PMC: return f == FE ? 0.0 : 2.0;
      default:     return 1.0;
   }
}
static int alt(int f) { return 1 - f; }

struct PlaneWave
{
   bool on = false;
   double spread = 1.0, mean1d = 0.0, freq = 0.0; // freq != 0 -> ModulatedGaussian
   double pol[3] = {1, 0, 0}, dir[3] = {0, 0, 1};
   int fieldtype = FE;
   // Function.h:361-409 (pol / dir normalised in the ctor, :330-339)
   double eval(const double *p, double t, int ft, int d) const
   {
      double polDir;
      double cr[3];
      if (fieldtype == FE)
      {
         if (ft == FE) { polDir = pol[d]; }
         else
         {
            cr[0] = dir[1]*pol[2] - dir[2]*pol[1];
            cr[1] = dir[2]*pol[0] - dir[0]*pol[2];
            cr[2] = dir[0]*pol[1] - dir[1]*pol[0];
            polDir = cr[d];
         }
      }
      else
      {
         if (ft == FH) { polDir = pol[d]; }
         else
         {
            cr[0] = pol[1]*dir[2] - pol[2]*dir[1];
            cr[1] = pol[2]*dir[0] - pol[0]*dir[2];
            cr[2] = pol[0]*dir[1] - pol[1]*dir[0];
            polDir = cr[d];
         }
      }
      // mfem::Vector operator* (dot product): sequential sum over 3 entries
      double phaseDelay = (p[0]*dir[0] + p[1]*dir[1] + p[2]*dir[2]) / 1.0; // speedOfLight = 1
      double u = phaseDelay - t;
      double g;
      if (freq == 0.0)
      {
         // Gaussian, dimension 1 (Function.h:71-76)
         g = exp(-pow(u - mean1d, 2) / (2.0 * pow(spread, 2)));
      }
      else
      {
         double arg = u - mean1d;   // ModulatedGaussian (Function.h:130-136)
         g = exp(-arg * arg / (2.0 * spread * spread)) * cos(2.0 * M_PI * freq * arg);
      }
      return g * polDir;
   }
};

struct Problem
{
   std::unique_ptr<Mesh> mesh;
   int order = 3;
   double alpha = 1.0;
   std::map<int, BC> bdr;                       // bdr attribute -> condition
   std::vector<int> tfsf_tags;                  // bdr attributes of TF/SF interior faces
   std::map<int, std::array<double, 3>> mat;    // element attribute -> eps, mu, sigma
   PlaneWave pw;
};

static std::vector<std::string> split(const std::string &s, char c)
{
   std::vector<std::string> r; std::stringstream ss(s); std::string it;
   while (std::getline(ss, it, c)) { r.push_back(it); }
   return r;
}

static std::unique_ptr<Mesh> makeMesh(const std::string &spec, int refine)
{
   std::unique_ptr<Mesh> m;
   auto p = split(spec, ':');
   if (p[0] == "cart1d")
   {
      int n = std::stoi(p[1]); double L = p.size() > 2 ? std::stod(p[2]) : 1.0;
      m.reset(new Mesh(Mesh::MakeCartesian1D(n, L)));
   }
   else if (p[0] == "cart2d")
   {
      int nx = std::stoi(p[1]), ny = std::stoi(p[2]);
      double sx = p.size() > 3 ? std::stod(p[3]) : 1.0, sy = p.size() > 4 ? std::stod(p[4]) : 1.0;
      m.reset(new Mesh(Mesh::MakeCartesian2D(nx, ny, Element::TRIANGLE, false, sx, sy)));
   }
   else if (p[0] == "cart3d")
   {
      int nx = std::stoi(p[1]);
      int ny = p.size() > 2 ? std::stoi(p[2]) : nx, nz = p.size() > 3 ? std::stoi(p[3]) : nx;
      double sx = p.size() > 4 ? std::stod(p[4]) : 1.0, sy = p.size() > 5 ? std::stod(p[5]) : 1.0,
             sz = p.size() > 6 ? std::stod(p[6]) : 1.0;
      m.reset(new Mesh(Mesh::MakeCartesian3D(nx, ny, nz, Element::TETRAHEDRON, sx, sy, sz)));
   }
   else
   {
      // driver.cpp:1176-1183: LoadFromFile(name, generate_edges=1, refine=0, fix_orientation=true)
      m.reset(new Mesh(Mesh::LoadFromFile(spec.c_str(), 1, 0, true)));
   }
   for (int r = 0; r < refine; r++) { m->UniformRefinement(); }   // driver.cpp:1261-1266
   return m;
}

// ----------------------------------------------------------------------------
// Operator assembly (restatement of DGOperatorFactory<FES>, serial BilinearForm)
// ----------------------------------------------------------------------------
struct Factory
{
   Problem &pd;
   FiniteElementSpace &fes;
   int dim, N;
   std::map<BC, Array<int>> bdrMarker;   // Model::getBoundaryToMarker (true boundary faces only)
   std::map<BC, Array<int>> ibMarker;    // Model::getInteriorBoundaryToMarker (PEC/PMC/SMA sheets inside the mesh)
   Array<int> ignoreMarker;              // buildInteriorIgnoreMarker (:373-389): faces the regular interior flux skips
   Array<int> tfsfMarker;

   Factory(Problem &p, FiniteElementSpace &f) : pd(p), fes(f)
   {
      Mesh &m = *fes.GetMesh();
      dim = m.Dimension(); N = fes.GetNDofs();
      int nattr = m.bdr_attributes.Size() ? m.bdr_attributes.Max() : 0;
      // driver.cpp:1012-1041: a boundary tag is INTERIOR when one of its boundary elements lies on an interior face
      std::set<int> interiorTag;
      for (int be = 0; be < m.GetNBE(); be++)
      {
         int a = m.GetBdrAttribute(be);
         if (pd.bdr.count(a) && m.FaceIsInterior(m.GetBdrElementFaceIndex(be))) { interiorTag.insert(a); }
      }
      for (auto &kv : pd.bdr)
      {
         if (kv.first > nattr) { continue; }
         auto &mk = interiorTag.count(kv.first) ? ibMarker[kv.second] : bdrMarker[kv.second];
         if (mk.Size() == 0) { mk.SetSize(nattr); mk = 0; }
         mk[kv.first - 1] = 1;
      }
      if (!ibMarker.empty())
      {
         ignoreMarker.SetSize(nattr); ignoreMarker = 0;
         for (auto &kv : ibMarker) for (int i = 0; i < nattr; i++) { if (kv.second[i] == 1) { ignoreMarker[i] = 1; } }
      }
      if (!pd.tfsf_tags.empty())
      {
         tfsfMarker.SetSize(nattr); tfsfMarker = 0;
         for (int t : pd.tfsf_tags) { if (t <= nattr) { tfsfMarker[t - 1] = 1; } }
      }
   }

   // :413-424 + Model::buildEpsMuPiecewiseVector
   std::unique_ptr<BilinearForm> MInv(int f)
   {
      Mesh &m = *fes.GetMesh();
      Vector pw(m.attributes.Max()); pw = 1.0;
      for (auto &kv : pd.mat) { if (kv.first <= pw.Size()) { pw[kv.first - 1] = kv.second[f]; } }
      PWConstCoefficient c(pw);
      auto r = std::make_unique<BilinearForm>(&fes);
      r->AddDomainIntegrator(new InverseIntegrator(new MassIntegrator(c)));
      r->Assemble(); r->Finalize();
      return r;
   }
   // :426-467 (straight-sided meshes: default rule)
   std::unique_ptr<BilinearForm> Deriv(int d)
   {
      auto r = std::make_unique<BilinearForm>(&fes);
      if (d < dim)
      {
         ConstantCoefficient one(1.0);
         r->AddDomainIntegrator(new DerivativeIntegrator(one, d));
      }
      r->Assemble(); r->Finalize();
      return r;
   }
   // :469-503
   std::unique_ptr<BilinearForm> ZeroNormal(int f)
   {
      auto r = std::make_unique<BilinearForm>(&fes);
      if (ignoreMarker.Size() > 0) { r->AddInteriorFaceIntegrator(new mx::MaxwellDGZeroNormalJumpIntegrator(pd.alpha), ignoreMarker); }
      else { r->AddInteriorFaceIntegrator(new mx::MaxwellDGZeroNormalJumpIntegrator(pd.alpha)); }
      for (auto &kv : bdrMarker)
      {
         double c = kv.first != BC_SMA ? bdrCoeff(kv.first, f) * pd.alpha : 1.0;
         r->AddBdrFaceIntegrator(new mx::MaxwellDGZeroNormalJumpIntegrator(c), kv.second);
      }
      r->Assemble(); r->Finalize();
      return r;
   }
   // :505-538
   std::unique_ptr<BilinearForm> OneNormal(int f, int x)
   {
      std::vector<Direction> dt{Direction(x)};
      auto r = std::make_unique<BilinearForm>(&fes);
      if (ignoreMarker.Size() > 0) { r->AddInteriorFaceIntegrator(new mx::MaxwellDGOneNormalJumpIntegrator(dt, 1.0), ignoreMarker); }
      else { r->AddInteriorFaceIntegrator(new mx::MaxwellDGOneNormalJumpIntegrator(dt, 1.0)); }
      for (auto &kv : bdrMarker)
      {
         double c = kv.first != BC_SMA ? bdrCoeff(kv.first, f) : 1.0;
         r->AddBdrFaceIntegrator(new mx::MaxwellDGOneNormalJumpIntegrator(dt, c), kv.second);
      }
      r->Assemble(); r->Finalize();
      return r;
   }
   // :540-573
   std::unique_ptr<BilinearForm> TwoNormal(int f, int d, int d2)
   {
      std::vector<Direction> dt{Direction(d), Direction(d2)};
      auto r = std::make_unique<BilinearForm>(&fes);
      if (ignoreMarker.Size() > 0) { r->AddInteriorFaceIntegrator(new mx::MaxwellDGTwoNormalJumpIntegrator(dt, pd.alpha), ignoreMarker); }
      else { r->AddInteriorFaceIntegrator(new mx::MaxwellDGTwoNormalJumpIntegrator(dt, pd.alpha)); }
      for (auto &kv : bdrMarker)
      {
         double c = kv.first != BC_SMA ? bdrCoeff(kv.first, f) * pd.alpha : 1.0;
         r->AddBdrFaceIntegrator(new mx::MaxwellDGTwoNormalJumpIntegrator(dt, c), kv.second);
      }
      r->Assemble(); r->Finalize();
      return r;
   }
   // interior-boundary forms (:575-675): MaxwellDGInteriorJumpIntegrator keeps the two self blocks only, i.e. each side
   // of the sheet sees a boundary face with the true-boundary coefficients (SMA: 1.0 regardless of alpha)
   std::unique_ptr<BilinearForm> IBZero(int f)
   {
      auto r = std::make_unique<BilinearForm>(&fes);
      for (auto &kv : ibMarker)
      {
         double c = kv.first != BC_SMA ? bdrCoeff(kv.first, f) * pd.alpha : 1.0;
         r->AddInternalBoundaryFaceIntegrator(new mx::MaxwellDGInteriorJumpIntegrator({}, c), kv.second);
      }
      r->Assemble(); r->Finalize(); return r;
   }
   std::unique_ptr<BilinearForm> IBOne(int f, int x)
   {
      std::vector<Direction> dt{Direction(x)};
      auto r = std::make_unique<BilinearForm>(&fes);
      for (auto &kv : ibMarker)
      {
         double c = kv.first != BC_SMA ? bdrCoeff(kv.first, f) : 1.0;
         r->AddInternalBoundaryFaceIntegrator(new mx::MaxwellDGInteriorJumpIntegrator(dt, c), kv.second);
      }
      r->Assemble(); r->Finalize(); return r;
   }
   std::unique_ptr<BilinearForm> IBTwo(int f, int d, int d2)
   {
      std::vector<Direction> dt{Direction(d), Direction(d2)};
      auto r = std::make_unique<BilinearForm>(&fes);
      for (auto &kv : ibMarker)
      {
         double c = kv.first != BC_SMA ? bdrCoeff(kv.first, f) * pd.alpha : 1.0;
         r->AddInternalBoundaryFaceIntegrator(new mx::MaxwellDGInteriorJumpIntegrator(dt, c), kv.second);
      }
      r->Assemble(); r->Finalize(); return r;
   }
   // sigma mass (:  buildSigmaMassOperator) — PW sigma per attribute
   std::unique_ptr<BilinearForm> SigmaMass()
   {
      Mesh &m = *fes.GetMesh();
      Vector pw(m.attributes.Max()); pw = 0.0;
      for (auto &kv : pd.mat) { if (kv.first <= pw.Size()) { pw[kv.first - 1] = kv.second[2]; } }
      PWConstCoefficient c(pw);
      auto r = std::make_unique<BilinearForm>(&fes);
      r->AddDomainIntegrator(new MassIntegrator(c));
      r->Assemble(); r->Finalize();
      return r;
   }
   // TF/SF source-face forms (:678-710): internal boundary faces under the marker
   std::unique_ptr<BilinearForm> SrcZero()
   {
      auto r = std::make_unique<BilinearForm>(&fes);
      r->AddInternalBoundaryFaceIntegrator(new mx::MaxwellDGZeroNormalJumpIntegrator(pd.alpha), tfsfMarker);
      r->Assemble(); r->Finalize(); return r;
   }
   std::unique_ptr<BilinearForm> SrcOne(int x)
   {
      std::vector<Direction> dt{Direction(x)};
      auto r = std::make_unique<BilinearForm>(&fes);
      r->AddInternalBoundaryFaceIntegrator(new mx::MaxwellDGOneNormalJumpIntegrator(dt, 1.0), tfsfMarker);
      r->Assemble(); r->Finalize(); return r;
   }
   std::unique_ptr<BilinearForm> SrcTwo(int d, int d2)
   {
      std::vector<Direction> dt{Direction(d), Direction(d2)};
      auto r = std::make_unique<BilinearForm>(&fes);
      r->AddInternalBoundaryFaceIntegrator(new mx::MaxwellDGTwoNormalJumpIntegrator(dt, pd.alpha), tfsfMarker);
      r->Assemble(); r->Finalize(); return r;
   }
};

struct Placement { std::unique_ptr<SparseMatrix> blk; int row, col; double sign; };

// mergeBlocksToCSR (:178-236): first-seen column order per row, duplicates accumulate
// in block order; then Threshold(1e-20).
static std::unique_ptr<SparseMatrix> mergeBlocks(std::vector<Placement> &blocks, int rows, int cols)
{
   std::vector<int> marker(cols, -1);
   int *I = Memory<int>(rows + 1); I[0] = 0;
   for (int r = 0; r < rows; r++)
   {
      int nnz = 0;
      for (auto &bp : blocks)
      {
         int lr = r - bp.row;
         if (lr < 0 || lr >= bp.blk->Height()) { continue; }
         const int n = bp.blk->RowSize(lr); const int *c = bp.blk->GetRowColumns(lr);
         for (int k = 0; k < n; k++) { int gc = c[k] + bp.col; if (marker[gc] != r) { marker[gc] = r; nnz++; } }
      }
      I[r + 1] = I[r] + nnz;
   }
   int *J = Memory<int>(I[rows]); real_t *A = Memory<real_t>(I[rows]);
   std::fill(marker.begin(), marker.end(), -1);
   int pos = 0;
   for (int r = 0; r < rows; r++)
   {
      for (auto &bp : blocks)
      {
         int lr = r - bp.row;
         if (lr < 0 || lr >= bp.blk->Height()) { continue; }
         const int n = bp.blk->RowSize(lr); const int *c = bp.blk->GetRowColumns(lr);
         const real_t *v = bp.blk->GetRowEntries(lr);
         for (int k = 0; k < n; k++)
         {
            int gc = c[k] + bp.col;
            if (marker[gc] < I[r]) { J[pos] = gc; A[pos] = v[k] * bp.sign; marker[gc] = pos; pos++; }
            else { A[marker[gc]] += v[k] * bp.sign; }
         }
      }
   }
   return std::make_unique<SparseMatrix>(I, J, A, rows, cols);
}

static SparseMatrix *prod(BilinearForm &a, BilinearForm &b) { return mfem::Mult(a.SpMat(), b.SpMat()); }

// buildGlobalOperator (:1428-1602), single rank (no face-neighbour columns)
static std::unique_ptr<SparseMatrix> buildGlobal(Factory &F)
{
   const int N = F.N, dim = F.dim;
   std::vector<Placement> blocks;
   std::unique_ptr<BilinearForm> MI[2] = {F.MInv(FE), F.MInv(FH)};
   auto place = [&](SparseMatrix *op, int fr, int dr, int fc, int dc, double s)
   { blocks.push_back({std::make_unique<SparseMatrix>(*op), (3 * fr + dr) * N, (3 * fc + dc) * N, s}); };
   // interior boundaries first (:1444-1484, collectGlobal{One,Zero,Two}NormalIBFIOperators :1214-1265): same block
   // placement and signs as the regular flux forms
   if (!F.ibMarker.empty())
   {
      for (int f : {FE, FH}) for (int x = 0; x < 3; x++)
      {
         if (x >= dim) { continue; }
         int y = (x + 1) % 3, z = (x + 2) % 3;
         auto B = F.IBOne(alt(f), x);
         std::unique_ptr<SparseMatrix> op(prod(*MI[f], *B));
         place(op.get(), f, y, alt(f), z, 1.0 - 2.0 * f);
         place(op.get(), f, z, alt(f), y, -1.0 + 2.0 * f);
      }
      for (int f : {FE, FH})
      {
         auto B = F.IBZero(f);
         std::unique_ptr<SparseMatrix> op(prod(*MI[f], *B));
         for (int d = 0; d < 3; d++) { place(op.get(), f, d, f, d, -1.0); }
      }
      for (int f : {FE, FH}) for (int d = 0; d < 3; d++)
      {
         if (d >= dim) { continue; }
         for (int d2 = 0; d2 < 3; d2++)
         {
            if (d2 >= dim) { continue; }
            auto B = F.IBTwo(f, d, d2);
            std::unique_ptr<SparseMatrix> op(prod(*MI[f], *B));
            place(op.get(), f, d, f, d2, 1.0);
         }
      }
   }
   // directional (:1268-1289)
   for (int f : {FE, FH}) for (int x = 0; x < 3; x++)
   {
      if (x >= dim) { continue; }
      int y = (x + 1) % 3, z = (x + 2) % 3;
      auto D = F.Deriv(x);
      std::unique_ptr<SparseMatrix> op(prod(*MI[f], *D));
      place(op.get(), f, z, alt(f), y, 1.0 - 2.0 * f);
      place(op.get(), f, y, alt(f), z, -1.0 + 2.0 * f);
   }
   // one-normal (:1306-1327)
   for (int f : {FE, FH}) for (int x = 0; x < 3; x++)
   {
      if (x >= dim) { continue; }
      int y = (x + 1) % 3, z = (x + 2) % 3;
      auto B = F.OneNormal(alt(f), x);
      std::unique_ptr<SparseMatrix> op(prod(*MI[f], *B));
      place(op.get(), f, y, alt(f), z, 1.0 - 2.0 * f);
      place(op.get(), f, z, alt(f), y, -1.0 + 2.0 * f);
   }
   // zero-normal (:1291-1304)
   for (int f : {FE, FH})
   {
      auto B = F.ZeroNormal(f);
      std::unique_ptr<SparseMatrix> op(prod(*MI[f], *B));
      for (int d = 0; d < 3; d++) { place(op.get(), f, d, f, d, -1.0); }
   }
   // two-normal (:1329-1347)
   for (int f : {FE, FH}) for (int d = 0; d < 3; d++)
   {
      if (d >= dim) { continue; }
      for (int d2 = 0; d2 < 3; d2++)
      {
         if (d2 >= dim) { continue; }
         auto B = F.TwoNormal(f, d, d2);
         std::unique_ptr<SparseMatrix> op(prod(*MI[f], *B));
         place(op.get(), f, d, f, d2, 1.0);
      }
   }
   // conductive (:1349-1361)
   {
      auto MS = F.SigmaMass();
      std::unique_ptr<SparseMatrix> op(prod(*MI[FE], *MS));
      for (int d = 0; d < 3; d++) { place(op.get(), FE, d, FE, d, -1.0); }
   }
   auto res = mergeBlocks(blocks, 6 * N, 6 * N);
   blocks.clear();
   res->Threshold(1e-20);
   return res;
}

// buildSourceFaceOperator(marker) (:1405-1425): one-normal, zero-normal, two-normal
static std::unique_ptr<SparseMatrix> buildTFSF(Factory &F)
{
   const int N = F.N, dim = F.dim;
   auto res = std::make_unique<SparseMatrix>(6 * N, 6 * N);
   std::unique_ptr<BilinearForm> MI[2] = {F.MInv(FE), F.MInv(FH)};
   auto load = [&](SparseMatrix *op, int fr, int dr, int fc, int dc, double s)
   {
      Array<int> cols; Vector vals;
      for (int r = 0; r < N; r++)
      {
         op->GetRow(r, cols, vals);
         for (int c = 0; c < cols.Size(); c++) { res->Add((3 * fr + dr) * N + r, (3 * fc + dc) * N + cols[c], vals[c] * s); }
      }
   };
   for (int f : {FE, FH}) for (int x = 0; x < 3; x++)
   {
      if (x >= dim) { continue; }
      int y = (x + 1) % 3, z = (x + 2) % 3;
      auto B = F.SrcOne(x);
      std::unique_ptr<SparseMatrix> op(prod(*MI[f], *B));
      load(op.get(), f, y, alt(f), z, 1.0 - 2.0 * f);
      load(op.get(), f, z, alt(f), y, -1.0 + 2.0 * f);
   }
   for (int f : {FE, FH})
   {
      auto B = F.SrcZero();
      std::unique_ptr<SparseMatrix> op(prod(*MI[f], *B));
      for (int d = 0; d < 3; d++) { load(op.get(), f, d, f, d, -1.0); }
   }
   for (int f : {FE, FH}) for (int d = 0; d < 3; d++)
   {
      if (d >= dim) { continue; }
      for (int d2 = 0; d2 < 3; d2++)
      {
         if (d2 >= dim) { continue; }
         auto B = F.SrcTwo(d, d2);
         std::unique_ptr<SparseMatrix> op(prod(*MI[f], *B));
         load(op.get(), f, d, f, d2, 1.0);
      }
   }
   res->Finalize();
   return res;
}

// ----------------------------------------------------------------------------
// TF/SF source vector: which elements touch a TF/SF face, which side they are on
// (SubMesher.cpp:677-771, 3-D rule), DOF coordinates, +-1/2 mask
// (SourcesManager.cpp:129-188).
// ----------------------------------------------------------------------------
struct TFSFSource
{
   std::vector<int> dof;       // parent dof ids (all dofs of TF/SF-adjacent elements)
   std::vector<double> xyz;    // 3 per dof
   std::vector<double> sign;   // +0.5 TF, -0.5 SF
   std::vector<int> elemSide;  // per mesh element: 0 none, 1 TF, 2 SF   (exported for the product)
};

static void nodeCoords(FiniteElementSpace &fes, std::vector<double> &xyz)
{
   const int N = fes.GetNDofs(); xyz.assign(3 * (size_t)N, 0.0);
   Array<int> dofs; Vector c;
   for (int e = 0; e < fes.GetNE(); e++)
   {
      const FiniteElement *fe = fes.GetFE(e);
      ElementTransformation *T = fes.GetElementTransformation(e);
      const IntegrationRule &ir = fe->GetNodes();
      fes.GetElementDofs(e, dofs);
      for (int i = 0; i < dofs.Size(); i++)
      {
         T->SetIntPoint(&ir.IntPoint(i)); T->Transform(ir.IntPoint(i), c);
         for (int k = 0; k < c.Size(); k++) { xyz[3 * (size_t)dofs[i] + k] = c[k]; }
      }
   }
}

static TFSFSource buildTFSFSource(Factory &F, const std::vector<double> &xyz)
{
   TFSFSource S; Mesh &m = *F.fes.GetMesh();
   S.elemSide.assign(m.GetNE(), 0);
   if (F.tfsfMarker.Size() == 0) { return S; }
   double ctr[3] = {0, 0, 0}; int nv = 0; std::set<int> counted; Array<int> verts;
   for (int be = 0; be < m.GetNBE(); be++)
   {
      if (F.tfsfMarker[m.GetBdrAttribute(be) - 1] != 1) { continue; }
      m.GetBdrElementVertices(be, verts);
      for (int i = 0; i < verts.Size(); i++)
         if (counted.insert(verts[i]).second) { const double *v = m.GetVertex(verts[i]); for (int d = 0; d < 3; d++) { ctr[d] += v[d]; } nv++; }
   }
   if (nv) { for (int d = 0; d < 3; d++) { ctr[d] /= double(nv); } }
   auto bary = [&](int e, double *b)
   {
      // getBarycenterOfElement: mean of the element's vertices
      Array<int> v; m.GetElementVertices(e, v); b[0] = b[1] = b[2] = 0;
      for (int i = 0; i < v.Size(); i++) { const double *p = m.GetVertex(v[i]); for (int d = 0; d < 3; d++) { b[d] += p[d]; } }
      for (int d = 0; d < 3; d++) { b[d] /= v.Size(); }
   };
   int seen1d = 0;
   for (int be = 0; be < m.GetNBE(); be++)
   {
      if (F.tfsfMarker[m.GetBdrAttribute(be) - 1] != 1) { continue; }
      int f = m.GetBdrElementFaceIndex(be), e1, e2; m.GetFaceElements(f, &e1, &e2);
      if (e2 < 0) { MFEM_ABORT("oracle: TF/SF tag on a true boundary face is not restated"); }
      double b1[3], b2[3]; bary(e1, b1); bary(e2, b2);
      double d1 = 0, d2 = 0;
      for (int d = 0; d < 3; d++) { d1 += (b1[d] - ctr[d]) * (b1[d] - ctr[d]); d2 += (b2[d] - ctr[d]) * (b2[d] - ctr[d]); }
      bool e1tf;
      if (m.Dimension() == 3) { e1tf = d1 < d2; }                       // centroid rule, SubMesher.cpp:677-771
      else if (m.Dimension() == 2)
      {
         // setIndividualTFSFAttributesForSubMeshing2D (:568-660): cross(Elem1 -> Elem2 barycentres, face tangent)_z >= 0 puts
         // Elem1 on the scattered-field side; the tangent runs along m.GetFace(f)'s vertices (buildTangent2D, :239-250)
         const int *fv = m.GetFace(f)->GetVertices();
         const double *v0 = m.GetVertex(fv[0]), *v1 = m.GetVertex(fv[1]);
         const double ori = (b2[0] - b1[0]) * (v1[1] - v0[1]) - (b2[1] - b1[1]) * (v1[0] - v0[0]);
         e1tf = !(ori >= 0.0);
      }
      else { e1tf = seen1d == 1; }                                       // assignIndividualTFSFAtts{One,Two}Point(s)1D, :476-547
      seen1d++;
      // SF marks override TF marks (sf_dof_set applied last, SourcesManager.cpp:173-187)
      auto mark = [&](int e, bool tf) { if (!tf) { S.elemSide[e] = 2; } else if (S.elemSide[e] == 0) { S.elemSide[e] = 1; } };
      mark(e1, e1tf); mark(e2, !e1tf);
   }
   Array<int> dofs;
   for (int e = 0; e < m.GetNE(); e++)
   {
      if (!S.elemSide[e]) { continue; }
      F.fes.GetElementDofs(e, dofs);
      for (int i = 0; i < dofs.Size(); i++)
      {
         S.dof.push_back(dofs[i]);
         for (int k = 0; k < 3; k++) { S.xyz.push_back(xyz[3 * (size_t)dofs[i] + k]); }
         S.sign.push_back(S.elemSide[e] == 1 ? 0.5 : -0.5);
      }
   }
   return S;
}

// ----------------------------------------------------------------------------
// The evolution operator (GlobalEvolution::Mult, single rank, no SGBC)
// ----------------------------------------------------------------------------
class GlobalOracle : public TimeDependentOperator
{
public:
   std::unique_ptr<SparseMatrix> A, Atfsf;
   TFSFSource src; PlaneWave pw; int N;
   mutable Vector work; mutable std::array<Vector, 6> sub;
   static constexpr double skip_threshold = 1e-8;   // GlobalEvolution.h:100
   mutable long nskipped = 0, napplied = 0;

   GlobalOracle(int n) : TimeDependentOperator(6 * n), N(n) {}
   void Mult(const Vector &in, Vector &out) const override
   {
      if (out.Size() != 6 * N) { out.SetSize(6 * N); }
      A->Mult(in, out);                                   // GlobalEvolution.cpp:811
      if (!Atfsf || src.dof.empty() || !pw.on) { return; }
      const double t = GetTime(); const int V = (int)src.dof.size();
      for (int c = 0; c < 6; c++) { sub[c].SetSize(V); }
      for (int i = 0; i < V; i++)                         // SourcesManager.cpp:204-231
         for (int ft = 0; ft < 2; ft++) for (int d = 0; d < 3; d++)
         { sub[3 * ft + d][i] = 0.0; sub[3 * ft + d][i] += pw.eval(&src.xyz[3 * (size_t)i], t, ft, d) * src.sign[i]; }
      double norm2 = 0.0;                                 // GlobalEvolution.cpp:584-598
      for (int c = 0; c < 6; c++) { double n = sub[c].Norml2(); norm2 += n * n; }
      if (norm2 < skip_threshold * skip_threshold) { nskipped++; return; }
      napplied++;
      if (work.Size() != 6 * N) { work.SetSize(6 * N); work = 0.0; }
      for (int i = 0; i < V; i++) for (int c = 0; c < 6; c++) { work[c * N + src.dof[i]] = sub[c][i]; }
      Atfsf->AddMult(work, out, -1.0);                    // :615
      for (int i = 0; i < V; i++) for (int c = 0; c < 6; c++) { work[c * N + src.dof[i]] = 0.0; }
   }
};

// ----------------------------------------------------------------------------
// I/O helpers
// ----------------------------------------------------------------------------
template <class T> static void dump(const std::string &path, const T *p, size_t n)
{
   FILE *f = fopen(path.c_str(), "wb");
   if (!f) { fprintf(stderr, "cannot write %s\n", path.c_str()); exit(2); }
   if (n) { fwrite(p, sizeof(T), n, f); }
   fclose(f);
}
static std::vector<double> slurp(const std::string &path)
{
   std::ifstream f(path, std::ios::binary | std::ios::ate);
   if (!f) { fprintf(stderr, "cannot read %s\n", path.c_str()); exit(2); }
   size_t n = f.tellg(); f.seekg(0); std::vector<double> v(n / 8);
   f.read((char *)v.data(), n); return v;
}

static void dumpMesh(Mesh &m, const std::string &dir)
{
   const int dim = m.Dimension(), nv = m.GetNV(), ne = m.GetNE(), nbe = m.GetNBE();
   std::vector<double> verts(3 * (size_t)nv, 0.0);
   for (int v = 0; v < nv; v++) for (int d = 0; d < m.SpaceDimension(); d++) { verts[3 * (size_t)v + d] = m.GetVertex(v)[d]; }
   std::vector<int> ev((dim + 1) * (size_t)ne), ea(ne), bv(dim * (size_t)nbe), ba(nbe);
   Array<int> v;
   for (int e = 0; e < ne; e++)
   {
      m.GetElementVertices(e, v);
      if (v.Size() != dim + 1) { MFEM_ABORT("oracle: simplex meshes only"); }
      for (int k = 0; k <= dim; k++) { ev[(dim + 1) * (size_t)e + k] = v[k]; }
      ea[e] = m.GetAttribute(e);
   }
   for (int b = 0; b < nbe; b++)
   {
      m.GetBdrElementVertices(b, v);
      for (int k = 0; k < dim; k++) { bv[dim * (size_t)b + k] = v[k]; }
      ba[b] = m.GetBdrAttribute(b);
   }
   dump(dir + "/verts.f64", verts.data(), verts.size());
   dump(dir + "/elems.i32", ev.data(), ev.size());
   dump(dir + "/elem_attr.i32", ea.data(), ea.size());
   dump(dir + "/bdr.i32", bv.data(), bv.size());
   dump(dir + "/bdr_attr.i32", ba.data(), ba.size());
}

static std::map<std::string, std::string> parseArgs(int argc, char **argv, int start)
{
   std::map<std::string, std::string> a;
   for (int i = start; i < argc; i++)
   {
      std::string k = argv[i];
      if (k.rfind("--", 0) != 0) { fprintf(stderr, "bad arg %s\n", k.c_str()); exit(2); }
      k = k.substr(2);
      if (i + 1 < argc && std::string(argv[i + 1]).rfind("--", 0) != 0) { a[k] = argv[++i]; }
      else { a[k] = "1"; }
   }
   return a;
}
static BC parseBC(const std::string &s)
{
   if (s == "pec" || s == "PEC") { return BC_PEC; }
   if (s == "pmc" || s == "PMC") { return BC_PMC; }
   if (s == "sma" || s == "SMA") { return BC_SMA; }
   fprintf(stderr, "unknown bc %s\n", s.c_str()); exit(2);
}

static Problem problemFromArgs(std::map<std::string, std::string> &a)
{
   Problem p;
   p.mesh = makeMesh(a.count("mesh") ? a["mesh"] : "cart3d:2", a.count("refine") ? std::stoi(a["refine"]) : 0);
   p.order = a.count("order") ? std::stoi(a["order"]) : 3;
   p.alpha = a.count("alpha") ? std::stod(a["alpha"]) : 1.0;
   if (a.count("bdr"))
      for (auto &it : split(a["bdr"], ',')) { auto kv = split(it, ':'); p.bdr[std::stoi(kv[0])] = parseBC(kv[1]); }
   if (a.count("bdr-all"))
   {
      BC bc = parseBC(a["bdr-all"]);
      for (int i = 0; i < p.mesh->bdr_attributes.Size(); i++) { if (!p.bdr.count(p.mesh->bdr_attributes[i])) { p.bdr[p.mesh->bdr_attributes[i]] = bc; } }
   }
   if (a.count("tfsf")) for (auto &it : split(a["tfsf"], ',')) { p.tfsf_tags.push_back(std::stoi(it)); p.bdr.erase(std::stoi(it)); }
   if (a.count("mat"))
      for (auto &it : split(a["mat"], ',')) { auto kv = split(it, ':'); p.mat[std::stoi(kv[0])] = {std::stod(kv[1]), std::stod(kv[2]), std::stod(kv[3])}; }
   if (a.count("pw"))
   {
      // spread:mean1d:freq:px,py,pz:kx,ky,kz   (mean1d = "auto" -> driver.cpp:576-589)
      auto q = split(a["pw"], ':');
      p.pw.on = true; p.pw.spread = std::stod(q[0]); p.pw.freq = std::stod(q[2]);
      auto pv = split(q[3], ','), kv = split(q[4], ',');
      double pn = 0, kn = 0;
      for (int d = 0; d < 3; d++) { p.pw.pol[d] = std::stod(pv[d]); p.pw.dir[d] = std::stod(kv[d]); }
      // Vector::Norml2 of a 3-vector; ctor divides component-wise (Function.h:335-338)
      { Vector P(p.pw.pol, 3), K(p.pw.dir, 3); pn = P.Norml2(); kn = K.Norml2(); }
      for (int d = 0; d < 3; d++) { p.pw.pol[d] /= pn; p.pw.dir[d] /= kn; }
      if (q[1] == "auto")
      {
         Mesh &m = *p.mesh; double minph = std::numeric_limits<double>::max(); Array<int> verts;
         for (int be = 0; be < m.GetNBE(); be++)
         {
            int at = m.GetBdrAttribute(be); bool has = false;
            for (int t : p.tfsf_tags) { has |= (t == at); }
            if (!has) { continue; }
            m.GetBdrElementVertices(be, verts);
            double c[3] = {0, 0, 0};
            for (int v = 0; v < verts.Size(); v++) for (int d = 0; d < m.Dimension(); d++) { c[d] += m.GetVertex(verts[v])[d]; }
            double ph = 0; for (int d = 0; d < m.Dimension(); d++) { c[d] /= double(verts.Size()); ph += c[d] * p.pw.dir[d]; }
            minph = std::min(minph, ph);
         }
         if (minph == std::numeric_limits<double>::max()) { minph = 0.0; }
         double mean1d = minph - 5.0 * p.pw.spread * std::sqrt(2.0);
         // buildGaussianPlanewave projects mean_vec = mean1d*d_hat back onto dir (driver.cpp:483-499)
         double mv[3], pm = 0; for (int d = 0; d < 3; d++) { mv[d] = mean1d * p.pw.dir[d]; }
         for (int d = 0; d < 3; d++) { pm += mv[d] * p.pw.dir[d]; }
         p.pw.mean1d = pm / 1.0;
      }
      else { p.pw.mean1d = std::stod(q[1]); }
   }
   return p;
}

static void initState(Vector &x, const std::string &spec, const std::vector<double> &xyz, int N, int dim)
{
   x.SetSize(6 * N); x = 0.0;
   auto q = split(spec, ':');
   if (q[0] == "random")
   {
      std::mt19937_64 g(q.size() > 1 ? std::stoull(q[1]) : 1); std::uniform_real_distribution<double> u(-1.0, 1.0);
      for (int i = 0; i < 6 * N; i++) { x[i] = u(g); }
   }
   else if (q[0] == "gauss")
   {
      // gauss:<E|H>:<comp>:<spread>:<fdim>:<cx,cy,cz>  — InitialField::eval (Sources.cpp:38-58) + Gaussian (Function.h:71-94)
      int f = q[1] == "H" ? 1 : 0, c = std::stoi(q[2]); double s = std::stod(q[3]); int fd = std::stoi(q[4]);
      auto cv = split(q[5], ','); double ctr[3] = {0, 0, 0}; for (size_t k = 0; k < cv.size() && k < 3; k++) { ctr[k] = std::stod(cv[k]); }
      for (int i = 0; i < N; i++)
      {
         double r2 = 0; for (int k = 0; k < fd; k++) { r2 += pow(xyz[3 * (size_t)i + k] - ctr[k], 2.0); }
         x[(3 * f + c) * N + i] = exp(-r2 / (2.0 * pow(s, 2.0)));
      }
   }
   else if (q[0] == "resonant")
   {
      // resonant:<comp>:<mx,my[,mz]> : E_comp = prod sin(m_k pi x_k)  (SinusoidalMode)
      int c = std::stoi(q[1]); auto mv = split(q[2], ',');
      for (int i = 0; i < N; i++)
      {
         double v = 1.0; for (size_t k = 0; k < mv.size(); k++) { v *= sin(std::stod(mv[k]) * M_PI * xyz[3 * (size_t)i + k]); }
         x[c * N + i] = v;
      }
   }
   else if (q[0] == "smooth")
   {
      // smooth : all six components non-zero, reproducible from the node coordinates alone (tests/conftest.py:smooth_state):
      // u_c = sin(1.3 x + 0.7 c + 0.2) cos(0.9 y - 0.4 c) + 0.5 sin(1.1 z + c)
      for (int c = 0; c < 6; c++)
         for (int i = 0; i < N; i++)
         {
            const double X = xyz[3 * (size_t)i], Y = dim > 1 ? xyz[3 * (size_t)i + 1] : 0.0, Z = dim > 2 ? xyz[3 * (size_t)i + 2] : 0.0;
            x[c * N + i] = sin(1.3 * X + 0.7 * c + 0.2) * cos(0.9 * Y - 0.4 * c) + 0.5 * sin(1.1 * Z + c);
         }
   }
   else if (q[0] == "file") { auto v = slurp(q[1]); MFEM_VERIFY((int)v.size() == 6 * N, "x0 size"); for (int i = 0; i < 6 * N; i++) { x[i] = v[i]; } }
   else if (q[0] != "zero") { fprintf(stderr, "unknown init %s\n", spec.c_str()); exit(2); }
}

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

// ----------------------------------------------------------------------------
static int cmdGen(std::map<std::string, std::string> &a, bool bench)
{
   Problem pd = problemFromArgs(a);
   Mesh &mesh = *pd.mesh; const int dim = mesh.Dimension();
   DG_FECollection fec(pd.order, dim, BasisType::GaussLobatto);           // Solver.cpp:86
   FiniteElementSpace fes(&mesh, &fec);
   const int N = fes.GetNDofs();
   std::string out = a.count("out") ? a["out"] : "";
   if (!out.empty()) { std::string c = "mkdir -p '" + out + "'"; if (system(c.c_str())) { return 2; } }

   double t0 = now();
   Factory F(pd, fes);
   GlobalOracle op(N);
   op.A = buildGlobal(F);
   std::vector<double> xyz; nodeCoords(fes, xyz);
   op.pw = pd.pw;
   if (!pd.tfsf_tags.empty()) { op.Atfsf = buildTFSF(F); op.src = buildTFSFSource(F, xyz); }
   else { op.src.elemSide.assign(mesh.GetNE(), 0); }
   double tAsm = now() - t0;

   const double dt = a.count("dt") ? std::stod(a["dt"]) : 1e-3;
   const int steps = a.count("steps") ? std::stoi(a["steps"]) : 1;
   double t = a.count("t0") ? std::stod(a["t0"]) : 0.0;
   const double tstart = t;
   Vector x; initState(x, a.count("init") ? a["init"] : "random:1", xyz, N, dim);

   if (!out.empty())
   {
      dumpMesh(mesh, out);
      dump(out + "/nodes.f64", xyz.data(), xyz.size());
      dump(out + "/x0.f64", x.GetData(), (size_t)6 * N);
      dump(out + "/tfsf_side.i32", op.src.elemSide.data(), op.src.elemSide.size());
      Vector k; op.SetTime(t); op.Mult(x, k);
      dump(out + "/k0.f64", k.GetData(), (size_t)6 * N);
      if (a.count("dump-csr"))
      {
         dump(out + "/A_I.i32", op.A->GetI(), (size_t)6 * N + 1);
         dump(out + "/A_J.i32", op.A->GetJ(), (size_t)op.A->NumNonZeroElems());
         dump(out + "/A_V.f64", op.A->GetData(), (size_t)op.A->NumNonZeroElems());
      }
   }
   if (bench)
   {
      const char *dev = a.count("device") ? a["device"].c_str() : "omp";
      static Device device(dev);                                       // launcher.cpp:61 (-d omp)
   }
   RK4Solver rk; rk.Init(op);                                          // Solver.cpp:41-47, 124-125
   const int warm = a.count("warmup") ? std::stoi(a["warmup"]) : 0;
   for (int s = 0; s < warm; s++) { double d = dt; rk.Step(x, t, d); }
   std::vector<int> snaps;
   if (a.count("snap")) for (auto &s : split(a["snap"], ',')) { snaps.push_back(std::stoi(s)); }
   double tr0 = now();
   for (int s = 1; s <= steps; s++)
   {
      double d = dt; rk.Step(x, t, d);                                 // Solver.cpp:535-551
      if (!out.empty())
         for (int q : snaps) if (q == s) { dump(out + "/x_step" + std::to_string(s) + ".f64", x.GetData(), (size_t)6 * N); }
   }
   double tRun = now() - tr0;
   // median of `repeats` timed windows of `steps` RK4 steps each, and the share of the window spent in the
   // operator application alone (4 SparseMatrix::Mult per step, sparsemat.cpp:921-931) against the 7 AXPY passes of RK4Solver
   const int repeats = bench && a.count("repeats") ? std::max(1, std::stoi(a["repeats"])) : 1;
   std::vector<double> runs{tRun};
   for (int r = 1; r < repeats; r++)
   {
      double q0 = now();
      for (int s = 0; s < steps; s++) { double d = dt; rk.Step(x, t, d); }
      runs.push_back(now() - q0);
   }
   std::vector<double> sorted = runs; std::sort(sorted.begin(), sorted.end());
   const double tMed = sorted[sorted.size() / 2];
   double tSpmv = 0.0;
   if (bench && a.count("spmv-share"))
   {
      Vector k(6 * N); op.SetTime(t);
      double q0 = now();
      for (int s = 0; s < 4 * steps; s++) { op.Mult(x, k); }
      tSpmv = now() - q0;
   }
   if (!out.empty()) { dump(out + "/x_final.f64", x.GetData(), (size_t)6 * N); }

   int threads = 1;
#ifdef _OPENMP
   threads = bench ? omp_get_max_threads() : 1;
#endif
   std::ostringstream js;
   js.precision(17);
   js << "{\"dim\": " << dim << ", \"order\": " << pd.order << ", \"alpha\": " << pd.alpha
      << ", \"ne\": " << mesh.GetNE() << ", \"nv\": " << mesh.GetNV() << ", \"nbe\": " << mesh.GetNBE()
      << ", \"np\": " << (mesh.GetNE() ? N / mesh.GetNE() : 0) << ", \"n\": " << N
      << ", \"nnz\": " << op.A->NumNonZeroElems() << ", \"nnz_tfsf\": " << (op.Atfsf ? op.Atfsf->NumNonZeroElems() : 0)
      << ", \"tfsf_dofs\": " << op.src.dof.size()
      << ", \"dt\": " << dt << ", \"t0\": " << tstart << ", \"steps\": " << steps << ", \"warmup\": " << warm
      << ", \"t_final\": " << t << ", \"norm_final\": " << x.Norml2()
      << ", \"tfsf_applied\": " << op.napplied << ", \"tfsf_skipped\": " << op.nskipped
      << ", \"pw\": {\"on\": " << (pd.pw.on ? "true" : "false") << ", \"spread\": " << pd.pw.spread << ", \"mean1d\": " << pd.pw.mean1d
      << ", \"freq\": " << pd.pw.freq << ", \"pol\": [" << pd.pw.pol[0] << ", " << pd.pw.pol[1] << ", " << pd.pw.pol[2]
      << "], \"dir\": [" << pd.pw.dir[0] << ", " << pd.pw.dir[1] << ", " << pd.pw.dir[2] << "]}"
      << ", \"assemble_s\": " << tAsm << ", \"run_s\": " << tMed << ", \"threads\": " << threads << ", \"repeats\": " << repeats << ", \"runs_s\": [";
   for (size_t r = 0; r < runs.size(); r++) { js << (r ? ", " : "") << runs[r]; }
   js << "], \"spmv_s\": " << tSpmv << ", \"spmv_share\": " << (tMed > 0 ? tSpmv / tMed : 0.0)
      << ", \"dof_updates_per_s\": " << (tMed > 0 ? 6.0 * N * 4.0 * steps / tMed : 0.0) << "}";
   std::cout << js.str() << std::endl;
   if (!out.empty()) { std::ofstream f(out + "/meta.json"); f << js.str() << "\n"; }
   return 0;
}

// known-answer check: the nine blocks of Hesthaven2DTest.cpp:234-553.
// The JSON is a flat list written by extract_known_answers.py:
//   name rows cols v00 v01 ...   (one record per line, whitespace separated)
static int cmdKnown(const std::string &path, const std::string &meshFile)
{
   Problem pd; pd.mesh.reset(new Mesh(Mesh::LoadFromFile(meshFile.c_str(), 1, 0)));
   pd.order = 1; pd.alpha = 1.0; pd.bdr[2] = BC_PEC;                   // GeomTagToBoundary{{2,PEC}}, EvolutionOptions default alpha
   DG_FECollection fec(1, 2, BasisType::GaussLobatto);
   FiniteElementSpace fes(pd.mesh.get(), &fec);
   Factory F(pd, fes);
   std::ifstream in(path); if (!in) { fprintf(stderr, "cannot read %s\n", path.c_str()); return 2; }
   std::string name; int rows, cols; double worst = 0; int bad = 0, n = 0;
   while (in >> name >> rows >> cols)
   {
      std::vector<double> ref(rows * (size_t)cols); for (auto &v : ref) { in >> v; }
      std::unique_ptr<BilinearForm> mi, b;
      if      (name == "2D_Operator_ZeroNormal_PEC")          { mi = F.MInv(FE); b = F.ZeroNormal(FE); }
      else if (name == "2D_Operator_OneNormal_nxEZ_HX_PEC")   { mi = F.MInv(FE); b = F.OneNormal(FH, 0); }
      else if (name == "2D_Operator_OneNormal_nyEZ_HY_PEC")   { mi = F.MInv(FE); b = F.OneNormal(FH, 1); }
      else if (name == "2D_Operator_OneNormal_nyHX_EZ_PEC")   { mi = F.MInv(FH); b = F.OneNormal(FE, 1); }
      else if (name == "2D_Operator_OneNormal_nxHY_EZ_PEC")   { mi = F.MInv(FH); b = F.OneNormal(FE, 0); }
      else if (name == "2D_Operator_TwoNormal_nxHXnx_HX_PEC") { mi = F.MInv(FH); b = F.TwoNormal(FH, 0, 0); }
      else if (name == "2D_Operator_TwoNormal_nxHXny_HY_PEC") { mi = F.MInv(FH); b = F.TwoNormal(FH, 0, 1); }
      else if (name == "2D_Operator_TwoNormal_nyHYnx_HY_PEC") { mi = F.MInv(FH); b = F.TwoNormal(FH, 1, 0); }
      else if (name == "2D_Operator_TwoNormal_nyHYny_HY_PEC") { mi = F.MInv(FH); b = F.TwoNormal(FH, 1, 1); }
      else { fprintf(stderr, "unknown known-answer %s\n", name.c_str()); return 2; }
      std::unique_ptr<SparseMatrix> op(prod(*mi, *b));
      std::unique_ptr<DenseMatrix> d(op->ToDenseMatrix());
      double md = 0;
      for (int i = 0; i < rows; i++) for (int j = 0; j < cols; j++) { md = std::max(md, std::abs((*d)(i, j) - ref[i * (size_t)cols + j])); }
      printf("%-40s max|diff| = %.3e %s\n", name.c_str(), md, md < 1e-8 ? "OK" : "FAIL");
      worst = std::max(worst, md); bad += !(md < 1e-8); n++;
   }
   printf("known-answers: %d checked, %d failed, worst %.3e\n", n, bad, worst);
   return (bad || n == 0) ? 1 : 0;
}

#ifdef DGTD_WITH_B200_SHELL
// ----------------------------------------------------------------------------
// `shell`: the product's drop-in shells (dgtd_b200/mfem_shell/B200Evolution.h) driven exactly like the reference drives
// its own operator — mfem::TimeDependentOperator::Mult, mfem::RK4Solver::Step — next to the reference-based GlobalOracle
// on the same mfem::FiniteElementSpace.  Needs a GPU (the product has no CPU path); this binary is the checker.
// ----------------------------------------------------------------------------
#include "B200Adaptor.h"
// Stand-ins for the reference's Model / SourcesManager / EvolutionOptions with the SAME accessor names and value types
// (src/components/Model.h:131-156, Material.h:51-53, src/solver/SourcesManager.h:34, src/components/Sources.h,
// src/math/Function.h:47-136, 328-409, src/evolution/EvolutionOptions.h:13-20, src/components/Types.h:48-56): the real
// classes need MPI (Model owns a ParMesh), which this image lacks, so the adaptor template is instantiated on these.
namespace mock {
enum class BdrCond { PEC, PMC, SMA, SurfaceCond, NearToFarField = 201, TotalFieldIn = 301, SGBC = 401 };
enum FieldType { E = 0, H = 1 };
struct Material
{
   double eps, mu, sigma;
   double getPermittivity() const { return eps; }
   double getPermeability() const { return mu; }
   double getConductivity() const { return sigma; }
};
struct Function { virtual ~Function() = default; };
struct Gaussian : Function
{
   double spread_; Vector mean_;
   Gaussian(double s, double m) : spread_(s), mean_(1) { mean_[0] = m; }
   double spread() const { return spread_; }
   const Vector &mean() const { return mean_; }
};
struct ModulatedGaussian : Function
{
   double spread_, freq_; Vector mean_;
   ModulatedGaussian(double s, double m, double f) : spread_(s), freq_(f), mean_(1) { mean_[0] = m; }
   double spread() const { return spread_; }
   const Vector &mean() const { return mean_; }
   double frequency() const { return freq_; }
};
struct EHFieldFunction { virtual ~EHFieldFunction() = default; };
struct Planewave : EHFieldFunction
{
   std::unique_ptr<Function> function_; Vector polarization_, propagation_; FieldType fieldtype_;
   Function *function() { return function_.get(); }
   const Vector &polarization() const { return polarization_; }
   const Vector &propagation() const { return propagation_; }
   FieldType fieldType() const { return fieldtype_; }
};
struct Source { virtual ~Source() = default; };
struct InitialField : Source {};
struct TotalField : Source
{
   std::unique_ptr<EHFieldFunction> function_;
   EHFieldFunction *function() { return function_.get(); }
};
struct SourcesManager { std::vector<std::unique_ptr<Source>> sources; };
struct EvolutionOptions { int op = 1; int order = 2; double alpha = 1.0; };   // EvolutionOperatorType::Global = 1
struct Model
{
   Mesh &mesh;
   std::map<int, BdrCond> bdr, intBdr;
   std::map<int, Material> mat;
   std::map<BdrCond, Array<int>> tfsf;
   explicit Model(Mesh &m) : mesh(m) {}
   const std::map<int, BdrCond> &getGeomTagToBoundaryCond() const { return bdr; }
   const std::map<int, BdrCond> &getGeomTagToIntBoundaryCond() const { return intBdr; }
   const std::map<int, Material> &getGeomTagToMaterial() const { return mat; }
   std::map<BdrCond, Array<int>> &getTotalFieldScatteredFieldToMarker() { return tfsf; }
   Mesh &getSerialMesh() { return mesh; }
};
using Types = maxwell::B200SourceTypes<TotalField, Planewave, Gaussian, ModulatedGaussian>;
}  // namespace mock
static double relL2(const Vector &a, const Vector &b)
{
   Vector d(a); d -= b; double nb = b.Norml2(); return d.Norml2() / (nb > 0 ? nb : 1.0);
}
static int cmdShell(std::map<std::string, std::string> &a)
{
   Problem pd = problemFromArgs(a);
   Mesh &mesh = *pd.mesh; const int dim = mesh.Dimension();
   DG_FECollection fec(pd.order, dim, BasisType::GaussLobatto);
   FiniteElementSpace fes(&mesh, &fec);
   const int N = fes.GetNDofs();
   Factory F(pd, fes);
   GlobalOracle op(N);
   op.A = buildGlobal(F);
   std::vector<double> xyz; nodeCoords(fes, xyz);
   op.pw = pd.pw;
   if (!pd.tfsf_tags.empty()) { op.Atfsf = buildTFSF(F); op.src = buildTFSFSource(F, xyz); }
   // the reference's objects as its driver would fill them (driver.cpp:1012-1041 splits the boundary tags into true and
   // interior ones; buildGaussianPlanewave / buildModulatedGaussianPlanewave, driver.cpp:483-519)
   mock::Model model(mesh);
   mock::SourcesManager srcs;
   mock::EvolutionOptions eo; eo.order = pd.order; eo.alpha = pd.alpha; eo.op = 1;
   {
      std::set<int> interiorTag;
      for (int be = 0; be < mesh.GetNBE(); be++)
         if (pd.bdr.count(mesh.GetBdrAttribute(be)) && mesh.FaceIsInterior(mesh.GetBdrElementFaceIndex(be))) { interiorTag.insert(mesh.GetBdrAttribute(be)); }
      for (auto &kv : pd.bdr)
      {
         mock::BdrCond c = kv.second == BC_PEC ? mock::BdrCond::PEC : kv.second == BC_PMC ? mock::BdrCond::PMC : mock::BdrCond::SMA;
         (interiorTag.count(kv.first) ? model.intBdr : model.bdr)[kv.first] = c;
      }
      for (auto &kv : pd.mat) { model.mat[kv.first] = mock::Material{kv.second[0], kv.second[1], kv.second[2]}; }
      if (!pd.tfsf_tags.empty())
      {
         Array<int> mk(mesh.bdr_attributes.Max()); mk = 0;
         for (int t : pd.tfsf_tags) { mk[t - 1] = 1; }
         model.tfsf[mock::BdrCond::TotalFieldIn] = mk;
      }
      if (pd.pw.on)
      {
         auto pw = std::make_unique<mock::Planewave>();
         if (pd.pw.freq == 0.0) { pw->function_ = std::make_unique<mock::Gaussian>(pd.pw.spread, pd.pw.mean1d); }
         else { pw->function_ = std::make_unique<mock::ModulatedGaussian>(pd.pw.spread, pd.pw.mean1d, pd.pw.freq); }
         pw->polarization_.SetSize(3); pw->propagation_.SetSize(3);
         for (int d = 0; d < 3; d++) { pw->polarization_[d] = pd.pw.pol[d]; pw->propagation_[d] = pd.pw.dir[d]; }
         pw->fieldtype_ = pd.pw.fieldtype == FE ? mock::E : mock::H;
         auto tf = std::make_unique<mock::TotalField>(); tf->function_ = std::move(pw);
         srcs.sources.push_back(std::make_unique<mock::InitialField>());
         srcs.sources.push_back(std::move(tf));
      }
   }
   maxwell::B200EvolutionFor<mock::Types> ev(fes, model, srcs, eo);      // the reference's constructor signature
   const double dt = a.count("dt") ? std::stod(a["dt"]) : 1e-3;
   const int steps = a.count("steps") ? std::stoi(a["steps"]) : 3;
   const double t0 = a.count("t0") ? std::stod(a["t0"]) : 0.0;
   Vector x0; initState(x0, a.count("init") ? a["init"] : "random:1", xyz, N, dim);
   // (1) TimeDependentOperator::Mult
   Vector kr, kb;                               // kb arrives unsized, as in GlobalEvolution.cpp:807-810
   op.SetTime(t0); op.Mult(x0, kr); ev.SetTime(t0); ev.Mult(x0, kb);
   const double eMult = relL2(kb, kr);
   // (2) mfem::RK4Solver driving each operator (Solver.cpp:41-47, 124-125, 544)
   Vector xr(x0), xm(x0), xf(x0);
   { RK4Solver rk; rk.Init(op); double t = t0; for (int s = 0; s < steps; s++) { double d = dt; rk.Step(xr, t, d); } }
   { RK4Solver rk; rk.Init(ev); double t = t0; for (int s = 0; s < steps; s++) { double d = dt; rk.Step(xm, t, d); } }
   // (3) the fused B200RK4Solver in RK4Solver's place, and its device-resident loop
   { maxwell::B200RK4Solver rk; rk.Init(ev); double t = t0; for (int s = 0; s < steps; s++) { double d = dt; rk.Step(xf, t, d); } }
   Vector xres(x0);
   { maxwell::B200RK4Solver rk; rk.Init(ev); double t = t0; rk.Upload(xres); rk.Run(t, dt, steps); rk.Download(xres); }
   // (3b) Solver::run semantics (final short step) and an asynchronous probe snapshot taken while the loop goes on
   Vector xun(x0), xrun(x0);
   const double tEnd = t0 + (steps - 0.5) * dt;
   { RK4Solver rk; rk.Init(op); double t = t0; while (t <= tEnd - 1e-8 * dt) { double d = std::min(dt, tEnd - t); rk.Step(xrun, t, d); } }
   double eGather = 0.0;
   {
      maxwell::B200RK4Solver rk; rk.Init(ev); double t = t0; rk.Upload(xun);
      std::vector<long long> dofs; for (long long i = 0; i < N; i += 7) { dofs.push_back(i); }
      maxwell::B200Gather probe(ev, dofs);
      probe.Launch();                                         // the initial state ...
      long long n = 0; const bool ok = rk.RunUntil(t, dt, tEnd, 1, &n);   // ... while the loop runs
      probe.Wait();
      if (!ok || n != steps) { eGather = 1.0; }
      const mfem::Vector &snap = probe.Data(); const long long nl = probe.NumLocal();
      for (int c = 0; c < 6; c++) for (long long i = 0; i < nl; i++) { eGather = std::max(eGather, std::fabs(snap[c * nl + i] - x0[c * N + probe.OwnedDofs()[i]])); }
      rk.Download(xun);
   }
   const double eUntil = relL2(xun, xrun);
   const double eRk = relL2(xm, xr), eFused = relL2(xf, xr), eRes = relL2(xres, xr);
   // (4) a foreign operator through B200RK4Solver must reproduce RK4Solver bit for bit
   Vector xg(x0);
   { maxwell::B200RK4Solver rk; rk.Init(op); double t = t0; for (int s = 0; s < steps; s++) { double d = dt; rk.Step(xg, t, d); } }
   const double eGen = relL2(xg, xr);
   printf("{\"n\": %d, \"steps\": %d, \"mult_rel_l2\": %.3e, \"mfem_rk4_on_b200_rel_l2\": %.3e, \"fused_rk4_rel_l2\": %.3e, "
          "\"resident_run_rel_l2\": %.3e, \"run_until_rel_l2\": %.3e, \"gather_abs\": %.3e, \"generic_fallback_rel_l2\": %.3e, \"tfsf_applied\": %ld, \"tfsf_skipped\": %ld}\n",
          N, steps, eMult, eRk, eFused, eRes, eUntil, eGather, eGen, op.napplied, op.nskipped);
   const double tol = 1e-10;                    // north_star: 1e-10 relative L2 per step
   return (eMult < tol && eRk < tol && eFused < tol && eRes < tol && eUntil < tol && eGather == 0.0 && eGen == 0.0) ? 0 : 1;
}
#endif

int main(int argc, char **argv)
{
   if (argc < 2)
   {
      fprintf(stderr, "usage: dgtd_ref gen|bench --mesh <file|cart3d:n|cart2d:nx:ny|cart1d:n> [--refine r] --order p --alpha a\n"
              "          [--bdr a:pec,b:sma] [--bdr-all pec] [--tfsf t1,t2] [--pw spread:mean|auto:freq:px,py,pz:kx,ky,kz]\n"
              "          [--mat a:eps:mu:sigma] [--init random:seed|gauss:E:c:s:fdim:cx,cy,cz|resonant:c:mx,my|smooth|file:path]\n"
              "          [--dt dt] [--steps k] [--warmup w] [--t0 t] [--snap 1,2] [--out dir] [--dump-csr]\n"
              "       dgtd_ref known-answers <records.txt> <Maxwell2D_K2.mesh>\n");
      return 2;
   }
   std::string cmd = argv[1];
   if (cmd == "known-answers") { return cmdKnown(argv[2], argv[3]); }
   auto a = parseArgs(argc, argv, 2);
   if (cmd == "gen") { return cmdGen(a, false); }
   if (cmd == "bench") { return cmdGen(a, true); }
#ifdef DGTD_WITH_B200_SHELL
   if (cmd == "shell") { return cmdShell(a); }
#endif
   fprintf(stderr, "unknown command %s\n", cmd.c_str());
   return 2;
}
