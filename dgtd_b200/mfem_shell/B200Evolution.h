// B200Evolution / B200RK4Solver — the two thin C++ shells that put libdgtd_b200.so behind the reference's own interfaces.
//
//   maxwell::B200Evolution : mfem::TimeDependentOperator     drop-in for maxwell::GlobalEvolution / HesthavenEvolution
//        (src/evolution/GlobalEvolution.h:19-22, HesthavenEvolution.h:20-21; selected in Solver::assignEvolutionOperator,
//         src/solver/Solver.cpp:22-39).  Mult(in, out) has the reference's contract: `in` is the 6N state
//         [Ex,Ey,Ez,Hx,Hy,Hz] (src/evolution/Fields.h:45-65), `out` may arrive unsized (GlobalEvolution.cpp:807-810),
//         time comes from GetTime().
//   maxwell::B200RK4Solver : mfem::ODESolver                  drop-in for mfem::RK4Solver (linalg/ode.cpp:109-136) as
//        chosen by Solver::assignODESolver (src/solver/Solver.cpp:41-47): Step(x, t, dt) advances the HOST vector x.
//        With a B200Evolution underneath the four stages run fused on the device (4 launches, no k vector, no AXPYs);
//        any other operator is handed to an owned mfem::RK4Solver, so it can replace RK4Solver unconditionally.
//
// Header-only, depends on <mfem.hpp> and include/dgtd_b200.h only (serial or parallel MFEM: it takes the
// mfem::FiniteElementSpace base that ParFiniteElementSpace derives from).  Errors become std::runtime_error like the
// reference's (src/solver/Solver.cpp:37, 74).  B200Adaptor.h fills B200Problem from the reference's own
// maxwell::Model / SourcesManager / EvolutionOptions and adds the constructor with the reference's signature.
#pragma once
#include <mfem.hpp>

#include <array>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include "dgtd_b200.h"

namespace maxwell {

// What the reference ctor reads from Model / EvolutionOptions / Sources (GlobalEvolution.cpp:32-364), as plain data.
struct B200Problem {
    int order = 3;                                        // EvolutionOptions::order
    double alpha = 1.0;                                   // EvolutionOptions::alpha (upwind_alpha)
    std::map<int, int> bdr;                               // boundary attribute -> DGTD_BC_PEC|PMC|SMA (Model::getGeomTagToBoundaryCond)
    std::vector<int> tfsfTags;                            // attributes of TF/SF interior faces (Model::getTotalFieldScatteredFieldToMarker)
    std::map<int, std::array<double, 3>> materials;       // element attribute -> eps, mu, sigma (Model::getGeomTagToMaterial)
    dgtd_planewave planewave{};                           // TotalField plane wave (Sources, math/Function.h:328-409)
    bool tfsfGate = true;                                 // true: `global` semantics (skip when ||s|| < 1e-8), false: `hesthaven`
    int device = 0;
    int rank = 0, nranks = 1;                             // one context per GPU/rank
    const int *partitioning = nullptr;                    // element -> rank (Mesh::GeneratePartitioning contract) or null
};

class B200Evolution : public mfem::TimeDependentOperator {
public:
    // fes: the space the reference hands to its operators (GlobalEvolution.h:19-22) — on one rank a FiniteElementSpace on the
    // whole mesh; in the MPI build the rank's ParFiniteElementSpace, and then `serialMesh` is the undivided mesh every rank
    // holds (Model::getSerialMesh, Model.h:131) together with pb.partitioning, the array the ParMesh was built from
    // (driver.cpp:1269-1277).  Vectors are sized like the reference's: 6 * fes.GetNDofs() (the rank's own dofs).
    B200Evolution(mfem::FiniteElementSpace &fes, const B200Problem &pb, mfem::Mesh *serialMesh = nullptr)
        : mfem::TimeDependentOperator(6 * fes.GetNDofs()), fes_(fes), parlocal_(pb.nranks > 1)   // GlobalEvolution.cpp:34
    {
        if (pb.nranks > 1 && (!serialMesh || !pb.partitioning))
            throw std::runtime_error("B200Evolution: a multi-rank operator needs the serial mesh and the partitioning.");
        mfem::Mesh &m = serialMesh ? *serialMesh : *fes.GetMesh();
        const int dim = m.Dimension(), nv = m.GetNV(), ne = m.GetNE(), nbe = m.GetNBE();
        std::vector<double> verts(3 * (size_t)nv, 0.0);
        for (int v = 0; v < nv; v++) for (int c = 0; c < m.SpaceDimension(); c++) verts[3 * (size_t)v + c] = m.GetVertex(v)[c];
        std::vector<int> elems((size_t)ne * (dim + 1)), eattr(ne), bdr((size_t)nbe * dim), battr(nbe);
        mfem::Array<int> vs;
        for (int e = 0; e < ne; e++) {
            if (m.GetElementGeometry(e) != (dim == 1 ? mfem::Geometry::SEGMENT : dim == 2 ? mfem::Geometry::TRIANGLE : mfem::Geometry::TETRAHEDRON))
                throw std::runtime_error("B200Evolution: simplex meshes only (segments, triangles, tetrahedra).");
            m.GetElementVertices(e, vs);
            for (int k = 0; k <= dim; k++) elems[(size_t)e * (dim + 1) + k] = vs[k];
            eattr[e] = m.GetAttribute(e);
        }
        for (int b = 0; b < nbe; b++) {
            m.GetBdrElementVertices(b, vs);
            for (int k = 0; k < dim; k++) bdr[(size_t)b * dim + k] = vs[k];
            battr[b] = m.GetBdrAttribute(b);
        }
        check(dgtd_mesh_from_arrays(dim, nv, verts.data(), ne, elems.data(), eattr.data(), nbe, bdr.data(), battr.data(), &mesh_));
        std::vector<int> ba, bc, ma; std::vector<double> mv;
        for (auto &kv : pb.bdr) { ba.push_back(kv.first); bc.push_back(kv.second); }
        for (auto &kv : pb.materials) { ma.push_back(kv.first); mv.insert(mv.end(), kv.second.begin(), kv.second.end()); }
        dgtd_options o{};
        o.order = pb.order; o.alpha = pb.alpha;
        o.n_bdr = (int)ba.size(); o.bdr_attr = ba.data(); o.bdr_cond = bc.data();
        o.n_tfsf = (int)pb.tfsfTags.size(); o.tfsf_attr = pb.tfsfTags.data();
        o.n_mat = (int)ma.size(); o.mat_attr = ma.data(); o.mat_eps_mu_sigma = mv.data();
        o.pw = pb.planewave; o.tfsf_gate = pb.tfsfGate ? 1 : 0; o.device = pb.device;
        o.rank = pb.rank; o.nranks = pb.nranks; o.partitioning = pb.partitioning;
        // a constructor that throws does not run the destructor: release what exists before leaving
        auto bail = [&](const std::string &msg) { dgtd_destroy(ctx_); ctx_ = nullptr; dgtd_mesh_destroy(mesh_); mesh_ = nullptr; throw std::runtime_error("B200Evolution: " + msg); };
        if (dgtd_create(mesh_, &o, &ctx_) != DGTD_OK) bail(dgtd_last_error());
        long long n = 0, nl = 0; int np = 0;
        if (dgtd_sizes(ctx_, &n, &np, nullptr, &nl) != DGTD_OK) bail(dgtd_last_error());
        if (6 * (parlocal_ ? nl : n) != Height() || np != fes.GetFE(0)->GetDof())
            bail("the finite element space is not the order-p L2 Gauss-Lobatto space the kernels assume.");
    }
    ~B200Evolution() override { dgtd_destroy(ctx_); dgtd_mesh_destroy(mesh_); }
    B200Evolution(const B200Evolution &) = delete;
    B200Evolution &operator=(const B200Evolution &) = delete;

    // out = f(GetTime(), in); host vectors in the reference layout
    void Mult(const mfem::Vector &in, mfem::Vector &out) const override
    {
        if (in.Size() != Height()) throw std::runtime_error("B200Evolution::Mult: input size does not match 6*NDofs.");
        if (out.Size() != Height()) out.SetSize(Height());
        if (parlocal_) check(dgtd_mult_parlocal(ctx_, GetTime(), in.HostRead(), out.HostWrite()));
        else check(dgtd_mult(ctx_, GetTime(), in.HostRead(), out.HostWrite(), 0));
    }
    // host state vectors sized like the reference's Fields::allDOFs on this rank (6 * fes.GetNDofs())
    void setState(const mfem::Vector &x) const { check(parlocal_ ? dgtd_set_state_parlocal(ctx_, x.HostRead()) : dgtd_set_state(ctx_, x.HostRead())); }
    void getState(mfem::Vector &x) const
    {
        if (x.Size() != Height()) x.SetSize(Height());
        check(parlocal_ ? dgtd_get_state_parlocal(ctx_, x.HostWrite()) : dgtd_get_state(ctx_, x.HostWrite()));
    }

    dgtd_ctx *context() const { return ctx_; }
    mfem::FiniteElementSpace &fes() const { return fes_; }
    // Fields::getNorml2 replacement for the per-step stability check (Solver.cpp:500-516), device reduction
    double residentNorml2() const { double s = 0; check(dgtd_norm2_local(ctx_, &s)); return std::sqrt(s); }

    static void check(int rc) { if (rc != DGTD_OK) throw std::runtime_error(std::string("dgtd_b200: ") + dgtd_last_error()); }

private:
    mfem::FiniteElementSpace &fes_;
    bool parlocal_ = false;               // multi-rank: vectors hold this rank's dofs in its ParMesh's element order
    dgtd_mesh *mesh_ = nullptr;
    dgtd_ctx *ctx_ = nullptr;
};

// Probe / surface-export snapshots without stopping the time loop (replaces the six TransferMaps + host copy per export step
// of RCSSurfaceExporter::transferFields, RCSSurfaceExporter.cpp:71-79, and the point/field probe reads of ProbesManager):
// Launch() queues a gather kernel and an asynchronous device-to-host copy of the six fields at a fixed list of scalar dofs
// (Fields numbering, element * Np + node), Wait() makes Data() readable: [Ex|Ey|Ez|Hx|Hy|Hz][NumLocal()].
class B200Gather {
public:
    B200Gather(B200Evolution &ev, const std::vector<long long> &dofs) : ev_(ev)
    {
        long long n = 0;
        B200Evolution::check(dgtd_gather_create(ev.context(), (long long)dofs.size(), dofs.data(), &g_, &n));
        owned_.resize((size_t)n); data_.SetSize((int)(6 * n));
        if (n) B200Evolution::check(dgtd_gather_dofs(g_, owned_.data()));
        data_.HostWrite();
    }
    ~B200Gather() { dgtd_gather_destroy(g_); }
    B200Gather(const B200Gather &) = delete;
    B200Gather &operator=(const B200Gather &) = delete;
    void Launch() { B200Evolution::check(dgtd_gather_launch(ev_.context(), g_, data_.HostWrite())); }
    void Wait() { B200Evolution::check(dgtd_gather_wait(ev_.context(), g_)); }
    long long NumLocal() const { return (long long)owned_.size(); }
    const std::vector<long long> &OwnedDofs() const { return owned_; }      // this rank's share of the list, output order
    const mfem::Vector &Data() const { return data_; }
private:
    B200Evolution &ev_;
    dgtd_gather *g_ = nullptr;
    std::vector<long long> owned_;
    mfem::Vector data_;
};

class B200RK4Solver : public mfem::ODESolver {
public:
    void Init(mfem::TimeDependentOperator &f) override
    {
        mfem::ODESolver::Init(f);
        b200_ = dynamic_cast<B200Evolution *>(&f);
        resident_ = false;
        if (!b200_) generic_.Init(f);     // any other operator: MFEM's own RK4Solver does the work
    }
    // mfem::ODESolver contract: x is a host vector; it is uploaded, advanced by one fused RK4 step and downloaded.
    void Step(mfem::Vector &x, mfem::real_t &t, mfem::real_t &dt) override
    {
        if (!b200_) { generic_.Step(x, t, dt); return; }
        b200_->setState(x);
        B200Evolution::check(dgtd_rk4_step(b200_->context(), t, dt));
        b200_->getState(x);
        t += dt;
        resident_ = true;
    }
    // Device-resident time loop (Solver::run body without the per-step host touches): upload once, run
    // nsteps, download when a probe/export is due.
    void Upload(const mfem::Vector &x) { need(); b200_->setState(x); resident_ = true; }
    void Run(mfem::real_t &t, mfem::real_t dt, int nsteps)
    {
        need(); if (!resident_) throw std::runtime_error("B200RK4Solver::Run: call Upload first.");
        B200Evolution::check(dgtd_rk4_run(b200_->context(), t, dt, nsteps));
        t += nsteps * dt;
    }
    // Solver::run (Solver.cpp:497-533) on the resident state: steps of min(dt, tFinal - t) until tFinal, the reference's
    // stability test on the state norm every `checkEvery` steps (0: never).  Returns false when the test fired.
    bool RunUntil(mfem::real_t &t, mfem::real_t dt, mfem::real_t tFinal, int checkEvery = 0, long long *nsteps = nullptr)
    {
        need(); if (!resident_) throw std::runtime_error("B200RK4Solver::RunUntil: call Upload first.");
        double tt = t; int unstable = 0;
        B200Evolution::check(dgtd_run_until(b200_->context(), &tt, dt, tFinal, checkEvery, nsteps, &unstable));
        t = tt;
        return unstable == 0;
    }
    void Download(mfem::Vector &x) { need(); b200_->getState(x); }

private:
    void need() const { if (!b200_) throw std::runtime_error("B200RK4Solver: the operator is not a B200Evolution."); }
    B200Evolution *b200_ = nullptr;
    bool resident_ = false;
    mfem::RK4Solver generic_;
};

}  // namespace maxwell
