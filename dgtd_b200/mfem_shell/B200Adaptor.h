// B200Adaptor — fills maxwell::B200Problem from the reference's own objects and gives B200Evolution the reference's
// constructor signature
//     (mfem::ParFiniteElementSpace&, Model&, SourcesManager&, EvolutionOptions&)      src/evolution/GlobalEvolution.h:19-22
// so that Solver::assignEvolutionOperator (src/solver/Solver.cpp:22-39) constructs it exactly like GlobalEvolution.
//
// Everything is a template over the reference's types: this header compiles against the reference tree (MPI build) as well
// as against stand-ins with the same accessor names (oracle/ref/dgtd_ref.cpp instantiates it with mock types in the shell
// checker, since Model / SourcesManager need MPI, which the build image lacks).  Accessors used, all existing in the
// reference unless marked (+):
//   EvolutionOptions  .order .alpha .op                                   src/evolution/EvolutionOptions.h:13-20
//   Model             getGeomTagToBoundaryCond()      map<tag, BdrCond>    src/components/Model.h:155-156
//                     getGeomTagToIntBoundaryCond()   map<tag, BdrCond>    Model.h:152-153  (PEC/PMC/SMA sheets inside the mesh)
//                     getGeomTagToMaterial()          map<tag, Material>   Model.h:141; Material::getPermittivity/Permeability/Conductivity (Material.h:51-53)
//                     getTotalFieldScatteredFieldToMarker()  map<BdrCond, Array<int>>   Model.h:148
//                     getSerialMesh()                                      Model.h:131
//   SourcesManager    .sources (iterable of unique_ptr<Source>)            src/solver/SourcesManager.h:34
//   TotalField        function()  (+)  -> EHFieldFunction*                 mirrors InitialField::function(), Sources.h:47
//   Planewave         polarization() propagation() fieldType() function()  (+)   members of math/Function.h:328-409
//   Gaussian          spread() mean()  (+)     ModulatedGaussian  spread() mean() frequency()  (+)    Function.h:47-136
// The six (+) one-line getters are the only additions the reference needs: its source classes keep their parameters private.
#pragma once
#include "B200Evolution.h"

namespace maxwell {

// BdrCond (src/components/Types.h:48-56) is an enum class {PEC, PMC, SMA, SurfaceCond, NearToFarField = 201, TotalFieldIn = 301,
// SGBC = 401}: only the first three are conditions of the evolution operator
template <class BdrCondT> inline int b200BoundaryCode(BdrCondT cond)
{
    switch (static_cast<int>(cond)) {
        case 0: return DGTD_BC_PEC;
        case 1: return DGTD_BC_PMC;
        case 2: return DGTD_BC_SMA;
        default: return -1;
    }
}

// The reference's source classes, named once by the integrator:
//   using B200RefTypes = B200SourceTypes<maxwell::TotalField, maxwell::Planewave, maxwell::Gaussian, maxwell::ModulatedGaussian>;
template <class TotalFieldT, class PlanewaveT, class GaussianT, class ModulatedGaussianT> struct B200SourceTypes {
    using TotalField = TotalFieldT; using Planewave = PlanewaveT; using Gaussian = GaussianT; using ModulatedGaussian = ModulatedGaussianT;
};

template <class Types, class ModelT, class SourcesManagerT, class EvolutionOptionsT>
B200Problem makeB200Problem(ModelT &model, SourcesManagerT &srcmngr, const EvolutionOptionsT &opts,
                            int rank = 0, int nranks = 1, const int *partitioning = nullptr, int device = 0)
{
    B200Problem pb;
    pb.order = opts.order; pb.alpha = opts.alpha;                         // EvolutionOptions.h:15-16
    pb.tfsfGate = static_cast<int>(opts.op) != 2;                         // Hesthaven (= 2) never skips the injection (HesthavenEvolution.cpp:97-124)
    pb.rank = rank; pb.nranks = nranks; pb.partitioning = partitioning; pb.device = device;
    for (const auto &kv : model.getGeomTagToBoundaryCond()) {
        const int code = b200BoundaryCode(kv.second);
        if (code > 0) pb.bdr[kv.first] = code;
    }
    for (const auto &kv : model.getGeomTagToIntBoundaryCond()) {          // the library sees from the mesh that these faces are interior
        const int code = b200BoundaryCode(kv.second);
        if (code > 0) pb.bdr[kv.first] = code;
        else if (static_cast<int>(kv.second) == 401) throw std::runtime_error("B200Evolution: SGBC interior boundaries are not supported.");
    }
    for (const auto &kv : model.getGeomTagToMaterial())
        pb.materials[kv.first] = {kv.second.getPermittivity(), kv.second.getPermeability(), kv.second.getConductivity()};
    for (const auto &kv : model.getTotalFieldScatteredFieldToMarker())   // marker[a] == 1: boundary attribute a + 1 is a TF/SF face
        for (int a = 0; a < kv.second.Size(); a++)
            if (kv.second[a] == 1) pb.tfsfTags.push_back(a + 1);
    for (const auto &src : srcmngr.sources) {
        auto *tf = dynamic_cast<typename Types::TotalField *>(src.get());
        if (!tf) continue;                                                // InitialField sources only set the initial state
        auto *pw = dynamic_cast<typename Types::Planewave *>(tf->function());
        if (!pw) throw std::runtime_error("B200Evolution: the TF/SF source must be a Planewave (dipoles are not supported).");
        if (pb.planewave.enabled) throw std::runtime_error("B200Evolution: one TF/SF plane wave per problem.");
        dgtd_planewave &w = pb.planewave;
        w.enabled = 1; w.fieldtype = static_cast<int>(pw->fieldType());   // E = 0, H = 1 (Types.h)
        for (int d = 0; d < 3; d++) { w.pol[d] = pw->polarization()[d]; w.dir[d] = pw->propagation()[d]; }
        if (auto *g = dynamic_cast<typename Types::Gaussian *>(pw->function())) { w.spread = g->spread(); w.mean1d = g->mean()[0]; w.freq = 0.0; }
        else if (auto *m = dynamic_cast<typename Types::ModulatedGaussian *>(pw->function())) { w.spread = m->spread(); w.mean1d = m->mean()[0]; w.freq = m->frequency(); }
        else throw std::runtime_error("B200Evolution: plane-wave profile must be Gaussian or ModulatedGaussian.");
    }
    return pb;
}

// B200Evolution with the reference's constructor signature.  `Types` names the reference's source classes (above).
// Multi-rank: one context per MPI rank on GPU `device`; `partitioning` is the array the driver handed to Model
// (driver.cpp:1269-1277) and the mesh is Model::getSerialMesh(), which every rank holds.
template <class Types> class B200EvolutionFor : public B200Evolution {
public:
    template <class ModelT, class SourcesManagerT, class EvolutionOptionsT>
    B200EvolutionFor(mfem::FiniteElementSpace &fes, ModelT &model, SourcesManagerT &srcmngr, EvolutionOptionsT &opts,
                     int rank = 0, int nranks = 1, const int *partitioning = nullptr, int device = 0)
        : B200Evolution(fes, makeB200Problem<Types>(model, srcmngr, opts, rank, nranks, partitioning, device),
                        nranks > 1 ? &model.getSerialMesh() : nullptr)
    {
    }
};

}  // namespace maxwell
