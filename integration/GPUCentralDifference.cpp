// GPUCentralDifference.cpp -- the reference-side binding of the drop-in boundary, as a translation
// unit that REPLACES 10-Integrators/02-CentralDifference/CentralDifference.cpp at link time: same class name, the
// reference's header CentralDifference.hpp untouched, every other reference source unmodified.  The reference's own Driver,
// DynamicAnalysis, Recorder and Mesh run as they are; the explicit step happens on the device through include/svlgpu.h.
//
//   make -C oracle refgpu      ->  oracle/_ref/SeismoVLAB_refgpu.exe  (reference objects - CentralDifference.o + this file + libsvlgpu.so)
//
// What the class does (the reference method each part replaces is named at the method):
//   Initialize        walks the reference's Mesh (nodes, elements, materials, masses, constraints, dampings, the loads of the
//                     active combination) and hands it to the C-ABI builder; svlgpu_finalize plays CentralDifference::Initialize
//   ComputeNewStep    svlgpu_step for step k, then U, V, A of all total dofs back into the Eigen members that
//                     DynamicAnalysis::UpdateDomain reads through GetDisplacements / GetVelocities / GetAccelerations
//   ComputeReactionForce   zero vector (the per-step reaction pass is dropped from the step; REACTION
//                     recorders are served by the device recorder of the stand-alone driver, svl_b200/host)
//
// Two things the reference does not expose through public getters -- the material an element was built with and the amplitude
// series of a load -- are read with GCC's -fno-access-control (set for THIS file only in oracle/Makefile).  A maintainer would
// add two getters instead (Element::GetMaterialPrototype(), Load::GetAmplitudes()): INTEGRATION.md.
// Scope of this binding: lin3DHexa8 / lin2DQuad4 with Elastic3DLinear / Elastic2DPlaneStrain / Plastic3DJ2 / PlasticPlaneStrainJ2,
// ZeroLength1D + Viscous1DLinear dashpots, point masses, EQUAL constraints, Rayleigh (mass-proportional) damping, concentrated
// point loads; anything else stops the analysis with a message (no silent fallback to the CPU path).
#include <cstdio>
#include <cstring>
#include <iostream>
#include <map>
#include <memory>
#include <string>
#include <vector>
#include <strings.h>

#include "CentralDifference.hpp"
#include "Definitions.hpp"
#include "lin3DHexa8.hpp"
#include "lin2DQuad4.hpp"
#include "ZeroLength1D.hpp"
#include "Plastic3DJ2.hpp"
#include "PlasticPlaneStrainJ2.hpp"
#include "Viscous1DLinear.hpp"
#include "svlgpu.h"

namespace {
struct GpuState {
    svlgpu_model *h = nullptr;
    std::shared_ptr<LoadCombo> combo;
    bool failed = false;
};
std::map<const CentralDifference *, GpuState> g_state;

bool stop_with(const char *what) {
    std::cout << "\x1B[31m ERROR: \x1B[0mGPU CentralDifference: " << what << "\n";
    return true;
}
bool stop_gpu() { return stop_with(svlgpu_last_error()); }
}  // namespace

// CentralDifference.cpp:6-27 (the Assembler is kept for ComputeProgressiveForce)
CentralDifference::CentralDifference(std::shared_ptr<Mesh> &mesh, double TimeStep, double mtol, double ktol, double ftol) :
Integrator(mesh), dt(TimeStep){
    U.resize(numberOfTotalDofs); U.fill(0.0);
    V.resize(numberOfTotalDofs); V.fill(0.0);
    A.resize(numberOfTotalDofs); A.fill(0.0);
    Up.resize(numberOfTotalDofs); Up.fill(0.0);
    theAssembler = std::make_unique<Assembler>();
    theAssembler->SetMassTolerance(mtol);
    theAssembler->SetForceTolerance(ftol);
    theAssembler->SetStiffnessTolerance(ktol);
    Fbar = theAssembler->ComputeProgressiveForceVector(mesh);
}

CentralDifference::~CentralDifference(){
    auto it = g_state.find(this);
    if (it != g_state.end()) { if (it->second.h) svlgpu_destroy(it->second.h); g_state.erase(it); }
}

// CentralDifference.cpp:35-71: M, C, Keff = M/dt^2 + C/2dt, Up = U - dt V + dt^2/2 A -- all inside svlgpu_finalize
void
CentralDifference::Initialize(std::shared_ptr<Mesh> &mesh){
    GpuState &S = g_state[this];
    if (S.h) { svlgpu_destroy(S.h); S.h = nullptr; }
    S.failed = true;
    svlgpu_model *h = svlgpu_create((int)nDimensions, MassFormulation ? 1 : 0);
    if (!h) { stop_gpu(); return; }
    S.h = h;

    // ---- nodes (Node.cpp:80-117), in ascending tag order = the reference's std::map order
    std::map<unsigned int, std::shared_ptr<Node> > &Nodes = mesh->GetNodes();
    std::map<unsigned int, int> nidx;
    std::vector<int32_t> ndof, tot, fre;
    std::vector<double> xyz, U0(numberOfTotalDofs, 0.0), V0(numberOfTotalDofs, 0.0), A0(numberOfTotalDofs, 0.0);
    for (auto &it : Nodes) {
        const std::shared_ptr<Node> &n = it.second;
        nidx[it.first] = (int)ndof.size();
        ndof.push_back((int32_t)n->GetNumberOfDegreeOfFreedom());
        const std::vector<int> &t = n->GetTotalDegreeOfFreedom(), &f = n->GetFreeDegreeOfFreedom();
        const Eigen::VectorXd &x = n->GetCoordinates(), &u = n->GetDisplacements(), &v = n->GetVelocities(), &a = n->GetAccelerations();
        for (unsigned int c = 0; c < nDimensions; c++) xyz.push_back(x(c));
        for (size_t c = 0; c < t.size(); c++) {
            tot.push_back(t[c]); fre.push_back(f[c]);
            U0[t[c]] = u(c); V0[t[c]] = v(c); A0[t[c]] = a(c);
        }
    }
    if (svlgpu_set_nodes(h, (int)ndof.size(), ndof.data(), xyz.data(), tot.data(), fre.data(), (int)numberOfTotalDofs, (int)numberOfFreeDofs)) { stop_gpu(); return; }
    for (auto &it : Nodes) {                                         // point masses (Assembler.cpp:622-657)
        Eigen::VectorXd ms = it.second->GetMass();
        if (ms.size() == 0) continue;
        std::vector<double> mv(ms.size());
        for (int c = 0; c < (int)ms.size(); c++) mv[c] = ms(c);
        const int32_t id = nidx[it.first];
        if (svlgpu_add_nodal_mass(h, 1, &id, mv.data())) { stop_gpu(); return; }
    }
    for (auto &it : mesh->GetConstraints()) {                        // Mesh.cpp:360-375
        const std::shared_ptr<Constraint> &c = it.second;
        std::vector<int32_t> master;
        for (unsigned int q : c->GetMasterInformation()) master.push_back((int32_t)q);
        const std::vector<double> fac = c->GetCombinationFactors();
        if (svlgpu_add_constraint(h, it.first, (int)c->GetSlaveInformation(), (int)master.size(), master.data(), fac.data())) { stop_gpu(); return; }
    }

    // ---- elements in ascending tag order (Assembler.cpp:251); one material entry per distinct parameter set
    std::map<std::vector<double>, int> matIndex;
    auto material_of = [&](int kind, std::vector<double> par) -> int {
        std::vector<double> key = par; key.insert(key.begin(), (double)kind);
        auto f = matIndex.find(key);
        if (f != matIndex.end()) return f->second;
        const int id = svlgpu_add_material(h, kind, par.data(), (int)par.size());
        matIndex[key] = id;
        return id;
    };
    std::map<unsigned int, int> eidx;
    std::vector<std::pair<int, std::pair<double, double> > > rayleigh;    // element index -> (am, ak)
    int ne = 0;
    for (auto &it : mesh->GetElements()) {
        Element *e = it.second.get();
        const std::string name = e->GetName();
        int kind = 0, mat = -1;
        std::vector<double> attr;
        std::shared_ptr<Damping> damp;
        const Material *mp = nullptr;
        if (name == "lin3DHexa8") { lin3DHexa8 *q = static_cast<lin3DHexa8 *>(e); kind = SVLGPU_LIN3DHEXA8; mp = q->theMaterial[0].get(); damp = q->theDamping; }
        else if (name == "lin2DQuad4") { lin2DQuad4 *q = static_cast<lin2DQuad4 *>(e); kind = SVLGPU_LIN2DQUAD4; mp = q->theMaterial[0].get(); damp = q->theDamping; attr.push_back(q->t); }
        else if (name == "ZeroLength1D") { ZeroLength1D *q = static_cast<ZeroLength1D *>(e); kind = SVLGPU_ZEROLENGTH1D; mp = q->theMaterial.get(); attr.push_back((double)q->theDirection); }
        else { stop_with(("element " + name + " is not on the device path").c_str()); return; }
        const std::string mname = const_cast<Material *>(mp)->GetName();
        if (mname == "Elastic3DLinear") mat = material_of(SVLGPU_ELASTIC3DLINEAR, {mp->GetElasticityModulus(), mp->GetPoissonRatio(), mp->GetDensity()});
        else if (mname == "Elastic2DPlaneStrain") mat = material_of(SVLGPU_ELASTIC2DPLANESTRAIN, {mp->GetElasticityModulus(), mp->GetPoissonRatio(), mp->GetDensity()});
        else if (mname == "Plastic3DJ2") { const Plastic3DJ2 *j = static_cast<const Plastic3DJ2 *>(mp); mat = material_of(SVLGPU_PLASTIC3DJ2, {j->K, j->G, j->Rho, j->H, j->beta, j->SigmaY}); }
        else if (mname == "PlasticPlaneStrainJ2") { const PlasticPlaneStrainJ2 *j = static_cast<const PlasticPlaneStrainJ2 *>(mp); mat = material_of(SVLGPU_PLASTICPLANESTRAINJ2, {j->K, j->G, j->Rho, j->H, j->beta, j->SigmaY}); }
        else if (mname == "Viscous1DLinear") mat = material_of(SVLGPU_VISCOUS1DLINEAR, {static_cast<const Viscous1DLinear *>(mp)->eta});
        else { stop_with(("material " + mname + " is not on the device path").c_str()); return; }
        if (mat < 0) { stop_gpu(); return; }
        std::vector<int32_t> conn;
        for (unsigned int nt : e->GetNodes()) conn.push_back(nidx[nt]);
        const int32_t mi = mat;
        if (svlgpu_add_elements(h, kind, 1, conn.data(), &mi, attr.empty() ? nullptr : attr.data(), (int)attr.size()) < 0) { stop_gpu(); return; }
        if (damp && strcasecmp(damp->GetName().c_str(), "Rayleigh") == 0 && kind != SVLGPU_ZEROLENGTH1D) {
            const std::vector<double> p = damp->GetParameters();     // lin3DHexa8.cpp:354-366: C_e = am M_e + ak K_e
            rayleigh.push_back({ne, {p[0], p[1]}});
        }
        eidx[it.first] = ne++;
    }
    for (auto &r : rayleigh) {
        const int32_t e = r.first;
        if (svlgpu_set_rayleigh(h, 1, &e, r.second.first, r.second.second)) { stop_gpu(); return; }
    }

    // ---- loads of the active combination (Assembler::ComputeExternalForceVector, Assembler.cpp:290-489)
    if (S.combo) {
        std::map<unsigned int, std::shared_ptr<Load> > &Loads = mesh->GetLoads();
        const std::vector<unsigned int> ids = S.combo->GetLoadCombination();
        const std::vector<double> fac = S.combo->GetLoadFactors();
        for (size_t q = 0; q < ids.size(); q++) {
            const std::shared_ptr<Load> &L = Loads[ids[q]];
            const unsigned int cls = L->GetClassification();
            if (cls != POINTLOAD_CONCENTRATED_CONSTANT && cls != POINTLOAD_CONCENTRATED_DYNAMIC) { stop_with("only concentrated point loads are bound in this translation unit"); return; }
            std::vector<int32_t> nodes;
            for (unsigned int nt : L->GetNodes()) nodes.push_back(nidx[nt]);
            double dir[3] = {0.0, 0.0, 0.0};
            for (int c = 0; c < (int)L->ForceDirection.size() && c < 3; c++) dir[c] = L->ForceDirection(c);
            if (svlgpu_add_point_load(h, (int)nodes.size(), nodes.data(), 3, dir, (int)L->ForceAmplitude.size(), L->ForceAmplitude.data(), fac[q])) { stop_gpu(); return; }
        }
    }
    if (svlgpu_set_initial_state(h, U0.data(), V0.data(), A0.data())) { stop_gpu(); return; }
    if (svlgpu_finalize(h, dt, 0)) { stop_gpu(); return; }
    for (unsigned int i = 0; i < numberOfTotalDofs; i++) { U(i) = U0[i]; V(i) = V0[i]; A(i) = A0[i]; }
    S.failed = false;
    svlgpu_counters c;
    svlgpu_get_counters(h, &c);
    std::cout << " GPU CentralDifference: " << c.n_elements << " elements on the device (lattice nodes " << c.n_block_nodes
              << ", neighbour-list nodes " << c.n_nbr_nodes << ", Gauss-point elements " << c.n_generic_elements << ")\n";
}

void
CentralDifference::SetLoadCombination(std::shared_ptr<LoadCombo> &combo){
    theAssembler->SetLoadCombination(combo);
    g_state[this].combo = combo;
}

void
CentralDifference::SetAlgorithm(std::shared_ptr<Algorithm> &algorithm){
    theAlgorithm = algorithm;                  // kept for interface parity: the solve is fused into the device step
}

const Eigen::VectorXd&
CentralDifference::GetDisplacements(){ return U; }
const Eigen::VectorXd&
CentralDifference::GetVelocities(){ return V; }
const Eigen::VectorXd&
CentralDifference::GetAccelerations(){ return A; }
const Eigen::VectorXd&
CentralDifference::GetPMLHistoryVector(){ return Ubar; }

// CentralDifference.cpp:123-152 + Linear.cpp:22-56 + Algorithm.cpp:18-56: effective force, diagonal solve, state update
bool
CentralDifference::ComputeNewStep(std::shared_ptr<Mesh>& /*mesh*/, unsigned int k){
    GpuState &S = g_state[this];
    if (S.failed || !S.h) return stop_with("the device model was not built");
    if (svlgpu_step(S.h, (int)k, (int)k + 1, 1)) return stop_gpu();
    std::vector<double> buf(numberOfTotalDofs);
    Eigen::VectorXd *dst[3] = {&U, &V, &A};
    for (int f = 0; f < 3; f++) {
        if (svlgpu_get_state(S.h, f, nullptr, 0, buf.data())) return stop_gpu();
        for (unsigned int i = 0; i < numberOfTotalDofs; i++) {
            if (buf[i] != buf[i]) return stop_with("NaN in the response");
            (*dst[f])(i) = buf[i];
        }
    }
    return false;
}

// CentralDifference.cpp:155-171: dropped from the per-step path
Eigen::VectorXd
CentralDifference::ComputeReactionForce(std::shared_ptr<Mesh>& /*mesh*/, unsigned int /*k*/){
    Eigen::VectorXd R(numberOfTotalDofs); R.fill(0.0);
    return R;
}

// CentralDifference.cpp:174-186 (phase hand-over: stays on the host, runs once per analysis)
Eigen::VectorXd
CentralDifference::ComputeProgressiveForce(std::shared_ptr<Mesh> &mesh, unsigned int k){
    Eigen::VectorXd Fext = theAssembler->ComputeExternalForceVector(mesh, k);
    Eigen::VectorXd Force = Fext + Fbar;
    return Force;
}

// the three entry points of the Linear algorithm are not called any more (ComputeNewStep does not go through it)
void
CentralDifference::ComputeSupportMotionVector(std::shared_ptr<Mesh>& /*mesh*/, Eigen::VectorXd &Feff, double /*factor*/, unsigned int /*k*/){ (void)Feff; }
void
CentralDifference::ComputeEffectiveForce(std::shared_ptr<Mesh>& /*mesh*/, Eigen::VectorXd &Feff, double /*factor*/, unsigned int /*k*/){ Feff.fill(0.0); }
void
CentralDifference::ComputeEffectiveStiffness(std::shared_ptr<Mesh>& /*mesh*/, Eigen::SparseMatrix<double> &Keff){ (void)Keff; }
