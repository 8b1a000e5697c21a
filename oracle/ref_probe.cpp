// TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// ref_probe.cpp -- thin extern "C" probes around the UNMODIFIED reference classes
// (compiled where they lie under /root/reference against oracle/shim, see
// oracle/Makefile target `ref`).  Linked into oracle/_ref/libsvlref_probe.so and
// used only in this container to (1) validate oracle/svl_oracle.c function by
// function and (2) generate the fixtures under tests/golden/.  No reference
// source is copied: this file only *calls* the reference's public class surface
// (Element.hpp:69-195, Material.hpp, Node.hpp).
#include <map>
#include <memory>
#include <vector>
#include <cstring>
#include <Eigen/Dense>
#include "Node.hpp"
#include "Load.hpp"
#include "Damping.hpp"
#include "Material.hpp"
#include "Elastic3DLinear.hpp"
#include "Elastic2DPlaneStrain.hpp"
#include "Plastic3DJ2.hpp"
#include "lin3DHexa8.hpp"
#include "lin2DQuad4.hpp"
#include "PML3DHexa8.hpp"
#include "PML2DQuad4.hpp"
#include "Definitions.hpp"

namespace {
struct Rig {
    std::map<unsigned int, std::shared_ptr<Node>> nodes;
    std::shared_ptr<Element> elem;
    std::shared_ptr<Damping> damp;
};

std::unique_ptr<Material> make_material(int kind, const double *p) {
    switch (kind) {
    case 1: return std::make_unique<Elastic3DLinear>(p[0], p[1], p[2]);
    case 2: return std::make_unique<Elastic2DPlaneStrain>(p[0], p[1], p[2]);
    case 3: return std::make_unique<Plastic3DJ2>(p[0], p[1], p[2], p[3], p[4], p[5]);
    }
    return nullptr;
}

// kind: 1 lin3DHexa8, 2 lin2DQuad4, 3 PML3DHexa8, 4 PML2DQuad4 (include/svlgpu.h)
Rig make_rig(int kind, const double *X, const double *U, int matkind, const double *mp,
             const double *attr, int lumped) {
    Rig r;
    const int nn = (kind == 1 || kind == 3) ? 8 : 4;
    const int nd = (kind == 1 || kind == 3) ? 3 : 2;
    const int ndof = kind == 1 ? 3 : kind == 2 ? 2 : kind == 3 ? 9 : 5;
    nDimensions = nd;
    MassFormulation = lumped != 0;
    std::vector<unsigned int> conn(nn);
    for (int i = 0; i < nn; i++) {
        Eigen::VectorXd c(nd);
        for (int k = 0; k < nd; k++) c(k) = X[nd * i + k];
        auto n = std::make_shared<Node>(ndof, c, false);
        std::vector<int> tot(ndof), fre(ndof);
        for (int k = 0; k < ndof; k++) tot[k] = fre[k] = ndof * i + k;
        n->SetTotalDegreeOfFreedom(tot);
        n->SetFreeDegreeOfFreedom(fre);
        Eigen::VectorXd u(ndof);
        u.fill(0.0);
        if (U) for (int k = 0; k < ndof; k++) u(k) = U[ndof * i + k];
        n->SetDisplacements(u);
        Eigen::VectorXd z(ndof);
        z.fill(0.0);
        n->SetIncrementalDisplacements(z);
        conn[i] = i + 1;
        r.nodes[i + 1] = n;
    }
    std::unique_ptr<Material> mat = make_material(matkind, mp);
    if (kind == 1) r.elem = std::make_shared<lin3DHexa8>(conn, mat, "GAUSS", 8);
    if (kind == 2) r.elem = std::make_shared<lin2DQuad4>(conn, mat, attr[0], "GAUSS", 4);
    if (kind == 3) r.elem = std::make_shared<PML3DHexa8>(conn, mat, std::vector<double>(attr, attr + 9), "GAUSS", 8);
    if (kind == 4) r.elem = std::make_shared<PML2DQuad4>(conn, mat, std::vector<double>(attr, attr + 8), "GAUSS", 4);
    r.elem->SetDomain(r.nodes);
    r.damp = std::make_shared<Damping>("Free", std::vector<double>());
    r.elem->SetDamping(r.damp);
    return r;
}

void copy_out(const Eigen::MatrixXd &A, double *out) {
    if (!out) return;
    for (int i = 0; i < A.rows(); i++)
        for (int j = 0; j < A.cols(); j++) out[i * A.cols() + j] = A(i, j);
}
} // namespace

extern "C" {

// UpdateState() with nodal displacements U, then ComputeInternalForces()
// (lin3DHexa8.cpp:86-107, 382-412).  f has ndof_elem entries.
int refprobe_internal_force(int kind, const double *X, const double *U, int matkind, const double *mp,
                            const double *attr, double *f) {
    Rig r = make_rig(kind, X, U, matkind, mp, attr, 1);
    r.elem->UpdateState();
    Eigen::VectorXd F = r.elem->ComputeInternalForces();
    for (int i = 0; i < F.size(); i++) f[i] = F(i);
    return (int)F.size();
}

// Element::Compute{Mass,Damping,Stiffness,PML}Matrix; any pointer may be NULL.
int refprobe_matrices(int kind, const double *X, int matkind, const double *mp, const double *attr,
                      int lumped, double *M, double *C, double *K, double *G) {
    Rig r = make_rig(kind, X, nullptr, matkind, mp, attr, lumped);
    if (M) copy_out(r.elem->ComputeMassMatrix(), M);
    if (C) copy_out(r.elem->ComputeDampingMatrix(), C);
    if (K) copy_out(r.elem->ComputeStiffnessMatrix(), K);
    if (G && (kind == 3)) copy_out(r.elem->ComputePMLMatrix(), G);
    return (int)r.elem->GetNumberOfDegreeOfFreedom();
}

// nsteps successive Material::UpdateState(eps,1) + CommitState() calls
// (Plastic3DJ2.cpp:206-259, 163-170); sig gets nsteps*6 stresses.
int refprobe_material_path(int matkind, const double *mp, int nsteps, int ncomp, const double *eps, double *sig) {
    std::unique_ptr<Material> mat = make_material(matkind, mp);
    for (int s = 0; s < nsteps; s++) {
        Eigen::VectorXd e(ncomp);
        for (int i = 0; i < ncomp; i++) e(i) = eps[ncomp * s + i];
        mat->UpdateState(e, 1);
        mat->CommitState();
        Eigen::VectorXd S = mat->GetStress();
        for (int i = 0; i < ncomp; i++) sig[ncomp * s + i] = S(i);
    }
    return 0;
}

// Element::ComputeDomainReductionForces (lin3DHexa8.cpp:660-718, lin2DQuad4.cpp:564-614).
// field: [nn][3*nd] rows (u,v,a) as in the .drm file for one time step, ext[nn] flags.
int refprobe_drm_force(int kind, const double *X, int matkind, const double *mp, const double *attr,
                       int lumped, const unsigned char *ext, const double *field, double *f) {
    Rig r = make_rig(kind, X, nullptr, matkind, mp, attr, lumped);
    const int nn = kind == 1 ? 8 : 4, nd = kind == 1 ? 3 : 2;
    auto load = std::make_shared<Load>(9 /*ELEMENTLOAD_DOMAIN_REDUCTION*/);
    for (int i = 0; i < nn; i++) {
        Eigen::MatrixXd S(1, 3 * nd);
        for (int c = 0; c < 3 * nd; c++) S(0, c) = field[3 * nd * i + c];
        if (ext[i]) S = -1.00 * S;                       // Driver.hpp:1714-1716
        load->AddDRMCondition(i + 1, ext[i] != 0);
        r.nodes[i + 1]->SetDomainReductionMotion(S);
    }
    Eigen::VectorXd F = r.elem->ComputeDomainReductionForces(load, 0);
    for (int i = 0; i < F.size(); i++) f[i] = F(i);
    return (int)F.size();
}

}
