/*
 * svl_oracle.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C, single-threaded-by-default CPU restatement of the arithmetic on
 * SeismoVLAB/SVL's explicit-dynamics hot path, written from the reference's
 * published algorithm (every function cites the reference file:line it
 * follows, paths relative to /root/reference/02-Run_Process/).  It exists to
 * CHECK the CUDA path: only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it.  The product path
 * (svl_b200/, libsvlgpu.so) never links or calls anything in oracle/.
 *
 * Parity pinning: validated against (a) the reference's own classes compiled
 * unmodified into oracle/_ref/libsvlref_probe.so (element vectors, matrices,
 * J2 return map), (b) NODE recorder histories written by the reference
 * executable oracle/_ref/SeismoVLAB.exe for CentralDifference, NewmarkBeta,
 * ExtendedNewmarkBeta and NewmarkBeta + NewtonRaphson runs, committed as
 * fixtures under tests/golden/ (generator: tests/golden/make_golden.py),
 * (c) the known answers of SURVEY.md App. B.5, (d) the reference's own
 * validation fixtures J05, J02, F02, F06, F03, F07 (OpenSees histories and
 * Gauss-point recorder files) and F11, J12 (their input files run through the
 * reference executable): tests/golden/fixtures/, tests/golden/J05/.
 */
#ifndef SVL_ORACLE_H
#define SVL_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* element kinds / material kinds: same numbers as include/svlgpu.h */
enum { SVLO_LIN3DHEXA8 = 1, SVLO_LIN2DQUAD4 = 2, SVLO_PML3DHEXA8 = 3, SVLO_PML2DQUAD4 = 4,
       SVLO_ZEROLENGTH1D = 5 /* attr[0] = direction; conn uses the first 2 slots */ };
enum { SVLO_ELASTIC3DLINEAR = 1, SVLO_ELASTIC2DPLANESTRAIN = 2, SVLO_PLASTIC3DJ2 = 3,
       SVLO_PLASTICPLANESTRAINJ2 = 4, SVLO_VISCOUS1DLINEAR = 5 /* par[0] = eta */ };

/* ---- element / material level ------------------------------------------- */
void svlo_elastic3d_C(double E, double nu, double C[36]);
void svlo_planestrain_C(double E, double nu, double C[9]);

/* X: 8x3 node coords (reference local order), U: 8x3 displacements           */
void svlo_hex8_strain(const double *X, const double *U, double eps[8][6]);
void svlo_hex8_force(const double *X, const double sig[8][6], double f[24]);
void svlo_hex8_mass(const double *X, double rho, int lumped, double M[24 * 24]);
void svlo_hex8_stiffness(const double *X, const double C[36], double K[24 * 24]);
void svlo_quad4_strain(const double *X, const double *U, double eps[4][3]);
void svlo_quad4_force(const double *X, double th, const double sig[4][3], double f[8]);
void svlo_quad4_mass(const double *X, double th, double rho, int lumped, double M[8 * 8]);
void svlo_quad4_stiffness(const double *X, double th, const double C[9], double K[8 * 8]);

/* Plastic3DJ2::UpdateState(eps,1) + CommitState.  par = {K,G,rho,H,beta,Sy};
 * state = {eps_p[6] (tensor shear), backstress[6], alpha}                     */
void svlo_j2_update(const double par[6], const double eps_eng[6], double state[13], double sig[6]);

/* PML element matrices, row-major (72x72 / 20x20).  par3 = {n,L,R,x0,y0,z0,nx,ny,nz};
 * par2 = {th,n,L,R,x0,y0,nx,ny}.  Any output pointer may be NULL.            */
void svlo_pml3d_matrices(const double *X, double E, double nu, double rho, const double par3[9],
                         double *M, double *C, double *K, double *G);
void svlo_pml2d_matrices(const double *X, double E, double nu, double rho, const double par2[8],
                         double *M, double *C, double *K);

/* DRM element force (lin3DHexa8.cpp:660-718 / lin2DQuad4.cpp:564-614), lumped
 * or consistent mass, no damping.  ext[i] = node i is exterior.  Uo/Vo/Ao are
 * the (already sign-flipped for exterior nodes) rows of the node fields.      */
void svlo_hex8_drm_force(const double *X, const double C[36], double rho, int lumped,
                         const uint8_t ext[8], const double Uo[24], const double Vo[24],
                         const double Ao[24], double f[24]);
void svlo_quad4_drm_force(const double *X, double th, const double C[9], double rho, int lumped,
                          const uint8_t ext[4], const double Uo[8], const double Vo[8],
                          const double Ao[8], double f[8]);

/* ---- whole-analysis level ------------------------------------------------ */
typedef struct svlo_model {
    int32_t ndim, lumped;
    int32_t n_nodes, n_total, n_free;
    const int32_t *node_ndof;      /* [n_nodes]                                  */
    const int32_t *node_ptr;       /* [n_nodes+1] into totaldof/freedof          */
    const int32_t *totaldof;       /* concat                                     */
    const int32_t *freedof;        /* concat: >=0 free, -1 fixed, < -1 constraint*/
    const double  *coords;         /* [n_nodes*ndim]                             */
    /* constraints */
    int32_t n_cons;
    const int32_t *cons_tag;       /* [n_cons] (< -1)                            */
    const int32_t *cons_slave;     /* [n_cons] slave TOTAL dof                   */
    const int32_t *cons_ptr;       /* [n_cons+1]                                 */
    const int32_t *cons_master;    /* master FREE dofs                           */
    const double  *cons_factor;
    /* nodal masses */
    int32_t n_mass;
    const int32_t *mass_node;      /* [n_mass]                                   */
    const double  *mass_val;       /* concat ndof of each listed node            */
    /* materials */
    int32_t n_mat;
    const int32_t *mat_kind;
    const double  *mat_par;        /* [n_mat*8]                                  */
    /* elements */
    int32_t n_elem;
    const int32_t *elem_kind;      /* [n_elem]                                   */
    const int32_t *elem_conn;      /* [n_elem*8] (quads use first 4)             */
    const int32_t *elem_mat;       /* [n_elem]                                   */
    const double  *elem_attr;      /* [n_elem*10]                                */
    /* Rayleigh damping per element: am (mass prop.), ak (stiffness prop.)       */
    const double  *elem_am;        /* [n_elem] or NULL                           */
    const double  *elem_ak;        /* [n_elem] or NULL                           */
    /* point loads */
    int32_t n_pload;
    const int32_t *pl_ptr;         /* [n_pload+1] into pl_nodes                  */
    const int32_t *pl_nodes;
    const double  *pl_dir;         /* [n_pload*3]                                */
    const int32_t *pl_nt;          /* [n_pload]  (1 = constant)                  */
    const int32_t *pl_sptr;        /* [n_pload+1] into pl_series                 */
    const double  *pl_series;
    const double  *pl_factor;      /* [n_pload]                                  */
    /* DRM loads (one load; field already as in .drm files, not sign-flipped)    */
    int32_t n_drm_elem, n_drm_node, drm_nt;
    const int32_t *drm_elem;
    const int32_t *drm_node;
    const uint8_t *drm_ext;        /* [n_drm_node]                               */
    const double  *drm_field;      /* [n_drm_node][drm_nt][3*ndim]               */
    double drm_factor;
    /* analysis */
    double dt, ftol, mtol;
    const double *U0, *V0, *A0;    /* [n_total] or NULL                          */
    /* support motions (Driver.hpp:509-563 UpdateSupportMotion; Node::GetSupportMotion, Node.cpp:228-247;
     * Assembler::ComputeSupportMotionIncrement, Assembler.cpp:493-533): one entry per (node, dof) that a SUPPORTMOTION
     * load of the combination lists; factor = that load's combination factor    */
    int32_t n_sup;
    const int32_t *sup_dof;        /* [n_sup] TOTAL dof                          */
    const int32_t *sup_ptr;        /* [n_sup+1] into sup_series                  */
    const double  *sup_series;     /* Xo of each entry (size 1 = CONSTANT)       */
    const double  *sup_factor;     /* [n_sup]                                    */
} svlo_model;

/* Runs DynamicAnalysis::Analyze with CentralDifference+Linear for k=1..nt-1
 * and writes, for every step, the values of `field` (0 disp,1 vel,2 accel,
 * 3 reaction = Integrator::ComputeReactionForce as DynamicAnalysis::UpdateDomain
 * stores it, DynamicAnalysis.cpp:130-150, CentralDifference.cpp:155-171) at
 * the total dofs rec_dofs[0..n_rec) into out[(nt-1)*n_rec].  Returns 0 on
 * success.  If Ufinal != NULL it receives the last U (n_total).              */
int svlo_run_central_difference(const svlo_model *m, int nt, int field, int n_rec,
                                const int32_t *rec_dofs, double *out, double *Ufinal,
                                int nthreads);

/* Same with NewmarkBeta (average acceleration) + Linear + a direct solve: 10-Integrators/03-Newmark/NewmarkBeta.cpp.
 * Linear materials only (the tangent is the elastic stiffness); returns 4 otherwise.                           */
int svlo_run_newmark(const svlo_model *m, int nt, int field, int n_rec,
                     const int32_t *rec_dofs, double *out, double *Ufinal, int nthreads);

/* ExtendedNewmarkBeta (NewmarkBeta + the PML history matrix G): 10-Integrators/03-Newmark/ExtendedNewmarkBeta.cpp.   */
int svlo_run_extended_newmark(const svlo_model *m, int nt, int field, int n_rec,
                              const int32_t *rec_dofs, double *out, double *Ufinal, int nthreads);

/* NewmarkBeta + NewtonRaphson with the materials' consistent tangents (09-Algorithms/02-Newton/NewtonRaphson.cpp);
 * tol / nmax / flag as in the JSON's algorithm block (cnvgtol / nstep / cnvgtest).                                    */
int svlo_run_newmark_newton(const svlo_model *m, double tol, int nmax, int flag, int nt, int field, int n_rec,
                            const int32_t *rec_dofs, double *out, double *Ufinal, int nthreads);
void svlo_j2_update_tangent(const double par[6], const double eps_eng[6], double state[13], double sig[6], double Ct[36]);

/* Assembler::ComputeInternalForceVector on a given displacement state (all
 * materials start from the virgin state, one UpdateState with U).            */
int svlo_internal_force(const svlo_model *m, const double *U, double *F);
/* lumped global mass diagonal (n_total)                                       */
int svlo_mass_diagonal(const svlo_model *m, double *Md);

#ifdef __cplusplus
}
#endif
#endif
