/*
 * svlgpu.h -- C ABI of the B200-native explicit-dynamics engine (libsvlgpu.so).
 *
 * This is the drop-in boundary for ONE hot path of SeismoVLAB/SVL: the per-step
 * internal-force evaluation + lumped-mass CentralDifference update (+ point /
 * DRM effective forces, + the PML block solve).  Every entry point names the
 * reference interface it replaces (paths relative to the reference's
 * 02-Run_Process/).  Plain pointers and sizes only; no C++ or torch types.
 *
 * Conventions (mirroring the reference):
 *   - return value: 0 = ok, non-zero = "stop" (the reference's `bool stop`,
 *     e.g. Integrator::ComputeNewStep, LinearSystem::SolveSystem); the text of
 *     the last failure is returned by svlgpu_last_error().
 *   - node / element / material arguments are 0-based indices in the order
 *     they were added (the facade maps the reference's tags to indices);
 *     "total dof" and "free dof" ids are the reference's own numbering
 *     (Node::GetTotalDegreeOfFreedom / GetFreeDegreeOfFreedom).
 *   - everything is FP64; ids are 32-bit (Element.hpp:248, Node.cpp:82-92).
 *   - host arrays are copied during the call, never retained.
 *   - one model per handle, calls on one handle are not thread-safe
 *     (the reference keeps process-wide globals, Definitions.cpp:4-44).
 *   - there is NO CPU fallback: finalize fails if no CUDA device is usable.
 */
#ifndef SVLGPU_H
#define SVLGPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct svlgpu_model svlgpu_model;

/* ---- enumerations ------------------------------------------------------ */
enum svlgpu_elem_kind {          /* Driver.hpp:1185-1305 element names        */
    SVLGPU_LIN3DHEXA8 = 1,       /* 04-Elements/10-Hexahedron/lin3DHexa8.cpp  */
    SVLGPU_LIN2DQUAD4 = 2,       /* 04-Elements/06-Quadrilateral/lin2DQuad4   */
    SVLGPU_PML3DHEXA8 = 3,       /* 04-Elements/10-Hexahedron/PML3DHexa8.cpp  */
    SVLGPU_PML2DQUAD4 = 4,       /* 04-Elements/06-Quadrilateral/PML2DQuad4   */
    SVLGPU_ZEROLENGTH1D = 5      /* 04-Elements/01-Zero/ZeroLength1D.cpp (Lysmer dashpots, Builder.py:1086-1131) */
};
enum svlgpu_mat_kind {           /* Driver.hpp:567-757 material names         */
    SVLGPU_ELASTIC3DLINEAR      = 1,  /* params: E, nu, rho                   */
    SVLGPU_ELASTIC2DPLANESTRAIN = 2,  /* params: E, nu, rho                   */
    SVLGPU_PLASTIC3DJ2          = 3,  /* params: K, G, rho, H, beta, SigmaY   */
    SVLGPU_PLASTICPLANESTRAINJ2 = 4,  /* params: K, G, rho, H, beta, SigmaY   */
    SVLGPU_VISCOUS1DLINEAR      = 5   /* params: eta (Driver.hpp:602-607)      */
};
enum svlgpu_field {              /* Recorder.cpp:239-269 "resp" values        */
    SVLGPU_DISP = 0, SVLGPU_VEL = 1, SVLGPU_ACCEL = 2, SVLGPU_REACTION = 3
};
enum svlgpu_gauss_field { SVLGPU_STRAIN = 0, SVLGPU_STRESS = 1, SVLGPU_STATE = 2 };

/* ---- model construction (replaces Driver.hpp UpdateMesh:1981-2046) ------ */

/* Global{ndim,massform}: Driver.hpp:1993-2008.  lumped!=0 <=> massform LUMPED. */
svlgpu_model *svlgpu_create(int ndim, int lumped);
void          svlgpu_destroy(svlgpu_model *m);
const char   *svlgpu_last_error(void);

/* Node table: Node.cpp:3-21,80-117; Driver.hpp:304-392.
 * ndof[i] dofs per node; totaldof/freedof are the concatenated per-node lists
 * (freedof: >=0 free id, -1 restrained, < -1 constraint tag).  coords: n*ndim. */
int svlgpu_set_nodes(svlgpu_model *m, int n, const int32_t *ndof, const double *coords,
                     const int32_t *totaldof, const int32_t *freedof,
                     int ntotal, int nfree);

/* Nodal masses: Driver.hpp:396-436, Assembler.cpp:622-657.  mass: concat ndof. */
int svlgpu_add_nodal_mass(svlgpu_model *m, int n, const int32_t *node, const double *mass);

/* Constraint: Constraint.cpp, Mesh.cpp:360-375.  tag < -1 as in freedof.       */
int svlgpu_add_constraint(svlgpu_model *m, int tag, int slave_total_dof, int nmaster,
                          const int32_t *master_free_dof, const double *factor);

/* Material prototype: Material.hpp; returns the material index (>=0) or -1.    */
int svlgpu_add_material(svlgpu_model *m, int kind, const double *params, int nparams);

/* Elements of one kind: Element.hpp:51-249.  conn: n*(8|4) node indices in the
 * reference's local order (lin3DHexa8.cpp:791-798).  attrs: n*nattr doubles:
 *   LIN3DHEXA8: nattr=0;  LIN2DQUAD4: [th];
 *   PML3DHEXA8: [n, L, R, x0(3), npml(3)]      (Driver.hpp:1288-1305)
 *   PML2DQUAD4: [th, n, L, R, x0(2), npml(2)]  (Driver.hpp:1203-1219)
 *   ZEROLENGTH1D: 2 nodes per element, [dir]   (Driver.hpp:1072-1078); with
 *     VISCOUS1DLINEAR it adds eta to the damping diagonal of its free dof
 *     (ZeroLength1D.cpp:212-231); both ends free would couple two dofs in
 *     Keff = M/dt^2 + C/2dt and is refused.
 * Elements are numbered in call order (ascending = the reference's std::map
 * iteration order, Assembler.cpp:251).  Returns first element index or -1.      */
int svlgpu_add_elements(svlgpu_model *m, int kind, int n, const int32_t *conn,
                        const int32_t *material, const double *attrs, int nattr);

/* Damping: Damping.cpp, lin3DHexa8.cpp:354-366.  Only FREE and the mass-
 * proportional part of RAYLEIGH keep the CentralDifference Keff diagonal: there
 * ak != 0 is refused at finalize; the Newmark integrator takes a uniform ak.     */
int svlgpu_set_rayleigh(svlgpu_model *m, int n, const int32_t *elems, double am, double ak);

/* Optional hint: nodes [node0, node0+nx*ny*nz) form a lattice numbered
 * x-fastest (Builder.py:134-141).  The planner verifies it and only then uses
 * the block-stencil kernel there; a wrong hint is ignored, never trusted.      */
int svlgpu_hint_structured_block(svlgpu_model *m, int node0, int nx, int ny, int nz);

/* Planner / solver options, before finalize.  name: "lattice_guess" (1: without hints, guess the
 * makeDomainVolume / makeDomainArea lattice from the connectivity -- verified like a hint; default 1),
 * "cuda_graph" (1: replay steps from a CUDA graph of 6 consecutive steps; default 0),
 * "pml_rtol" (relative residual of the PML block solve, default 1e-14), "keep_gauss" (1: keep Gauss-point
 * strain / stress for svlgpu_get_gauss, default 0), "ftol" (Assembler.cpp:262 filter of the PML element
 * forces, default 1e-12), "integrator" (0: CentralDifference, 10-Integrators/02-CentralDifference, the
 * default; 1: NewmarkBeta + Linear, 10-Integrators/03-Newmark/NewmarkBeta.cpp:64-133 with Linear.cpp:22-56:
 * linear materials, lumped mass, Rayleigh damping with both coefficients, dashpots (several ranks: interface sums inside
 * the K operator and all-reduced dot products are written but have not run on hardware yet); the sparse
 * factorisation of Keff = K + 4/dt^2 M + 2/dt C is replaced by matrix-free conjugate gradients),
 * "newmark_rtol" (relative residual of that solve, default 1e-13), "nbr_classes" (1: nodes outside the lattice blocks
 * whose incident elements are all linear and whose assembled row of K occurs at >= 64 nodes are advanced from pre-summed
 * row blocks + an explicit neighbour list instead of the Gauss-point kernels + element-force arena; verified numerically,
 * never trusted; default 1), "renumber" (1: a model without a lattice block -- no hint, and the lattice guess does not apply -- is
 * renumbered inside the library along a Morton curve through the node coordinates, so that neighbours in space are
 * neighbours in HBM; invisible to the caller, who only ever names total dofs; default 1), "reaction_collective" (1 on EVERY rank of a
 * partitioned model that has a REACTION recorder on any rank: the reaction pass sums the partial forces of interface nodes
 * over the ranks, so every rank must issue that exchange on every step, also a rank that records nothing; default 0), "pml_collective" (1 on EVERY rank of a
 * partitioned model that has PML elements anywhere, also on ranks without one: the PML block solve exchanges the
 * unknowns on shared nodes and all-reduces its dot products, so all ranks must issue the same collectives;
 * svlgpu_add_halo lists may then contain 9- / 5-dof PML nodes; default 0).          */
int svlgpu_set_option(svlgpu_model *m, const char *name, double value);

/* ---- loads (replaces Assembler::ComputeExternalForceVector:290-489) ------ */

/* POINTLOAD CONCENTRATED {CONSTANT|TIMESERIES}: Assembler.cpp:316-350,
 * Load.cpp:67-76.  nt==1 => constant.  factor = LoadCombo factor.             */
int svlgpu_add_point_load(svlgpu_model *m, int nnodes, const int32_t *nodes, int ndir,
                          const double *dir, int nt, const double *series, double factor);

/* ELEMENTLOAD GENERALWAVE (DRM): Assembler.cpp:460-478, lin3DHexa8.cpp:660-718,
 * Driver.hpp:1669-1721.  field: [nnodes][nt][3*ndim] = u,v,a rows exactly as in
 * the .drm files (NOT yet sign-flipped); exterior[i] = the file's cond flag.   */
int svlgpu_add_drm_load(svlgpu_model *m, int nelems, const int32_t *elems, int nnodes,
                        const int32_t *nodes, const uint8_t *exterior, int nt,
                        const double *field, double factor);

/* Same load with the incident field evaluated on the device instead of being
 * tabulated: vertically propagating plane wave
 * u(x,t) = amp * pol * ricker_disp(t - (x-x_ref).dir/c), see DESIGN.md.        */
int svlgpu_add_drm_planewave(svlgpu_model *m, int nelems, const int32_t *elems, int nnodes,
                             const int32_t *nodes, const uint8_t *exterior,
                             const double *dir, const double *pol, const double *xref,
                             double c, double f0, double t0, double amp, double factor);

/* Support motion of one restrained dof: Supports{node: type, value | file, dof} (Driver.hpp:509-563,
 * Mesh::SetSupportMotion -> Node.cpp:132-134) listed by a SUPPORTMOTION load of the combination (Driver.hpp:1725-1735,
 * factor = its combination factor).  series = Xo (nt == 1: CONSTANT).  Each step the dof moves by
 * factor * (g(k) - g(k-1)), g(k) = Xo[k] if k < nt else Xo[0] (Assembler::ComputeSupportMotionIncrement,
 * Assembler.cpp:493-533; Node::GetSupportMotion, Node.cpp:228-247; CentralDifference.cpp:135,189-202).  Like the
 * reference, the elements see the moved support one step late (Algorithm.cpp:23: UpdateStatesIncrements gets T dU only).
 * CentralDifference only; the dof must be restrained (freedof -1) and its node must not carry a PML or ZeroLength1D
 * element (their Keff rows are not diagonal); one entry per dof.                                                   */
int svlgpu_add_support_motion(svlgpu_model *m, int node, int dof, int nt, const double *series, double factor);

/* ---- recorders (Recorder.cpp:73-105,239-269) ----------------------------- */
/* NODE recorder of `field` at `nodes`; returns recorder id.  Rows are kept on
 * the device (one per step) and fetched with svlgpu_read_recorder.
 * SVLGPU_REACTION (Recorder.cpp:258): rows of Integrator::ComputeReactionForce as DynamicAnalysis::UpdateDomain
 * stores them (DynamicAnalysis.cpp:130-150, CentralDifference.cpp:155-171, Assembler.cpp:272-287,568-619): at the dofs
 * of FIXED nodes  F_int + C V + M A (elements and point masses) - F_ext(k),  zero rows for free nodes.  The reference
 * evaluates this for the whole mesh every step; here it costs one extra force pass per RECORDED step.
 * CentralDifference, models without PML elements.                                                                   */
int svlgpu_add_node_recorder(svlgpu_model *m, int field, int nnodes, const int32_t *nodes,
                             int max_rows);

/* ---- analysis (CentralDifference.cpp, Linear.cpp, DynamicAnalysis.cpp) ---- */

/* CentralDifference::Initialize (CentralDifference.cpp:35-71): builds lumped
 * M, C, Keff = M/dt^2 + C/2dt, Up = U - dt V + dt^2/2 A, uploads everything,
 * plans the kernels.  device = CUDA ordinal.  ftol: Assembler.cpp:262 filter
 * (kept for API parity; the device path does not drop small entries, see
 * DESIGN.md).                                                                  */
int svlgpu_finalize(svlgpu_model *m, double dt, int device);

/* Initial conditions per total dof (Node::SetDisplacements etc.), before
 * finalize.  Any of the pointers may be NULL (=0).                             */
int svlgpu_set_initial_state(svlgpu_model *m, const double *U, const double *V, const double *A);

/* DynamicAnalysis::Analyze loop body for k = k_begin .. k_end-1
 * (DynamicAnalysis.cpp:36-57): ComputeNewStep + CommitState + recorder row.
 * k is the load-sample index (starts at 1 in the reference).  Asynchronous on
 * the model's stream unless sync != 0.                                         */
int svlgpu_step(svlgpu_model *m, int k_begin, int k_end, int sync);
int svlgpu_sync(svlgpu_model *m);

/* Same, but the load amplitudes of each time-series point load for step k are
 * passed from the host (nloads doubles, pinned or pageable) and the newest row
 * of recorder `rec` is copied back into row_out: the per-step host round trip
 * of the reference's Integrator::ComputeNewStep + Recorder::WriteResponse.     */
int svlgpu_step_host(svlgpu_model *m, int k, const double *amplitudes, int nloads,
                     int rec, double *row_out, int row_len);

/* Integrator::GetDisplacements/Velocities/Accelerations (Integrator.hpp):
 * gathers `n` total dofs (dofs==NULL: all ntotal in order) into out.           */
int svlgpu_get_state(svlgpu_model *m, int field, const int32_t *dofs, int n, double *out);

/* Assembler::ComputeInternalForceVector (Assembler.cpp:239-269) for the current
 * state: ntotal doubles.                                                       */
int svlgpu_internal_force(svlgpu_model *m, double *F);

/* Lumped global mass diagonal (Assembler::ComputeMassMatrix, :47-67), ntotal.  */
int svlgpu_get_mass_diagonal(svlgpu_model *m, double *Mdiag);

/* Element::GetStrain/GetStress at Gauss points (lin3DHexa8.cpp:150-200):
 * out[nelem][ngauss][ncomp].                                                   */
int svlgpu_get_gauss(svlgpu_model *m, int field, int nelem, const int32_t *elems, double *out);

/* rows [r0,r1) of a NODE recorder, r1-r0 rows of (sum ndof of listed nodes).   */
int svlgpu_read_recorder(svlgpu_model *m, int rec, int r0, int r1, double *out);
int svlgpu_recorder_rows(svlgpu_model *m, int rec);
int svlgpu_recorder_width(svlgpu_model *m, int rec);

/* ---- introspection / measurement ---------------------------------------- */
typedef struct svlgpu_counters {
    int64_t n_elements, n_nodes, n_total_dofs;
    int64_t n_block_nodes;       /* nodes advanced by the block-stencil kernel     */
    int64_t n_generic_nodes;     /* nodes advanced by the gather kernel            */
    int64_t n_generic_elements;  /* elements evaluated Gauss point by Gauss point  */
    int64_t n_elem_classes, n_node_classes;
    int64_t launches_per_step;   /* kernels launched by one svlgpu_step iteration  */
    int64_t total_launches;      /* since finalize                                 */
    int64_t device_bytes;        /* HBM allocated by the handle                    */
    double  last_step_ms;        /* CUDA-event time of the last svlgpu_step call   */
    double  stencil_ms;          /* ... of which block-stencil kernel (if timed)   */
    int64_t n_pml_elements;      /* PML elements (block solve)       */
    int64_t n_pml_unknowns;      /* dofs of the non-diagonal block of Keff         */
    int64_t pml_solves, pml_iterations;   /* Krylov solves / iterations so far (PML block BiCGStab + Newmark CG) */
    int64_t n_nbr_nodes, n_nbr_classes;   /* non-lattice nodes advanced through pre-summed rows + neighbour lists  */
} svlgpu_counters;
int svlgpu_get_counters(svlgpu_model *m, svlgpu_counters *out);

/* When on, every svlgpu_step brackets its dominant kernel with CUDA events on
 * the launching stream so that bench.py can report the roofline live.          */
int svlgpu_set_kernel_timing(svlgpu_model *m, int on);
/* average duration (ms) and launch count of kernel `which` since last reset:
 * 0 = block stencil (dominant class), 1 = Gauss-point element force, 2 = generic node update,
 * 3 = point loads, 4 = block stencil (shell gather), 5 = DRM.  reset!=0 clears after reading.              */
int svlgpu_kernel_time(svlgpu_model *m, int which, double *avg_ms, int64_t *launches, int reset);

/* Device micro-benchmarks for the roofline record: FP64 FMA throughput (TFLOP/s, 2 flop per DFMA,
 * dependent-chain-free kernel at full occupancy) and streaming-copy bandwidth (GB/s, read + write bytes of a 1 GiB
 * copy).  Either pointer may be NULL.  No reference counterpart: measurement only.                              */
int svlgpu_measure_peaks(int device, double *fp64_tflops, double *copy_gbs);

/* raw device pointers for zero-copy interop (torch.from_blob / NCCL plumbing):
 * which: 0 = U_n, 1 = U_{n-1}, 2 = scratch U_{n+1}; length ntotal doubles.     */
int svlgpu_device_ptr(svlgpu_model *m, int which, void **ptr, int64_t *len);

/* ---- multi-GPU halo (replaces MumpsSolver.cpp:56,161) ----- */
/* One process per GPU / partition file.  Interface nodes shared with rank `peer`
 * (interface nodes are duplicated in every partition that touches them,
 * 01-Pre_Process/Core/SeismoVLAB.py:354-358): both sides must list the same
 * nodes in the same order (ascending global tag).  Call before finalize, one
 * list per peer.  Each step the partial (internal - external) forces of the
 * listed nodes are exchanged with NCCL send/recv on a second stream, overlapped
 * with the bulk kernels, and every replica applies the same rank-ordered sum. */
int svlgpu_add_halo(svlgpu_model *m, int peer, int nnodes, const int32_t *nodes);
/* ncclGetUniqueId on rank 0 (128 bytes); the launcher broadcasts it.            */
int svlgpu_nccl_unique_id(void *out128);
/* after finalize, on every rank: joins the communicator and replaces the
 * partial lumped mass / damping of the interface dofs by their global sums.     */
int svlgpu_comm_init(svlgpu_model *m, const void *id128, int rank, int nranks);

#ifdef __cplusplus
}
#endif
#endif /* SVLGPU_H */
