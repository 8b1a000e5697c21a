// model.h -- host-side model (what the C ABI builder calls fill in) and the device plan.
#pragma once
#include <cstdint>
#include <string>
#include <utility>
#include <vector>
#include <cuda_runtime.h>
#include "../../include/svlgpu.h"

namespace svl {

struct Material { int kind; double p[8]; };

struct PointLoad {
    std::vector<int32_t> nodes;
    double dir[3];
    std::vector<double> series;      // size 1 = constant
    double factor;
};

struct DrmLoad {
    std::vector<int32_t> elems, nodes;
    std::vector<uint8_t> ext;
    int nt = 0;
    std::vector<double> field;       // [nnodes][nt][3*ndim], empty when analytic
    bool analytic = false;
    double dir[3], pol[3], xref[3], c = 0, f0 = 0, t0 = 0, amp = 0;
    double factor = 1.0;
};

struct Recorder {
    int field = 0;
    std::vector<int32_t> nodes;
    int max_rows = 0;
    int32_t *d_dofs = nullptr;       // internal dof ids, `width` entries
    double *d_rows = nullptr;        // [max_rows][width]
    int width = 0, rows = 0;
    // REACTION recorders (DynamicAnalysis.cpp:130-150): per recorded dof the lumped mass and damping diagonal (uploaded at the
    // first use: svlgpu_comm_init replaces interface values by their global sums), whether the node is fixed (rows of free
    // nodes stay zero), and the off-diagonal damping couplings of ZeroLength1D dashpots as a CSR (other dof, coefficient)
    std::vector<int32_t> h_dofs, h_cptr, h_cdof;
    std::vector<double> h_ccoef, h_cdg;      // h_cdg: dashpot diagonal on restrained ends (not part of the Keff diagonal)
    std::vector<uint8_t> h_fixed;
    double *d_rmass = nullptr, *d_rcd = nullptr, *d_ccoef = nullptr;
    uint8_t *d_fixed = nullptr;
    int32_t *d_cptr = nullptr, *d_cdof = nullptr;
};

// support motion of one restrained dof (Driver.hpp:509-563, Node.cpp:132-134,228-247)
struct SupportMotion { int node = 0, dof = 0; std::vector<double> series; double factor = 1.0; };
struct SupportDev {
    int n = 0;
    int32_t *d_dof = nullptr;        // [n] internal dof, ascending
    int32_t *d_ptr = nullptr;        // [n+1] into d_series
    double *d_series = nullptr, *d_fac = nullptr;
};

struct BlockHint { int node0, nx, ny, nz; };

// One verified lattice block advanced by the block-stencil kernel.
struct Block {
    int node0 = 0, nx = 0, ny = 0, nz = 0;   // node lattice dims
    int ndim = 3;                    // 3 (hex8) or 2 (quad4, nz == 1)
    long long dof0 = 0;              // internal dof of lattice node 0
    int ncls = 0;                    // node classes incl. class 0 (= not handled here)
    uint8_t *d_cls = nullptr;        // [nx*ny*nz] class per node
    double *d_tbl = nullptr;         // [ncls][stride]
    int64_t n_stencil_nodes = 0;
    // 3-D: dominant classes (constant-bank coefficient kernel) + list of the remaining stencil nodes
    struct Dom {
        int cls = 0, slot = 0;
        bool ortho = false;
        int bi0 = 0, bj0 = 0, bk0 = 0, bk1 = 0;      // box of the nodes of this class
        int bi1 = 0, bj1 = 0;
        bool pure = false;                           // the class fills its box and all 27 neighbours exist -> TMA kernel
        bool sym = false;                            // stencil even / odd in the offsets (k_stencil3_v4 SYM)
        bool v4 = true;                              // barrier-free renaming kernel (SVLGPU_STENCIL_V=3 selects the older one)
        bool sep = false;                            // k_stencil3_sep: the table is a sum of tensor products of 1-D stencils
        double sepc[9] = {}, sepe[3] = {};           // fitted coefficients c[a][axis], e[xy, xz, yz] (planner)
        int rows = 4;                                // lattice rows per thread (4 or 6)
        bool nobar = false;                          // k_stencil3_v4 without the per-plane CTA barrier (SVLGPU_STENCIL_NOBAR)
        int tiles_x = 0, tiles_y = 0, zchunks = 0, kz = 0;
        double tbl[276];
        int64_t nodes = 0;
    };
    std::vector<Dom> doms;
    int32_t *d_glist = nullptr;      // remaining stencil nodes sorted by class (halo-free listing)
    int n_glist = 0;
    int32_t *d_shell_list = nullptr; // those nodes minus the interface nodes, cut into one-class chunks of kShellChunk
    uint8_t *d_shell_cls = nullptr;  // (padded with -1) + class of each chunk: the step pass of k_stencil3_shell
    int n_shell_chunks = 0;
    int32_t *d_shell_all_list = nullptr;   // all of them: the force-only pass (Assembler::ComputeInternalForceVector)
    uint8_t *d_shell_all_cls = nullptr;
    int n_shell_all = 0;
    std::vector<uint8_t> h_cls;      // host copy of the class ids (planner scratch)
};
constexpr int kShellNPT = 2;         // nodes per thread of k_stencil3_shell
constexpr int kShellChunk = 128 * kShellNPT;
constexpr int kDomNW = 4;            // warps per CTA of k_stencil3_dom
constexpr int kDomR = 4;             // lattice rows per thread

// per-class coefficient table strides (doubles)
constexpr int kTbl3Stride = 276;     // 27 x (9 coefficients + 1 pad) + kinv[3] + km[3]
constexpr int kTbl2Stride = 40;      // 9 x (2x2) + kinv[2] + km[2]

// kernel timers: 0 dominant stencil, 1 Gauss-point elements, 2 generic node gather, 3 point loads, 4 shell classes,
// 5 DRM, 6 PML element products, 7 PML gathers / right-hand side, 8 PML Krylov vector updates + scatter
constexpr int kNumTimers = 9;
struct KernelTimer {
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    double total_ms = 0.0;
    int64_t launches = 0;
    bool pending = false;
};

struct GenericSet {                  // Gauss-point path: elements of one kind
    int kind = 0, n = 0, npe = 8, ndofn = 3, ngp = 8;
    int32_t *d_conn = nullptr;       // [n][npe] node ids
    int32_t *d_mat = nullptr;        // [n] material index
    double *d_th = nullptr;          // [n] thickness (quad)
    double *d_fe = nullptr;          // [n][npe][ndofn] (inside the fe arena)
    double *d_state = nullptr;       // [13][n*ngp] J2 state (only when the set has J2 elements)
    double *d_gp = nullptr;          // [2*ncomp][n*ngp] strain | stress at Gauss points
    int32_t *d_ecls = nullptr;       // [n] gradient-table class (when few classes cover the set)
    double *d_gtab = nullptr;        // [ncls][ngp][npe*ndim + 1] shape-function gradients + w|J|
    std::vector<int32_t> elems;      // global element ids
    bool has_j2 = false;
};

// Non-lattice nodes whose incident elements are all linear and whose assembled row of K repeats across the mesh
// ("neighbour-list node classes", k_nbr_nodes): pre-summed row blocks per class + an explicit neighbour list per node.
constexpr int kNbrSlots = 32;        // neighbour nodes per row, the node itself included (27 in a regular hex8 mesh)
constexpr int kNbrNPT = 2;           // nodes per thread
constexpr int kNbrChunk = 128 * kNbrNPT;
struct NbrDev {
    int n_nodes = 0, n_chunks = 0, n_cls = 0, stride = 0;   // stride: doubles per class table
    double *d_tbl = nullptr;         // [n_cls][kNbrSlots * nd * nd + 2 * nd]: blocks [slot][b][a], then 1/Keff, Kminus
    int32_t *d_cls_nn = nullptr;     // [n_cls] slots in use
    int32_t *d_chunk_cls = nullptr;  // [n_chunks]
    int64_t *d_chunk_off = nullptr;  // [n_chunks] start of the chunk's neighbour lists in d_nbr
    int32_t *d_dof0 = nullptr;       // [n_chunks][kNbrChunk] internal dof0 of the node, -1 = padding
    int32_t *d_nbr = nullptr;        // per chunk [nn][kNbrChunk]: internal dof0 of the neighbour in each slot (slot-major: coalesced)
};

struct DrmDev {
    int n_nodes = 0, n_all = 0, nt = 0, nf = 0;   // rows with entries / all DRM nodes
    int32_t *d_node_dof0 = nullptr;  // internal dof0 of each row node
    uint8_t *d_ext = nullptr;
    double *d_field = nullptr;       // [nnodes][nt][nf]
    int32_t *d_row_ptr = nullptr;    // CSR over DRM nodes
    int32_t *d_col_blk = nullptr;    // per entry: (local DRM node index of the column node, index into the dictionary of
                                     // unique K blocks) as int2
    double *d_blk = nullptr;         // [nblk][ndim][4 | 2] unique K blocks (row node <- col node), rows padded for wide loads
    double *d_uo[2] = {nullptr, nullptr};   // [n_all][4 | 2] incident displacement of step k in buffer k & 1 (padded)
    double *d_F[2] = {nullptr, nullptr};    // [n_nodes][ndim] row forces of step k in buffer k & 1
    int buf_k[2] = {-1, -1};
    cudaEvent_t ev_ready[2] = {nullptr, nullptr};
    bool ev_valid[2] = {false, false};      // ev_ready[b] was recorded outside a capture and not yet superseded
    double *d_xyz = nullptr;         // node coordinates (analytic mode)
    bool analytic = false;
    double dir[3], pol[3], xref[3], c = 0, f0 = 0, t0 = 0, amp = 0, factor = 1;
    int32_t *d_target = nullptr;     // per row: first slot in halo.d_hF (interface node) or -1
    double *d_rkinv = nullptr;       // [n_nodes][ndim] 1 / Keff of the row dofs, compact (k_drm_apply reads it coalesced)
    // analytic plane wave: blocks pre-contracted with the polarisation, one scalar wave value per DRM node (k_drm_pw)
    double *d_wdict = nullptr;       // [nblk][4 | 2]  B . pol
    double *d_sc = nullptr;          // [n_all] (x - xref) . dir / c
    double *d_sval[2] = {nullptr, nullptr};   // [n_all] +-amp * ricker of step k in buffer k & 1
    bool inline_apply = false;       // k_drm_pw_apply: force + application in one kernel on the main stream (no interface rows)
    bool fused = false;              // k_drm_pw_fused: wave value, row force and application in one launch (default for plane waves)
};

// Multi-GPU: interface nodes shared with other ranks.  Every rank computes the partial
// (internal - external) force of its own elements at the interface nodes, exchanges the partials with
// the ranks sharing each node, and every replica applies the same rank-ordered sum.
struct HaloPeer {
    int peer = 0;
    std::vector<int32_t> nodes;        // shared soil nodes (ndim dofs each), ascending global id on both sides
    int offset = 0;
    // shared nodes of the PML block (9- / 5-dof PML nodes and the soil nodes tied to them) are exchanged per unknown
    // inside the block solve: unknown indices in the order of the peer's mirror list, first entry in the send buffer
    std::vector<int32_t> unk;
    int unk_offset = 0;
};
struct HaloDev {
    bool active = false;
    int n_if = 0, nd = 3;              // interface nodes, dofs per node exchanged (= ndim)
    int n_entries = 0;                 // sum of the peers' node counts
    int32_t *d_if_dof0 = nullptr;      // [n_if] internal dof0
    double *d_hF = nullptr;            // [n_if][nd] own partial force
    int32_t *d_send_map = nullptr;     // [n_entries] interface index of each send entry
    double *d_send = nullptr, *d_recv = nullptr;   // [n_entries][nd]
    int32_t *d_fix_ptr = nullptr, *d_fix_src = nullptr;   // per interface node: sources in rank order (-1 = own)
    // generic (Gauss-point path) interface nodes: subset CSR into the element-force arena
    int n_gen = 0;
    int32_t *d_g_ndof = nullptr, *d_g_ptr = nullptr, *d_g_target = nullptr;
    int64_t *d_g_slot = nullptr;
    // lattice interface nodes, per block: lattice-local ids + interface index
    struct Lat {
        int block = 0, n = 0;
        int32_t *d_list = nullptr, *d_target = nullptr;
        // 3-D: the same nodes as one-class chunks for k_stencil3_shell (interface index per entry, -1 = padding)
        int n_chunks = 0;
        int32_t *d_chunk_list = nullptr, *d_chunk_target = nullptr;
        uint8_t *d_chunk_cls = nullptr;
    };
    std::vector<Lat> lats;
    cudaStream_t comm_stream = nullptr;
    cudaEvent_t e_ready = nullptr, e_done = nullptr;
    void *comm = nullptr;              // ncclComm_t
    int rank = 0, nranks = 1;
};

// PML block: the dofs of the PML nodes plus the soil dofs they are tied to form the part of
// Keff = M/dt^2 + C/2dt that is not diagonal; it is solved every step by a matrix-free,
// Jacobi-scaled BiCGStab (pml.cu).  All element matrices live in per-class tables.
constexpr int kPmlChunk3 = 64;       // elements of one class per CTA of k_pml_elem_rg (= 8 NEG in pml.cu): 3-D, 4 CTAs/SM
constexpr int kPmlChunk2 = 128;      // 2-D
struct PmlDev {
    bool present = false;
    int nde = 72;                    // dofs per element (72 | 20)
    int n_elem = 0, n_cls = 0, nc = 0;
    double *d_A = nullptr;           // [n_cls][nde*nde] Keff_e, stored transposed ([col][row])
    double *d_K = nullptr;           // [n_cls][...] K_e
    double *d_Km = nullptr;          // [n_cls][...] M/dt^2 - C/2dt
    // pattern-sparse twins [n_cls][npe][sp_q][ndofn][npe] + one-class element groups (k_pml_elem_rg); sp_q == 0: dense kernel
    double *d_sA = nullptr, *d_sK = nullptr, *d_sKm = nullptr;
    int sp_q = 0, n_chunks = 0;
    int8_t sp_pat[9][4] = {};
    int32_t *d_chunk_cls = nullptr, *d_chunk_elem = nullptr;
    int32_t *d_ecls = nullptr;       // [n_elem]
    int32_t *d_edof = nullptr;       // [n_elem][nde] internal dof of every element dof
    int32_t *d_ecd = nullptr;        // [n_elem][nde] unknown index (slaves -> their master's) or -1 (restrained)
    double *d_ye = nullptr;          // [n_elem][nde] element products
    int32_t *d_c_dof = nullptr;      // [nc] internal dof that carries the unknown
    int32_t *d_c_ptr = nullptr, *d_c_slot = nullptr;   // gather lists into d_ye, ascending element order
    int32_t *d_c_hf = nullptr;       // [nc] slot in halo.d_hF for soil (master) dofs, else -1
    double *d_diag = nullptr;        // [nc] soil diagonal of Keff (0 for pure PML dofs)
    double *d_kms = nullptr;         // [nc] soil M/dt^2 - C/2dt
    double *d_w = nullptr;           // [nc] row scaling: sign(diag) / sqrt|diag(Keff)|
    double *d_sc = nullptr;          // [nc] column scaling 1 / sqrt|diag(Keff)|  (unknown y = x / sc)
    double *d_x = nullptr, *d_b = nullptr, *d_bext = nullptr, *d_r = nullptr, *d_rh = nullptr, *d_p = nullptr,
           *d_v = nullptr, *d_s = nullptr, *d_t = nullptr;
    double *d_xp = nullptr;          // the increment of the step before the last one (starting guess: linear extrapolation)
    bool extrapolate = true;
    int n_sc = 0;                    // dofs written back (unknown carriers + slaves)
    int32_t *d_sc_dof = nullptr, *d_sc_c = nullptr;
    double *d_part = nullptr;        // reduction partials
    double *h_scal = nullptr;        // pinned: residual / rhs norms read back by the host
    double rtol = 1e-14, ftol = 1e-12;   // ftol: Assembler.cpp:262 (Driver.hpp:1806 default)
    int max_iter = 2000, last_iters = 8;
    int64_t total_iters = 0, solves = 0;
    // multi-GPU: unknowns on nodes shared with other ranks are replicated; every operator application and
    // the right-hand side exchange their partial sums (rank-ordered, so all replicas hold the same bits), the dot products
    // count every unknown once (d_own) and are all-reduced
    bool collective = false;         // option "pml_collective": this rank joins the reductions even without PML unknowns
    bool multi = false;              // communicator joined (halo_comm_init)
    int n_xe = 0, n_xu = 0;          // send entries (sum over the peers) / distinct shared unknowns
    int32_t *d_x_send = nullptr;     // [n_xe] unknown of each send entry
    double *d_x_sbuf = nullptr, *d_x_rbuf = nullptr;     // [n_xe]
    int32_t *d_x_unk = nullptr;      // [n_xu] shared unknowns, ascending
    int32_t *d_x_ptr = nullptr, *d_x_src = nullptr;      // per shared unknown: sources in rank order (-1 = own, else rbuf index)
    double *d_own = nullptr;         // [nc] 1 where this rank is the lowest one that holds the unknown, else 0
    double *d_raw = nullptr;         // [nc] operator / right-hand-side values before the exchange and the row scaling
    std::vector<double> h_diag;      // this rank's part of diag(Keff) per unknown (summed over the ranks at comm_init)
};

// NewmarkBeta + Linear on the device (newmark.cu)
struct NewmarkDev {
    bool present = false;
    double *d_V = nullptr, *d_A = nullptr;           // velocity / acceleration state (internal dof order)
    double *d_mass = nullptr, *d_cd = nullptr;       // lumped mass / damping diagonal
    double *d_mask = nullptr;                        // 1 free, 0 restrained
    double *d_dd = nullptr, *d_dinv = nullptr;       // D = 4/dt^2 M + 2/dt C and mask / D
    double *d_b = nullptr, *d_x = nullptr, *d_r = nullptr, *d_p = nullptr, *d_q = nullptr;
    double *d_part = nullptr, *h_scal = nullptr;
    double rtol = 1e-13;
    double ak = 0.0;                                 // uniform stiffness-proportional Rayleigh coefficient (C = am M + ak K)
    int max_iter = 5000, last_iters = 8;
    int64_t total_iters = 0, solves = 0;
    // several ranks: replicated interface dofs count once in the dot products; the communicator is joined
    double *d_own = nullptr;
    bool multi = false;
};

}  // namespace svl

struct svlgpu_model {
    // ---- builder state (host) ----
    int ndim = 3, lumped = 1;
    int n_nodes = 0, n_total = 0, n_free = 0;
    std::vector<int32_t> node_ndof, node_ptr, totaldof, freedof;
    std::vector<double> coords;
    std::vector<std::pair<int32_t, std::vector<double>>> masses;
    struct Constraint { int tag, slave; std::vector<int32_t> master; std::vector<double> factor; };
    std::vector<Constraint> constraints;
    std::vector<svl::Material> materials;
    std::vector<int32_t> elem_kind, elem_conn /*8 per elem*/, elem_mat;
    std::vector<double> elem_attr /*10 per elem, allocated only once an element carries attributes*/, elem_am, elem_ak;
    double attr(long long e, int a) const { return elem_attr.empty() ? 0.0 : elem_attr[10 * e + a]; }
    std::vector<svl::PointLoad> ploads;
    std::vector<svl::DrmLoad> drms;
    std::vector<svl::Recorder> recorders;
    std::vector<svl::SupportMotion> supports;
    svl::SupportDev sup;
    bool has_reaction_rec = false;
    bool opt_reaction_collective = false;           // several ranks: join the reaction pass's interface exchange on every step
    std::vector<svl::BlockHint> hints;
    std::vector<double> U0, V0, A0;
    bool opt_lattice_guess = true, opt_keep_gauss = false;
    int opt_graph = -1;                             // -1: environment decides
    int opt_integrator = 0;                         // 0 CentralDifference, 1 NewmarkBeta (+ Linear)
    svl::NewmarkDev nm;

    // ---- plan / device state ----
    bool finalized = false;
    int device = 0;
    double dt = 0.0;
    cudaStream_t stream = nullptr;
    cudaStream_t side[2] = {nullptr, nullptr};       // DRM forces one step ahead / shell classes beside the bulk kernels
    cudaEvent_t ev_fork = nullptr, ev_fork2 = nullptr, ev_join = nullptr;
    bool overlap = true;
    // CUDA-graph replay of steps
    int32_t *d_kctl = nullptr;                       // device step control {k, recorder row}
    int dev_k = -1, k_of_step = 0;
    bool use_graph = true, graph_capturing = false;
    void *graph_exec = nullptr;                      // cudaGraphExec_t of kGraphSteps consecutive steps
    int graph_k0 = 0, graph_cur = 0;
    int64_t graph_launches = 0;
    int64_t device_bytes = 0;
    std::vector<void *> allocs;

    int n_int = 0;                                  // internal dofs (= sum of node ndof)
    std::vector<int32_t> int_of_total;              // total dof -> internal dof (= node_ptr[node] + c)
    int32_t *d_int_of_total = nullptr, *d_node_ptr = nullptr;
    double *d_U[3] = {nullptr, nullptr, nullptr};   // rotating buffers
    int cur = 0, prev = 1, next = 2;
    double *d_kinv = nullptr, *d_km = nullptr;      // per internal dof: 1/Keff, Kminus (0 if not free)
    int *d_flag = nullptr;                          // non-finite state flag (state_is_finite)
    int steps_since_check = 0;
    double *d_fscratch = nullptr;                   // force-only passes (svlgpu_internal_force, reactions): never a state buffer
    std::vector<double> h_mass, h_cdiag;            // lumped mass / damping diagonal per internal dof

    std::vector<svl::Block> blocks;
    std::vector<svl::GenericSet> gsets;
    double *d_coords = nullptr;                     // [n_nodes][ndim]
    double *d_matpar = nullptr;                     // [n_mat][8]
    int32_t *d_matkind = nullptr;
    int n_gnodes = 0;
    int32_t *d_gn_dof0 = nullptr, *d_gn_ndof = nullptr, *d_gn_ptr = nullptr;
    int64_t *d_gn_slot = nullptr;                   // offsets (in doubles) into the fe arena
    double *d_fe_arena = nullptr;

    // nodal loads: CSR over loaded dofs
    int n_ploads = 0, n_pl_dofs = 0;
    int32_t *d_pl_dof = nullptr, *d_pl_ptr = nullptr, *d_pl_load = nullptr;
    double *d_pl_coef = nullptr, *d_pl_series = nullptr;
    int32_t *d_pl_soff = nullptr, *d_pl_nt = nullptr;
    double *d_pl_amp = nullptr;                     // per-load amplitude of the current step (host-fed)
    double *h_pl_amp = nullptr, *h_row = nullptr;   // pinned staging
    int h_row_len = 0;
    bool shell_lowreg = false;                      // SVLGPU_SHELL_LOWREG: 96-register shell kernel (co-resides with the stencil)
    int mirror_rec = -1;                            // recorder whose next row k_record also writes to h_row (step_host)
    bool host_step = false;                         // inside svlgpu_step_host: one step per call, host round trip on the critical path
    const double *step_amp = nullptr;               // host-fed load amplitudes of the step being recorded (reaction pass)

    std::vector<svl::DrmDev> drm_dev;

    // multi-GPU halo
    std::vector<svl::HaloPeer> halo_peers;
    svl::HaloDev halo;
    std::vector<int32_t> if_of_node;                // node -> interface index or -1 (host, plan time)
    int32_t *d_pl_target = nullptr;                 // per loaded dof: slot in halo.d_hF, -2-c for PML unknown c, or -1
    svl::PmlDev pml;
    svl::NbrDev nbr;
    bool opt_nbr_classes = true, opt_renumber = true;

    // counters / timing
    int64_t total_launches = 0, launches_per_step = 0;
    int64_t n_block_nodes = 0, n_generic_elements = 0, n_elem_classes = 0, n_node_classes = 0;
    bool kernel_timing = false;
    svl::KernelTimer timers[svl::kNumTimers];
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    double last_step_ms = 0.0;
    int steps_done = 0;
};

namespace svl {
void set_error(const std::string &s);
int plan_and_upload(svlgpu_model *m);                                       // planner.cu
int run_steps(svlgpu_model *m, int k0, int k1, const double *dev_amp);      // kernels.cu
int compute_internal_force(svlgpu_model *m, double *F_host);
int gather_state(svlgpu_model *m, int field, const int32_t *dofs, int n, double *out);
int state_is_finite(svlgpu_model *m);
void timer_flush(svlgpu_model *m);
void timer_begin(svlgpu_model *m, int which);
void timer_end(svlgpu_model *m, int which);
int configure_kernels();
size_t stencil3_smem(int nw, int r);
bool stencil_entry_nonzero(int di, int b, int dj, int s, int a);
int stencil_entry_sym_index(int di, int b, int dj, int s, int a);
bool stencil_entry_sym_negated(int di, int b, int dj, int s, int a);
void forget_const_owner(svlgpu_model *m);
void graph_destroy(svlgpu_model *m);
// halo.cu
int halo_plan(svlgpu_model *m);                        // after the node lists are known (planner)
int halo_comm_init(svlgpu_model *m, const void *id128, int rank, int nranks);
int halo_unique_id(void *out128);
int halo_async_begin(svlgpu_model *m);                 // interface pass on the comm stream, beside the bulk kernels
int halo_async_exchange(svlgpu_model *m);
int halo_exchange_begin(svlgpu_model *m);              // main stream: hF complete -> comm stream: pack + NCCL
int halo_exchange_end(svlgpu_model *m, const double *U, const double *Up, double *Un, int mode);
int halo_lattice_force(svlgpu_model *m, const double *U);   // partial forces of lattice interface nodes -> hF
int halo_generic_force(svlgpu_model *m);                    // ... of generic interface nodes -> hF
void halo_destroy(svlgpu_model *m);
int pmlx_setup(svlgpu_model *m);                                                     // lists + global diag(Keff); from halo_comm_init
int pmlx_exchange(svlgpu_model *m, cudaStream_t st);                                 // raw[shared] <- rank-ordered sum over the holders
int pmlx_allreduce(svlgpu_model *m, double *a, size_t na, double *b, size_t nb, cudaStream_t st);   // in place, sum
// pml.cu
int pml_rescale(svlgpu_model *m);                                                    // w, sc from the summed diagonal in d_raw
int pml_step(svlgpu_model *m, const double *U, const double *Up, double *Un);
int pml_internal_force(svlgpu_model *m, const double *U, double *F);
void pml_destroy(svlgpu_model *m);
int pml_configure();
// newmark.cu / kernels.cu
int newmark_plan(svlgpu_model *m);
int newmark_step(svlgpu_model *m, int k, const double *dev_amp);
int newmark_set_initial(svlgpu_model *m, const double *V_int, const double *A_int);
int newmark_comm_setup(svlgpu_model *m, const double *mglob, const double *cglob);     // from halo_comm_init
int halo_vec_load(svlgpu_model *m, const double *src);                                  // hF <- src at the interface dofs (main stream)
int halo_vec_sum(svlgpu_model *m, double *dst);                                         // exchange hF; dst <- rank-ordered sum at the interface dofs
int external_forces_interface(svlgpu_model *m, int k, const double *dev_amp);           // hF -= Fext(k) at the interface dofs
void newmark_destroy(svlgpu_model *m);
int operator_K(svlgpu_model *m, const double *x, double *out);
int external_forces_raw(svlgpu_model *m, int k, const double *dev_amp, double *b);
int record_rows(svlgpu_model *m, bool devk);
}  // namespace svl
