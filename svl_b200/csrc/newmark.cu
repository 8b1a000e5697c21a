// newmark.cu -- NewmarkBeta (average acceleration) + Linear on the device.
//
// Reference: 10-Integrators/03-Newmark/NewmarkBeta.cpp (:21-36 Initialize, :64-79 ComputeNewStep, :106-121
// ComputeEffectiveForce, :124-133 ComputeEffectiveStiffness) driven by 09-Algorithms/01-Linear/Linear.cpp:22-56:
//     Keff dU = Fext(k) + Fbar - Fint(U_n) + M (4/dt V + A) + C V,   Keff = K + 4/dt^2 M + 2/dt C
//     U += dU;  A <- 4/dt^2 dU - 4/dt V - A;  V <- 2/dt dU - V
// The reference assembles Keff as a sparse matrix EVERY step and factors / back-substitutes it (EigenSolver / MUMPS /
// PETSc).  Here Keff is never formed: K p is the matrix-free internal-force pass of the explicit path applied to the
// search direction (block-stencil kernels on lattices, Gauss-point kernels elsewhere), M and C are the lumped
// diagonals, and the system is solved by conjugate gradients preconditioned with D = 4/dt^2 M + 2/dt C (so the
// iteration count depends on (omega_max dt)^2 only).  Scalars and the convergence flag live on the device; reductions
// have a fixed shape (deterministic).  Scope: linear materials, lumped mass, Rayleigh damping (mass-proportional part
// on the diagonal, a uniform stiffness-proportional part folded into the K operator), dashpots (diagonal C),
// restrained dofs.  Several GPUs (written, not yet run on hardware): interface values of K p are summed over the ranks
// with the force halo lists, dot products count every dof on its lowest rank and are all-reduced (newmark_comm_setup).
#include <algorithm>
#include <cstring>
#include "model.h"

namespace svl {

#define CUDA_OK(x)                                                                          \
    do {                                                                                    \
        cudaError_t e_ = (x);                                                               \
        if (e_ != cudaSuccess) {                                                            \
            set_error(std::string(#x) + ": " + cudaGetErrorString(e_));                     \
            return 1;                                                                       \
        }                                                                                   \
    } while (0)

constexpr int kNmBlocks = 1184;      // 8 per SM (the vector kernels are streaming / latency-bound); every reduction writes kNmBlocks partials
constexpr int kNmThreads = 256;
// partial-sum slots and scalar cells behind them
enum { N_BB = 0, N_RZ0 = 1, N_RZ1 = 2, N_PAP = 3, N_RR = 4, N_NSLOT = 5 };
constexpr int kNmFlag = N_NSLOT * kNmBlocks + 8;      // sticky "converged" flag
constexpr int kNmOut = N_NSLOT * kNmBlocks + 10;      // {rr, bb} of the last test

__device__ __forceinline__ double nm_block_sum(double v) {
    __shared__ double sh[kNmThreads / 32];
    for (int o = 16; o; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    double r = 0.0;
    if (threadIdx.x < 32) {
        r = (threadIdx.x < kNmThreads / 32) ? sh[threadIdx.x] : 0.0;
        for (int o = 16; o; o >>= 1) r += __shfl_down_sync(0xffffffffu, r, o);
    }
    __syncthreads();
    return r;      // valid in thread 0
}
__device__ __forceinline__ double nm_total(const double *part, int slot) {
    __shared__ double tot;
    double v = 0.0;
    if (threadIdx.x < 32) {
        for (int i = threadIdx.x; i < kNmBlocks; i += 32) v += part[slot * kNmBlocks + i];
        for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (threadIdx.x == 0) tot = v;
    }
    __syncthreads();
    const double r = tot;
    __syncthreads();
    return r;
}

// b = mask (-Fint + M (4/dt V + A) + C V)   (NewmarkBeta.cpp:116-118 with dU = 0); external forces are added afterwards
__global__ void __launch_bounds__(kNmThreads) k_nm_rhs(int n, double c4, const double *F, const double *mass, const double *cd,
                                                        const double *mask, const double *V, const double *A, double *b, double *part) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        b[i] = mask[i] * (mass[i] * (c4 * V[i] + A[i]) + cd[i] * V[i] - F[i]);
    if (blockIdx.x == 0 && threadIdx.x == 0) part[kNmFlag] = 0.0;
}
// start of the solve: x = 0 (Linear.cpp:25), r = mask b, z = r / D, p = z; partials bb = (r, r), rz = (r, z)
__global__ void __launch_bounds__(kNmThreads) k_nm_init(int n, const double *mask, const double *dinv, const double *b, double *x,
                                                         double *r, double *p, double *part, const double *own) {
    double a0 = 0.0, a1 = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const double rv = mask[i] * b[i], zv = rv * dinv[i];
        x[i] = 0.0; r[i] = rv; p[i] = zv;
        const double o = own ? own[i] : 1.0;              // several ranks: a replicated dof counts on its lowest rank
        a0 = fma(o * rv, rv, a0); a1 = fma(o * rv, zv, a1);
    }
    const double s0 = nm_block_sum(a0), s1 = nm_block_sum(a1);
    if (threadIdx.x == 0) {
        part[N_BB * kNmBlocks + blockIdx.x] = s0; part[N_RR * kNmBlocks + blockIdx.x] = s0;
        part[N_RZ0 * kNmBlocks + blockIdx.x] = s1;
    }
}
// w = U - ak V: the right-hand side holds -Fint + C V with C = am M + ak K (lin3DHexa8.cpp:360-366), and for linear
// materials -K U + ak K V = -K (U - ak V): one operator application serves both terms
__global__ void __launch_bounds__(kNmThreads) k_nm_w(int n, double ak, const double *U, const double *V, double *w) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) w[i] = fma(-ak, V[i], U[i]);
}
// q = mask (ck K p + D p) with K p already in q, ck = 1 + 2 ak / dt; partial (p, q)
__global__ void __launch_bounds__(kNmThreads) k_nm_ap(int n, double ck, const double *mask, const double *dd, const double *p, double *q,
                                                       double *part, const double *own) {
    if (part[kNmFlag] != 0.0) return;
    double a0 = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const double pv = p[i], qv = mask[i] * (ck * q[i] + dd[i] * pv);
        q[i] = qv;
        a0 = fma((own ? own[i] : 1.0) * pv, qv, a0);
    }
    const double s0 = nm_block_sum(a0);
    if (threadIdx.x == 0) part[N_PAP * kNmBlocks + blockIdx.x] = s0;
}
// x += alpha p, r -= alpha q; partials rr = (r, r), rz_next = (r, r / D)
__global__ void __launch_bounds__(kNmThreads) k_nm_xr(int n, int rz_slot, const double *dinv, const double *p, const double *q,
                                                       double *x, double *r, double *part, const double *own) {
    if (part[kNmFlag] != 0.0) return;
    const double pap = nm_total(part, N_PAP), rz = nm_total(part, rz_slot);
    const double alpha = pap != 0.0 ? rz / pap : 0.0;
    const int nslot = (rz_slot == N_RZ0) ? N_RZ1 : N_RZ0;
    double a0 = 0.0, a1 = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        x[i] = fma(alpha, p[i], x[i]);
        const double rv = fma(-alpha, q[i], r[i]);
        r[i] = rv;
        const double o = own ? own[i] : 1.0;
        a0 = fma(o * rv, rv, a0); a1 = fma(o * (rv * dinv[i]), rv, a1);
    }
    const double s0 = nm_block_sum(a0), s1 = nm_block_sum(a1);
    if (threadIdx.x == 0) { part[N_RR * kNmBlocks + blockIdx.x] = s0; part[nslot * kNmBlocks + blockIdx.x] = s1; }
}
// convergence test (one warp) + p = z + beta p
__global__ void k_nm_flag(double *part, double tol2) {
    if (part[kNmFlag] != 0.0) return;
    const double rr = nm_total(part, N_RR), bb = nm_total(part, N_BB);
    if (threadIdx.x == 0) {
        part[kNmOut] = rr; part[kNmOut + 1] = bb;
        if (rr <= tol2 * bb + 1e-300 || !(rr == rr)) part[kNmFlag] = 1.0;
    }
}
__global__ void __launch_bounds__(kNmThreads) k_nm_p(int n, int rz_old, const double *dinv, const double *r, double *p,
                                                      const double *part) {
    if (part[kNmFlag] != 0.0) return;
    const int rz_new = (rz_old == N_RZ0) ? N_RZ1 : N_RZ0;
    const double o = nm_total(part, rz_old), nw = nm_total(part, rz_new);
    const double beta = o != 0.0 ? nw / o : 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) p[i] = fma(beta, p[i], r[i] * dinv[i]);
}
// NewmarkBeta.cpp:73-76
__global__ void __launch_bounds__(kNmThreads) k_nm_update(int n, double dt, const double *x, const double *U, double *Un, double *V,
                                                           double *A) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const double du = x[i], v = V[i], a = A[i];
        Un[i] = U[i] + du;
        A[i] = 4.0 / dt / dt * du - 4.0 / dt * v - a;
        V[i] = 2.0 / dt * du - v;
    }
}

// several ranks: lumped mass / damping summed over the ranks at the interface dofs -> D and 1 / D
__global__ void k_nm_refresh(int n, double dt, const double *mglob, const double *cglob, const double *mask, double *mass, double *cd,
                             double *dd, double *dinv) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double d = 4.0 / dt / dt * mglob[i] + 2.0 / dt * cglob[i];
    mass[i] = mglob[i]; cd[i] = cglob[i]; dd[i] = d;
    dinv[i] = (mask[i] != 0.0 && d > 0.0) ? 1.0 / d : 0.0;
}

int newmark_plan(svlgpu_model *m) {
    NewmarkDev &N = m->nm;
    const int n = m->n_int;
    const double dt = m->dt;
    if (m->pml.present) { set_error("Newmark: PML elements need ExtendedNewmarkBeta (history term G), not built"); return 1; }
    if (!m->constraints.empty()) { set_error("Newmark: constrained dofs are not supported on the device path"); return 1; }
    for (auto &mat : m->materials)
        if (mat.kind == SVLGPU_PLASTIC3DJ2 || mat.kind == SVLGPU_PLASTICPLANESTRAINJ2) {
            set_error("Newmark + Linear on the device is limited to linear materials (the tangent is the elastic stiffness)");
            return 1;
        }
    std::vector<double> mask(n), dd(n), dinv(n);
    for (int q = 0; q < n; q++) {
        mask[q] = m->freedof[q] >= 0 ? 1.0 : 0.0;
        dd[q] = 4.0 / dt / dt * m->h_mass[q] + 2.0 / dt * m->h_cdiag[q];
        // several ranks: an interface dof may get all of its mass from other ranks (newmark_comm_setup refreshes D and checks)
        if (mask[q] != 0.0 && !(dd[q] > 0.0) && m->halo_peers.empty()) { set_error("Newmark: free dof without mass (the D-preconditioner needs a positive diagonal)"); return 1; }
        dinv[q] = mask[q] != 0.0 ? 1.0 / dd[q] : 0.0;
    }
    auto up = [&](const std::vector<double> &h, double **d) -> int {
        CUDA_OK(cudaMalloc(d, sizeof(double) * (n + 2)));
        m->allocs.push_back(*d);
        CUDA_OK(cudaMemcpy(*d, h.data(), sizeof(double) * n, cudaMemcpyHostToDevice));
        return 0;
    };
    if (up(mask, &N.d_mask) || up(dd, &N.d_dd) || up(dinv, &N.d_dinv) || up(m->h_mass, &N.d_mass) || up(m->h_cdiag, &N.d_cd)) return 1;
    double **vecs[] = {&N.d_V, &N.d_A, &N.d_b, &N.d_x, &N.d_r, &N.d_p, &N.d_q};
    for (double **v : vecs) {
        CUDA_OK(cudaMalloc(v, sizeof(double) * (n + 2)));
        m->allocs.push_back(*v);
        CUDA_OK(cudaMemset(*v, 0, sizeof(double) * (n + 2)));
    }
    CUDA_OK(cudaMalloc(&N.d_part, sizeof(double) * 8192));
    m->allocs.push_back(N.d_part);
    CUDA_OK(cudaMemset(N.d_part, 0, sizeof(double) * 8192));
    CUDA_OK(cudaMallocHost(&N.h_scal, sizeof(double) * 4));
    N.present = true;
    return 0;
}
// after the communicator exists (halo_comm_init): mglob / cglob = the lumped diagonals with the interface dofs summed over
// the ranks (null on a rank without interface nodes); ownership mask of the dot products
int newmark_comm_setup(svlgpu_model *m, const double *mglob, const double *cglob) {
    NewmarkDev &N = m->nm;
    if (!N.present) return 0;
    const int n = m->n_int, rank = m->halo.rank;
    std::vector<double> own(n, 1.0);
    for (auto &hp : m->halo_peers)
        if (hp.peer < rank)
            for (int node : hp.nodes)
                for (int q = m->node_ptr[node]; q < m->node_ptr[node + 1]; q++) own[q] = 0.0;
    CUDA_OK(cudaMalloc(&N.d_own, sizeof(double) * (n + 2)));
    m->allocs.push_back(N.d_own);
    CUDA_OK(cudaMemcpy(N.d_own, own.data(), sizeof(double) * n, cudaMemcpyHostToDevice));
    if (mglob && cglob) {
        k_nm_refresh<<<(n + 255) / 256, 256, 0, m->stream>>>(n, m->dt, mglob, cglob, N.d_mask, N.d_mass, N.d_cd, N.d_dd, N.d_dinv);
        std::vector<double> dd(n);
        CUDA_OK(cudaMemcpyAsync(dd.data(), N.d_dd, sizeof(double) * n, cudaMemcpyDeviceToHost, m->stream));
        CUDA_OK(cudaStreamSynchronize(m->stream));
        for (int q = 0; q < n; q++)
            if (m->freedof[q] >= 0 && !(dd[q] > 0.0)) { set_error("Newmark: free dof without mass (the D-preconditioner needs a positive diagonal)"); return 1; }
    }
    N.multi = true;
    return 0;
}
void newmark_destroy(svlgpu_model *m) {
    if (m->nm.h_scal) cudaFreeHost(m->nm.h_scal);
    m->nm.h_scal = nullptr;
}
int newmark_set_initial(svlgpu_model *m, const double *V_int, const double *A_int) {
    if (V_int) CUDA_OK(cudaMemcpy(m->nm.d_V, V_int, sizeof(double) * m->n_int, cudaMemcpyHostToDevice));
    if (A_int) CUDA_OK(cudaMemcpy(m->nm.d_A, A_int, sizeof(double) * m->n_int, cudaMemcpyHostToDevice));
    return 0;
}

// one NewmarkBeta step (DynamicAnalysis.cpp:36-57 loop body with NewmarkBeta::ComputeNewStep)
int newmark_step(svlgpu_model *m, int k, const double *dev_amp) {
    NewmarkDev &N = m->nm;
    cudaStream_t st = m->stream;
    const int n = m->n_int;
    const double dt = m->dt, tol2 = N.rtol * N.rtol;
    const double *U = m->d_U[m->cur];
    double *Un = m->d_U[m->next];
    m->k_of_step = k;
    const bool mg = N.multi;                              // several ranks: every rank issues the same collectives
    const double *own = mg ? N.d_own : nullptr;
    // Fint(U_n) = K U_n: the explicit path's force-only pass (Assembler::ComputeInternalForceVector)
    if (N.ak != 0.0) {
        k_nm_w<<<kNmBlocks, kNmThreads, 0, st>>>(n, N.ak, U, N.d_V, N.d_r);
        m->total_launches++;
        if (operator_K(m, N.d_r, N.d_q)) return 1;
    } else if (operator_K(m, U, N.d_q)) return 1;
    if (mg) {
        // interface dofs: this rank's part of K U minus the external forces it was handed, summed over the holders
        if (halo_vec_load(m, N.d_q) || external_forces_interface(m, k, dev_amp) || halo_vec_sum(m, N.d_q)) return 1;
    }
    k_nm_rhs<<<kNmBlocks, kNmThreads, 0, st>>>(n, 4.0 / dt, N.d_q, N.d_mass, N.d_cd, N.d_mask, N.d_V, N.d_A, N.d_b, N.d_part);
    if (external_forces_raw(m, k, dev_amp, N.d_b)) return 1;           // b += Fext(k)  (Assembler::ComputeExternalForceVector)
    k_nm_init<<<kNmBlocks, kNmThreads, 0, st>>>(n, N.d_mask, N.d_dinv, N.d_b, N.d_x, N.d_r, N.d_p, N.d_part, own);
    if (mg && pmlx_allreduce(m, N.d_part, (size_t)N_NSLOT * kNmBlocks, nullptr, 0, st)) return 1;   // slots not written here are rewritten before use
    k_nm_flag<<<1, 32, 0, st>>>(N.d_part, tol2);
    m->total_launches += 3;
    int it = 0, rz = N_RZ0;
    bool done = false;
    int batch = std::max(2, std::min(N.last_iters, N.max_iter));
    while (!done) {
        for (int q = 0; q < batch; q++, it++) {
            if (operator_K(m, N.d_p, N.d_q)) return 1;
            if (mg && (halo_vec_load(m, N.d_q) || halo_vec_sum(m, N.d_q))) return 1;
            k_nm_ap<<<kNmBlocks, kNmThreads, 0, st>>>(n, 1.0 + 2.0 * N.ak / dt, N.d_mask, N.d_dd, N.d_p, N.d_q, N.d_part, own);
            if (mg && pmlx_allreduce(m, N.d_part + N_PAP * kNmBlocks, kNmBlocks, nullptr, 0, st)) return 1;
            k_nm_xr<<<kNmBlocks, kNmThreads, 0, st>>>(n, rz, N.d_dinv, N.d_p, N.d_q, N.d_x, N.d_r, N.d_part, own);
            if (mg && pmlx_allreduce(m, N.d_part + N_RR * kNmBlocks, kNmBlocks,
                                     N.d_part + ((rz == N_RZ0) ? N_RZ1 : N_RZ0) * kNmBlocks, kNmBlocks, st)) return 1;
            k_nm_flag<<<1, 32, 0, st>>>(N.d_part, tol2);
            k_nm_p<<<kNmBlocks, kNmThreads, 0, st>>>(n, rz, N.d_dinv, N.d_r, N.d_p, N.d_part);
            rz = (rz == N_RZ0) ? N_RZ1 : N_RZ0;
            m->total_launches += 4;
        }
        CUDA_OK(cudaMemcpyAsync(N.h_scal, N.d_part + kNmOut, 2 * sizeof(double), cudaMemcpyDeviceToHost, st));
        CUDA_OK(cudaStreamSynchronize(st));
        const double rr = N.h_scal[0], bb = N.h_scal[1];
        if (!(rr == rr)) { set_error("Newmark: the Krylov solve broke down (NaN residual)"); return 1; }
        if (rr <= tol2 * bb + 1e-300) done = true;
        else if (it >= N.max_iter) { set_error("Newmark: the Krylov solve did not converge (LinearSystem::SolveSystem stop)"); return 1; }
        batch = 2;
    }
    N.last_iters = std::max(2, it - 1);
    N.total_iters += it; N.solves++;
    k_nm_update<<<kNmBlocks, kNmThreads, 0, st>>>(n, dt, N.d_x, U, Un, N.d_V, N.d_A);
    m->total_launches++;
    if (record_rows(m, false)) return 1;
    const int old_prev = m->prev;
    m->prev = m->cur; m->cur = m->next; m->next = old_prev;
    m->steps_done++;
    m->dev_k = -1;                                  // the device step counter is not maintained on this path
    CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace svl
