// pml.cu -- the non-diagonal block of the "explicit" step.
//
// PML3DHexa8 / PML2DQuad4 ignore the lumped-mass flag (PML3DHexa8.cpp:214-285, PML2DQuad4.cpp:286-289), so
// Keff = M/dt^2 + C/2dt couples the dofs of the PML nodes and of the soil nodes tied to them.  The reference factors
// this sparse symmetric indefinite matrix once (EigenSolver.cpp:19-60, SimplicialLDLT) or runs PETSc BiCGStab to 1e-12
// and back-substitutes every step; here the block is solved every step, matrix-free, by BiCGStab on the symmetrically
// scaled, sign-flipped system  (J S Keff S) y = J S b,  x = S y,  S = |diag(Keff)|^-1/2,  J = sign(diag(Keff))  (the stress
// rows have a negative diagonal; flipping them makes the operator positive real, and the two-sided scaling balances
// displacement and stress unknowns, which differ by ~10 orders of magnitude).  Element products use per-class tables of Keff_e; the right-hand side
//   b = T'(Fext - Fint + (M/dt^2 - C/2dt)(U_n - U_{n-1}))        (CentralDifference.cpp:217-220)
// uses the class tables of K_e and Kminus_e (PML3DHexa8::ComputeInternalForces is K_e u_e, :716-743).
// Everything is deterministic: gathers run in ascending element order, reductions have a fixed shape.
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

constexpr int kRedBlocks = 1184;     // 8 per SM (full occupancy: the gathers are latency-bound); every reduction writes kRedBlocks partials
constexpr int kRedThreads = 256;
// partial-sum slots
enum { S_BB = 0, S_RHO = 1, S_RHV = 2, S_TS = 3, S_TT = 4, S_RR0 = 5, S_RR1 = 6, S_NSLOT = 7 };
constexpr int kFlagAt = S_NSLOT * kRedBlocks + 16;   // converged flag, behind the partial slots and the 8 carried scalars

__device__ __forceinline__ double block_sum(double v) {
    __shared__ double sh[kRedThreads / 32];
    for (int o = 16; o; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    double r = 0.0;
    if (threadIdx.x < 32) {
        r = (threadIdx.x < kRedThreads / 32) ? sh[threadIdx.x] : 0.0;
        for (int o = 16; o; o >>= 1) r += __shfl_down_sync(0xffffffffu, r, o);
    }
    __syncthreads();
    return r;      // valid in thread 0
}
// every thread gets the full sum of one slot (fixed order -> identical in all blocks)
__device__ __forceinline__ double slot_total(const double *part, int slot) {
    __shared__ double tot;
    double v = 0.0;
    if (threadIdx.x < 32) {
        for (int i = threadIdx.x; i < kRedBlocks; i += 32) v += part[slot * kRedBlocks + i];
        for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (threadIdx.x == 0) tot = v;
    }
    __syncthreads();
    const double r = tot;
    __syncthreads();
    return r;
}

// ---- element products, one thread per (element, row) -----------------------------------------------------------
// mode 0: ye = T1 x1                                   (operator application; x1 indexed by unknown, scaled by xs)
// mode 1: ye = T2 (x1 - x2) - filt(T1 x1)              (right-hand side: T1 = K_e, T2 = Kminus_e, x1 = U_n, x2 = U_{n-1})
// mode 2: ye = filt(T1 x1)                             (internal force K_e u_e)
// filt drops components with |f| <= ftol exactly like Assembler::ComputeInternalForceVector (Assembler.cpp:262); the
// stress rows of the PML are scaled ~1e-10, so unlike in the solids this filter is visible at the 1e-7 level.
template <int NDE, int EPB>
__global__ void __launch_bounds__(NDE * EPB) k_pml_elem(int n_elem, int mode, const int32_t *__restrict__ ecls,
                                                       const int32_t *__restrict__ idx, const double *__restrict__ T1,
                                                       const double *__restrict__ x1, const double *__restrict__ T2,
                                                       const double *__restrict__ x2, double *__restrict__ ye,
                                                       const double *__restrict__ xs, double ftol, const double *part,
                                                       int rr_slot, double tol2) {
    __shared__ double sx1[EPB][NDE], sx2[EPB][NDE];
    if (part && part[kFlagAt] != 0.0) return;             // solver already converged
    const int le = threadIdx.x / NDE, i = threadIdx.x - le * NDE;
    const int e = blockIdx.x * EPB + le;
    const bool act = e < n_elem;
    if (act) {
        const int q = idx[(size_t)e * NDE + i];
        const double a = q >= 0 ? (xs ? x1[q] * xs[q] : x1[q]) : 0.0;
        sx1[le][i] = a;
        if (mode == 1) sx2[le][i] = a - x2[q];
    }
    __syncthreads();
    if (!act) return;
    const size_t base = (size_t)ecls[e] * NDE * NDE + i;
    double acc = 0.0, acc2 = 0.0;
    if (mode == 1) {
#pragma unroll 4
        for (int j = 0; j < NDE; j++) {
            acc = fma(T1[base + (size_t)j * NDE], sx1[le][j], acc);
            acc2 = fma(T2[base + (size_t)j * NDE], sx2[le][j], acc2);
        }
        acc = acc2 - (fabs(acc) > ftol ? acc : 0.0);
    } else {
#pragma unroll 4
        for (int j = 0; j < NDE; j++) acc = fma(T1[base + (size_t)j * NDE], sx1[le][j], acc);
        if (mode == 2 && !(fabs(acc) > ftol)) acc = 0.0;
    }
    ye[(size_t)e * NDE + i] = acc;
}

// ---- element products, class-blocked, pattern-sparse and register-tiled ---------------------------------------------
// The 9x9 (3-D) / 5x5 (2-D) node-pair blocks of M, C, K carry 33 / 13 structural non-zeros (one
// diagonal term per row, the s11-s22-s33 compliance block and the u-sigma gradient coupling), so a row of the element
// matrix holds at most Q = 4 / 3 entries per neighbour node, and -- this is what the tiling uses -- the COLUMN pattern of a
// row depends on its component r only, not on its node j.  One CTA takes up to EB = 8 NEG elements of ONE class:
//   * the class table Ts[k][q][r][j] (18 KB) and the gathered element vectors xs[dof][element] (75 KB) sit in shared memory;
//   * thread (r, eg) owns the NPE rows (j, r) of 8 elements: per (neighbour node k, entry q) it reads NPE table values and
//     8 vector values with 128-bit shared loads and issues 8 NPE DFMAs -- one shared load per 8 DFMAs, where the dense
//     kernel above needs a global and a shared load per DFMA (LSU-bound) and the previous chunked kernel re-read the
//     table from L2 for every 8 elements (profiles/r1n, r1o: 11.7 of 15.1 ms per step at 200^3 + PML);
//   * results go back through shared memory so that the element-force arena is written with coalesced stores.
// The retained terms are summed in the dense kernel's order (ascending column), the skipped ones are exact zeros, so all
// three kernels return the same bits.
//   mode 0: ye = T x        mode 2: ye = filt(T x)        mode 3: ye = T (x1 - x2) - ye   (second pass of the right-hand side)
constexpr int kPmlG = 8;
struct PmlRg {
    int mode;
    const int32_t *grp_cls;        // [n_groups]
    const int32_t *grp_elem;       // [n_groups][EB] element index or -1
    const int32_t *idx;            // [n_elem][NDE] gather index of every element dof (or -1)
    const double *T;               // class tables [cls][NPE][Q][NN][NPE]
    const double *x1, *x2, *xs;
    double *ye;
    double ftol;
    const double *part;            // convergence flag (or null)
    int8_t pat[9][4];              // column component of entry q of a row of component r (padding: coefficient 0)
};
template <int NN, int NPE, int Q, int NEG>
__global__ void __launch_bounds__(NN * NEG, (NEG == 8) ? 4 : 2) k_pml_elem_rg(const PmlRg a) {
    constexpr int NDE = NN * NPE, EB = NEG * kPmlG, ROW = EB + 2, NT = NN * NEG, TSZ = NPE * Q * NN * NPE;
    extern __shared__ __align__(16) double sm[];
    double *xs = sm;                       // [NDE][ROW]: element e = eg + NEG t sits at (t / 2) * 2 NEG + 2 eg + (t & 1)
    double *Ts = sm + NDE * ROW;           // [NPE][Q][NN][NPE]
    __shared__ int spat[9 * 4];
    if (a.part && a.part[kFlagAt] != 0.0) return;         // solver already converged
    const int tid = threadIdx.x;
    if (tid < 36) spat[tid] = a.pat[tid >> 2][tid & 3];
    const int32_t *ge = a.grp_elem + (size_t)blockIdx.x * EB;
    {
        const double *Tg = a.T + (size_t)a.grp_cls[blockIdx.x] * TSZ;
        for (int i = tid; i < TSZ; i += NT) Ts[i] = Tg[i];
    }
    auto slot = [](int e) { const int eg = e % NEG, t = e / NEG; return (t >> 1) * (2 * NEG) + 2 * eg + (t & 1); };
    // gather, 8 (element, dof) pairs per thread at a time: the index loads, then the value loads, go out together (the
    // phase is pure latency: two dependent global loads per pair)
    static_assert((EB * NDE) % (8 * NT) == 0, "gather batches must tile the group");
    for (int p0 = tid; p0 < EB * NDE; p0 += 8 * NT) {
        int zz[8], qq[8];
#pragma unroll
        for (int b = 0; b < 8; b++) zz[b] = ge[(p0 + b * NT) / NDE];
#pragma unroll
        for (int b = 0; b < 8; b++) {
            const int p = p0 + b * NT, e = p / NDE, d = p - e * NDE;
            qq[b] = zz[b] >= 0 ? a.idx[(size_t)zz[b] * NDE + d] : -1;
        }
        double vv[8], ss[8], ww[8];
#pragma unroll
        for (int b = 0; b < 8; b++) {
            vv[b] = qq[b] >= 0 ? a.x1[qq[b]] : 0.0;
            ss[b] = (qq[b] >= 0 && a.xs) ? a.xs[qq[b]] : 1.0;
            ww[b] = (qq[b] >= 0 && a.mode == 3) ? a.x2[qq[b]] : 0.0;
        }
#pragma unroll
        for (int b = 0; b < 8; b++) {
            const int p = p0 + b * NT, e = p / NDE, d = p - e * NDE;
            double v = a.xs ? vv[b] * ss[b] : vv[b];
            if (a.mode == 3) v -= ww[b];
            xs[d * ROW + slot(e)] = v;
        }
    }
    __syncthreads();
    const int r = tid / NEG, eg = tid - r * NEG;
    int colo[Q];                           // row offset of the column component of entry q
#pragma unroll
    for (int q = 0; q < Q; q++) colo[q] = spat[r * 4 + q] * ROW + 2 * eg;
    double acc[NPE][kPmlG];
#pragma unroll
    for (int j = 0; j < NPE; j++)
#pragma unroll
        for (int t = 0; t < kPmlG; t++) acc[j][t] = 0.0;
#pragma unroll
    for (int k = 0; k < NPE; k++) {
#pragma unroll
        for (int q = 0; q < Q; q++) {
            double tv[NPE], xv[kPmlG];
            const double2 *tp = reinterpret_cast<const double2 *>(Ts + ((k * Q + q) * NN + r) * NPE);
#pragma unroll
            for (int j = 0; j < NPE / 2; j++) { const double2 w = tp[j]; tv[2 * j] = w.x; tv[2 * j + 1] = w.y; }
            const double2 *xp = reinterpret_cast<const double2 *>(xs + k * NN * ROW + colo[q]);
#pragma unroll
            for (int u = 0; u < kPmlG / 2; u++) { const double2 w = xp[u * NEG]; xv[2 * u] = w.x; xv[2 * u + 1] = w.y; }
#pragma unroll
            for (int j = 0; j < NPE; j++)
#pragma unroll
                for (int t = 0; t < kPmlG; t++) acc[j][t] = fma(tv[j], xv[t], acc[j][t]);
        }
    }
    __syncthreads();                       // all reads of xs done: reuse it for the results
#pragma unroll
    for (int j = 0; j < NPE; j++) {
        double2 *op = reinterpret_cast<double2 *>(xs + (j * NN + r) * ROW + 2 * eg);
#pragma unroll
        for (int u = 0; u < kPmlG / 2; u++) op[u * NEG] = make_double2(acc[j][2 * u], acc[j][2 * u + 1]);
    }
    __syncthreads();
    for (int p0 = tid; p0 < EB * NDE; p0 += 8 * NT) {
        double old[8];
        int zz[8];
#pragma unroll
        for (int b = 0; b < 8; b++) {
            const int p = p0 + b * NT, e = p / NDE, d = p - e * NDE;
            zz[b] = ge[e];
            old[b] = (a.mode == 3 && zz[b] >= 0) ? a.ye[(size_t)zz[b] * NDE + d] : 0.0;
        }
#pragma unroll
        for (int b = 0; b < 8; b++) {
            const int p = p0 + b * NT, e = p / NDE, d = p - e * NDE;
            if (zz[b] < 0) continue;
            double v = xs[d * ROW + slot(e)];
            if (a.mode == 2 && !(fabs(v) > a.ftol)) v = 0.0;
            if (a.mode == 3) v -= old[b];
            a.ye[(size_t)zz[b] * NDE + d] = v;
        }
    }
}
template <int NN, int NPE, int Q, int NEG> constexpr size_t pml_rg_smem() {
    return sizeof(double) * ((size_t)NN * NPE * (NEG * kPmlG + 2) + (size_t)NPE * Q * NN * NPE);
}

// ---- right-hand side: b~ = W (bext + gather(ye) + soil part);  partial ||b~||^2 ---------------------------------
__global__ void __launch_bounds__(kRedThreads) k_pml_rhs(int nc, const int32_t *ptr, const int32_t *slot, const double *ye,
                                                         const int32_t *c_dof, const int32_t *c_hf, const double *kms,
                                                         const double *hF, const double *U, const double *Up,
                                                         double *bext, const double *w, double *b, double *part,
                                                         double *raw) {
    double acc = 0.0;
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < nc; c += gridDim.x * blockDim.x) {
        double v = bext[c];
        bext[c] = 0.0;
        for (int q = ptr[c]; q < ptr[c + 1]; q++) v += ye[slot[q]];
        const int hs = c_hf[c];
        if (hs >= 0) { const int d = c_dof[c]; v += kms[c] * (U[d] - Up[d]) - hF[hs]; }
        if (raw) { raw[c] = v; continue; }                  // several ranks: this rank's part; k_pml_post finishes it
        v *= w[c];
        b[c] = v;
        acc = fma(v, v, acc);
    }
    if (raw) return;
    const double s = block_sum(acc);
    if (threadIdx.x == 0) {
        part[S_BB * kRedBlocks + blockIdx.x] = s;
        if (blockIdx.x == 0) part[kFlagAt] = 0.0;
    }
}

// ---- out = W (dsoil * in + gather(ye)); optional dots with up to two vectors -----------------------------------
// mode 0: r = b - out, rh = r, p = r (initial residual), partial rho = rr = (r,r)
// mode 1: v = out, partial (rh, v)
// mode 2: t = out, partials (t, s), (t, t)
__global__ void __launch_bounds__(kRedThreads) k_pml_gather(int nc, int mode, const int32_t *ptr, const int32_t *slot,
                                                            const double *ye, const double *dsoil, const double *w,
                                                            const double *sc, const double *in, double *out, const double *b, double *r,
                                                            double *rh, double *p, const double *s, double *part,
                                                            int rr_slot, double tol2, double *raw) {
    if (mode != 0 && part[kFlagAt] != 0.0) return;
    double a0 = 0.0, a1 = 0.0;
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < nc; c += gridDim.x * blockDim.x) {
        double v = dsoil[c] * (sc[c] * in[c]);
        for (int q = ptr[c]; q < ptr[c + 1]; q++) v += ye[slot[q]];
        if (raw) { raw[c] = v; continue; }                  // several ranks: this rank's part; k_pml_post finishes it
        v *= w[c];
        if (mode == 0) {
            const double rv = b[c] - v;
            r[c] = rv; rh[c] = rv; p[c] = rv;
            a0 = fma(rv, rv, a0);
        } else if (mode == 1) {
            out[c] = v;
            a0 = fma(rh[c], v, a0);
        } else {
            out[c] = v;
            a0 = fma(v, s[c], a0);
            a1 = fma(v, v, a1);
        }
    }
    if (raw) return;
    const double s0 = block_sum(a0);
    const double s1 = block_sum(a1);
    if (threadIdx.x == 0) {
        if (mode == 0) { part[S_RHO * kRedBlocks + blockIdx.x] = s0; part[rr_slot * kRedBlocks + blockIdx.x] = s0; }
        else if (mode == 1) part[S_RHV * kRedBlocks + blockIdx.x] = s0;
        else { part[S_TS * kRedBlocks + blockIdx.x] = s0; part[S_TT * kRedBlocks + blockIdx.x] = s1; }
    }
}

// ---- several ranks: second half of k_pml_rhs / k_pml_gather after the exchange of the shared unknowns ----------------
// raw holds the complete sums (identical bits on every rank that holds the unknown); the dot products count an unknown on
// its lowest rank only (own = 1 there, 0 elsewhere) and are all-reduced by the caller.
// mode -1: b = W raw, partial ||b||^2, clears the converged flag;   modes 0, 1, 2: as k_pml_gather
__global__ void __launch_bounds__(kRedThreads) k_pml_post(int nc, int mode, const double *raw, const double *w, const double *own,
                                                          double *out, double *b, double *r, double *rh, double *p,
                                                          const double *s, double *part, int rr_slot) {
    if (mode > 0 && part[kFlagAt] != 0.0) return;
    double a0 = 0.0, a1 = 0.0;
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < nc; c += gridDim.x * blockDim.x) {
        const double v = raw[c] * w[c], o = own[c];
        if (mode == -1) {
            b[c] = v;
            a0 = fma(o * v, v, a0);
        } else if (mode == 0) {
            const double rv = b[c] - v;
            r[c] = rv; rh[c] = rv; p[c] = rv;
            a0 = fma(o * rv, rv, a0);
        } else if (mode == 1) {
            out[c] = v;
            a0 = fma(o * rh[c], v, a0);
        } else {
            out[c] = v;
            a0 = fma(o * v, s[c], a0);
            a1 = fma(o * v, v, a1);
        }
    }
    const double s0 = block_sum(a0);
    const double s1 = block_sum(a1);
    if (threadIdx.x == 0) {
        if (mode == -1) { part[S_BB * kRedBlocks + blockIdx.x] = s0; if (blockIdx.x == 0) part[kFlagAt] = 0.0; }
        else if (mode == 0) { part[S_RHO * kRedBlocks + blockIdx.x] = s0; part[rr_slot * kRedBlocks + blockIdx.x] = s0; }
        else if (mode == 1) part[S_RHV * kRedBlocks + blockIdx.x] = s0;
        else { part[S_TS * kRedBlocks + blockIdx.x] = s0; part[S_TT * kRedBlocks + blockIdx.x] = s1; }
    }
}
// w = sign(d) / sqrt|d|, sc = 1 / sqrt|d| from the diagonal of Keff summed over the ranks (in raw)
__global__ void k_pml_rescale(int nc, const double *diag, double *w, double *sc) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nc) return;
    const double d = diag[c], s = 1.0 / sqrt(fabs(d));
    sc[c] = s;
    w[c] = d > 0.0 ? s : -s;
}

// scalars carried between kernels: [0] rho of the previous iteration, [1] alpha, [2] omega, [3] rho of this iteration
// p = r + beta (p - omega v)
__global__ void __launch_bounds__(kRedThreads) k_bicg_p(int nc, int first, const double *r, double *p, const double *v,
                                                        double *scal, const double *part, int rr_slot, double tol2) {
    if (part[kFlagAt] != 0.0) return;
    const double rho = slot_total(part, S_RHO);
    if (!first) {
        const double beta = (rho / scal[0]) * (scal[1] / scal[2]);
        const double omega = scal[2];
        for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < nc; c += gridDim.x * blockDim.x)
            p[c] = r[c] + beta * (p[c] - omega * v[c]);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) scal[3] = rho;
}
// s = r - alpha v
__global__ void __launch_bounds__(kRedThreads) k_bicg_s(int nc, const double *r, const double *v, double *s, double *scal,
                                                        const double *part, int rr_slot, double tol2) {
    if (part[kFlagAt] != 0.0) return;
    const double rhv = slot_total(part, S_RHV);
    const double alpha = scal[3] / rhv;
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < nc; c += gridDim.x * blockDim.x) s[c] = r[c] - alpha * v[c];
    if (blockIdx.x == 0 && threadIdx.x == 0) { scal[4] = alpha; }
}
// x += alpha p + omega s; r = s - omega t; partial rho_next = (rh, r), rr = (r, r) into the other rr slot
__global__ void __launch_bounds__(kRedThreads) k_bicg_x(int nc, double *x, const double *p, const double *s, const double *t,
                                                        double *r, const double *rh, double *scal, double *part,
                                                        int rr_slot, double tol2, const double *own) {
    if (part[kFlagAt] != 0.0) return;
    const int nslot = (rr_slot == S_RR0) ? S_RR1 : S_RR0;
    const double tt = slot_total(part, S_TT), ts = slot_total(part, S_TS);
    const double omega = tt > 0.0 ? ts / tt : 0.0;
    const double alpha = scal[4];
    double a0 = 0.0, a1 = 0.0;
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < nc; c += gridDim.x * blockDim.x) {
        const double sv = s[c];
        x[c] += alpha * p[c] + omega * sv;
        const double rv = sv - omega * t[c];
        r[c] = rv;
        const double o = own ? own[c] : 1.0;              // several ranks: an unknown counts on its lowest rank only
        a0 = fma(o * rh[c], rv, a0);
        a1 = fma(o * rv, rv, a1);
    }
    const double s0 = block_sum(a0);
    const double s1 = block_sum(a1);
    if (threadIdx.x == 0) {
        part[S_RHO * kRedBlocks + blockIdx.x] = s0;
        part[nslot * kRedBlocks + blockIdx.x] = s1;
        if (blockIdx.x == 0) { scal[0] = scal[3]; scal[1] = alpha; scal[2] = omega; }
    }
}
// convergence test after every residual update: one warp sums the partials once, every later kernel of the solve reads
// one word (a per-CTA re-summation of the partials cost 86 us per launch on 38 000 CTAs: profiles/r1n).  out = {rr, bb}
// of the last test that ran; the flag is sticky until the next right-hand side clears it.
__global__ void k_pml_flag(double *part, int rr_slot, double tol2, double *out) {
    if (part[kFlagAt] != 0.0) return;
    const double rr = slot_total(part, rr_slot), bb = slot_total(part, S_BB);
    if (threadIdx.x == 0) {
        out[0] = rr; out[1] = bb;
        if (rr <= tol2 * bb + 1e-280 || !(rr == rr)) part[kFlagAt] = 1.0;
    }
}
// U_{n+1} = U_n + T dU on every dof of the block (slaves take their master's increment, Mesh.cpp:360-375)
__global__ void k_pml_scatter(int n, const int32_t *dof, const int32_t *cix, const double *x, const double *sc,
                              const double *U, double *Un) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const int d = dof[t], c = cix[t];
    Un[d] = c >= 0 ? U[d] + sc[c] * x[c] : U[d];          // restrained dofs keep their value (Mesh.cpp:354-357)
}
__global__ void k_pml_fint_add(int n, const int32_t *edof, const double *ye, double *F) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    atomicAdd(&F[edof[t]], ye[t]);
}

template <int NDE, int EPB>
static void launch_elem(svlgpu_model *m, int mode, const int32_t *idx, const double *T1, const double *x1, const double *T2,
                        const double *x2, const double *xs, const double *part, int rr_slot, double tol2) {
    PmlDev &P = m->pml;
    k_pml_elem<NDE, EPB><<<(P.n_elem + EPB - 1) / EPB, NDE * EPB, 0, m->stream>>>(P.n_elem, mode, P.d_ecls, idx, T1, x1, T2, x2,
                                                                                 P.d_ye, xs, P.ftol, part, rr_slot, tol2);
    m->total_launches++;
}
static void elem_products(svlgpu_model *m, int mode, const int32_t *idx, const double *T1, const double *x1, const double *T2,
                          const double *x2, const double *xs, const double *part, int rr_slot, double tol2) {
    PmlDev &P = m->pml;
    if (P.n_elem == 0) return;                            // a rank that only holds replicas of PML unknowns
    timer_begin(m, 6);
    struct End { svlgpu_model *m; ~End() { timer_end(m, 6); } } end_{m};
    if (P.sp_q > 0) {
        // class-blocked register-tiled kernel: T1 / T2 name the dense tables; use their tiled twins
        auto sp = [&](const double *T) { return T == P.d_A ? P.d_sA : T == P.d_K ? P.d_sK : T == P.d_Km ? P.d_sKm : nullptr; };
        PmlRg a;
        a.grp_cls = P.d_chunk_cls; a.grp_elem = P.d_chunk_elem; a.idx = idx;
        a.x1 = x1; a.x2 = x2; a.xs = xs; a.ye = P.d_ye; a.ftol = P.ftol; a.part = part;
        std::memcpy(a.pat, P.sp_pat, sizeof(a.pat));
        auto launch = [&](int md, const double *T) {
            a.mode = md; a.T = T;
            if (P.nde == 72) k_pml_elem_rg<9, 8, 4, kPmlChunk3 / 8><<<P.n_chunks, 9 * (kPmlChunk3 / 8), pml_rg_smem<9, 8, 4, kPmlChunk3 / 8>(), m->stream>>>(a);
            else k_pml_elem_rg<5, 4, 3, 16><<<P.n_chunks, 5 * 16, pml_rg_smem<5, 4, 3, 16>(), m->stream>>>(a);
            m->total_launches++;
        };
        if (mode == 1) { launch(2, sp(T1)); launch(3, sp(T2)); }      // ye = Kminus (x1 - x2) - filt(K x1)
        else launch(mode, sp(T1));
        return;
    }
    if (P.nde == 72) launch_elem<72, 4>(m, mode, idx, T1, x1, T2, x2, xs, part, rr_slot, tol2);
    else launch_elem<20, 8>(m, mode, idx, T1, x1, T2, x2, xs, part, rr_slot, tol2);
}

// Several ranks: finish a right-hand side / operator application whose local part sits in d_raw -- exchange the shared
// unknowns (rank-ordered sums: identical bits on every holder), scale, update the vectors, all-reduce the dot products.
static int post_exchange(svlgpu_model *m, int mode, double *out, const double *s, int rr) {
    PmlDev &P = m->pml;
    cudaStream_t st = m->stream;
    if (pmlx_exchange(m, st)) return 1;
    timer_begin(m, 7);
    k_pml_post<<<kRedBlocks, kRedThreads, 0, st>>>(P.nc, mode, P.d_raw, P.d_w, P.d_own, out, P.d_b, P.d_r, P.d_rh, P.d_p, s, P.d_part, rr);
    timer_end(m, 7);
    m->total_launches++;
    double *part = P.d_part;
    switch (mode) {
    case -1: return pmlx_allreduce(m, part + S_BB * kRedBlocks, kRedBlocks, nullptr, 0, st);
    case 0: return pmlx_allreduce(m, part + S_RHO * kRedBlocks, kRedBlocks, part + rr * kRedBlocks, kRedBlocks, st);
    case 1: return pmlx_allreduce(m, part + S_RHV * kRedBlocks, kRedBlocks, nullptr, 0, st);
    default: return pmlx_allreduce(m, part + S_TS * kRedBlocks, 2 * kRedBlocks, nullptr, 0, st);      // S_TS, S_TT are adjacent
    }
}

// starting guess of the block solve: the increment extrapolated linearly from the last two steps, x0 = 2 x_n - x_{n-1}
// (the increment of a resolved wave field is smooth in time: the residual of this guess is O((w dt)^2) of ||b|| instead of
// O(w dt) for x_n alone, which saves BiCGStab iterations); element-wise, so replicas on several ranks stay bit-identical
__global__ void k_pml_extrapolate(int n, double *x, double *xp) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const double xn = x[t];
    x[t] = 2.0 * xn - xp[t];
    xp[t] = xn;
}

int pml_step(svlgpu_model *m, const double *U, const double *Up, double *Un) {
    PmlDev &P = m->pml;
    const bool mg = P.multi;                              // several ranks: every rank runs the same sequence of collectives
    if (!P.present || (!P.nc && !mg)) return 0;
    cudaStream_t st = m->stream;
    const double tol2 = P.rtol * P.rtol;
    double *scal = P.d_part + S_NSLOT * kRedBlocks;       // 8 carried scalars behind the partial slots
    double *raw = mg ? P.d_raw : nullptr;
    // right-hand side
    elem_products(m, 1, P.d_edof, P.d_K, U, P.d_Km, Up, nullptr, nullptr, 0, 0.0);
    timer_begin(m, 7);
    k_pml_rhs<<<kRedBlocks, kRedThreads, 0, st>>>(P.nc, P.d_c_ptr, P.d_c_slot, P.d_ye, P.d_c_dof, P.d_c_hf, P.d_kms, m->halo.d_hF,
                                                  U, Up, P.d_bext, P.d_w, P.d_b, P.d_part, raw);
    timer_end(m, 7);
    if (mg && post_exchange(m, -1, nullptr, nullptr, 0)) return 1;
    // initial residual with the previous increment (SVLGPU_PML_EXTRAP: the linearly extrapolated one) as the starting guess
    if (P.extrapolate && P.solves >= 2 && P.nc) {
        k_pml_extrapolate<<<(P.nc + 255) / 256, 256, 0, st>>>(P.nc, P.d_x, P.d_xp);
        m->total_launches++;
    } else if (P.extrapolate && P.nc) {
        CUDA_OK(cudaMemcpyAsync(P.d_xp, P.d_x, sizeof(double) * P.nc, cudaMemcpyDeviceToDevice, st));
    }
    elem_products(m, 0, P.d_ecd, P.d_A, P.d_x, nullptr, nullptr, P.d_sc, nullptr, 0, 0.0);
    int rr = S_RR0;
    timer_begin(m, 7);
    k_pml_gather<<<kRedBlocks, kRedThreads, 0, st>>>(P.nc, 0, P.d_c_ptr, P.d_c_slot, P.d_ye, P.d_diag, P.d_w, P.d_sc, P.d_x, nullptr, P.d_b,
                                                     P.d_r, P.d_rh, P.d_p, nullptr, P.d_part, rr, tol2, raw);
    timer_end(m, 7);
    if (mg && post_exchange(m, 0, nullptr, nullptr, rr)) return 1;
    k_pml_flag<<<1, 32, 0, st>>>(P.d_part, rr, tol2, scal + 6);
    m->total_launches += 3;
    int it = 0;
    bool done = false;
    int batch = std::max(2, std::min(P.last_iters, P.max_iter));
    while (!done) {
        for (int q = 0; q < batch; q++, it++) {
            timer_begin(m, 8);
            k_bicg_p<<<kRedBlocks, kRedThreads, 0, st>>>(P.nc, it == 0, P.d_r, P.d_p, P.d_v, scal, P.d_part, rr, tol2);
            timer_end(m, 8);
            elem_products(m, 0, P.d_ecd, P.d_A, P.d_p, nullptr, nullptr, P.d_sc, P.d_part, rr, tol2);
            timer_begin(m, 7);
            k_pml_gather<<<kRedBlocks, kRedThreads, 0, st>>>(P.nc, 1, P.d_c_ptr, P.d_c_slot, P.d_ye, P.d_diag, P.d_w, P.d_sc, P.d_p, P.d_v,
                                                             nullptr, nullptr, P.d_rh, nullptr, nullptr, P.d_part, rr, tol2, raw);
            timer_end(m, 7);
            if (mg && post_exchange(m, 1, P.d_v, nullptr, rr)) return 1;
            timer_begin(m, 8);
            k_bicg_s<<<kRedBlocks, kRedThreads, 0, st>>>(P.nc, P.d_r, P.d_v, P.d_s, scal, P.d_part, rr, tol2);
            timer_end(m, 8);
            elem_products(m, 0, P.d_ecd, P.d_A, P.d_s, nullptr, nullptr, P.d_sc, P.d_part, rr, tol2);
            timer_begin(m, 7);
            k_pml_gather<<<kRedBlocks, kRedThreads, 0, st>>>(P.nc, 2, P.d_c_ptr, P.d_c_slot, P.d_ye, P.d_diag, P.d_w, P.d_sc, P.d_s, P.d_t,
                                                             nullptr, nullptr, nullptr, nullptr, P.d_s, P.d_part, rr, tol2, raw);
            timer_end(m, 7);
            if (mg && post_exchange(m, 2, P.d_t, P.d_s, rr)) return 1;
            timer_begin(m, 8);
            k_bicg_x<<<kRedBlocks, kRedThreads, 0, st>>>(P.nc, P.d_x, P.d_p, P.d_s, P.d_t, P.d_r, P.d_rh, scal, P.d_part, rr, tol2,
                                                         mg ? P.d_own : nullptr);
            timer_end(m, 8);
            rr = (rr == S_RR0) ? S_RR1 : S_RR0;
            // k_bicg_x wrote the partials of the next rho and of ||r||^2 (into the slot rr now names)
            if (mg && pmlx_allreduce(m, P.d_part + S_RHO * kRedBlocks, kRedBlocks, P.d_part + rr * kRedBlocks, kRedBlocks, st)) return 1;
            k_pml_flag<<<1, 32, 0, st>>>(P.d_part, rr, tol2, scal + 6);
            m->total_launches += 6;
        }
        CUDA_OK(cudaMemcpyAsync(P.h_scal, scal + 6, 2 * sizeof(double), cudaMemcpyDeviceToHost, st));
        CUDA_OK(cudaStreamSynchronize(st));
        const double rrv = P.h_scal[0], bb = P.h_scal[1];
        if (!(rrv == rrv)) { set_error("PML block solve broke down (NaN residual)"); return 1; }
        if (rrv <= tol2 * bb + 1e-280) done = true;
        else if (it >= P.max_iter) { set_error("PML block solve did not converge (LinearSystem::SolveSystem stop)"); return 1; }
        batch = 2;
    }
    P.last_iters = std::max(2, it - 1);
    P.total_iters += it; P.solves++;
    if (P.n_sc) {
        timer_begin(m, 8);
        k_pml_scatter<<<(P.n_sc + 255) / 256, 256, 0, st>>>(P.n_sc, P.d_sc_dof, P.d_sc_c, P.d_x, P.d_sc, U, Un);
        timer_end(m, 8);
        m->total_launches++;
    }
    CUDA_OK(cudaGetLastError());
    return 0;
}

// several ranks, once: w and sc from the diagonal of Keff summed over the holders of each unknown (pmlx_setup left it in d_raw)
int pml_rescale(svlgpu_model *m) {
    PmlDev &P = m->pml;
    if (P.nc) k_pml_rescale<<<(P.nc + 255) / 256, 256, 0, m->stream>>>(P.nc, P.d_raw, P.d_w, P.d_sc);
    CUDA_OK(cudaGetLastError());
    return 0;
}

// adds the PML part of Assembler::ComputeInternalForceVector (K_e u_e per element) to F (internal dof order)
int pml_internal_force(svlgpu_model *m, const double *U, double *F) {
    PmlDev &P = m->pml;
    if (!P.present || !P.n_elem) return 0;
    elem_products(m, 2, P.d_edof, P.d_K, U, nullptr, nullptr, nullptr, nullptr, 0, 0.0);
    const int n = P.n_elem * P.nde;
    k_pml_fint_add<<<(n + 255) / 256, 256, 0, m->stream>>>(n, P.d_edof, P.d_ye, F);
    CUDA_OK(cudaGetLastError());
    return 0;
}

int pml_configure() {
    CUDA_OK(cudaFuncSetAttribute(k_pml_elem_rg<9, 8, 4, kPmlChunk3 / 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pml_rg_smem<9, 8, 4, kPmlChunk3 / 8>()));
    CUDA_OK(cudaFuncSetAttribute(k_pml_elem_rg<5, 4, 3, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pml_rg_smem<5, 4, 3, 16>()));
    return 0;
}

void pml_destroy(svlgpu_model *m) {
    if (m->pml.h_scal) cudaFreeHost(m->pml.h_scal);
    m->pml.h_scal = nullptr;
}

}  // namespace svl
