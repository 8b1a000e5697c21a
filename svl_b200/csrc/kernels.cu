// kernels.cu -- sm_100a FP64 kernels of the explicit step (see DESIGN.md for the data layout).
//
//   k_stencil3<NW>   block-stencil force + CentralDifference update for lattice blocks of
//                    lin3DHexa8/Elastic3DLinear (node-class pre-summed 27 x 3x3 rows of K)
//   k_stencil2       same for lin2DQuad4/Elastic2DPlaneStrain lattices (9 x 2x2)
//   k_gen_hex8/quad4 Gauss-point element force (elastic or J2 return map), one thread per
//                    Gauss point, warp-shuffle reduce-scatter of B^T sigma to the 8 (4) nodes
//   k_gen_nodes      atomic-free node gather (ascending element order) + update
//   k_nodal_loads    point loads,  k_drm  DRM effective forces,  k_record  NODE recorder rows
//
// Reference semantics reproduced: CentralDifference.cpp:123-152,205-221 (update),
// Assembler.cpp:239-269 (scatter order), lin3DHexa8.cpp:86-107,382-412 (element force).
#include <cstdio>
#include <cstring>
#include <algorithm>
#include <type_traits>
#include "model.h"
#include "elem_math.h"

namespace svl {

#define CUDA_OK(x)                                                                          \
    do {                                                                                    \
        cudaError_t e_ = (x);                                                               \
        if (e_ != cudaSuccess) {                                                            \
            set_error(std::string(#x) + ": " + cudaGetErrorString(e_));                     \
            return 1;                                                                       \
        }                                                                                   \
    } while (0)

// ------------------------------------------------------------------------------------------
// cp.async helpers (LDGSTS; 8-byte granules because a lattice row starts on an 8 B boundary)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async8(double *smem_dst, const double *gsrc, bool valid) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    int sz = valid ? 8 : 0;                       // src-size 0 => zero fill
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(s), "l"(gsrc), "r"(sz));
}
// same, ordered against the surrounding shared-memory accesses of the issuing thread
__device__ __forceinline__ void cp_async8m(double *smem_dst, const double *gsrc) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(s), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

// ------------------------------------------------------------------------------------------
// 3-D block stencil, dominant node class.
//
// A lattice block has one (or a few) node classes that cover almost all nodes (the interior of
// a homogeneous region).  Their 27 x 3x3 pre-summed coefficient rows live in __constant__
// memory and are used as constant-bank operands of the DFMAs: no register, no LSU traffic.
// The kernel computes every node of its box as if it belonged to the dominant class and stores
// only those that do; the thin shells of other classes are done by k_stencil3_gather.
//
// Work decomposition: CTA = 32 (x) x NW*R (y) nodes, marching over kz planes.  Plane kk of U_n
// is staged in shared memory by cp.async (3-deep ring); each thread owns R consecutive rows and
// scatters plane kk into the accumulators of output planes kk-1, kk, kk+1 (slots 0,1,2), so
// every U value is read from shared memory once per (di,b) and each output node is finished
// when its upper neighbour plane has been consumed.
// ------------------------------------------------------------------------------------------
constexpr int kDomSlots = 4;
__constant__ double cK[kDomSlots][kTbl3Stride];

struct Dom3 {
    const double *U;      // U_n      (internal dof array)
    const double *Up;     // U_{n-1}
    double *Un;           // U_{n+1}  (mode 0)  or  F_int (mode 1)
    const uint8_t *cls;   // [nx*ny*nz]
    long long dof0;       // internal dof of lattice node (0,0,0)
    int nx, ny, nz;
    int bi0, bj0, bk0, bk1;   // box origin and z end (exclusive) of the nodes of this class
    int bi1, bj1;             // x / y ends (exclusive); used by the pure-box kernel
    int tiles_x, tiles_y, kz;
    int dom;              // class id to store
    int mode;
};

// structural zero of the stencil of a rectangular cell with isotropic material: the a != b
// block entry vanishes unless the neighbour offset is non-zero along both axes a and b
__host__ __device__ constexpr bool stencil_nz(int di, int b, int dj, int s, int a) {
    if (a == b) return true;
    const int d0 = di - 1, d1 = dj - 1, d2 = 1 - s;
    const int da = (a == 0) ? d0 : (a == 1) ? d1 : d2;
    const int db = (b == 0) ? d0 : (b == 1) ? d1 : d2;
    return da != 0 && db != 0;
}

template <int NW, int R, int SLOT, bool ORTHO>
__global__ void __launch_bounds__(NW * 32, (NW <= 4) ? 3 : 1) k_stencil3_dom(const Dom3 p) {
    constexpr int TY = NW * R;           // tile rows
    constexpr int TYH = TY + 2;          // + halo
    constexpr int ROWP = 36;             // 34 columns (+halo) padded to a 16 B multiple
    constexpr int PLANE = 3 * TYH * ROWP;
    constexpr int ROWD = 102;            // doubles per tile row in global memory (34 nodes x 3)
    extern __shared__ __align__(16) double pl[];

    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int item = blockIdx.x;
    const int txi = item % p.tiles_x; item /= p.tiles_x;
    const int tyi = item % p.tiles_y; item /= p.tiles_y;
    const int i0 = p.bi0 + txi * 32, j0 = p.bj0 + tyi * TY;
    const int k0 = p.bk0 + item * p.kz, k1 = min(k0 + p.kz, p.bk1);
    const int gi = i0 + lane, gjb = j0 + w * R;

    auto load_plane = [&](int k, double *buf) {
        const bool kin = (k >= 0) && (k < p.nz);
        for (int t = threadIdx.x; t < TYH * ROWD; t += NW * 32) {
            const int row = t / ROWD, d = t - row * ROWD;
            const int ii = d / 3, c = d - 3 * ii;
            const int x = i0 - 1 + ii, y = j0 - 1 + row;
            const bool ok = kin && (x >= 0) && (x < p.nx) && (y >= 0) && (y < p.ny);
            const double *src = ok ? p.U + p.dof0 + 3ll * (x + (long long)p.nx * (y + (long long)p.ny * k)) + c : p.U;
            cp_async8(buf + (c * TYH + row) * ROWP + ii, src, ok);
        }
    };

    double *ups = pl + 3 * PLANE + (w * 32 + lane) * (R * 3);   // this thread's U_{n-1} slots

    double acc[3][R][3];
    double ucen[R][3];               // U_n of this thread's nodes on the previous plane
#pragma unroll
    for (int s = 0; s < 3; s++)
#pragma unroll
        for (int r = 0; r < R; r++)
#pragma unroll
            for (int a = 0; a < 3; a++) { acc[s][r][a] = 0.0; ucen[r][a] = 0.0; }

    // class of this thread's nodes, fetched one plane ahead of its use
    auto my_cls = [&](int r, int k) -> bool {
        if (gi >= p.nx || gjb + r >= p.ny || k < k0 || k >= k1) return false;
        return p.cls[gi + (long long)p.nx * (gjb + r + (long long)p.ny * k)] == p.dom;
    };
    bool mine[R], mine_nx[R];
#pragma unroll
    for (int r = 0; r < R; r++) { mine[r] = false; mine_nx[r] = false; }   // planes k0-2, k0-1: not ours

    load_plane(k0 - 1, pl);
    cp_async_commit();

    int idx = 0;
    for (int kk = k0 - 1; kk <= k1; kk++, idx++) {
        const double *buf = pl + (idx % 3) * PLANE;
        if (kk + 1 <= k1) load_plane(kk + 1, pl + ((idx + 1) % 3) * PLANE);
        // U_{n-1} of the nodes finished at the end of this iteration (plane kk-1): own slots only,
        // so cp.async.wait_group alone orders them (no barrier needed)
        if (p.mode == 0) {
#pragma unroll
            for (int r = 0; r < R; r++)
                if (mine[r]) {
                    const double *q = p.Up + p.dof0 + 3ll * (gi + (long long)p.nx * (gjb + r + (long long)p.ny * (kk - 1)));
                    cp_async8(ups + 3 * r + 0, q + 0, true);
                    cp_async8(ups + 3 * r + 1, q + 1, true);
                    cp_async8(ups + 3 * r + 2, q + 2, true);
                }
        }
        cp_async_commit();
        bool mine_n2[R];
#pragma unroll
        for (int r = 0; r < R; r++) mine_n2[r] = my_cls(r, kk + 1);
        cp_async_wait<1>();
        __syncthreads();

#pragma unroll
        for (int di = 0; di < 3; di++) {
#pragma unroll
            for (int b = 0; b < 3; b++) {
                const double *ub = buf + (b * TYH + w * R) * ROWP + lane + di;
                double u[R + 2];
#pragma unroll
                for (int q = 0; q < R + 2; q++) u[q] = ub[q * ROWP];
#pragma unroll
                for (int dj = 0; dj < 3; dj++)
#pragma unroll
                    for (int s = 0; s < 3; s++)
#pragma unroll
                        for (int a = 0; a < 3; a++) {
                            if (ORTHO && !stencil_nz(di, b, dj, s, a)) continue;
                            const double c = cK[SLOT][((di * 3 + b) * 3 + dj) * 10 + s * 3 + a];
#pragma unroll
                            for (int r = 0; r < R; r++) acc[s][r][a] = fma(c, u[r + dj], acc[s][r][a]);
                        }
            }
        }
        // ---- nodes of plane kk-1 are complete: CentralDifference update (CentralDifference.cpp:138-148)
        cp_async_wait<0>();
#pragma unroll
        for (int r = 0; r < R; r++) {
            if (mine[r]) {
                const long long d0 = p.dof0 + 3ll * (gi + (long long)p.nx * (gjb + r + (long long)p.ny * (kk - 1)));
#pragma unroll
                for (int a = 0; a < 3; a++) {
                    if (p.mode == 0) {
                        const double un = ucen[r][a];
                        const double du = (cK[SLOT][273 + a] * (un - ups[3 * r + a]) - acc[0][r][a]) * cK[SLOT][270 + a];
                        p.Un[d0 + a] = un + du;
                    } else {
                        p.Un[d0 + a] = acc[0][r][a];
                    }
                }
            }
        }
        // ---- rotate: plane kk becomes "previous"
#pragma unroll
        for (int r = 0; r < R; r++) {
#pragma unroll
            for (int a = 0; a < 3; a++) {
                acc[0][r][a] = acc[1][r][a];
                acc[1][r][a] = acc[2][r][a];
                acc[2][r][a] = 0.0;
                // plane kk stays in the ring until the prefetch of iteration kk+2, i.e. past the next barrier
                ucen[r][a] = buf[(a * TYH + w * R + r + 1) * ROWP + lane + 1];
            }
            mine[r] = mine_nx[r];
            mine_nx[r] = mine_n2[r];
        }
    }
}

// ------------------------------------------------------------------------------------------
// 3-D block stencil, dominant node class, v3: same arithmetic and work decomposition as k_stencil3_dom, but
//   * the planes of U_n are brought in by the TMA unit: one 1-D bulk copy (cp.async.bulk, UBLKCP) per tile row,
//     issued by TYH threads and completed on an mbarrier (3-stage ring, prefetch distance 2 planes) -- instead of
//     ~14 8-byte LDGSTS with index arithmetic per thread and plane.  Rows keep their global [node][component]
//     interleaving in shared memory (stride-3 doubles across lanes is bank-conflict free for 64-bit accesses); a
//     row whose first element sits on an odd double is copied from one element earlier and read with a per-row
//     shift of one element (bulk copies need 16-byte aligned addresses and sizes).
//   * the class of a node is not looked up: the planner guarantees that the class fills its bounding box and
//     that all 27 neighbours of its nodes exist (interior class), so "inside the box" is the store predicate and
//     out-of-lattice halo cells never contribute to a stored node.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity) {
    unsigned done;
    do {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void bulk_g2s(double *dst, const double *src, unsigned bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

constexpr int kT3Row = 106;          // doubles per staged row: 34 nodes x 3 + alignment slack, even

template <int R, bool ORTHO, int SLOT>
__device__ __forceinline__ void stencil3_plane(const double *pl, const int (&rowp)[R + 2], double (&A)[3][R][3]) {
#pragma unroll
    for (int di = 0; di < 3; di++) {
#pragma unroll
        for (int b = 0; b < 3; b++) {
            double u[R + 2];
#pragma unroll
            for (int q = 0; q < R + 2; q++) u[q] = pl[rowp[q] + 3 * di + b];
#pragma unroll
            for (int dj = 0; dj < 3; dj++)
#pragma unroll
                for (int s = 0; s < 3; s++)
#pragma unroll
                    for (int a = 0; a < 3; a++) {
                        if (ORTHO && !stencil_nz(di, b, dj, s, a)) continue;
                        const double c = cK[SLOT][((di * 3 + b) * 3 + dj) * 10 + s * 3 + a];
#pragma unroll
                        for (int r = 0; r < R; r++) A[s][r][a] = fma(c, u[r + dj], A[s][r][a]);
                    }
        }
    }
}

template <int NW, int R, int SLOT, bool ORTHO>
__global__ void __launch_bounds__(NW * 32, (NW <= 4) ? 3 : 1) k_stencil3_tma(const Dom3 p) {
    constexpr int TY = NW * R, TYH = TY + 2, NS = 4;       // 4 stages: plane kk-1 stays readable while kk+2 is in flight
    constexpr int PLANE = TYH * kT3Row;
    extern __shared__ __align__(16) double pl[];
    double *ups_all = pl + NS * PLANE;
    uint64_t *bar = reinterpret_cast<uint64_t *>(ups_all + NW * 32 * R * 3);

    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int item = blockIdx.x;
    const int txi = item % p.tiles_x; item /= p.tiles_x;
    const int tyi = item % p.tiles_y; item /= p.tiles_y;
    const int i0 = p.bi0 + txi * 32, j0 = p.bj0 + tyi * TY;
    const int k0 = p.bk0 + item * p.kz, k1 = min(k0 + p.kz, p.bk1);
    const int gi = i0 + lane, gjb = j0 + w * R;

    if (threadIdx.x == 0) {
        for (int s = 0; s < NS; s++) mbar_init(&bar[s], TYH);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();

    // row span in doubles relative to the lattice row start, clamped to the row
    const int e0 = 3 * (i0 - 1);
    const int ea = max(e0, 0), eb = min(3 * (i0 + 33), 3 * p.nx);
    const int q0 = ((ea - e0) + (int)(p.dof0 & 1) + ea) & 1;            // shift parity of row y = 0, k = 0
    const int altx = p.nx & 1, alty = p.ny & 1;

    auto issue_plane = [&](int k, int stage) {                            // threads 0 .. TYH-1: one row each
        if (threadIdx.x >= TYH) return;
        const int row = threadIdx.x, y = j0 - 1 + row;
        if (k < 0 || k >= p.nz || y < 0 || y >= p.ny || eb <= ea) { mbar_arrive(&bar[stage]); return; }
        const long long g = p.dof0 + 3ll * p.nx * (y + (long long)p.ny * k) + ea;
        const int par = (int)(g & 1);
        int len = eb - ea + par;
        len += len & 1;
        const int dsti = (ea - par - e0) + ((ea - par - e0) & 1);
        mbar_arrive_expect_tx(&bar[stage], (unsigned)len * 8u);
        bulk_g2s(pl + stage * PLANE + row * kT3Row + dsti, p.U + (g - par), (unsigned)len * 8u, &bar[stage]);
    };

    // U_{n-1} slots of this thread: [slot][thread] so that a warp touches consecutive 8-byte words (conflict free)
    constexpr int NT = NW * 32;
    double *ups = ups_all + threadIdx.x;
    double A[3][R][3];
#pragma unroll
    for (int s = 0; s < 3; s++)
#pragma unroll
        for (int r = 0; r < R; r++)
#pragma unroll
            for (int a = 0; a < 3; a++) A[s][r][a] = 0.0;

    const bool col_in = (gi >= p.bi0) && (gi < p.bi1);
    bool rowin[R];
#pragma unroll
    for (int r = 0; r < R; r++) rowin[r] = col_in && (gjb + r >= p.bj0) && (gjb + r < p.bj1);

    issue_plane(k0 - 1, 0);
    issue_plane(k0, 1);

    const int nit = k1 - k0 + 2;
    for (int it = 0; it < nit; it++) {
        const int kk = k0 - 1 + it;
        __syncthreads();                                   // everyone is done with the stage that is refilled now
        if (kk + 2 <= k1) issue_plane(kk + 2, (it + 2) % NS);
        const bool fin = (kk - 1 >= k0) && (kk - 1 < k1);  // plane kk-1 is finished by this iteration
        if (p.mode == 0 && fin) {
#pragma unroll
            for (int r = 0; r < R; r++)
                if (rowin[r]) {
                    const double *q = p.Up + p.dof0 + 3ll * (gi + (long long)p.nx * (gjb + r + (long long)p.ny * (kk - 1)));
                    cp_async8(ups + (3 * r + 0) * NT, q + 0, true);
                    cp_async8(ups + (3 * r + 1) * NT, q + 1, true);
                    cp_async8(ups + (3 * r + 2) * NT, q + 2, true);
                }
        }
        cp_async_commit();
        const int shk = (q0 + altx * ((alty * kk) & 1)) & 1;
        int rowp[R + 2];
#pragma unroll
        for (int q = 0; q < R + 2; q++) {
            const int y = gjb - 1 + q;
            rowp[q] = (it % NS) * PLANE + (w * R + q) * kT3Row + 3 * lane + ((shk + altx * (y & 1)) & 1);
        }
        mbar_wait(&bar[it % NS], (unsigned)((it / NS) & 1));
        stencil3_plane<R, ORTHO, SLOT>(pl, rowp, A);
        cp_async_wait<0>();
        constexpr int J = 0;                               // accumulator of output plane kk-1
        const int shp = (q0 + altx * ((alty * (kk - 1)) & 1)) & 1;
#pragma unroll
        for (int r = 0; r < R; r++) {
            if (fin && rowin[r]) {
                const long long d0 = p.dof0 + 3ll * (gi + (long long)p.nx * (gjb + r + (long long)p.ny * (kk - 1)));
                // U_n of this node: still staged (plane kk-1 lives in the previous stage until the next refill)
                const double *uc = pl + ((it + NS - 1) % NS) * PLANE + (w * R + r + 1) * kT3Row + 3 * lane + 3 +
                                   ((shp + altx * ((gjb + r) & 1)) & 1);
#pragma unroll
                for (int a = 0; a < 3; a++) {
                    if (p.mode == 0) {
                        const double un = uc[a];
                        p.Un[d0 + a] = un + (cK[SLOT][273 + a] * (un - ups[(3 * r + a) * NT]) - A[J][r][a]) * cK[SLOT][270 + a];
                    } else {
                        p.Un[d0 + a] = A[J][r][a];
                    }
                }
            }
#pragma unroll
            for (int a = 0; a < 3; a++) { A[0][r][a] = A[1][r][a]; A[1][r][a] = A[2][r][a]; A[2][r][a] = 0.0; }
        }
    }
}

// ------------------------------------------------------------------------------------------
// 3-D block stencil, dominant node class, v4.  Same arithmetic and tile shape as k_stencil3_tma; what changed is
// everything around the DFMAs (profiles/r1j: only half of the issued instructions were DFMA, 0.8 barrier stalls per issue):
//   * no CTA-wide barrier in the plane loop: a stage is handed back to the TMA-issuing lanes through an "empty"
//     mbarrier (one arrival per warp), so only those lanes ever wait for the slowest warp and the others run ahead
//     up to the prefetch distance;
//   * the three accumulator planes rotate by renaming (the loop is unrolled by the period 3) instead of 72 register moves;
//   * SYM: the interior class of a homogeneous rectangular-cell lattice has a stencil that is even in every offset for
//     a == b and odd in the offsets along a and b for a != b (tensor products of the 1-D mass / stiffness / gradient
//     matrices), so 153 coefficients collapse to 36 magnitudes -> 4x fewer constant loads (the sign is a free operand
//     modifier of DFMA).  The planner checks the symmetry numerically before choosing this variant;
//   * running 64-bit offsets instead of per-row index products.
// ------------------------------------------------------------------------------------------
__host__ __device__ constexpr int stencil_sym_idx(int di, int b, int dj, int s, int a) {
    int d[3] = {di - 1, dj - 1, 1 - s};
    if (a == b) {
        for (int c = 0; c < 3; c++) d[c] = d[c] != 0 ? -1 : 0;
    } else {
        const int c = 3 - a - b;
        d[c] = d[c] != 0 ? -1 : 0;
        d[a] = -1; d[b] = -1;
    }
    return (((d[0] + 1) * 3 + b) * 3 + (d[1] + 1)) * 10 + (1 - d[2]) * 3 + a;
}
__host__ __device__ constexpr bool stencil_sym_neg(int di, int b, int dj, int s, int a) {
    if (a == b) return false;
    const int d[3] = {di - 1, dj - 1, 1 - s};
    return d[a] * d[b] < 0;
}

template <int R, bool SYM, int SLOT, int ROT>
__device__ __forceinline__ void stencil3_plane4(const double *pl, const int (&rowp)[R + 2], double (&A)[3][R][3]) {
#pragma unroll
    for (int di = 0; di < 3; di++) {
#pragma unroll
        for (int b = 0; b < 3; b++) {
            double u[R + 2];
#pragma unroll
            for (int q = 0; q < R + 2; q++) u[q] = pl[rowp[q] + 3 * di + b];
#pragma unroll
            for (int dj = 0; dj < 3; dj++)
#pragma unroll
                for (int s = 0; s < 3; s++)
#pragma unroll
                    for (int a = 0; a < 3; a++) {
                        if (!stencil_nz(di, b, dj, s, a)) continue;
                        // the coefficient stays a uniform-register / constant-bank operand; the sign rides on u
                        const double c = cK[SLOT][SYM ? stencil_sym_idx(di, b, dj, s, a) : ((di * 3 + b) * 3 + dj) * 10 + s * 3 + a];
                        if (SYM && stencil_sym_neg(di, b, dj, s, a)) {
#pragma unroll
                            for (int r = 0; r < R; r++) A[(s + ROT) % 3][r][a] = fma(-u[r + dj], c, A[(s + ROT) % 3][r][a]);
                        } else {
#pragma unroll
                            for (int r = 0; r < R; r++) A[(s + ROT) % 3][r][a] = fma(u[r + dj], c, A[(s + ROT) % 3][r][a]);
                        }
                    }
        }
    }
}

template <int NW, int R, int SLOT, bool SYM, bool NOBAR>
__global__ void __launch_bounds__(NW * 32, (NW * R <= 16) ? 3 : 2) k_stencil3_v4(const Dom3 p) {
    // NOBAR: prefetch distance 1 plane, a stage is refilled one full iteration after its release (all threads poll
    // the "empty" mbarrier, which has normally completed long before); otherwise distance 2 and a CTA barrier per plane.
    // (A wait executed by the producer lanes only would be cheaper still, but any blocking operation under a
    // thread-dependent branch makes ptxas give up the uniform datapath for the coefficient loads: LDC instead of LDCU.)
    constexpr int DIST = NOBAR ? 1 : 2;
    constexpr int TY = NW * R, TYH = TY + 2, NS = 4, NT = NW * 32;
    constexpr int PLANE = TYH * kT3Row;
    static_assert(TYH <= 32, "the TMA-issuing lanes must sit in one warp");
    extern __shared__ __align__(16) double pl[];
    double *ups_all = pl + NS * PLANE;
    uint64_t *full = reinterpret_cast<uint64_t *>(ups_all + NT * R * 3);
    uint64_t *empty = full + NS;

    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int item = blockIdx.x;
    const int txi = item % p.tiles_x; item /= p.tiles_x;
    const int tyi = item % p.tiles_y; item /= p.tiles_y;
    const int i0 = p.bi0 + txi * 32, j0 = p.bj0 + tyi * TY;
    const int k0 = p.bk0 + item * p.kz, k1 = min(k0 + p.kz, p.bk1);
    const int gi = i0 + lane, gjb = j0 + w * R;

    if (threadIdx.x == 0) {
        for (int s = 0; s < NS; s++) { mbar_init(&full[s], TYH); mbar_init(&empty[s], NW); }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();

    const int e0 = 3 * (i0 - 1);
    const int ea = max(e0, 0), eb = min(3 * (i0 + 33), 3 * p.nx);
    const int q0 = ((ea - e0) + (int)(p.dof0 & 1) + ea) & 1;            // shift parity of row y = 0, k = 0
    const int altx = p.nx & 1, alty = p.ny & 1;

    // producer lanes 0 .. TYH-1 of warp 0: one row of plane k each, into `stage`
    const int prow_y = j0 - 1 + (int)threadIdx.x;
    const bool prow_ok = (threadIdx.x < TYH) && prow_y >= 0 && prow_y < p.ny && eb > ea;
    auto issue_plane = [&](int k, int stage) {
        if (!prow_ok || k < 0 || k >= p.nz) { mbar_arrive(&full[stage]); return; }
        const long long g = p.dof0 + 3ll * p.nx * (prow_y + (long long)p.ny * k) + ea;
        const int par = (int)(g & 1);
        int len = eb - ea + par;
        len += len & 1;
        const int dsti = (ea - par - e0) + ((ea - par - e0) & 1);
        mbar_arrive_expect_tx(&full[stage], (unsigned)len * 8u);
        bulk_g2s(pl + stage * PLANE + (int)threadIdx.x * kT3Row + dsti, p.U + (g - par), (unsigned)len * 8u, &full[stage]);
    };

    double *ups = ups_all + threadIdx.x;
    double A[3][R][3];
#pragma unroll
    for (int s = 0; s < 3; s++)
#pragma unroll
        for (int r = 0; r < R; r++)
#pragma unroll
            for (int a = 0; a < 3; a++) A[s][r][a] = 0.0;

    const bool col_in = (gi >= p.bi0) && (gi < p.bi1);
    bool rowin[R];
#pragma unroll
    for (int r = 0; r < R; r++) rowin[r] = col_in && (gjb + r >= p.bj0) && (gjb + r < p.bj1);

    if (threadIdx.x < TYH) {
        issue_plane(k0 - 1, 0);
        if (DIST == 2) issue_plane(k0, 1);
    }

    const long long pstride = 3ll * p.nx * p.ny;
    const int rstride = 3 * p.nx;
    long long off = p.dof0 + 3ll * (gi + (long long)p.nx * (gjb + (long long)p.ny * (k0 - 2)));   // node (gi, gjb, kk-1)
    const int nit = k1 - k0 + 2;

    auto body = [&](const int it, auto rot_tag) {
        constexpr int ROT = decltype(rot_tag)::value;
        const int kk = k0 - 1 + it;
        if (NOBAR) {
            // stage (it+1)%NS held plane it-3, released by every warp at the end of iteration it-2
            if (it >= 3 && kk + 1 <= k1) mbar_wait(&empty[(it + 1) % NS], (unsigned)(((it - 3) / NS) & 1));
        } else {
            __syncthreads();                               // everyone is done with the stage that is refilled now
        }
        if (threadIdx.x < TYH && kk + DIST <= k1) issue_plane(kk + DIST, (it + DIST) % NS);
        const bool fin = (kk - 1 >= k0) && (kk - 1 < k1);  // plane kk-1 is finished by this iteration
        if (p.mode == 0 && fin) {
#pragma unroll
            for (int r = 0; r < R; r++)
                if (rowin[r]) {
                    const double *q = p.Up + (off + r * rstride);
                    cp_async8(ups + (3 * r + 0) * NT, q + 0, true);
                    cp_async8(ups + (3 * r + 1) * NT, q + 1, true);
                    cp_async8(ups + (3 * r + 2) * NT, q + 2, true);
                }
        }
        cp_async_commit();
        const int shk = (q0 + altx * ((alty * kk) & 1)) & 1;
        int rowp[R + 2];
#pragma unroll
        for (int q = 0; q < R + 2; q++) {
            const int y = gjb - 1 + q;
            rowp[q] = (it % NS) * PLANE + (w * R + q) * kT3Row + 3 * lane + ((shk + altx * (y & 1)) & 1);
        }
        mbar_wait(&full[it % NS], (unsigned)((it / NS) & 1));
        stencil3_plane4<R, SYM, SLOT, ROT>(pl, rowp, A);
        cp_async_wait<0>();
        constexpr int J = ROT % 3;                         // accumulator of output plane kk-1
        const int shp = (q0 + altx * ((alty * (kk - 1)) & 1)) & 1;
#pragma unroll
        for (int r = 0; r < R; r++) {
            if (fin && rowin[r]) {
                // U_n of this node: still staged (plane kk-1 lives in the previous stage until it is released below)
                const double *uc = pl + ((it + NS - 1) % NS) * PLANE + (w * R + r + 1) * kT3Row + 3 * lane + 3 +
                                   ((shp + altx * ((gjb + r) & 1)) & 1);
                double *o = p.Un + (off + r * rstride);
#pragma unroll
                for (int a = 0; a < 3; a++) {
                    if (p.mode == 0) {
                        const double un = uc[a];
                        o[a] = un + (cK[SLOT][273 + a] * (un - ups[(3 * r + a) * NT]) - A[J][r][a]) * cK[SLOT][270 + a];
                    } else {
                        o[a] = A[J][r][a];
                    }
                }
            }
#pragma unroll
            for (int a = 0; a < 3; a++) A[J][r][a] = 0.0;   // becomes the accumulator of output plane kk+2
        }
        off += pstride;
        if (NOBAR && it >= 1) {                            // hand the stage of plane kk-1 back to the producer lanes
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[(it + NS - 1) % NS]);
        }
    };
    for (int it = 0; it < nit; it += 3) {
        body(it, std::integral_constant<int, 0>{});
        if (it + 1 < nit) body(it + 1, std::integral_constant<int, 1>{});
        if (it + 2 < nit) body(it + 2, std::integral_constant<int, 2>{});
    }
}

// ------------------------------------------------------------------------------------------
// 3-D block stencil, dominant node class, separable form ("v5").  Same tiles, TMA ring and update as k_stencil3_v4;
// the 27 x (3x3) row of K is not applied entry by entry (153 DFMA per node) but direction by direction.  For the interior
// class of a homogeneous lattice of rectangular cells the row is a sum of tensor products of the 1-D stencils
// m = [1 4 1] (mass), s = [-1 2 -1] (stiffness), d = [-1 0 1] (gradient):
//     K_aa = sum_axis c[a][axis] * (s along axis, m along the two others)        K_ab = e[ab] * (d along a, d along b, m along the third)
// (tests/test_oracle_cpu.py::test_interior_hex8_stencil_is_a_sum_of_tensor_products).  The planner FITS c and e to the
// class table it assembled from the element matrices and uses this kernel only if the tensor-product table reproduces
// every entry to 1e-13 of the largest (no material or geometry formula is trusted).  Per input plane a thread
//   1. applies m, s, d along x to the three components of rows j-1 .. j+R (operands are the 9 consecutive doubles
//      u(i-1..i+1) of the staged row): 12 FP64 instructions per row, (R+2)/R rows per node;
//   2. applies m, s, d along y to those 9 fields, only the 15 products the formula needs: 23 instructions per node;
//   3. groups the results by their z operator -- A (through m_z), B (through s_z), C (through d_z) -- and scatters
//      P = A - B, 4A + 2B, +-C into the accumulators of output planes kk-1, kk, kk+1: 28 instructions per node.
// ~78 FP64 instructions per node with the update instead of 162, two accumulator planes instead of three (the plane kk+1
// accumulator is initialised, not accumulated, so it takes the registers of the plane that was just finished: the loop
// is unrolled by the period 2).  The summation order differs from the 27-point form (and from the reference's element
// loop) at the 1e-16 level per operation; parity bar 1e-10 (tests/test_gpu_parity.py).
// ------------------------------------------------------------------------------------------
struct Sep3 {
    double c[3][3];      // c[a][axis]
    double e[3];         // xy, xz, yz
    double kinv[3], km[3];
};

template <int NW, int R>
__global__ void __launch_bounds__(NW * 32, 3) k_stencil3_sep(const Dom3 p, const Sep3 cf) {
    constexpr int TY = NW * R, TYH = TY + 2, NS = 4, NT = NW * 32;
    constexpr int PLANE = TYH * kT3Row;
    static_assert(TYH <= 32, "the TMA-issuing lanes must sit in one warp");
    extern __shared__ __align__(16) double pl[];
    double *ups_all = pl + NS * PLANE;
    uint64_t *full = reinterpret_cast<uint64_t *>(ups_all + NT * R * 3);

    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int item = blockIdx.x;
    const int txi = item % p.tiles_x; item /= p.tiles_x;
    const int tyi = item % p.tiles_y; item /= p.tiles_y;
    const int i0 = p.bi0 + txi * 32, j0 = p.bj0 + tyi * TY;
    const int k0 = p.bk0 + item * p.kz, k1 = min(k0 + p.kz, p.bk1);
    const int gi = i0 + lane, gjb = j0 + w * R;

    if (threadIdx.x == 0) {
        for (int s = 0; s < NS; s++) mbar_init(&full[s], TYH);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();

    const int e0 = 3 * (i0 - 1);
    const int ea = max(e0, 0), eb = min(3 * (i0 + 33), 3 * p.nx);
    const int q0 = ((ea - e0) + (int)(p.dof0 & 1) + ea) & 1;            // shift parity of row y = 0, k = 0
    const int altx = p.nx & 1, alty = p.ny & 1;

    const int prow_y = j0 - 1 + (int)threadIdx.x;
    const bool prow_ok = (threadIdx.x < TYH) && prow_y >= 0 && prow_y < p.ny && eb > ea;
    auto issue_plane = [&](int k, int stage) {
        if (!prow_ok || k < 0 || k >= p.nz) { mbar_arrive(&full[stage]); return; }
        const long long g = p.dof0 + 3ll * p.nx * (prow_y + (long long)p.ny * k) + ea;
        const int par = (int)(g & 1);
        int len = eb - ea + par;
        len += len & 1;
        const int dsti = (ea - par - e0) + ((ea - par - e0) & 1);
        mbar_arrive_expect_tx(&full[stage], (unsigned)len * 8u);
        bulk_g2s(pl + stage * PLANE + (int)threadIdx.x * kT3Row + dsti, p.U + (g - par), (unsigned)len * 8u, &full[stage]);
    };

    double *ups = ups_all + threadIdx.x;
    double A[2][R][3];
#pragma unroll
    for (int s = 0; s < 2; s++)
#pragma unroll
        for (int r = 0; r < R; r++)
#pragma unroll
            for (int a = 0; a < 3; a++) A[s][r][a] = 0.0;

    const bool col_in = (gi >= p.bi0) && (gi < p.bi1);
    bool rowin[R];
#pragma unroll
    for (int r = 0; r < R; r++) rowin[r] = col_in && (gjb + r >= p.bj0) && (gjb + r < p.bj1);

    if (threadIdx.x < TYH) {
        issue_plane(k0 - 1, 0);
        issue_plane(k0, 1);
    }

    const long long pstride = 3ll * p.nx * p.ny;
    const int rstride = 3 * p.nx;
    long long off = p.dof0 + 3ll * (gi + (long long)p.nx * (gjb + (long long)p.ny * (k0 - 2)));   // node (gi, gjb, kk-1)
    const int nit = k1 - k0 + 2;

    auto body = [&](const int it, auto rot_tag) {
        constexpr int ROT = decltype(rot_tag)::value;     // A[ROT]: output plane kk-1 (finished here, then re-used for kk+1)
        const int kk = k0 - 1 + it;
        __syncthreads();                                   // everyone is done with the stage that is refilled now
        if (threadIdx.x < TYH && kk + 2 <= k1) issue_plane(kk + 2, (it + 2) % NS);
        const bool fin = (kk - 1 >= k0) && (kk - 1 < k1);  // plane kk-1 is finished by this iteration
        const bool finn = p.mode == 0 && (kk >= k0) && (kk < k1);   // plane kk will be finished by the next one
        const int shk = (q0 + altx * ((alty * kk) & 1)) & 1;
        const int shp = (q0 + altx * ((alty * (kk - 1)) & 1)) & 1;
        const double *plk = pl + (it % NS) * PLANE + (w * R) * kT3Row + 3 * lane;
        const double *plp = pl + ((it + NS - 1) % NS) * PLANE + (w * R + 1) * kT3Row + 3 * lane + 3;
        mbar_wait(&full[it % NS], (unsigned)((it / NS) & 1));
        cp_async_wait<0>();

        // x-operator fields of three consecutive rows: X[row % 3][component][m, s, d]
        double X[3][3][3];
#pragma unroll
        for (int q = 0; q < R + 2; q++) {
            {
                const double *s = plk + q * kT3Row + ((shk + altx * ((gjb - 1 + q) & 1)) & 1);
#pragma unroll
                for (int b = 0; b < 3; b++) {
                    const double um = s[b], u0 = s[3 + b], up = s[6 + b];
                    const double t = um + up;
                    X[q % 3][b][0] = fma(4.0, u0, t);
                    X[q % 3][b][1] = fma(2.0, u0, -t);
                    X[q % 3][b][2] = up - um;
                }
            }
            if (q < 2) continue;
            const int r = q - 2;
            const int lo = (q + 1) % 3, mi = (q + 2) % 3, hi = q % 3;
            // y operators: my / sy / dy of the x fields that the formula needs
            double tt;
#define SVL_MY(b, f) (tt = X[lo][b][f] + X[hi][b][f], fma(4.0, X[mi][b][f], tt))
#define SVL_DY(b, f) (X[hi][b][f] - X[lo][b][f])
            const double my_sx_ux = SVL_MY(0, 1), my_sx_uy = SVL_MY(1, 1), my_sx_uz = SVL_MY(2, 1);
            const double t_mx = X[lo][0][0] + X[hi][0][0];
            const double my_mx_ux = fma(4.0, X[mi][0][0], t_mx), sy_mx_ux = fma(2.0, X[mi][0][0], -t_mx);
            const double t_my = X[lo][1][0] + X[hi][1][0];
            const double my_mx_uy = fma(4.0, X[mi][1][0], t_my), sy_mx_uy = fma(2.0, X[mi][1][0], -t_my);
            const double t_mz = X[lo][2][0] + X[hi][2][0];
            const double my_mx_uz = fma(4.0, X[mi][2][0], t_mz), sy_mx_uz = fma(2.0, X[mi][2][0], -t_mz);
            const double dy_mx_uy = SVL_DY(1, 0), dy_mx_uz = SVL_DY(2, 0);
            const double dy_dx_ux = SVL_DY(0, 2), dy_dx_uy = SVL_DY(1, 2);
            const double my_dx_ux = SVL_MY(0, 2), my_dx_uz = SVL_MY(2, 2);
#undef SVL_MY
#undef SVL_DY
            // grouped by z operator: A through m_z, B through s_z, C through d_z
            const double Ax = fma(cf.c[0][0], my_sx_ux, fma(cf.c[0][1], sy_mx_ux, cf.e[0] * dy_dx_uy));
            const double Ay = fma(cf.c[1][0], my_sx_uy, fma(cf.c[1][1], sy_mx_uy, cf.e[0] * dy_dx_ux));
            const double Az = fma(cf.c[2][0], my_sx_uz, cf.c[2][1] * sy_mx_uz);
            const double Px = fma(-cf.c[0][2], my_mx_ux, Ax), Py = fma(-cf.c[1][2], my_mx_uy, Ay), Pz = fma(-cf.c[2][2], my_mx_uz, Az);
            // output plane kk-1: last contribution (+P + C), then the CentralDifference update (CentralDifference.cpp:138-148)
            double F[3];
            F[0] = fma(cf.e[1], my_dx_uz, A[ROT][r][0] + Px);
            F[1] = fma(cf.e[2], dy_mx_uz, A[ROT][r][1] + Py);
            F[2] = fma(cf.e[2], dy_mx_uy, fma(cf.e[1], my_dx_ux, A[ROT][r][2] + Pz));
            // output plane kk: 4 A + 2 B
            A[ROT ^ 1][r][0] = fma(2.0 * cf.c[0][2], my_mx_ux, fma(4.0, Ax, A[ROT ^ 1][r][0]));
            A[ROT ^ 1][r][1] = fma(2.0 * cf.c[1][2], my_mx_uy, fma(4.0, Ay, A[ROT ^ 1][r][1]));
            A[ROT ^ 1][r][2] = fma(2.0 * cf.c[2][2], my_mx_uz, fma(4.0, Az, A[ROT ^ 1][r][2]));
            // output plane kk+1: first contribution (+P - C), into the registers of the plane that was just finished
            A[ROT][r][0] = fma(-cf.e[1], my_dx_uz, Px);
            A[ROT][r][1] = fma(-cf.e[2], dy_mx_uz, Py);
            A[ROT][r][2] = fma(-cf.e[2], dy_mx_uy, fma(-cf.e[1], my_dx_ux, Pz));
            if (fin && rowin[r]) {
                // U_n of this node: still staged (plane kk-1 lives in the previous stage until the next refill)
                const double *uc = plp + r * kT3Row + ((shp + altx * ((gjb + r) & 1)) & 1);
                double *o = p.Un + (off + r * rstride);
#pragma unroll
                for (int a = 0; a < 3; a++) {
                    if (p.mode == 0) {
                        const double un = uc[a];
                        o[a] = un + (cf.km[a] * (un - ups[(3 * r + a) * NT]) - F[a]) * cf.kinv[a];
                    } else {
                        o[a] = F[a];
                    }
                }
            }
            // U_{n-1} of the node one plane up (finished by the next iteration) into the slots this thread has just read:
            // the slots are private to the thread, so program order is all the ordering the prefetch needs, and it gets a
            // whole plane of compute to cover the DRAM latency
            if (finn && rowin[r]) {
                const double *g = p.Up + (off + pstride + r * rstride);
                cp_async8m(ups + (3 * r + 0) * NT, g + 0);
                cp_async8m(ups + (3 * r + 1) * NT, g + 1);
                cp_async8m(ups + (3 * r + 2) * NT, g + 2);
            }
        }
        cp_async_commit();
        off += pstride;
    };
    for (int it = 0; it < nit; it += 2) {
        body(it, std::integral_constant<int, 0>{});
        if (it + 1 < nit) body(it + 1, std::integral_constant<int, 1>{});
    }
}

// ------------------------------------------------------------------------------------------
// 3-D block stencil, remaining node classes (faces / edges / corners / material interfaces):
// one thread per listed node, coefficients from the per-class table (warp-uniform -> L1
// broadcast because the list is sorted by class), neighbours through L1/L2.
// ------------------------------------------------------------------------------------------
struct Gat3 {
    const double *U, *Up;
    double *Un;
    const uint8_t *cls;
    const double *tbl;      // [ncls][276]
    const int32_t *list;    // lattice-local node ids, sorted by class
    long long dof0;
    int n, nx, ny, nz, mode;
    const int32_t *target;  // halo pass: interface index of each listed node (else null)
    double *hF;             // halo pass: [n_if][3] partial internal force
};
__global__ void __launch_bounds__(128) k_stencil3_gather(const Gat3 p) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= p.n) return;
    const int q = p.list[t];
    const int i = q % p.nx, j = (q / p.nx) % p.ny, k = q / (p.nx * p.ny);
    const double *T = p.tbl + (size_t)p.cls[q] * kTbl3Stride;
    double F[3] = {0.0, 0.0, 0.0};
    for (int s = 0; s < 3; s++) {
        const int z = k + 1 - s;                  // slot s <-> neighbour plane offset dk = 1 - s
        if (z < 0 || z >= p.nz) continue;
        for (int dj = 0; dj < 3; dj++) {
            const int y = j + dj - 1;
            if (y < 0 || y >= p.ny) continue;
#pragma unroll
            for (int di = 0; di < 3; di++) {
                const int x = i + di - 1;
                if (x < 0 || x >= p.nx) continue;
                const double *u = p.U + p.dof0 + 3ll * (x + (long long)p.nx * (y + (long long)p.ny * z));
#pragma unroll
                for (int b = 0; b < 3; b++) {
                    const double ub = u[b];
                    const double *c = T + ((di * 3 + b) * 3 + dj) * 10 + s * 3;
                    F[0] = fma(c[0], ub, F[0]);
                    F[1] = fma(c[1], ub, F[1]);
                    F[2] = fma(c[2], ub, F[2]);
                }
            }
        }
    }
    if (p.hF) {
        double *o = p.hF + 3ll * p.target[t];
        o[0] = F[0]; o[1] = F[1]; o[2] = F[2];
        return;
    }
    const long long d0 = p.dof0 + 3ll * q;
#pragma unroll
    for (int a = 0; a < 3; a++) {
        if (p.mode == 0) {
            const double un = p.U[d0 + a];
            p.Un[d0 + a] = un + (T[273 + a] * (un - p.Up[d0 + a]) - F[a]) * T[270 + a];
        } else {
            p.Un[d0 + a] = F[a];
        }
    }
}

// Shell classes, step pass: the list is cut into chunks of kShellChunk nodes of ONE class (padded with -1 by the
// planner), one CTA per chunk.  The class table is staged in shared memory once and every thread advances kShellNPT
// nodes, so a coefficient is read once per kShellNPT DFMAs instead of once per DFMA (the per-node kernel above is
// bound by those loads: profiles/r1j).  Same neighbour and summation order as k_stencil3_gather.
struct Shell3 {
    const double *U, *Up;
    double *Un;
    const double *tbl;        // [ncls][276]
    const int32_t *list;      // [n_chunks][kShellChunk] lattice-local node ids or -1
    const uint8_t *chunk_cls; // [n_chunks]
    long long dof0;
    int nx, ny, nz, mode;
    const int32_t *target;    // halo pass: interface index of each list entry (else null)
    double *hF;               // halo pass: [n_if][3] partial internal force
};
template <bool LOWREG>
__global__ void __launch_bounds__(128, LOWREG ? 8 : 6) k_stencil3_shell(const Shell3 p) {
    __shared__ double T[kTbl3Stride];
    __shared__ unsigned slot_mask;       // bit (s * 3 + dj) * 3 + di: the class couples to that neighbour at all
    {
        const double *Tg = p.tbl + (size_t)p.chunk_cls[blockIdx.x] * kTbl3Stride;
        for (int i = threadIdx.x; i < kTbl3Stride; i += 128) T[i] = Tg[i];
    }
    __syncthreads();
    if (threadIdx.x < 32) {
        // boundary classes have no neighbour on one or more sides (9 of 27 slots on a face, 15 on an edge, 19 at a corner):
        // their blocks are exactly zero, so the CTA skips them -- indices, loads and DFMAs (adding 0 * u changes nothing)
        bool nz = false;
        if (threadIdx.x < 27) {
            const int s = threadIdx.x / 9, dj = (threadIdx.x / 3) % 3, di = threadIdx.x % 3;
#pragma unroll
            for (int b = 0; b < 3; b++)
#pragma unroll
                for (int a = 0; a < 3; a++) nz = nz || T[((di * 3 + b) * 3 + dj) * 10 + s * 3 + a] != 0.0;
        }
        const unsigned mk = __ballot_sync(0xffffffffu, nz);
        if (threadIdx.x == 0) slot_mask = mk;
    }
    __syncthreads();
    const unsigned mask = slot_mask;
    int q[kShellNPT], ci[kShellNPT], cj[kShellNPT], ck[kShellNPT];
    double F[kShellNPT][3];
    // a class whose three dofs are restrained (1 / Keff == 0: e.g. the fixed base of a soil box) keeps its displacement:
    // the update below would add (..) * 0 to U_n, so the 27-point row is not evaluated at all in the step pass
    if (p.mode == 0 && !p.hF && T[270] == 0.0 && T[271] == 0.0 && T[272] == 0.0) {
#pragma unroll
        for (int n = 0; n < kShellNPT; n++) {
            const int qq = p.list[(size_t)blockIdx.x * kShellChunk + n * 128 + threadIdx.x];
            if (qq < 0) continue;
            const long long d0 = p.dof0 + 3ll * qq;
            p.Un[d0] = p.U[d0]; p.Un[d0 + 1] = p.U[d0 + 1]; p.Un[d0 + 2] = p.U[d0 + 2];
        }
        return;
    }
#pragma unroll
    for (int n = 0; n < kShellNPT; n++) {
        q[n] = p.list[(size_t)blockIdx.x * kShellChunk + n * 128 + threadIdx.x];
        const int qq = max(q[n], 0);
        ci[n] = qq % p.nx; cj[n] = (qq / p.nx) % p.ny; ck[n] = qq / (p.nx * p.ny);
        if (q[n] < 0) ci[n] = -4;                       // padding: every neighbour is "outside"
        F[n][0] = F[n][1] = F[n][2] = 0.0;
    }
    for (int s = 0; s < 3; s++) {
        for (int dj = 0; dj < 3; dj++) {
            if (!((mask >> ((s * 3 + dj) * 3)) & 7u)) continue;
#pragma unroll
            for (int di = 0; di < 3; di++) {
                if (!((mask >> ((s * 3 + dj) * 3 + di)) & 1u)) continue;
                const double *u[kShellNPT];
#pragma unroll
                for (int n = 0; n < kShellNPT; n++) {
                    const int x = ci[n] + di - 1, y = cj[n] + dj - 1, z = ck[n] + 1 - s;
                    const bool ok = x >= 0 && x < p.nx && y >= 0 && y < p.ny && z >= 0 && z < p.nz;
                    u[n] = ok ? p.U + p.dof0 + 3ll * (x + (long long)p.nx * (y + (long long)p.ny * z)) : nullptr;
                }
#pragma unroll
                for (int b = 0; b < 3; b++) {
                    const double *c = T + ((di * 3 + b) * 3 + dj) * 10 + s * 3;
                    const double c0 = c[0], c1 = c[1], c2 = c[2];
#pragma unroll
                    for (int n = 0; n < kShellNPT; n++) {
                        const double ub = u[n] ? u[n][b] : 0.0;
                        F[n][0] = fma(c0, ub, F[n][0]);
                        F[n][1] = fma(c1, ub, F[n][1]);
                        F[n][2] = fma(c2, ub, F[n][2]);
                    }
                }
            }
        }
    }
    if (p.hF) {
#pragma unroll
        for (int n = 0; n < kShellNPT; n++) {
            if (q[n] < 0) continue;
            double *o = p.hF + 3ll * p.target[(size_t)blockIdx.x * kShellChunk + n * 128 + threadIdx.x];
            o[0] = F[n][0]; o[1] = F[n][1]; o[2] = F[n][2];
        }
        return;
    }
#pragma unroll
    for (int n = 0; n < kShellNPT; n++) {
        if (q[n] < 0) continue;
        const long long d0 = p.dof0 + 3ll * q[n];
#pragma unroll
        for (int a = 0; a < 3; a++) {
            if (p.mode == 0) {
                const double un = p.U[d0 + a];
                p.Un[d0 + a] = un + (T[273 + a] * (un - p.Up[d0 + a]) - F[n][a]) * T[270 + a];
            } else {
                p.Un[d0 + a] = F[n][a];
            }
        }
    }
}

// halo pass of a 2-D lattice: partial internal force of the listed (interface) nodes
struct Gat2 {
    const double *U;
    const uint8_t *cls;
    const double *tbl;
    const int32_t *list, *target;
    double *hF;
    long long dof0;
    int n, nx, ny;
};
__global__ void __launch_bounds__(128) k_stencil2_gather(const Gat2 p) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= p.n) return;
    const int q = p.list[t];
    const int i = q % p.nx, j = q / p.nx;
    const double *T = p.tbl + (size_t)p.cls[q] * kTbl2Stride;
    const double2 *U2 = reinterpret_cast<const double2 *>(p.U + p.dof0);
    double f0 = 0.0, f1 = 0.0;
    for (int dj = -1; dj <= 1; dj++)
        for (int di = -1; di <= 1; di++) {
            const int x = i + di, y = j + dj;
            double2 u = make_double2(0.0, 0.0);
            if (x >= 0 && x < p.nx && y >= 0 && y < p.ny) u = U2[x + (long long)p.nx * y];
            const double *k = T + ((dj + 1) * 3 + (di + 1)) * 4;
            f0 = fma(k[0], u.x, f0); f0 = fma(k[1], u.y, f0);
            f1 = fma(k[2], u.x, f1); f1 = fma(k[3], u.y, f1);
        }
    p.hF[2ll * p.target[t]] = f0;
    p.hF[2ll * p.target[t] + 1] = f1;
}

// ------------------------------------------------------------------------------------------
// 2-D block stencil (quad4 lattices): one thread per node, neighbours through L1/L2
// ------------------------------------------------------------------------------------------
struct Blk2 {
    const double *U, *Up;
    double *Un;
    const uint8_t *cls;
    const double *tbl;     // [ncls][40]: 9 neighbours x (2x2) + kinv[2] + km[2]
    long long dof0;
    int ncls, nx, ny, mode;
};

__global__ void __launch_bounds__(256) k_stencil2(const Blk2 p) {
    extern __shared__ __align__(16) double sm[];
    for (int t = threadIdx.x; t < p.ncls * kTbl2Stride; t += blockDim.x) sm[t] = p.tbl[t];
    __syncthreads();
    const int i = blockIdx.x * 64 + (threadIdx.x & 63);
    const int j = blockIdx.y * 4 + (threadIdx.x >> 6);
    if (i >= p.nx || j >= p.ny) return;
    const int c = p.cls[i + (long long)p.nx * j];
    if (!c) return;
    const double *t = sm + c * kTbl2Stride;
    const double2 *U2 = reinterpret_cast<const double2 *>(p.U + p.dof0);
    double f0 = 0.0, f1 = 0.0;
    double2 uc = make_double2(0.0, 0.0);
#pragma unroll
    for (int dj = -1; dj <= 1; dj++)
#pragma unroll
        for (int di = -1; di <= 1; di++) {
            const int x = i + di, y = j + dj;
            double2 u = make_double2(0.0, 0.0);
            if (x >= 0 && x < p.nx && y >= 0 && y < p.ny) u = U2[x + (long long)p.nx * y];
            if (di == 0 && dj == 0) uc = u;
            const double *k = t + ((dj + 1) * 3 + (di + 1)) * 4;
            f0 = fma(k[0], u.x, f0); f0 = fma(k[1], u.y, f0);
            f1 = fma(k[2], u.x, f1); f1 = fma(k[3], u.y, f1);
        }
    const long long n = i + (long long)p.nx * j;
    double2 *out = reinterpret_cast<double2 *>(p.Un + p.dof0) + n;
    if (p.mode == 0) {
        const double2 up = reinterpret_cast<const double2 *>(p.Up + p.dof0)[n];
        double2 r;
        r.x = uc.x + (t[38] * (uc.x - up.x) - f0) * t[36];
        r.y = uc.y + (t[39] * (uc.y - up.y) - f1) * t[37];
        *out = r;
    } else {
        *out = make_double2(f0, f1);
    }
}

// ------------------------------------------------------------------------------------------
// generic Gauss-point element force: one thread per Gauss point
// ------------------------------------------------------------------------------------------
struct GenArgs {
    int n;                     // elements in the set
    const int32_t *conn;       // [n][npe]
    const int32_t *mat;        // [n]
    const double *th;          // [n] (quad)
    const int32_t *matkind;    // per material
    const double *matpar;      // [nmat][8]
    const double *coords;      // [nnodes][ndim]
    const int32_t *node_ptr;   // internal dof0 per node
    const double *U;
    double *fe;                // [n][npe][ndofn]
    double *state;             // [13][n*ngp] or null
    double *gp;                // [2][ncomp][n*ngp] strain|stress or null
    int commit;                // 1: store the updated J2 state
    const int32_t *ecls;       // [n] class of each element (TAB kernels)
    const double *gtab;        // [ncls][ngp][npe*ndim + 1]: shape-function gradients + w|J| per Gauss point
};

__device__ __forceinline__ double shfl_xor_d(double v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }

// TAB: congruent elements share one table of shape-function gradients and w|J| per Gauss point (what the
// reference recomputes from the node coordinates 48 + 32 times per element and step);
// otherwise (distorted meshes: every element its own class) the gradients come from the coordinates.
template <bool TAB>
__global__ void __launch_bounds__(128) k_gen_hex8(const GenArgs a) {
    __shared__ double sx[4][4][8][TAB ? 3 : 6];  // [warp][elem in warp][node][(X(3)) U(3)]
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int el = lane >> 3, g = lane & 7;
    int e = tid >> 3;
    const bool act = e < a.n;
    if (!act) e = a.n - 1;
    {
        const int node = a.conn[(long long)e * 8 + g];
        const double *u = a.U + a.node_ptr[node];
        double *s = sx[w][el][g];
        if (!TAB) {
            const double *x = a.coords + 3ll * node;
            s[0] = x[0]; s[1] = x[1]; s[2] = x[2]; s[3] = u[0]; s[4] = u[1]; s[5] = u[2];
        } else {
            s[0] = u[0]; s[1] = u[1]; s[2] = u[2];
        }
    }
    __syncwarp();
    double Ue[8][3];
    double d[8][3];
    double wd;
    if (!TAB) {
        double X[8][3];
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const double *s = sx[w][el][i];
            X[i][0] = s[0]; X[i][1] = s[1]; X[i][2] = s[2];
            Ue[i][0] = s[3]; Ue[i][1] = s[4]; Ue[i][2] = s[5];
        }
        wd = hex8_grad(X, g, d, nullptr);                 // weight 1 * |det J|
    } else {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const double *s = sx[w][el][i];
            Ue[i][0] = s[0]; Ue[i][1] = s[1]; Ue[i][2] = s[2];
        }
        const double *t = a.gtab + ((size_t)a.ecls[e] * 8 + g) * 25;
#pragma unroll
        for (int i = 0; i < 8; i++) { d[i][0] = __ldg(t + 3 * i); d[i][1] = __ldg(t + 3 * i + 1); d[i][2] = __ldg(t + 3 * i + 2); }
        wd = __ldg(t + 24);
    }
    double ep[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int i = 0; i < 8; i++) {                         // eps = B u  (lin3DHexa8.cpp:721-741,845-850)
        ep[0] = fma(d[i][0], Ue[i][0], ep[0]);
        ep[1] = fma(d[i][1], Ue[i][1], ep[1]);
        ep[2] = fma(d[i][2], Ue[i][2], ep[2]);
        ep[3] += d[i][1] * Ue[i][0] + d[i][0] * Ue[i][1];
        ep[4] += d[i][2] * Ue[i][1] + d[i][1] * Ue[i][2];
        ep[5] += d[i][2] * Ue[i][0] + d[i][0] * Ue[i][2];
    }
    const int mi = a.mat[e];
    const double *mp = a.matpar + 8 * mi;
    double sg[6];
    const long long ngp = 8ll * a.n, q = 8ll * e + g;
    if (a.matkind[mi] == SVLGPU_PLASTIC3DJ2) {
        J2Par jp = {mp[0], mp[1], mp[3], mp[4], mp[5]};
        double st[13];
#pragma unroll
        for (int i = 0; i < 13; i++) st[i] = a.state[i * ngp + q];
        const bool yielded = j2_return_map(jp, ep, st, sg);
        if (a.commit && act && yielded) {          // an elastic step leaves the state untouched: no write
#pragma unroll
            for (int i = 0; i < 13; i++) a.state[i * ngp + q] = st[i];
        }
    } else {
        iso_stress3(iso_from_E_nu(mp[0], mp[1]), ep, sg);
    }
    if (a.gp && act) {
#pragma unroll
        for (int i = 0; i < 6; i++) { a.gp[i * ngp + q] = ep[i]; a.gp[(6 + i) * ngp + q] = sg[i]; }
    }
    // f_gp = w |J| B^T sigma  (lin3DHexa8.cpp:408), then reduce-scatter over the 8 Gauss points:
    // after the three exchanges lane g holds the 3 force components of node g.
    double f[8][3];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        f[i][0] = wd * (d[i][0] * sg[0] + d[i][1] * sg[3] + d[i][2] * sg[5]);
        f[i][1] = wd * (d[i][1] * sg[1] + d[i][0] * sg[3] + d[i][2] * sg[4]);
        f[i][2] = wd * (d[i][2] * sg[2] + d[i][1] * sg[4] + d[i][0] * sg[5]);
    }
    double h4[4][3], h2[2][3], h1[3];
    const bool b4 = g & 4, b2 = g & 2, b1 = g & 1;
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const double keep = b4 ? f[4 + i][c] : f[i][c];
            const double send = b4 ? f[i][c] : f[4 + i][c];
            h4[i][c] = keep + shfl_xor_d(send, 4);
        }
#pragma unroll
    for (int i = 0; i < 2; i++)
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const double keep = b2 ? h4[2 + i][c] : h4[i][c];
            const double send = b2 ? h4[i][c] : h4[2 + i][c];
            h2[i][c] = keep + shfl_xor_d(send, 2);
        }
#pragma unroll
    for (int c = 0; c < 3; c++) {
        const double keep = b1 ? h2[1][c] : h2[0][c];
        const double send = b1 ? h2[0][c] : h2[1][c];
        h1[c] = keep + shfl_xor_d(send, 1);
    }
    if (act) {
        double *o = a.fe + 3ll * q;
        o[0] = h1[0]; o[1] = h1[1]; o[2] = h1[2];
    }
}

// Class-table variant with a small register footprint (high occupancy hides the conn -> U gather latency):
// phase 1, lane = Gauss point g: eps = B_g u_e, material update, w|J| sigma_g staged in shared memory;
// phase 2, lane = node i: f_i = sum_g B_g,i^T (w|J| sigma_g), Gauss points in the reference's order (lin3DHexa8.cpp:397-409).
__global__ void __launch_bounds__(128, 5) k_gen_hex8_tab(const GenArgs a) {
    __shared__ double su[4][4][8][3];            // [warp][elem in warp][node][U(3)]
    __shared__ double ss[4][4][8][6];            // [warp][elem in warp][gp][w|J| sigma(6)]
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int el = lane >> 3, g = lane & 7;
    int e = tid >> 3;
    const bool act = e < a.n;
    if (!act) e = a.n - 1;
    {
        const int node = a.conn[(long long)e * 8 + g];
        const double *u = a.U + a.node_ptr[node];
        double *s = su[w][el][g];
        s[0] = u[0]; s[1] = u[1]; s[2] = u[2];
    }
    __syncwarp();
    const double *tab = a.gtab + (size_t)a.ecls[e] * 200;
    const double *t = tab + g * 25;
    double ep[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int i = 0; i < 8; i++) {                         // eps = B u  (lin3DHexa8.cpp:721-741,845-850)
        const double d0 = __ldg(t + 3 * i), d1 = __ldg(t + 3 * i + 1), d2 = __ldg(t + 3 * i + 2);
        const double *u = su[w][el][i];
        const double u0 = u[0], u1 = u[1], u2 = u[2];
        ep[0] = fma(d0, u0, ep[0]);
        ep[1] = fma(d1, u1, ep[1]);
        ep[2] = fma(d2, u2, ep[2]);
        ep[3] += d1 * u0 + d0 * u1;
        ep[4] += d2 * u1 + d1 * u2;
        ep[5] += d2 * u0 + d0 * u2;
    }
    const int mi = a.mat[e];
    const double *mp = a.matpar + 8 * mi;
    double sg[6];
    const long long ngp = 8ll * a.n, q = 8ll * e + g;
    if (a.matkind[mi] == SVLGPU_PLASTIC3DJ2) {
        J2Par jp = {mp[0], mp[1], mp[3], mp[4], mp[5]};
        double st[13];
#pragma unroll
        for (int i = 0; i < 13; i++) st[i] = a.state[i * ngp + q];
        const bool yielded = j2_return_map(jp, ep, st, sg);
        if (a.commit && act && yielded) {          // an elastic step leaves the state untouched: no write
#pragma unroll
            for (int i = 0; i < 13; i++) a.state[i * ngp + q] = st[i];
        }
    } else {
        iso_stress3(iso_from_E_nu(mp[0], mp[1]), ep, sg);
    }
    if (a.gp && act) {
#pragma unroll
        for (int i = 0; i < 6; i++) { a.gp[i * ngp + q] = ep[i]; a.gp[(6 + i) * ngp + q] = sg[i]; }
    }
    {
        const double wd = __ldg(t + 24);
        double *s = ss[w][el][g];
#pragma unroll
        for (int i = 0; i < 6; i++) s[i] = wd * sg[i];
    }
    __syncwarp();
    // phase 2: this lane is node i = g of its element
    double f0 = 0.0, f1 = 0.0, f2 = 0.0;
#pragma unroll
    for (int gp = 0; gp < 8; gp++) {
        const double *d = tab + gp * 25 + 3 * g;
        const double d0 = __ldg(d), d1 = __ldg(d + 1), d2 = __ldg(d + 2);
        const double *s = ss[w][el][gp];
        f0 += d0 * s[0] + d1 * s[3] + d2 * s[5];
        f1 += d1 * s[1] + d0 * s[3] + d2 * s[4];
        f2 += d2 * s[2] + d1 * s[4] + d0 * s[5];
    }
    if (act) {
        double *o = a.fe + 3ll * q;
        o[0] = f0; o[1] = f1; o[2] = f2;
    }
}

template <bool TAB>
__global__ void __launch_bounds__(128) k_gen_quad4(const GenArgs a) {
    __shared__ double sx[4][8][4][4];            // [warp][elem in warp][node][X(2) U(2)]
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int el = lane >> 2, g = lane & 3;
    int e = tid >> 2;
    const bool act = e < a.n;
    if (!act) e = a.n - 1;
    {
        const int node = a.conn[(long long)e * 4 + g];
        const double *u = a.U + a.node_ptr[node];
        double *s = sx[w][el][g];
        if (!TAB) { const double *x = a.coords + 2ll * node; s[0] = x[0]; s[1] = x[1]; }
        s[2] = u[0]; s[3] = u[1];
    }
    __syncwarp();
    double X[4][2], Ue[4][2];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const double *s = sx[w][el][i];
        if (!TAB) { X[i][0] = s[0]; X[i][1] = s[1]; }
        Ue[i][0] = s[2]; Ue[i][1] = s[3];
    }
    double d[4][2];
    double wd;
    if (!TAB) wd = a.th[e] * quad4_grad(X, g, d, nullptr);
    else {
        const double *t = a.gtab + ((size_t)a.ecls[e] * 4 + g) * 9;
#pragma unroll
        for (int i = 0; i < 4; i++) { d[i][0] = __ldg(t + 2 * i); d[i][1] = __ldg(t + 2 * i + 1); }
        wd = __ldg(t + 8);                              // thickness * |det J|
    }
    double ep[3] = {0, 0, 0};
#pragma unroll
    for (int i = 0; i < 4; i++) {                // lin2DQuad4.cpp:705-708
        ep[0] = fma(d[i][0], Ue[i][0], ep[0]);
        ep[1] = fma(d[i][1], Ue[i][1], ep[1]);
        ep[2] += d[i][1] * Ue[i][0] + d[i][0] * Ue[i][1];
    }
    const int mi = a.mat[e];
    const double *mp = a.matpar + 8 * mi;
    double sg[3];
    const long long ngp = 4ll * a.n, q = 4ll * e + g;
    if (a.matkind[mi] == SVLGPU_PLASTICPLANESTRAINJ2) {
        // PlasticPlaneStrainJ2.cpp:227-278: the 3-D return map on the embedded tensor [0, e11, e22, 0, e12/2, 0] (:235),
        // stress read back from slots 1, 2, 4 (:118-122); the 13 state values keep that embedding
        J2Par jp = {mp[0], mp[1], mp[3], mp[4], mp[5]};
        const double e6[6] = {0.0, ep[0], ep[1], 0.0, ep[2], 0.0};
        double st[13], s6[6];
#pragma unroll
        for (int i = 0; i < 13; i++) st[i] = a.state[i * ngp + q];
        const bool yielded = j2_return_map(jp, e6, st, s6);
        if (a.commit && act && yielded) {
#pragma unroll
            for (int i = 0; i < 13; i++) a.state[i * ngp + q] = st[i];
        }
        sg[0] = s6[1]; sg[1] = s6[2]; sg[2] = s6[4];
    } else {
        iso_stress2(iso_from_E_nu(mp[0], mp[1]), ep, sg);
    }
    if (a.gp && act) {
#pragma unroll
        for (int i = 0; i < 3; i++) { a.gp[i * ngp + q] = ep[i]; a.gp[(3 + i) * ngp + q] = sg[i]; }
    }
    double f[4][2];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        f[i][0] = wd * (d[i][0] * sg[0] + d[i][1] * sg[2]);
        f[i][1] = wd * (d[i][1] * sg[1] + d[i][0] * sg[2]);
    }
    double h2[2][2], h1[2];
    const bool b2 = g & 2, b1 = g & 1;
#pragma unroll
    for (int i = 0; i < 2; i++)
#pragma unroll
        for (int c = 0; c < 2; c++) {
            const double keep = b2 ? f[2 + i][c] : f[i][c];
            const double send = b2 ? f[i][c] : f[2 + i][c];
            h2[i][c] = keep + shfl_xor_d(send, 2);
        }
#pragma unroll
    for (int c = 0; c < 2; c++) {
        const double keep = b1 ? h2[1][c] : h2[0][c];
        const double send = b1 ? h2[0][c] : h2[1][c];
        h1[c] = keep + shfl_xor_d(send, 1);
    }
    if (act) {
        double *o = a.fe + 2ll * q;
        o[0] = h1[0]; o[1] = h1[1];
    }
}

// ------------------------------------------------------------------------------------------
// generic nodes: gather element contributions in ascending element order, then update
// ------------------------------------------------------------------------------------------
struct GNodeArgs {
    int n;
    const int32_t *dof0, *ndof, *ptr;
    const long long *slot;
    const double *fe;
    const double *U, *Up, *kinv, *km;
    double *Un;
    int mode;
    const int32_t *target;   // halo pass: interface index per listed node (dof0 unused then)
    double *hF;
    int hnd;
};
__global__ void __launch_bounds__(256) k_gen_nodes(const GNodeArgs a) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= a.n) return;
    const int d0 = a.hF ? 0 : a.dof0[t], nd = a.ndof[t];
    double F[9];
#pragma unroll
    for (int c = 0; c < 9; c++) F[c] = 0.0;
    for (int q = a.ptr[t]; q < a.ptr[t + 1]; q++) {
        const double *f = a.fe + a.slot[q];
#pragma unroll
        for (int c = 0; c < 9; c++)
            if (c < nd) F[c] += f[c];
    }
    if (a.hF) {
        for (int c = 0; c < a.hnd; c++) a.hF[(long long)a.target[t] * a.hnd + c] = F[c];
        return;
    }
#pragma unroll
    for (int c = 0; c < 9; c++)
        if (c < nd) {
            if (a.mode == 0) {
                const double un = a.U[d0 + c];
                const double du = (a.km[d0 + c] * (un - a.Up[d0 + c]) - F[c]) * a.kinv[d0 + c];
                a.Un[d0 + c] = un + du;
            } else {
                a.Un[d0 + c] = F[c];
            }
        }
}

// ------------------------------------------------------------------------------------------
// Non-lattice nodes with a repeating row of K (planner section D2): pre-summed row blocks per node class + explicit
// neighbour list per node.  One CTA = one chunk of kNbrChunk nodes of ONE class: the class table is staged in shared
// memory, every thread advances kNbrNPT nodes (one coefficient load per kNbrNPT DFMAs), the neighbour indices of a slot
// are stored slot-major so that a warp reads them coalesced; displacements come through L1 / L2.  243 DFMA per hex8
// node instead of ~2 800 in the Gauss-point kernels, and no element-force arena (192 B written + read per element).
// ------------------------------------------------------------------------------------------
struct NbrArgs {
    const double *U, *Up;
    double *Un;
    const double *tbl;
    const int32_t *cls_nn, *chunk_cls, *dof0, *nbr;
    const long long *chunk_off;
    int stride, mode;
};
template <int ND, int UNR>
__global__ void __launch_bounds__(128, 6) k_nbr_nodes(const NbrArgs p) {
    __shared__ double T[kNbrSlots * ND * ND + 2 * ND];
    const int c = p.chunk_cls[blockIdx.x];
    const int nn = p.cls_nn[c];
    {
        const double *Tg = p.tbl + (size_t)c * p.stride;
        for (int i = threadIdx.x; i < nn * ND * ND; i += 128) T[i] = Tg[i];
        if (threadIdx.x < 2 * ND) T[kNbrSlots * ND * ND + threadIdx.x] = Tg[kNbrSlots * ND * ND + threadIdx.x];
    }
    __syncthreads();
    int d0[kNbrNPT];
    double F[kNbrNPT][ND];
#pragma unroll
    for (int n = 0; n < kNbrNPT; n++) {
        d0[n] = p.dof0[(size_t)blockIdx.x * kNbrChunk + n * 128 + threadIdx.x];
#pragma unroll
        for (int a = 0; a < ND; a++) F[n][a] = 0.0;
    }
    const int32_t *nb = p.nbr + p.chunk_off[blockIdx.x] + threadIdx.x;
    // UNR slots per trip: the index loads of the next slots go out before the displacement loads of this one return (the kernel
    // is bound by gather latency, profiles/r3n: 25 % DRAM, 15 % FP64, L1 hit rate 86 %)
#pragma unroll UNR
    for (int s = 0; s < nn; s++) {
        int idx[kNbrNPT];
#pragma unroll
        for (int n = 0; n < kNbrNPT; n++) idx[n] = nb[(size_t)s * kNbrChunk + n * 128];
#pragma unroll
        for (int b = 0; b < ND; b++) {
            double cf[ND];
#pragma unroll
            for (int a = 0; a < ND; a++) cf[a] = T[(s * ND + b) * ND + a];
#pragma unroll
            for (int n = 0; n < kNbrNPT; n++) {
                const double ub = idx[n] >= 0 ? p.U[idx[n] + b] : 0.0;
#pragma unroll
                for (int a = 0; a < ND; a++) F[n][a] = fma(cf[a], ub, F[n][a]);
            }
        }
    }
#pragma unroll
    for (int n = 0; n < kNbrNPT; n++) {
        if (d0[n] < 0) continue;
#pragma unroll
        for (int a = 0; a < ND; a++) {
            if (p.mode == 0) {
                const double un = p.U[d0[n] + a];
                p.Un[d0[n] + a] = un + (T[kNbrSlots * ND * ND + ND + a] * (un - p.Up[d0[n] + a]) - F[n][a]) * T[kNbrSlots * ND * ND + a];
            } else {
                p.Un[d0[n] + a] = F[n][a];
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// nodal loads (Assembler.cpp:316-350 point loads): U_{n+1}[d] += kinv[d] * sum_l coef_l amp_l(k)
// ------------------------------------------------------------------------------------------
struct PLArgs {
    int n;                        // loaded dofs
    const int32_t *dof, *ptr;     // CSR over loaded dofs
    const int32_t *load;          // load index of each entry
    const double *coef;           // factor * dir
    const double *series;         // concatenated amplitude series
    const int32_t *soff, *snt;    // per load
    const double *amp;            // per-load amplitude of this step (host-fed) or null
    const double *kinv;
    double *Un;
    int k;
    const int32_t *kctl;          // device step control {k, recorder row}: overrides k when the step is replayed from a graph
    const int32_t *target;        // per loaded dof: slot in hF (interface dof), -2-c (PML unknown c) or -1
    double *hF, *bext;
    int phase;                    // 0: interface / PML dofs (before the exchange / block solve), 1: all other dofs
    double sign;                  // phase 1: +1, or -1 for the reaction pass (R = F_int - F_ext)
};
__global__ void k_nodal_loads(const PLArgs a) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= a.n) return;
    const int k = a.kctl ? a.kctl[0] : a.k;
    double F = 0.0;
    for (int q = a.ptr[t]; q < a.ptr[t + 1]; q++) {
        const int l = a.load[q];
        double amp;
        if (a.amp) amp = a.amp[l];
        else amp = (a.snt[l] == 1) ? a.series[a.soff[l]] : ((k < a.snt[l]) ? a.series[a.soff[l] + k] : 0.0);
        F += a.coef[q] * amp;
    }
    const int tg = a.target ? a.target[t] : -1;
    if (a.phase == 0) {
        if (tg >= 0) a.hF[tg] -= F;
        else if (tg <= -2) a.bext[-2 - tg] += F;
        return;
    }
    if (tg != -1) return;
    const int d = a.dof[t];
    a.Un[d] += a.sign * (a.kinv ? a.kinv[d] : 1.0) * F;
}

// ------------------------------------------------------------------------------------------
// DRM effective forces (lin3DHexa8.cpp:660-718): per DRM node, sum of pre-assembled
// boundary<->exterior stiffness blocks times the (sign-flipped for exterior) incident field.
// ------------------------------------------------------------------------------------------
struct DrmArgs {
    int n, nn, ndim, nt, nf, k, analytic;
    const int32_t *dof0, *ptr;
    const int2 *cb;               // per entry: (local DRM node index of the column node, id of the unique K block)
    int us;                       // doubles per node in uo (ndim padded to 4 / 2)
    const uint8_t *ext;
    const double *dict;           // unique K blocks [nblk][ndim*ndim]
    const double *field;          // [nn][nt][nf]
    const double *xyz;            // [nn][ndim]
    double *uo;                   // [nn][ndim] incident displacement of this step (sign-flipped if exterior)
    double dir[3], pol[3], xref[3], c, f0, t0, amp, factor, dt;
    const double *kinv;
    double *Un;
    const int32_t *target;        // per row: first slot in hF (interface node) or -1; null without halos
    double *hF;
    int phase;
    const int32_t *kctl;          // device step control: the field is evaluated for step kctl[0] + koff
    int koff;
};
__device__ __forceinline__ double ricker_disp(double tau, double f0) {
    // Ricker displacement pulse (1 - 2b) e^{-b}, b = (pi f0 tau)^2  (PlaneWave.py:222-223)
    const double b = (M_PI * f0 * tau) * (M_PI * f0 * tau);
    return (1.0 - 2.0 * b) * exp(-b);
}
// incident field of step k at every DRM node (Node::GetDomainReductionMotion, Node.cpp:222-225;
// exterior rows negated as in Driver.hpp:1714-1716)
__global__ void k_drm_field(const DrmArgs a) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= a.nn) return;
    const int k = a.kctl ? a.kctl[0] + a.koff : a.k;
    const double sgn = a.ext[t] ? -1.0 : 1.0;
    if (a.analytic) {
        double s = 0.0;
        for (int c = 0; c < a.ndim; c++) s += (a.xyz[(long long)t * a.ndim + c] - a.xref[c]) * a.dir[c];
        const double val = a.amp * ricker_disp(k * a.dt - a.t0 - s / a.c, a.f0);
        for (int c = 0; c < a.ndim; c++) a.uo[(long long)t * a.us + c] = sgn * (val * a.pol[c]);
    } else {
        if (k >= a.nt) { for (int c = 0; c < a.ndim; c++) a.uo[(long long)t * a.us + c] = 0.0; return; }
        const double *row = a.field + ((long long)t * a.nt + k) * a.nf;
        for (int c = 0; c < a.ndim; c++) a.uo[(long long)t * a.us + c] = sgn * row[c];
    }
}
// forces of the DRM rows for one step into a compact buffer F[row][ndim] (one thread per row and component);
// they depend on the step index only, never on the state, so they are computed one step ahead on a side stream.
// The kernel is bound by L1 requests (profiles/r1y: 54 % LSU wavefronts, 7 sectors per request), so the operands are
// laid out for wide loads: (column node, block id) pairs as int2, incident displacements and dictionary rows padded to
// 4 doubles (3-D: one 128-bit + one 64-bit load each; 2-D: one 128-bit load) -- 5 requests per entry instead of 8.
// (A one-thread-per-row ELL variant was measured slower: fewer threads in flight, profiles/r1l.)
template <int ND>
__global__ void __launch_bounds__(128) k_drm(const DrmArgs a, double *F) {
    constexpr int NS = (ND == 3) ? 4 : 2;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= a.n * ND) return;
    const int row = t / ND, r = t - row * ND;
    double f = 0.0;
    const int q1 = a.ptr[row + 1];
    int q = a.ptr[row];
    // four entries at a time: their index and operand loads are independent and go out together; the sum keeps the
    // ascending entry order
    for (; q + 4 <= q1; q += 4) {
        int2 cb[4];
#pragma unroll
        for (int z = 0; z < 4; z++) cb[z] = a.cb[q + z];
        double2 u2[4], b2[4];
        double u3[4], b3[4];
#pragma unroll
        for (int z = 0; z < 4; z++) {
            const double *u = a.uo + (long long)cb[z].x * NS;
            const double *B = a.dict + ((long long)cb[z].y * ND + r) * NS;
            u2[z] = *reinterpret_cast<const double2 *>(u); b2[z] = *reinterpret_cast<const double2 *>(B);
            if (ND == 3) { u3[z] = u[2]; b3[z] = B[2]; }
        }
#pragma unroll
        for (int z = 0; z < 4; z++) {
            f += b2[z].x * u2[z].x;
            f += b2[z].y * u2[z].y;
            if (ND == 3) f += b3[z] * u3[z];
        }
    }
    for (; q < q1; q++) {
        const int2 cb = a.cb[q];
        const double *u = a.uo + (long long)cb.x * NS;
        const double *B = a.dict + ((long long)cb.y * ND + r) * NS;
        for (int c = 0; c < ND; c++) f += B[c] * u[c];
    }
    F[t] = a.factor * f;
}
// Analytic plane wave: u_j(t) = +-amp ricker(t - t0 - s_j / c) pol, so the 3x3 (2x2) block products collapse to one scalar
// per entry: F_i = factor sum_j (B_ij pol) val_j with the blocks contracted with pol once at plan time.  One thread per
// DRM node evaluates the wave value; one thread per ROW then needs 3 requests per entry (index pair, scalar, contracted
// block) for ND products instead of 5 requests for one.  Entries are summed in the same ascending order.
__global__ void k_drm_field_pw(int nn, const uint8_t *ext, const double *sc, double amp, double f0, double t0, double dt, int k,
                               const int32_t *kctl, int koff, double *sval) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nn) return;
    const int kk = kctl ? kctl[0] + koff : k;
    const double val = amp * ricker_disp(kk * dt - t0 - sc[t], f0);
    sval[t] = ext[t] ? -val : val;
}
template <int ND>
__global__ void __launch_bounds__(128) k_drm_pw(int n, const int32_t *ptr, const int2 *cb, const double *sval, const double *wdict,
                                               double factor, double *F) {
    constexpr int NS = (ND == 3) ? 4 : 2;
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= n) return;
    double f[ND];
#pragma unroll
    for (int r = 0; r < ND; r++) f[r] = 0.0;
    const int q1 = ptr[row + 1];
    int q = ptr[row];
    for (; q + 4 <= q1; q += 4) {
        int2 e[4];
#pragma unroll
        for (int z = 0; z < 4; z++) e[z] = cb[q + z];
        double sv[4];
        double2 w2[4];
        double w3[4];
#pragma unroll
        for (int z = 0; z < 4; z++) {
            sv[z] = sval[e[z].x];
            const double *w = wdict + (long long)e[z].y * NS;
            w2[z] = *reinterpret_cast<const double2 *>(w);
            if (ND == 3) w3[z] = w[2];
        }
#pragma unroll
        for (int z = 0; z < 4; z++) {
            f[0] += w2[z].x * sv[z];
            f[1] += w2[z].y * sv[z];
            if (ND == 3) f[2] += w3[z] * sv[z];
        }
    }
    for (; q < q1; q++) {
        const int2 e = cb[q];
        const double sv = sval[e.x];
        const double *w = wdict + (long long)e.y * NS;
#pragma unroll
        for (int r = 0; r < ND; r++) f[r] += w[r] * sv;
    }
#pragma unroll
    for (int r = 0; r < ND; r++) F[(long long)row * ND + r] = factor * f[r];
}
// k_drm_pw with the application folded in (models without interface / PML rows): U_{n+1}[row dofs] += sign F / Keff straight
// from the registers -- no F buffer round trip, no k_drm_apply launch, and no side-stream prefetch (which never overlapped
// with the stencil kernel anyway).  Same products in the same order as k_drm_pw + k_drm_apply.
template <int ND>
__global__ void __launch_bounds__(128) k_drm_pw_apply(int n, const int32_t *ptr, const int2 *cb, const double *sval, const double *wdict,
                                                     double factor, const int32_t *dof0, const double *rkinv, double sign, double *Un) {
    constexpr int NS = (ND == 3) ? 4 : 2;
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= n) return;
    double f[ND];
#pragma unroll
    for (int r = 0; r < ND; r++) f[r] = 0.0;
    const int q1 = ptr[row + 1];
    int q = ptr[row];
    for (; q + 4 <= q1; q += 4) {
        int2 e[4];
#pragma unroll
        for (int z = 0; z < 4; z++) e[z] = cb[q + z];
        double sv[4];
        double2 w2[4];
        double w3[4];
#pragma unroll
        for (int z = 0; z < 4; z++) {
            sv[z] = sval[e[z].x];
            const double *w = wdict + (long long)e[z].y * NS;
            w2[z] = *reinterpret_cast<const double2 *>(w);
            if (ND == 3) w3[z] = w[2];
        }
#pragma unroll
        for (int z = 0; z < 4; z++) {
            f[0] += w2[z].x * sv[z];
            f[1] += w2[z].y * sv[z];
            if (ND == 3) f[2] += w3[z] * sv[z];
        }
    }
    for (; q < q1; q++) {
        const int2 e = cb[q];
        const double sv = sval[e.x];
        const double *w = wdict + (long long)e.y * NS;
#pragma unroll
        for (int r = 0; r < ND; r++) f[r] += w[r] * sv;
    }
    const int d = dof0[row];
#pragma unroll
    for (int r = 0; r < ND; r++) Un[d + r] += sign * (rkinv ? rkinv[(long long)row * ND + r] : 1.0) * (factor * f[r]);
}
// The same, fused (SVLGPU_DRM_FUSE; measured slower at 320^3 -- ten FP64 exp per row cost more than two launches and a buffer
// round trip save -- so off by default): wave value per ENTRY, row force, and its application in one launch -- phase 0: rows on interface nodes (hF -= F, before the exchange);
// phase 1: all other rows (U_{n+1} += sign F / Keff).  The forces depend on the step index only, so evaluating them where
// they are applied costs nothing extra; three launches and two buffers per step become one launch (the side-stream
// prefetch never overlapped with the register-saturating stencil kernel anyway: DESIGN.md section 5).  Same arithmetic
// and summation order as k_drm_field_pw + k_drm_pw + k_drm_apply: bit-identical forces.
template <int ND>
__global__ void __launch_bounds__(128) k_drm_pw_fused(int n, int phase, const int32_t *ptr, const int2 *cb, const uint8_t *ext, const double *sc,
                                                     const double *wdict, double amp, double f0, double t0, double tnow, double factor,
                                                     const int32_t *dof0, const int32_t *target, const double *rkinv, double sign,
                                                     double *Un, double *hF) {
    constexpr int NS = (ND == 3) ? 4 : 2;
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= n) return;
    const int tg = target ? target[row] : -1;
    if ((phase == 0) != (tg >= 0)) return;
    double f[ND];
#pragma unroll
    for (int r = 0; r < ND; r++) f[r] = 0.0;
    const int q1 = ptr[row + 1];
    int q = ptr[row];
    for (; q + 4 <= q1; q += 4) {
        int2 e[4];
#pragma unroll
        for (int z = 0; z < 4; z++) e[z] = cb[q + z];
        double sv[4];
        double2 w2[4];
        double w3[4];
#pragma unroll
        for (int z = 0; z < 4; z++) {
            const double val = amp * ricker_disp(tnow - t0 - sc[e[z].x], f0);
            sv[z] = ext[e[z].x] ? -val : val;
            const double *w = wdict + (long long)e[z].y * NS;
            w2[z] = *reinterpret_cast<const double2 *>(w);
            if (ND == 3) w3[z] = w[2];
        }
#pragma unroll
        for (int z = 0; z < 4; z++) {
            f[0] += w2[z].x * sv[z];
            f[1] += w2[z].y * sv[z];
            if (ND == 3) f[2] += w3[z] * sv[z];
        }
    }
    for (; q < q1; q++) {
        const int2 e = cb[q];
        const double val = amp * ricker_disp(tnow - t0 - sc[e.x], f0);
        const double sv = ext[e.x] ? -val : val;
        const double *w = wdict + (long long)e.y * NS;
#pragma unroll
        for (int r = 0; r < ND; r++) f[r] += w[r] * sv;
    }
    if (tg >= 0) {
#pragma unroll
        for (int r = 0; r < ND; r++) hF[tg + r] -= factor * f[r];
        return;
    }
    const int d = dof0[row];
#pragma unroll
    for (int r = 0; r < ND; r++) Un[d + r] += sign * (rkinv ? rkinv[(long long)row * ND + r] : 1.0) * (factor * f[r]);
}
// phase 0: rows on interface nodes (hF -= F, before the exchange); phase 1: all other rows (U_{n+1} += F / Keff)
__global__ void k_drm_apply(int n, int nd, int phase, const int32_t *dof0, const int32_t *target, const double *F,
                            const double *kinv, double *Un, double *hF, double sign) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * nd) return;
    const int row = t / nd, r = t - row * nd;
    const int tg = target ? target[row] : -1;
    if ((phase == 0) != (tg >= 0)) return;
    if (tg >= 0) { hF[tg + r] -= F[t]; return; }
    const int d = dof0[row] + r;
    Un[d] += sign * (kinv ? kinv[t] : 1.0) * F[t];          // kinv: compact per row dof (DrmDev::d_rkinv)
}

// ------------------------------------------------------------------------------------------
// Support motion (Assembler.cpp:493-533, Node.cpp:228-247).  The state buffers hold, at a moving support dof, the value the
// ELEMENTS see: like the reference, one step behind the true displacement (Algorithm.cpp:23 passes T dU only to the element
// update, CentralDifference.cpp:135 adds the increment afterwards).  With W_k the buffer read by step k:
// W_{k+1}[s] = W_k[s] + dg(k-1) (k_support), and the true history of a support dof is U(k) = W_{k+1} + dg(k),
// U(k-1) = W_{k+1}, U(k-2) = W_k -- what the recorders and getters report.
struct SupArgs {
    int n;
    const int32_t *dof, *ptr;     // ascending internal dofs, CSR into series
    const double *series, *fac;
};
__device__ __forceinline__ double sup_dg(const SupArgs &s, int q, int j) {      // factor (g(j) - g(j-1)), 0 for j < 1
    if (j < 1) return 0.0;
    const double *x = s.series + s.ptr[q];
    const int sz = s.ptr[q + 1] - s.ptr[q];
    const double g1 = (j < sz) ? x[j] : x[0], g0 = (j - 1 < sz) ? x[j - 1] : x[0];
    return s.fac[q] * (g1 - g0);
}
__device__ __forceinline__ int sup_find(const SupArgs &s, int d) {
    int lo = 0, hi = s.n - 1;
    while (lo <= hi) {
        const int mid = (lo + hi) >> 1, v = s.dof[mid];
        if (v == d) return mid;
        if (v < d) lo = mid + 1; else hi = mid - 1;
    }
    return -1;
}
// true U(k), U(k-1), U(k-2) of dof d after step k (buffers: Un = newest, U, Up)
__device__ __forceinline__ void state3(const SupArgs &s, int k, int d, const double *Un, const double *U, const double *Up,
                                       double &un, double &u, double &up) {
    un = Un[d]; u = U[d]; up = Up[d];
    if (s.n) {
        const int q = sup_find(s, d);
        if (q >= 0) { up = u; u = un; un = un + sup_dg(s, q, k); }
    }
}
__global__ void k_support(const SupArgs s, int k, double *Un) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < s.n) Un[s.dof[t]] += sup_dg(s, t, k - 1);
}

// ------------------------------------------------------------------------------------------
// NODE recorder row (Recorder.cpp:239-269); V, A as in CentralDifference.cpp:141-144
// ------------------------------------------------------------------------------------------
__global__ void k_record(int n, const int32_t *dofs, const double *Un, const double *U, const double *Up,
                         double dt, int field, double *row, const int32_t *kctl, int max_rows, double *mirror,
                         const double *Vs, const double *As, const SupArgs sup, int k) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    if (kctl) {                                   // graph replay: `row` is the recorder base, the row index lives on the device
        if (kctl[1] >= max_rows) return;
        row += (size_t)kctl[1] * n;
    }
    const int d = dofs[t];
    double un, u, up;
    state3(sup, k, d, Un, U, Up, un, u, up);
    double v;
    if (field == SVLGPU_DISP) v = un;
    else if (Vs) v = (field == SVLGPU_VEL) ? Vs[d] : As[d];      // Newmark keeps V, A as state (NewmarkBeta.cpp:75-76)
    else if (field == SVLGPU_VEL) v = 1.0 / 2.0 / dt * (un - up);
    else v = 1.0 / dt / dt * ((un - u) - u + up);
    row[t] = v;
    if (mirror) mirror[t] = v;                    // svlgpu_step_host: the row also goes straight to mapped pinned host memory
}

// REACTION recorder row: R = F_int + C V + M A - F_ext at the dofs of fixed nodes, zero elsewhere (DynamicAnalysis.cpp:130-150,
// CentralDifference.cpp:155-171).  Fs = F_int - F_ext of the newest state (force-only pass); M, C are the lumped diagonals
// (Assembler.cpp:568-619 sums M_e A_e and C_e V_e of every element at a fixed node: with lumped mass and mass-proportional
// damping these are diagonal), plus the off-diagonal couplings of ZeroLength1D dashpots.
struct ReacArgs {
    int n;
    const int32_t *dofs;
    const double *Fs, *Un, *U, *Up;
    const double *mass, *cd, *ccoef;
    const uint8_t *fixed;
    const int32_t *cptr, *cdof;
    double dt;
    int k;
    SupArgs sup;
};
__global__ void k_reaction(const ReacArgs a, double *row, double *mirror) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= a.n) return;
    double r = 0.0;
    if (a.fixed[t]) {
        const int d = a.dofs[t];
        double un, u, up;
        state3(a.sup, a.k, d, a.Un, a.U, a.Up, un, u, up);
        const double V = 1.0 / 2.0 / a.dt * (un - up), A = 1.0 / a.dt / a.dt * ((un - u) - u + up);
        r = a.Fs[d] + a.cd[t] * V + a.mass[t] * A;
        for (int q = a.cptr[t]; q < a.cptr[t + 1]; q++) {
            const int d2 = a.cdof[q];
            state3(a.sup, a.k, d2, a.Un, a.U, a.Up, un, u, up);
            r += a.ccoef[q] * (1.0 / 2.0 / a.dt * (un - up));
        }
    }
    row[t] = r;
    if (mirror) mirror[t] = r;
}

__global__ void k_gather(int n, const int32_t *dofs, const int32_t *int_of_total, const double *Un,
                         const double *U, const double *Up, double dt, int field, double *out, const double *Vs,
                         const double *As, const SupArgs sup, int k) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const int d = int_of_total[dofs ? dofs[t] : t];
    if (Vs && field != SVLGPU_DISP) { out[t] = (field == SVLGPU_VEL) ? Vs[d] : As[d]; return; }
    // state after the last step: Un = U_{n+1} (current), U = U_n, Up = U_{n-1}
    double un, u, up;
    state3(sup, k, d, Un, U, Up, un, u, up);
    double v;
    if (field == SVLGPU_DISP) v = un;
    else if (field == SVLGPU_VEL) v = 1.0 / 2.0 / dt * (un - up);
    else v = 1.0 / dt / dt * ((un - u) - u + up);
    out[t] = v;
}

// device step control block {k, recorder row}: advanced at the end of every step so that a captured graph of
// steps can be replayed without any host-side argument
__global__ void k_advance(int32_t *kctl) { kctl[0]++; kctl[1]++; }
// step index of the DRM forces that are about to be computed ahead into buffer b (slot 2 + b): written on the main
// stream before the side stream forks, so the side-stream kernels never race with k_advance
__global__ void k_setk(int32_t *kctl, int slot, int off) { kctl[slot] = kctl[0] + off; }

// ------------------------------------------------------------------------------------------
// host drivers
// ------------------------------------------------------------------------------------------
void timer_begin(svlgpu_model *m, int which) {
    if (!m->kernel_timing) return;
    KernelTimer &t = m->timers[which];
    if (!t.e0) { cudaEventCreate(&t.e0); cudaEventCreate(&t.e1); }
    if (t.pending) {
        float ms = 0;
        cudaEventSynchronize(t.e1);
        cudaEventElapsedTime(&ms, t.e0, t.e1);
        t.total_ms += ms; t.pending = false;
    }
    cudaEventRecord(t.e0, m->stream);
}
void timer_end(svlgpu_model *m, int which) {
    if (!m->kernel_timing) return;
    KernelTimer &t = m->timers[which];
    cudaEventRecord(t.e1, m->stream);
    t.pending = true;
    t.launches++;
}
void timer_flush(svlgpu_model *m) {
    for (int i = 0; i < kNumTimers; i++) {
        KernelTimer &t = m->timers[i];
        if (t.pending) {
            float ms = 0;
            cudaEventSynchronize(t.e1);
            cudaEventElapsedTime(&ms, t.e0, t.e1);
            t.total_ms += ms; t.pending = false;
        }
    }
}

size_t stencil3_tma_smem(int nw, int r) { return (4ull * (nw * r + 2) * kT3Row + (size_t)nw * 32 * r * 3) * sizeof(double) + 64; }
size_t stencil3_smem(int nw, int r) { return (3ull * 3 * (nw * r + 2) * 36 + (size_t)nw * 32 * r * 3) * sizeof(double); }

static svlgpu_model *g_const_owner = nullptr;
static int upload_dom_tables(svlgpu_model *m) {
    if (g_const_owner == m) return 0;
    CUDA_OK(cudaDeviceSynchronize());          // another model's kernels may still read the tables
    for (auto &b : m->blocks)
        for (auto &d : b.doms)
            CUDA_OK(cudaMemcpyToSymbolAsync(cK, d.tbl, sizeof(double) * kTbl3Stride, sizeof(double) * kTbl3Stride * d.slot,
                                            cudaMemcpyHostToDevice, m->stream));
    g_const_owner = m;
    return 0;
}
void forget_const_owner(svlgpu_model *m) { if (g_const_owner == m) g_const_owner = nullptr; }

template <int NW, int R, int SLOT>
static void launch_dom_so(const Dom3 &p, bool ortho, unsigned grid, cudaStream_t st) {
    const size_t sm = stencil3_smem(NW, R);
    if (ortho) k_stencil3_dom<NW, R, SLOT, true><<<grid, NW * 32, sm, st>>>(p);
    else k_stencil3_dom<NW, R, SLOT, false><<<grid, NW * 32, sm, st>>>(p);
}
template <int NW, int R, int SLOT>
static void launch_tma_so(const Dom3 &p, bool ortho, unsigned grid, cudaStream_t st) {
    const size_t sm = stencil3_tma_smem(NW, R);
    if (ortho) k_stencil3_tma<NW, R, SLOT, true><<<grid, NW * 32, sm, st>>>(p);
    else k_stencil3_tma<NW, R, SLOT, false><<<grid, NW * 32, sm, st>>>(p);
}
template <int NW, int R>
static void launch_tma(const Dom3 &p, int slot, bool ortho, unsigned grid, cudaStream_t st) {
    switch (slot) {
    case 0: launch_tma_so<NW, R, 0>(p, ortho, grid, st); break;
    case 1: launch_tma_so<NW, R, 1>(p, ortho, grid, st); break;
    case 2: launch_tma_so<NW, R, 2>(p, ortho, grid, st); break;
    default: launch_tma_so<NW, R, 3>(p, ortho, grid, st); break;
    }
}
size_t stencil3_v4_smem(int nw, int r) { return stencil3_tma_smem(nw, r); }
template <int NW, int R, int SLOT>
static void launch_v4_so(const Dom3 &p, bool nobar, unsigned grid, cudaStream_t st) {
    const size_t sm = stencil3_v4_smem(NW, R);
    if (nobar) k_stencil3_v4<NW, R, SLOT, true, true><<<grid, NW * 32, sm, st>>>(p);
    else k_stencil3_v4<NW, R, SLOT, true, false><<<grid, NW * 32, sm, st>>>(p);
}
template <int NW, int R>
static void launch_v4(const Dom3 &p, int slot, bool nobar, unsigned grid, cudaStream_t st) {
    switch (slot) {
    case 0: launch_v4_so<NW, R, 0>(p, nobar, grid, st); break;
    case 1: launch_v4_so<NW, R, 1>(p, nobar, grid, st); break;
    case 2: launch_v4_so<NW, R, 2>(p, nobar, grid, st); break;
    default: launch_v4_so<NW, R, 3>(p, nobar, grid, st); break;
    }
}
template <int NW, int R>
static void launch_dom(const Dom3 &p, int slot, bool ortho, unsigned grid, cudaStream_t st) {
    switch (slot) {
    case 0: launch_dom_so<NW, R, 0>(p, ortho, grid, st); break;
    case 1: launch_dom_so<NW, R, 1>(p, ortho, grid, st); break;
    case 2: launch_dom_so<NW, R, 2>(p, ortho, grid, st); break;
    default: launch_dom_so<NW, R, 3>(p, ortho, grid, st); break;
    }
}

// 1. generic Gauss-point elements (element forces into the arena; J2 state update + commit)
static int launch_generic_elements(svlgpu_model *m, const double *U, int commit) {
    for (auto &gs : m->gsets) {
        if (!gs.n) continue;
        GenArgs a;
        a.n = gs.n; a.conn = gs.d_conn; a.mat = gs.d_mat; a.th = gs.d_th; a.matkind = m->d_matkind;
        a.matpar = m->d_matpar; a.coords = m->d_coords; a.node_ptr = m->d_node_ptr; a.U = U;
        a.fe = gs.d_fe; a.state = gs.d_state; a.gp = gs.d_gp; a.commit = commit;
        a.ecls = gs.d_ecls; a.gtab = gs.d_gtab;
        timer_begin(m, 1);
        if (gs.kind == SVLGPU_LIN3DHEXA8) {
            const unsigned grid = (unsigned)((8ll * gs.n + 127) / 128);
            if (gs.d_gtab) k_gen_hex8_tab<<<grid, 128, 0, m->stream>>>(a);
            else k_gen_hex8<false><<<grid, 128, 0, m->stream>>>(a);
        } else {
            const unsigned grid = (unsigned)((4ll * gs.n + 127) / 128);
            if (gs.d_gtab) k_gen_quad4<true><<<grid, 128, 0, m->stream>>>(a);
            else k_gen_quad4<false><<<grid, 128, 0, m->stream>>>(a);
        }
        timer_end(m, 1);
        m->total_launches++;
    }
    return 0;
}

// 2. + 3. lattice blocks and generic nodes: force gather + CentralDifference update (mode 0) or force only (mode 1)
static int launch_node_update(svlgpu_model *m, const double *U, const double *Up, double *Un, int mode) {
    if (upload_dom_tables(m)) return 1;
    bool shell_on_side = false;
    // The shell classes run on side stream 1 beside the bulk kernel -- except in the per-step host call (svlgpu_step_host):
    // there the fork / join costs four more API calls per step on the critical path of the host round trip (measured e2e
    // 5.54 -> 5.81e10 without the side streams)
    const bool side_ok = m->overlap && !m->kernel_timing && !m->host_step;
    if (side_ok) {                                   // side stream 1 may start once U_n is final
        CUDA_OK(cudaEventRecord(m->ev_fork, m->stream));
        CUDA_OK(cudaStreamWaitEvent(m->side[1], m->ev_fork, 0));
    }
    for (auto &b : m->blocks) {
        if (!b.n_stencil_nodes) continue;
        if (b.ndim == 3) {
            for (auto &d : b.doms) {
                Dom3 p;
                p.U = U; p.Up = Up; p.Un = Un; p.cls = b.d_cls; p.dof0 = b.dof0;
                p.nx = b.nx; p.ny = b.ny; p.nz = b.nz;
                p.bi0 = d.bi0; p.bj0 = d.bj0; p.bk0 = d.bk0; p.bk1 = d.bk1;
                p.bi1 = d.bi1; p.bj1 = d.bj1;
                p.tiles_x = d.tiles_x; p.tiles_y = d.tiles_y; p.kz = d.kz; p.dom = d.cls; p.mode = mode;
                const unsigned grid = (unsigned)(d.tiles_x * d.tiles_y * d.zchunks);
                timer_begin(m, 0);
                if (d.sep) {
                    Sep3 cf;
                    for (int a = 0; a < 3; a++) {
                        for (int x = 0; x < 3; x++) cf.c[a][x] = d.sepc[3 * a + x];
                        cf.e[a] = d.sepe[a]; cf.kinv[a] = d.tbl[270 + a]; cf.km[a] = d.tbl[273 + a];
                    }
                    k_stencil3_sep<kDomNW, kDomR><<<grid, kDomNW * 32, stencil3_tma_smem(kDomNW, kDomR), m->stream>>>(p, cf);
                } else if (d.pure && d.sym && d.v4 && d.rows == 6) launch_v4<kDomNW, 6>(p, d.slot, d.nobar, grid, m->stream);
                else if (d.pure && d.sym && d.v4) launch_v4<kDomNW, kDomR>(p, d.slot, d.nobar, grid, m->stream);
                else if (d.pure) launch_tma<kDomNW, kDomR>(p, d.slot, d.ortho, grid, m->stream);
                else launch_dom<kDomNW, kDomR>(p, d.slot, d.ortho, grid, m->stream);
                timer_end(m, 0);
                m->total_launches++;
            }
            if (mode == 0 ? b.n_shell_chunks : b.n_shell_all) {
                // the shell classes write nodes no other kernel of the step writes: run them beside the bulk kernels
                Shell3 p;
                p.U = U; p.Up = Up; p.Un = Un; p.tbl = b.d_tbl;
                p.list = mode == 0 ? b.d_shell_list : b.d_shell_all_list;
                p.chunk_cls = mode == 0 ? b.d_shell_cls : b.d_shell_all_cls;
                const int nchunks = mode == 0 ? b.n_shell_chunks : b.n_shell_all;
                p.dof0 = b.dof0; p.nx = b.nx; p.ny = b.ny; p.nz = b.nz; p.mode = mode; p.target = nullptr; p.hF = nullptr;
                cudaStream_t st = side_ok ? m->side[1] : m->stream;
                timer_begin(m, 4);
                if (m->shell_lowreg) k_stencil3_shell<true><<<nchunks, 128, 0, st>>>(p);
                else k_stencil3_shell<false><<<nchunks, 128, 0, st>>>(p);
                timer_end(m, 4);
                if (st != m->stream) shell_on_side = true;
                m->total_launches++;
            }
        } else {
            Blk2 p;
            p.U = U; p.Up = Up; p.Un = Un; p.cls = b.d_cls; p.tbl = b.d_tbl; p.dof0 = b.dof0;
            p.ncls = b.ncls; p.nx = b.nx; p.ny = b.ny; p.mode = mode;
            dim3 grid((b.nx + 63) / 64, (b.ny + 3) / 4);
            timer_begin(m, 0);
            k_stencil2<<<grid, 256, (size_t)b.ncls * kTbl2Stride * sizeof(double), m->stream>>>(p);
            timer_end(m, 0);
            m->total_launches++;
        }
    }
    if (m->nbr.n_chunks) {
        NbrArgs a;
        a.U = U; a.Up = Up; a.Un = Un; a.tbl = m->nbr.d_tbl; a.cls_nn = m->nbr.d_cls_nn; a.chunk_cls = m->nbr.d_chunk_cls;
        a.dof0 = m->nbr.d_dof0; a.nbr = m->nbr.d_nbr; a.chunk_off = (const long long *)m->nbr.d_chunk_off;
        a.stride = m->nbr.stride; a.mode = mode;
        timer_begin(m, 2);
        // 3 slots per trip: 0.409 -> 0.390 ms (160^3, lattice order) and 0.660 -> 0.569 ms (random numbering), profiles/r3o
        static const int unr = getenv("SVLGPU_NBR_UNROLL") ? atoi(getenv("SVLGPU_NBR_UNROLL")) : 3;
        if (m->ndim == 3) { if (unr == 3) k_nbr_nodes<3, 3><<<m->nbr.n_chunks, 128, 0, m->stream>>>(a); else k_nbr_nodes<3, 1><<<m->nbr.n_chunks, 128, 0, m->stream>>>(a); }
        else { if (unr == 3) k_nbr_nodes<2, 3><<<m->nbr.n_chunks, 128, 0, m->stream>>>(a); else k_nbr_nodes<2, 1><<<m->nbr.n_chunks, 128, 0, m->stream>>>(a); }
        timer_end(m, 2);
        m->total_launches++;
    }
    if (m->n_gnodes) {
        GNodeArgs a;
        a.n = m->n_gnodes; a.dof0 = m->d_gn_dof0; a.ndof = m->d_gn_ndof; a.ptr = m->d_gn_ptr;
        a.slot = (const long long *)m->d_gn_slot; a.fe = m->d_fe_arena; a.U = U; a.Up = Up;
        a.kinv = m->d_kinv; a.km = m->d_km; a.Un = Un; a.mode = mode;
        a.target = nullptr; a.hF = nullptr; a.hnd = 0;
        timer_begin(m, 2);
        k_gen_nodes<<<(m->n_gnodes + 255) / 256, 256, 0, m->stream>>>(a);
        timer_end(m, 2);
        m->total_launches++;
    }
    if (shell_on_side) {
        CUDA_OK(cudaEventRecord(m->ev_join, m->side[1]));
        CUDA_OK(cudaStreamWaitEvent(m->stream, m->ev_join, 0));
    }
    CUDA_OK(cudaGetLastError());
    return 0;
}

// partial internal force of the interface nodes -> halo.d_hF (lattice nodes through the class tables,
// generic nodes through the element-force arena)
int halo_lattice_force(svlgpu_model *m, const double *U) {
    HaloDev &h = m->halo;
    for (auto &l : h.lats) {
        if (!l.n) continue;
        const Block &b = m->blocks[l.block];
        if (b.ndim == 3 && l.n_chunks) {
            Shell3 p;
            p.U = U; p.Up = nullptr; p.Un = nullptr; p.tbl = b.d_tbl; p.list = l.d_chunk_list; p.chunk_cls = l.d_chunk_cls;
            p.dof0 = b.dof0; p.nx = b.nx; p.ny = b.ny; p.nz = b.nz; p.mode = 1; p.target = l.d_chunk_target; p.hF = h.d_hF;
            k_stencil3_shell<false><<<l.n_chunks, 128, 0, m->stream>>>(p);
        } else if (b.ndim == 3) {
            Gat3 p;
            p.U = U; p.Up = nullptr; p.Un = nullptr; p.cls = b.d_cls; p.tbl = b.d_tbl; p.list = l.d_list;
            p.dof0 = b.dof0; p.n = l.n; p.nx = b.nx; p.ny = b.ny; p.nz = b.nz; p.mode = 1;
            p.target = l.d_target; p.hF = h.d_hF;
            k_stencil3_gather<<<(l.n + 127) / 128, 128, 0, m->stream>>>(p);
        } else {
            Gat2 p;
            p.U = U; p.cls = b.d_cls; p.tbl = b.d_tbl; p.list = l.d_list; p.target = l.d_target; p.hF = h.d_hF;
            p.dof0 = b.dof0; p.n = l.n; p.nx = b.nx; p.ny = b.ny;
            k_stencil2_gather<<<(l.n + 127) / 128, 128, 0, m->stream>>>(p);
        }
        m->total_launches++;
    }
    CUDA_OK(cudaGetLastError());
    return 0;
}
int halo_generic_force(svlgpu_model *m) {
    HaloDev &h = m->halo;
    if (!h.n_gen) return 0;
    GNodeArgs a;
    a.n = h.n_gen; a.dof0 = nullptr; a.ndof = h.d_g_ndof; a.ptr = h.d_g_ptr; a.slot = (const long long *)h.d_g_slot;
    a.fe = m->d_fe_arena; a.U = nullptr; a.Up = nullptr; a.kinv = nullptr; a.km = nullptr; a.Un = nullptr; a.mode = 1;
    a.target = h.d_g_target; a.hF = h.d_hF; a.hnd = h.nd;
    k_gen_nodes<<<(h.n_gen + 255) / 256, 256, 0, m->stream>>>(a);
    m->total_launches++;
    CUDA_OK(cudaGetLastError());
    return 0;
}

static SupArgs sup_args(const svlgpu_model *m) {
    SupArgs s;
    s.n = m->sup.n; s.dof = m->sup.d_dof; s.ptr = m->sup.d_ptr; s.series = m->sup.d_series; s.fac = m->sup.d_fac;
    return s;
}

static int launch_external(svlgpu_model *m, int k, const double *dev_amp, double *Un, int phase, bool scaled = true, double sign = 1.0);
static int launch_generic_elements(svlgpu_model *m, const double *U, int commit);
static int launch_node_update(svlgpu_model *m, const double *U, const double *Up, double *Un, int mode);

// F_int(newest state) - F_ext(k) of the whole mesh into the scratch vector, interface dofs summed over the ranks:
// what Integrator::ComputeReactionForce assembles (CentralDifference.cpp:155-171), at the cost of one force-only pass
static int reaction_pass(svlgpu_model *m, int k, const double *dev_amp) {
    if (!m->d_fscratch) {
        CUDA_OK(cudaMalloc(&m->d_fscratch, sizeof(double) * m->n_int));
        m->allocs.push_back(m->d_fscratch);
        m->device_bytes += (int64_t)sizeof(double) * m->n_int;
    }
    double *Fs = m->d_fscratch;
    const double *Un = m->d_U[m->next];
    CUDA_OK(cudaMemsetAsync(Fs, 0, sizeof(double) * m->n_int, m->stream));
    if (launch_generic_elements(m, Un, 0)) return 1;             // stresses of the newest state, nothing committed
    if (launch_node_update(m, Un, nullptr, Fs, 1)) return 1;
    if (launch_external(m, k, dev_amp, Fs, 1, false, -1.0)) return 1;
    if (m->halo.active && !m->halo_peers.empty()) {
        if (halo_vec_load(m, Fs)) return 1;                        // hF <- partial F_int at the interface dofs
        if (external_forces_interface(m, k, dev_amp)) return 1;    // hF -= F_ext handed to this rank
        if (halo_vec_sum(m, Fs)) return 1;                         // rank-ordered sum: the same bits on every holder
    }
    return 0;
}

int record_rows(svlgpu_model *m, bool devk) {
    int ri = -1;
    bool reaction_ready = false;
    const SupArgs sup = sup_args(m);
    if (m->opt_reaction_collective && m->halo.active && !m->halo_peers.empty()) {
        // some rank records reactions: the pass exchanges interface partial forces, so every rank runs it on every step
        if (reaction_pass(m, m->k_of_step, m->step_amp)) return 1;
        reaction_ready = true;
    }
    for (auto &r : m->recorders) {
        ri++;
        if (r.rows >= r.max_rows || !r.width) continue;
        double *row = devk ? r.d_rows : r.d_rows + (size_t)r.rows * r.width;
        double *mirror = (ri == m->mirror_rec && !devk) ? m->h_row : nullptr;
        if (r.field == SVLGPU_REACTION) {
            if (!r.d_rmass) {                          // first use: after svlgpu_comm_init the diagonals are the global sums
                std::vector<double> ms(r.width), cd(r.width);
                for (int t = 0; t < r.width; t++) { ms[t] = m->h_mass[r.h_dofs[t]]; cd[t] = m->h_cdiag[r.h_dofs[t]] + r.h_cdg[t]; }
                auto up = [&](const void *src, size_t bytes) -> void * {
                    void *p = nullptr;
                    if (cudaMalloc(&p, std::max<size_t>(bytes, 8)) != cudaSuccess) return nullptr;
                    cudaMemcpyAsync(p, src, bytes, cudaMemcpyHostToDevice, m->stream);
                    cudaStreamSynchronize(m->stream);
                    m->allocs.push_back(p);
                    return p;
                };
                r.d_rmass = (double *)up(ms.data(), sizeof(double) * r.width);
                r.d_rcd = (double *)up(cd.data(), sizeof(double) * r.width);
                r.d_fixed = (uint8_t *)up(r.h_fixed.data(), r.width);
                r.d_cptr = (int32_t *)up(r.h_cptr.data(), sizeof(int32_t) * r.h_cptr.size());
                r.d_cdof = (int32_t *)up(r.h_cdof.data(), sizeof(int32_t) * r.h_cdof.size());
                r.d_ccoef = (double *)up(r.h_ccoef.data(), sizeof(double) * r.h_ccoef.size());
                if (!r.d_rmass || !r.d_rcd || !r.d_fixed || !r.d_cptr || !r.d_cdof || !r.d_ccoef) { set_error("out of device memory (REACTION recorder tables)"); return 1; }
            }
            if (!reaction_ready) { if (reaction_pass(m, m->k_of_step, m->step_amp)) return 1; reaction_ready = true; }
            ReacArgs a;
            a.n = r.width; a.dofs = r.d_dofs; a.Fs = m->d_fscratch; a.Un = m->d_U[m->next]; a.U = m->d_U[m->cur]; a.Up = m->d_U[m->prev];
            a.mass = r.d_rmass; a.cd = r.d_rcd; a.ccoef = r.d_ccoef; a.fixed = r.d_fixed; a.cptr = r.d_cptr; a.cdof = r.d_cdof;
            a.dt = m->dt; a.k = m->k_of_step; a.sup = sup;
            k_reaction<<<(r.width + 127) / 128, 128, 0, m->stream>>>(a, row, mirror);
        } else {
            k_record<<<(r.width + 127) / 128, 128, 0, m->stream>>>(r.width, r.d_dofs, m->d_U[m->next], m->d_U[m->cur],
                                                                     m->d_U[m->prev], m->dt, r.field, row,
                                                                     devk ? m->d_kctl : nullptr, r.max_rows, mirror,
                                                                     m->nm.present ? m->nm.d_V : nullptr, m->nm.present ? m->nm.d_A : nullptr,
                                                                     sup, m->k_of_step);
        }
        r.rows++;
        m->total_launches++;
    }
    return 0;
}

// incident field + DRM row forces of step k into buffer k & 1, on stream st
static int drm_compute(svlgpu_model *m, DrmDev &d, int k, cudaStream_t st) {
    DrmArgs a;
    a.n = d.n_nodes; a.nn = d.n_all; a.ndim = m->ndim; a.nt = d.nt; a.nf = d.nf; a.k = k; a.analytic = d.analytic;
    a.dof0 = d.d_node_dof0; a.ptr = d.d_row_ptr; a.cb = (const int2 *)d.d_col_blk; a.us = (m->ndim == 3) ? 4 : 2; a.ext = d.d_ext;
    a.dict = d.d_blk; a.field = d.d_field; a.xyz = d.d_xyz; a.uo = d.d_uo[k & 1];
    for (int c = 0; c < 3; c++) { a.dir[c] = d.dir[c]; a.pol[c] = d.pol[c]; a.xref[c] = d.xref[c]; }
    a.c = d.c; a.f0 = d.f0; a.t0 = d.t0; a.amp = d.amp; a.factor = d.factor; a.dt = m->dt;
    a.kinv = nullptr; a.Un = nullptr; a.target = nullptr; a.hF = nullptr; a.phase = 0;
    a.kctl = m->graph_capturing ? m->d_kctl + 2 + (k & 1) : nullptr; a.koff = 0;
    if (d.analytic && d.d_wdict) {
        k_drm_field_pw<<<(a.nn + 255) / 256, 256, 0, st>>>(a.nn, d.d_ext, d.d_sc, d.amp, d.f0, d.t0, m->dt, k, a.kctl, 0, d.d_sval[k & 1]);
        if (m->ndim == 3) k_drm_pw<3><<<(a.n + 127) / 128, 128, 0, st>>>(a.n, d.d_row_ptr, (const int2 *)d.d_col_blk, d.d_sval[k & 1], d.d_wdict, d.factor, d.d_F[k & 1]);
        else k_drm_pw<2><<<(a.n + 127) / 128, 128, 0, st>>>(a.n, d.d_row_ptr, (const int2 *)d.d_col_blk, d.d_sval[k & 1], d.d_wdict, d.factor, d.d_F[k & 1]);
        d.buf_k[k & 1] = k;
        m->total_launches += 2;
        CUDA_OK(cudaGetLastError());
        return 0;
    }
    k_drm_field<<<(a.nn + 127) / 128, 128, 0, st>>>(a);
    if (m->ndim == 3) k_drm<3><<<(a.n * 3 + 127) / 128, 128, 0, st>>>(a, d.d_F[k & 1]);
    else k_drm<2><<<(a.n * 2 + 127) / 128, 128, 0, st>>>(a, d.d_F[k & 1]);
    d.buf_k[k & 1] = k;
    m->total_launches += 2;
    CUDA_OK(cudaGetLastError());
    return 0;
}
// forces of step k + 1 while step k runs (side stream; buffer (k+1)&1 was last read by step k-1)
static int drm_prefetch(svlgpu_model *m, int knext) {
    for (auto &d : m->drm_dev) {
        if (!d.n_nodes || (!d.analytic && knext >= d.nt)) continue;
        if (d.fused && !m->graph_capturing) continue;          // evaluated where it is applied (k_drm_pw_fused)
        if (d.inline_apply && m->host_step && !(m->halo.active || m->pml.present) && !m->graph_capturing) continue;   // k_drm_pw_apply
        const int b = knext & 1;
        if (drm_compute(m, d, knext, m->side[0])) return 1;
        CUDA_OK(cudaEventRecord(d.ev_ready[b], m->side[0]));
        d.ev_valid[b] = true;
    }
    return 0;
}

// external forces of step k (Assembler::ComputeExternalForceVector).  phase 0: contributions to interface
// dofs, subtracted from the partial force that is about to be exchanged; phase 1: everything else,
// applied to U_{n+1} directly (the solve is diagonal there).
// `scaled`: forces are divided by the diagonal Keff on the way into U_{n+1} (CentralDifference); otherwise they are added raw
static int launch_external(svlgpu_model *m, int k, const double *dev_amp, double *Un, int phase, bool scaled, double sign) {
    const double *kinv = scaled ? m->d_kinv : nullptr;
    const bool halo = m->halo.active || m->pml.present;
    if (phase == 0 && !halo) return 0;
    if (m->n_pl_dofs) {
        PLArgs a;
        a.n = m->n_pl_dofs; a.dof = m->d_pl_dof; a.ptr = m->d_pl_ptr; a.load = m->d_pl_load;
        a.coef = m->d_pl_coef; a.series = m->d_pl_series; a.soff = m->d_pl_soff; a.snt = m->d_pl_nt;
        a.amp = dev_amp; a.kinv = kinv; a.Un = Un; a.k = k;
        a.target = halo ? m->d_pl_target : nullptr; a.hF = m->halo.d_hF; a.bext = m->pml.d_bext; a.phase = phase;
        a.kctl = m->graph_capturing ? m->d_kctl : nullptr;
        a.sign = sign;
        timer_begin(m, 3);
        k_nodal_loads<<<(a.n + 127) / 128, 128, 0, m->stream>>>(a);
        timer_end(m, 3);
        m->total_launches++;
    }
    for (auto &d : m->drm_dev) {
        if (!d.analytic && k >= d.nt) continue;
        if (!d.n_nodes) continue;
        const int b = k & 1;
        timer_begin(m, 5);
        if (d.fused && !m->graph_capturing) {
            const int32_t *tgt = halo ? d.d_target : nullptr;
            const double *rk = kinv ? d.d_rkinv : nullptr;
            if (m->ndim == 3) k_drm_pw_fused<3><<<(d.n_nodes + 127) / 128, 128, 0, m->stream>>>(d.n_nodes, phase, d.d_row_ptr, (const int2 *)d.d_col_blk, d.d_ext, d.d_sc, d.d_wdict, d.amp, d.f0, d.t0, k * m->dt, d.factor, d.d_node_dof0, tgt, rk, sign, Un, m->halo.d_hF);
            else k_drm_pw_fused<2><<<(d.n_nodes + 127) / 128, 128, 0, m->stream>>>(d.n_nodes, phase, d.d_row_ptr, (const int2 *)d.d_col_blk, d.d_ext, d.d_sc, d.d_wdict, d.amp, d.f0, d.t0, k * m->dt, d.factor, d.d_node_dof0, tgt, rk, sign, Un, m->halo.d_hF);
            timer_end(m, 5);
            m->total_launches++;
            continue;
        }
        // Bulk stepping (svlgpu_step over many steps) keeps the two-kernel form prefetched on a side stream: part of it hides in
        // the stencil kernel's tail (0.513 vs 0.528 ms per step).  The per-step host call takes the inline form: fewer launches
        // and no stream fork on the critical path of the host round trip (e2e 6.00 -> 6.20e10), profiles/r3r.
        if (d.inline_apply && m->host_step && !halo && !m->graph_capturing) {
            // wave values of step k, then force + application in one kernel (k_drm_pw_apply); phase 0 has returned above (!halo)
            k_drm_field_pw<<<(d.n_all + 255) / 256, 256, 0, m->stream>>>(d.n_all, d.d_ext, d.d_sc, d.amp, d.f0, d.t0, m->dt, k, nullptr, 0, d.d_sval[0]);
            const double *rk = kinv ? d.d_rkinv : nullptr;
            if (m->ndim == 3) k_drm_pw_apply<3><<<(d.n_nodes + 127) / 128, 128, 0, m->stream>>>(d.n_nodes, d.d_row_ptr, (const int2 *)d.d_col_blk, d.d_sval[0], d.d_wdict, d.factor, d.d_node_dof0, rk, sign, Un);
            else k_drm_pw_apply<2><<<(d.n_nodes + 127) / 128, 128, 0, m->stream>>>(d.n_nodes, d.d_row_ptr, (const int2 *)d.d_col_blk, d.d_sval[0], d.d_wdict, d.factor, d.d_node_dof0, rk, sign, Un);
            timer_end(m, 5);
            m->total_launches += 2;
            continue;
        }
        if (d.buf_k[b] != k) { if (drm_compute(m, d, k, m->stream)) return 1; }        // not precomputed: do it now
        else if (d.ev_valid[b]) cudaStreamWaitEvent(m->stream, d.ev_ready[b], 0);
        k_drm_apply<<<(d.n_nodes * m->ndim + 127) / 128, 128, 0, m->stream>>>(d.n_nodes, m->ndim, phase, d.d_node_dof0,
                                                                              halo ? d.d_target : nullptr, d.d_F[b], kinv ? d.d_rkinv : nullptr, Un,
                                                                              m->halo.d_hF, sign);
        timer_end(m, 5);
        m->total_launches++;
    }
    return 0;
}

// one step (DynamicAnalysis.cpp:36-57 loop body) enqueued on the model's streams
static int step_once(svlgpu_model *m, int k, const double *dev_amp) {
    const bool xchg = !m->halo_peers.empty();                 // interface nodes shared with other ranks
    const bool halo = m->halo.active || m->pml.present;       // interface nodes exist (peers and / or PML ties)
    const double *U = m->d_U[m->cur], *Up = m->d_U[m->prev];
    double *Un = m->d_U[m->next];
    m->k_of_step = k;
    m->step_amp = dev_amp;
    bool drm_ahead = false;                                   // some DRM load still uses the two-step (field, force) kernels
    for (auto &d : m->drm_dev)
        drm_ahead = drm_ahead || !((d.fused || (d.inline_apply && m->host_step && !(m->halo.active || m->pml.present))) && !m->graph_capturing);
    if (m->overlap && drm_ahead && !m->kernel_timing && !m->host_step) {
        // DRM forces of step k+1 are computed on side stream 0 while this step runs; its buffer was last read
        // by step k-1, which is complete on the main stream at this point.  They are enqueued BEFORE the bulk
        // kernels: enqueued after them they only get SM slots when the stencil drains and the step serialises
        // (measured: 0.693 -> 0.781 ms per step, profiles/r1s).
        if (m->graph_capturing) { k_setk<<<1, 1, 0, m->stream>>>(m->d_kctl, 2 + ((k + 1) & 1), 1); m->total_launches++; }
        CUDA_OK(cudaEventRecord(m->ev_fork2, m->stream));
        CUDA_OK(cudaStreamWaitEvent(m->side[0], m->ev_fork2, 0));
        if (drm_prefetch(m, k + 1)) return 1;
    }
    if (launch_generic_elements(m, U, 1)) return 1;
    if (halo) {
        const bool async_if = xchg && m->overlap && !m->kernel_timing && !m->pml.present && !m->graph_capturing &&
                              !getenv("SVLGPU_HALO_SYNC");
        if (async_if) {
            // interface pass on the comm stream: the bulk kernels below do not wait for it
            if (halo_async_begin(m)) return 1;
            cudaStream_t main_stream = m->stream;
            m->stream = m->halo.comm_stream;
            const int rc = halo_lattice_force(m, U) || halo_generic_force(m) || launch_external(m, k, dev_amp, Un, 0);
            m->stream = main_stream;
            if (rc || halo_async_exchange(m)) return 1;
        } else {
            // interface partial forces first, so that their exchange overlaps the bulk of the step
            if (halo_lattice_force(m, U) || halo_generic_force(m)) return 1;
            if (launch_external(m, k, dev_amp, Un, 0)) return 1;
            if (xchg && halo_exchange_begin(m)) return 1;
        }
    }
    if (launch_node_update(m, U, Up, Un, 0)) return 1;
    if (xchg && halo_exchange_end(m, U, Up, Un, 0)) return 1;
    if (m->pml.present && pml_step(m, U, Up, Un)) return 1;
    if (launch_external(m, k, dev_amp, Un, 1)) return 1;
    if (m->sup.n) {                                           // W_{k+1}[s] = W_k[s] + dg(k-1), see k_support
        k_support<<<(m->sup.n + 127) / 128, 128, 0, m->stream>>>(sup_args(m), k, Un);
        m->total_launches++;
    }
    m->step_amp = dev_amp;
    if (record_rows(m, m->graph_capturing)) return 1;
    if (m->use_graph) {                                       // the device-side step counter only serves graph replay
        k_advance<<<1, 1, 0, m->stream>>>(m->d_kctl);
        m->total_launches++;
        m->dev_k = k + 1;
    }
    // rotate: U_{n-1} <- U_n <- U_{n+1}
    const int old_prev = m->prev;
    m->prev = m->cur; m->cur = m->next; m->next = old_prev;
    m->steps_done++;
    return 0;
}

// Steps are replayed from a CUDA graph of kGraphSteps consecutive steps (one period of the 3 rotating state
// buffers and the 2 DRM buffers): per-step launch latency is what limits small partitions (8 GPUs on 10^8 DOF).
constexpr int kGraphSteps = 6;
static bool graph_usable(const svlgpu_model *m, const double *dev_amp) {
    return m->use_graph && !dev_amp && !m->kernel_timing && !m->pml.present && !m->sup.n && !m->has_reaction_rec && !m->opt_reaction_collective;
}
void graph_destroy(svlgpu_model *m) {
    if (m->graph_exec) cudaGraphExecDestroy((cudaGraphExec_t)m->graph_exec);
    m->graph_exec = nullptr;
}

int run_steps(svlgpu_model *m, int k0, int k1, const double *dev_amp) {
    const int64_t before = m->total_launches;
    if (!m->halo_peers.empty() && !m->halo.active) { set_error("step: halos were declared but svlgpu_comm_init was not called"); return 1; }
    if (upload_dom_tables(m)) return 1;
    int k = k0;
    if (m->nm.present) {                                      // NewmarkBeta + Linear (newmark.cu)
        for (; k < k1; k++) if (newmark_step(m, k, dev_amp)) return 1;
        if (k1 > k0) m->launches_per_step = (m->total_launches - before) / (k1 - k0);
        return 0;
    }
    while (k < k1) {
        if (m->use_graph && m->dev_k != k) {                  // (re)synchronise the device step control block
            const int32_t h[2] = {k, m->recorders.empty() ? 0 : m->recorders[0].rows};
            CUDA_OK(cudaMemcpyAsync(m->d_kctl, h, sizeof(h), cudaMemcpyHostToDevice, m->stream));
            CUDA_OK(cudaStreamSynchronize(m->stream));
            m->dev_k = k;
        }
        bool rows_uniform = true;                             // one device row counter serves all recorders
        for (auto &r : m->recorders) rows_uniform = rows_uniform && r.rows == m->recorders[0].rows && r.rows + kGraphSteps <= r.max_rows;
        const bool can_graph = graph_usable(m, dev_amp) && rows_uniform && k + kGraphSteps <= k1;
        if (can_graph && (m->graph_exec ? ((k - m->graph_k0) % kGraphSteps == 0 && m->graph_cur == m->cur) : (k1 - k >= 2 * kGraphSteps))) {
            // the forces of step k must sit in DRM buffer k & 1, ordered before the graph on the main stream
            for (auto &d : m->drm_dev) {
                if (!d.n_nodes) continue;
                if (d.buf_k[k & 1] != k) { if (drm_compute(m, d, k, m->stream)) return 1; }
                else if (d.ev_valid[k & 1]) CUDA_OK(cudaStreamWaitEvent(m->stream, d.ev_ready[k & 1], 0));
                d.ev_valid[0] = d.ev_valid[1] = false;       // from here on main-stream order covers both buffers
            }
        }
        if (can_graph && m->graph_exec && (k - m->graph_k0) % kGraphSteps == 0 && m->graph_cur == m->cur) {
            CUDA_OK(cudaGraphLaunch((cudaGraphExec_t)m->graph_exec, m->stream));
            for (auto &r : m->recorders) r.rows += kGraphSteps;
            for (auto &d : m->drm_dev) { d.buf_k[k & 1] = k + kGraphSteps; d.buf_k[(k + 1) & 1] = k + kGraphSteps - 1; d.ev_valid[0] = d.ev_valid[1] = false; }
            m->steps_done += kGraphSteps; m->dev_k = k + kGraphSteps;
            m->total_launches += m->graph_launches;
            k += kGraphSteps;
            continue;
        }
        if (can_graph && !m->graph_exec && k1 - k >= 2 * kGraphSteps) {
            // capture kGraphSteps steps; the step index and the recorder row come from the device control block
            const int64_t l0 = m->total_launches;
            cudaGraph_t graph = nullptr;
            m->graph_capturing = true;
            cudaError_t ce = cudaStreamBeginCapture(m->stream, cudaStreamCaptureModeThreadLocal);
            int rc = ce != cudaSuccess;
            for (int q = 0; q < kGraphSteps && !rc; q++) rc = step_once(m, k + q, nullptr);
            if (!rc && m->overlap && !m->drm_dev.empty())     // join the last DRM prefetch into the origin stream
                for (auto &d : m->drm_dev) if (d.ev_valid[(k + kGraphSteps) & 1]) cudaStreamWaitEvent(m->stream, d.ev_ready[(k + kGraphSteps) & 1], 0);
            ce = cudaStreamEndCapture(m->stream, &graph);
            m->graph_capturing = false;
            if (rc || ce != cudaSuccess || !graph) { set_error(std::string("graph capture failed: ") + cudaGetErrorString(ce)); return 1; }
            cudaGraphExec_t exec = nullptr;
            ce = cudaGraphInstantiate(&exec, graph, 0);
            cudaGraphDestroy(graph);
            if (ce != cudaSuccess) { set_error(std::string("graph instantiate failed: ") + cudaGetErrorString(ce)); return 1; }
            m->graph_exec = exec; m->graph_k0 = k; m->graph_cur = m->cur;      // 6 steps return the rotation to where it began
            m->graph_launches = m->total_launches - l0;
            for (auto &d : m->drm_dev) d.ev_valid[0] = d.ev_valid[1] = false;   // events recorded in capture are not waitable outside
            // the capture advanced the host mirrors as if the steps had run: run them now
            CUDA_OK(cudaGraphLaunch(exec, m->stream));
            k += kGraphSteps;
            continue;
        }
        if (step_once(m, k, dev_amp)) return 1;
        k++;
    }
    CUDA_OK(cudaGetLastError());
    if (k1 > k0) m->launches_per_step = (m->total_launches - before) / (k1 - k0);
    return 0;
}

// NaN / Inf anywhere in the newest displacement vector -> flag (the reference does not check; SURVEY.md 8(b): must surface as stop)
__global__ void __launch_bounds__(256) k_finite_check(const double *U, long long n, int *flag) {
    bool bad = false;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const double v = U[i];
        bad = bad || !(fabs(v) <= 1.7976931348623157e308);
    }
    if (__syncthreads_or(bad) && threadIdx.x == 0) *flag = 1;
}
// returns 1 (and sets the error text) if the state holds a non-finite value; one streaming read of U_n
int state_is_finite(svlgpu_model *m) {
    if (!m->d_flag) {
        CUDA_OK(cudaMalloc(&m->d_flag, sizeof(int)));
        m->allocs.push_back(m->d_flag);
    }
    CUDA_OK(cudaMemsetAsync(m->d_flag, 0, sizeof(int), m->stream));
    k_finite_check<<<148 * 8, 256, 0, m->stream>>>(m->d_U[m->cur], (long long)m->n_int, m->d_flag);
    int h = 0;
    CUDA_OK(cudaMemcpyAsync(&h, m->d_flag, sizeof(int), cudaMemcpyDeviceToHost, m->stream));
    CUDA_OK(cudaStreamSynchronize(m->stream));
    if (h) { set_error("NaN / Inf in the displacement vector: the analysis diverged (time step above the stability limit?)"); return 1; }
    return 0;
}

// out = K x for any vector x in the internal dof layout: the force-only pass of the explicit path
// (Gauss-point elements + lattice stencils + generic node gather); used by the Newmark Krylov solve
int operator_K(svlgpu_model *m, const double *x, double *out) {
    if (launch_generic_elements(m, x, 0)) return 1;
    return launch_node_update(m, x, nullptr, out, 1);
}
// b += Fext(k) on every loaded dof (point loads, DRM), no Keff scaling
int external_forces_raw(svlgpu_model *m, int k, const double *dev_amp, double *b) {
    return launch_external(m, k, dev_amp, b, 1, false);
}

// hF -= Fext(k) at the interface dofs (loads are handed to one partition only, so they travel with the exchange)
int external_forces_interface(svlgpu_model *m, int k, const double *dev_amp) {
    return launch_external(m, k, dev_amp, nullptr, 0, false);
}

// Assembler::ComputeInternalForceVector for the current displacement state
int compute_internal_force(svlgpu_model *m, double *F_host) {
    // a scratch vector of its own: the three rotating state buffers stay intact (d_U[next] holds U_{n-1}, which the
    // VEL / ACCEL getters and recorders read)
    if (!m->d_fscratch) {
        CUDA_OK(cudaMalloc(&m->d_fscratch, sizeof(double) * m->n_int));
        m->allocs.push_back(m->d_fscratch);
        m->device_bytes += (int64_t)sizeof(double) * m->n_int;
    }
    double *tmp = m->d_fscratch;
    CUDA_OK(cudaMemsetAsync(tmp, 0, sizeof(double) * m->n_int, m->stream));
    if (launch_generic_elements(m, m->d_U[m->cur], 0)) return 1;
    if (launch_node_update(m, m->d_U[m->cur], m->d_U[m->prev], tmp, 1)) return 1;
    if (pml_internal_force(m, m->d_U[m->cur], tmp)) return 1;
    std::vector<double> h(m->n_int);
    CUDA_OK(cudaMemcpyAsync(h.data(), tmp, sizeof(double) * m->n_int, cudaMemcpyDeviceToHost, m->stream));
    CUDA_OK(cudaStreamSynchronize(m->stream));
    for (int t = 0; t < m->n_total; t++) F_host[t] = h[m->int_of_total[t]];
    return 0;
}

int gather_state(svlgpu_model *m, int field, const int32_t *dofs, int n, double *out) {
    int32_t *d_dofs = nullptr;
    double *d_out = nullptr;
    if (dofs) {
        CUDA_OK(cudaMalloc(&d_dofs, sizeof(int32_t) * n));
        CUDA_OK(cudaMemcpyAsync(d_dofs, dofs, sizeof(int32_t) * n, cudaMemcpyHostToDevice, m->stream));
    }
    CUDA_OK(cudaMalloc(&d_out, sizeof(double) * n));
    // after a step the newest state sits in `cur`; U_n in `prev`, U_{n-1} in `next`
    k_gather<<<(n + 255) / 256, 256, 0, m->stream>>>(n, d_dofs, m->d_int_of_total, m->d_U[m->cur], m->d_U[m->prev],
                                                      m->d_U[m->next], m->dt, (m->steps_done || m->nm.present) ? field : SVLGPU_DISP, d_out,
                                                      m->nm.present ? m->nm.d_V : nullptr, m->nm.present ? m->nm.d_A : nullptr,
                                                      sup_args(m), m->steps_done ? m->k_of_step : 0);
    CUDA_OK(cudaMemcpyAsync(out, d_out, sizeof(double) * n, cudaMemcpyDeviceToHost, m->stream));
    CUDA_OK(cudaStreamSynchronize(m->stream));
    cudaFree(d_dofs); cudaFree(d_out);
    return 0;
}

template <int SLOT> static int cfg_slot() {
    CUDA_OK(cudaFuncSetAttribute(k_stencil3_dom<kDomNW, kDomR, SLOT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)stencil3_smem(kDomNW, kDomR)));
    CUDA_OK(cudaFuncSetAttribute(k_stencil3_dom<kDomNW, kDomR, SLOT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)stencil3_smem(kDomNW, kDomR)));
    CUDA_OK(cudaFuncSetAttribute(k_stencil3_tma<kDomNW, kDomR, SLOT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)stencil3_tma_smem(kDomNW, kDomR)));
    CUDA_OK(cudaFuncSetAttribute(k_stencil3_tma<kDomNW, kDomR, SLOT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)stencil3_tma_smem(kDomNW, kDomR)));
    CUDA_OK(cudaFuncSetAttribute(k_stencil3_v4<kDomNW, kDomR, SLOT, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)stencil3_v4_smem(kDomNW, kDomR)));
    CUDA_OK(cudaFuncSetAttribute(k_stencil3_v4<kDomNW, kDomR, SLOT, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)stencil3_v4_smem(kDomNW, kDomR)));
    CUDA_OK(cudaFuncSetAttribute(k_stencil3_v4<kDomNW, 6, SLOT, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)stencil3_v4_smem(kDomNW, 6)));
    CUDA_OK(cudaFuncSetAttribute(k_stencil3_v4<kDomNW, 6, SLOT, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)stencil3_v4_smem(kDomNW, 6)));
    return 0;
}
int configure_kernels() {
    if (cfg_slot<0>() || cfg_slot<1>() || cfg_slot<2>() || cfg_slot<3>()) return 1;
    CUDA_OK(cudaFuncSetAttribute(k_stencil3_sep<kDomNW, kDomR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)stencil3_tma_smem(kDomNW, kDomR)));
    CUDA_OK(cudaFuncSetAttribute(k_stencil2, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    if (pml_configure()) return 1;
    return 0;
}
bool stencil_entry_nonzero(int di, int b, int dj, int s, int a) { return stencil_nz(di, b, dj, s, a); }
int stencil_entry_sym_index(int di, int b, int dj, int s, int a) { return stencil_sym_idx(di, b, dj, s, a); }
bool stencil_entry_sym_negated(int di, int b, int dj, int s, int a) { return stencil_sym_neg(di, b, dj, s, a); }

}  // namespace svl
