// halo.cu -- multi-GPU interface exchange: replaces the per-step MPI_Reduce of the full
// right-hand side and MPI_Bcast of the full solution (11-Solvers/01-Direct/MumpsSolver.cpp:56,161) by an
// exchange of the partial forces of the interface nodes only.  One process per GPU; NCCL send/recv with every
// rank that shares nodes with this one, on a dedicated stream so that the exchange overlaps the bulk kernels.
// NCCL is bound with dlopen so that the library uses whichever libnccl.so.2 the host process has loaded
// (torch's bundled copy under torchrun) and still loads on machines without NCCL for single-GPU use.
#include <dlfcn.h>
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

// ---- the few NCCL entry points used, resolved at run time ---------------------------------------
typedef struct { char internal[128]; } nccl_uid;
struct Nccl {
    void *lib = nullptr;
    int (*GetUniqueId)(nccl_uid *) = nullptr;
    int (*CommInitRank)(void **, int, nccl_uid, int) = nullptr;
    int (*CommDestroy)(void *) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    int (*Send)(const void *, size_t, int, int, void *, cudaStream_t) = nullptr;
    int (*Recv)(void *, size_t, int, int, void *, cudaStream_t) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, void *, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
};
static Nccl g_nccl;
static const int kNcclDouble = 8;   // ncclFloat64 (nccl.h)
static const int kNcclSum = 0;      // ncclSum

static int nccl_load() {
    if (g_nccl.lib) return 0;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *n : names) {
        g_nccl.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (g_nccl.lib) break;
    }
    if (!g_nccl.lib) { set_error("multi-GPU: libnccl.so.2 not found"); return 1; }
#define SYM(field, name)                                                              \
    *(void **)(&g_nccl.field) = dlsym(g_nccl.lib, name);                              \
    if (!g_nccl.field) { set_error(std::string("multi-GPU: NCCL symbol missing: ") + name); return 1; }
    SYM(GetUniqueId, "ncclGetUniqueId") SYM(CommInitRank, "ncclCommInitRank") SYM(CommDestroy, "ncclCommDestroy")
    SYM(GroupStart, "ncclGroupStart") SYM(GroupEnd, "ncclGroupEnd") SYM(Send, "ncclSend") SYM(Recv, "ncclRecv")
    SYM(AllReduce, "ncclAllReduce") SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
    return 0;
}
#define NCCL_OK(x)                                                                           \
    do {                                                                                     \
        int r_ = (x);                                                                        \
        if (r_ != 0) { set_error(std::string(#x) + ": " + g_nccl.GetErrorString(r_)); return 1; } \
    } while (0)

// ---- kernels -------------------------------------------------------------------------------------
__global__ void k_halo_pack(int n, int nd, const int32_t *map, const double *hF, double *send) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * nd) return;
    const int e = t / nd, c = t - e * nd;
    send[t] = hF[(long long)map[e] * nd + c];
}
// every replica of an interface node sums the partial forces in ascending rank order and applies the
// CentralDifference update (CentralDifference.cpp:138-148, 217) with the globally summed lumped mass
__global__ void k_halo_fix(int n_if, int nd, const int32_t *dof0, const int32_t *ptr, const int32_t *src,
                           const double *hF, const double *recv, const double *U, const double *Up, double *Un,
                           const double *kinv, const double *km, int mode) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_if * nd) return;
    const int i = t / nd, c = t - i * nd;
    double F = 0.0;
    for (int q = ptr[i]; q < ptr[i + 1]; q++) {
        const int s = src[q];
        F += (s < 0) ? hF[t] : recv[(long long)s * nd + c];
    }
    const int d = dof0[i] + c;
    if (mode == 0) {
        const double un = U[d];
        Un[d] = un + (km[d] * (un - Up[d]) - F) * kinv[d];
    } else {
        Un[d] = F;
    }
}
__global__ void k_halo_load(int n_if, int nd, const int32_t *dof0, const double *src, double *hF) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_if * nd) return;
    hF[t] = src[dof0[t / nd] + (t % nd)];
}
__global__ void k_halo_sum_to(int n_if, int nd, const int32_t *dof0, const int32_t *ptr, const int32_t *src,
                              const double *hF, const double *recv, double *dst) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_if * nd) return;
    const int i = t / nd, c = t - i * nd;
    double F = 0.0;
    for (int q = ptr[i]; q < ptr[i + 1]; q++) {
        const int s = src[q];
        F += (s < 0) ? hF[t] : recv[(long long)s * nd + c];
    }
    dst[dof0[i] + c] = F;
}
__global__ void k_halo_coeffs(int n_if, int nd, const int32_t *dof0, const double *mass, const double *cd,
                              const uint8_t *isfree, double dt, double *kinv, double *km) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_if * nd) return;
    const int d = dof0[t / nd] + (t % nd);
    if (!isfree[d]) { kinv[d] = 0.0; km[d] = 0.0; return; }
    const double keff = 1.0 / dt / dt * mass[d] + 1.0 / 2.0 / dt * cd[d];
    kinv[d] = 1.0 / keff;
    km[d] = 1.0 / dt / dt * mass[d] - 1.0 / 2.0 / dt * cd[d];
}

// ---- exchange ---------------------------------------------------------------------------------------
static int exchange_on(svlgpu_model *m, cudaStream_t st) {
    HaloDev &h = m->halo;
    const int nd = h.nd;
    if (h.n_entries) k_halo_pack<<<(h.n_entries * nd + 255) / 256, 256, 0, st>>>(h.n_entries, nd, h.d_send_map, h.d_hF, h.d_send);
    NCCL_OK(g_nccl.GroupStart());
    for (auto &hp : m->halo_peers) {
        const size_t cnt = hp.nodes.size() * (size_t)nd;
        if (!cnt) continue;                              // a peer that shares PML nodes only (exchanged inside the block solve)
        NCCL_OK(g_nccl.Send(h.d_send + (size_t)hp.offset * nd, cnt, kNcclDouble, hp.peer, h.comm, st));
        NCCL_OK(g_nccl.Recv(h.d_recv + (size_t)hp.offset * nd, cnt, kNcclDouble, hp.peer, h.comm, st));
    }
    NCCL_OK(g_nccl.GroupEnd());
    m->total_launches += 1;
    return 0;
}

int halo_exchange_begin(svlgpu_model *m) {
    HaloDev &h = m->halo;
    CUDA_OK(cudaEventRecord(h.e_ready, m->stream));
    CUDA_OK(cudaStreamWaitEvent(h.comm_stream, h.e_ready, 0));
    if (exchange_on(m, h.comm_stream)) return 1;
    CUDA_OK(cudaEventRecord(h.e_done, h.comm_stream));
    return 0;
}

// Asynchronous variant: the whole interface pass (partial forces of the interface nodes, external forces acting on them,
// pack, NCCL) is enqueued on the high-priority comm stream, so the bulk kernels on the main stream start at once and the
// interface pass runs beside them instead of in front of them.  halo_async_begin forks the comm stream off the main
// stream (U_n and the element-force arena are final there); the caller then enqueues the interface kernels on
// halo.comm_stream and calls halo_async_exchange.
int halo_async_begin(svlgpu_model *m) {
    HaloDev &h = m->halo;
    CUDA_OK(cudaEventRecord(h.e_ready, m->stream));
    CUDA_OK(cudaStreamWaitEvent(h.comm_stream, h.e_ready, 0));
    return 0;
}
int halo_async_exchange(svlgpu_model *m) {
    HaloDev &h = m->halo;
    if (exchange_on(m, h.comm_stream)) return 1;
    CUDA_OK(cudaEventRecord(h.e_done, h.comm_stream));
    return 0;
}

int halo_exchange_end(svlgpu_model *m, const double *U, const double *Up, double *Un, int mode) {
    HaloDev &h = m->halo;
    CUDA_OK(cudaStreamWaitEvent(m->stream, h.e_done, 0));
    if (h.n_if) {
        k_halo_fix<<<(h.n_if * h.nd + 255) / 256, 256, 0, m->stream>>>(h.n_if, h.nd, h.d_if_dof0, h.d_fix_ptr, h.d_fix_src, h.d_hF,
                                                                       h.d_recv, U, Up, Un, m->d_kinv, m->d_km, mode);
        m->total_launches++;
    }
    CUDA_OK(cudaGetLastError());
    return 0;
}

// any vector in the internal dof layout: interface values summed over the ranks that hold the node (Newmark Krylov solve)
int halo_vec_load(svlgpu_model *m, const double *src) {
    HaloDev &h = m->halo;
    if (h.n_if) k_halo_load<<<(h.n_if * h.nd + 255) / 256, 256, 0, m->stream>>>(h.n_if, h.nd, h.d_if_dof0, src, h.d_hF);
    m->total_launches++;
    return 0;
}
int halo_vec_sum(svlgpu_model *m, double *dst) {
    HaloDev &h = m->halo;
    if (exchange_on(m, m->stream)) return 1;
    if (h.n_if) k_halo_sum_to<<<(h.n_if * h.nd + 255) / 256, 256, 0, m->stream>>>(h.n_if, h.nd, h.d_if_dof0, h.d_fix_ptr, h.d_fix_src, h.d_hF,
                                                                                  h.d_recv, dst);
    m->total_launches++;
    CUDA_OK(cudaGetLastError());
    return 0;
}

int halo_unique_id(void *out128) {
    if (nccl_load()) return 1;
    nccl_uid id;
    NCCL_OK(g_nccl.GetUniqueId(&id));
    std::memcpy(out128, &id, 128);
    return 0;
}

// joins the communicator, builds the rank-ordered summation lists and replaces the partial lumped mass /
// damping of the interface dofs by their global sums (Assembler.cpp:47-67 summed over partitions)
int halo_comm_init(svlgpu_model *m, const void *id128, int rank, int nranks) {
    HaloDev &h = m->halo;
    if (!m->finalized) { set_error("comm_init: finalize the model first"); return 1; }
    if (nccl_load()) return 1;
    CUDA_OK(cudaSetDevice(m->device));
    nccl_uid id;
    std::memcpy(&id, id128, 128);
    NCCL_OK(g_nccl.CommInitRank(&h.comm, nranks, id, rank));
    h.rank = rank; h.nranks = nranks;
    {
        int lo = 0, hi = 0;                            // the exchange must not queue behind the bulk kernels' CTAs
        CUDA_OK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        CUDA_OK(cudaStreamCreateWithPriority(&h.comm_stream, cudaStreamNonBlocking, getenv("SVLGPU_NO_PRIO") ? lo : hi));
    }
    CUDA_OK(cudaEventCreateWithFlags(&h.e_ready, cudaEventDisableTiming));
    CUDA_OK(cudaEventCreateWithFlags(&h.e_done, cudaEventDisableTiming));
    // summation order: ascending rank, own contribution at its place
    std::vector<std::vector<std::pair<int, int>>> srcs(h.n_if);
    for (int i = 0; i < h.n_if; i++) srcs[i].push_back({rank, -1});
    for (auto &hp : m->halo_peers) {
        if (hp.peer == rank || hp.peer < 0 || hp.peer >= nranks) { set_error("comm_init: bad halo peer rank"); return 1; }
        for (size_t j = 0; j < hp.nodes.size(); j++) srcs[m->if_of_node[hp.nodes[j]]].push_back({hp.peer, hp.offset + (int)j});
    }
    std::vector<int32_t> ptr(h.n_if + 1, 0), src;
    for (int i = 0; i < h.n_if; i++) {
        std::sort(srcs[i].begin(), srcs[i].end());
        for (auto &p : srcs[i]) src.push_back(p.second);
        ptr[i + 1] = (int32_t)src.size();
    }
    int32_t *dptr = nullptr, *dsrc = nullptr;
    CUDA_OK(cudaMalloc(&dptr, sizeof(int32_t) * ptr.size()));
    CUDA_OK(cudaMalloc(&dsrc, sizeof(int32_t) * std::max<size_t>(1, src.size())));
    m->allocs.push_back(dptr); m->allocs.push_back(dsrc);
    CUDA_OK(cudaMemcpy(dptr, ptr.data(), sizeof(int32_t) * ptr.size(), cudaMemcpyHostToDevice));
    CUDA_OK(cudaMemcpy(dsrc, src.data(), sizeof(int32_t) * src.size(), cudaMemcpyHostToDevice));
    h.d_fix_ptr = dptr; h.d_fix_src = dsrc;
    h.active = true;

    // global lumped mass / damping at the interface dofs
    if (h.n_if) {
        const int nd = h.nd, nthr = h.n_if * nd;
        double *dm = nullptr, *dc = nullptr;
        uint8_t *dfree = nullptr;
        std::vector<uint8_t> isfree(m->n_int);
        for (int q = 0; q < m->n_int; q++) isfree[q] = m->freedof[q] >= 0;
        CUDA_OK(cudaMalloc(&dm, sizeof(double) * m->n_int));
        CUDA_OK(cudaMalloc(&dc, sizeof(double) * m->n_int));
        CUDA_OK(cudaMalloc(&dfree, m->n_int));
        CUDA_OK(cudaMemcpy(dm, m->h_mass.data(), sizeof(double) * m->n_int, cudaMemcpyHostToDevice));
        CUDA_OK(cudaMemcpy(dc, m->h_cdiag.data(), sizeof(double) * m->n_int, cudaMemcpyHostToDevice));
        CUDA_OK(cudaMemcpy(dfree, isfree.data(), m->n_int, cudaMemcpyHostToDevice));
        for (double *arr : {dm, dc}) {
            k_halo_load<<<(nthr + 255) / 256, 256, 0, m->stream>>>(h.n_if, nd, h.d_if_dof0, arr, h.d_hF);
            if (exchange_on(m, m->stream)) return 1;
            k_halo_sum_to<<<(nthr + 255) / 256, 256, 0, m->stream>>>(h.n_if, nd, h.d_if_dof0, h.d_fix_ptr, h.d_fix_src, h.d_hF, h.d_recv, arr);
        }
        k_halo_coeffs<<<(nthr + 255) / 256, 256, 0, m->stream>>>(h.n_if, nd, h.d_if_dof0, dm, dc, dfree, m->dt, m->d_kinv, m->d_km);
        std::vector<double> cglob(m->n_int);
        CUDA_OK(cudaMemcpyAsync(m->h_mass.data(), dm, sizeof(double) * m->n_int, cudaMemcpyDeviceToHost, m->stream));
        CUDA_OK(cudaMemcpyAsync(cglob.data(), dc, sizeof(double) * m->n_int, cudaMemcpyDeviceToHost, m->stream));
        CUDA_OK(cudaStreamSynchronize(m->stream));
        // the planner let interface dofs pass without local mass (it may all sit on other ranks): the global sum must have some
        for (auto &hp : m->halo_peers)
            for (int node : hp.nodes)
                for (int q = m->node_ptr[node]; q < m->node_ptr[node + 1]; q++)
                    if (isfree[q] && !(1.0 / m->dt / m->dt * m->h_mass[q] + 1.0 / 2.0 / m->dt * cglob[q] > 0.0)) {
                        cudaFree(dm); cudaFree(dc); cudaFree(dfree);
                        set_error("free dof without mass: Keff is singular (EigenSolver.cpp:52-55)");
                        return 1;
                    }
        m->h_cdiag = cglob;                            // REACTION recorders read the global diagonals (uploaded at first use)
        const int rc = newmark_comm_setup(m, dm, dc);
        cudaFree(dm); cudaFree(dc); cudaFree(dfree);
        if (rc) return 1;
    } else {
        // ranks without interface nodes still take part in nothing: no peers, no NCCL calls -- except the all-reduces of a
        // Krylov solve (Newmark, PML block), which every rank of the communicator issues
        if (newmark_comm_setup(m, nullptr, nullptr)) return 1;
    }
    CUDA_OK(cudaGetLastError());
    return pmlx_setup(m);
}

// ---- PML block on several ranks (pml.cu): exchange of the shared unknowns, all-reduce of the dot products --------------
__global__ void k_pmlx_pack(int n, const int32_t *unk, const double *raw, double *sbuf) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) sbuf[t] = raw[unk[t]];
}
// every holder of a shared unknown adds the parts in ascending rank order (its own at its place): same bits everywhere
__global__ void k_pmlx_sum(int n, const int32_t *unk, const int32_t *ptr, const int32_t *src, double *raw, const double *rbuf) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const int c = unk[t];
    const double mine = raw[c];
    double v = 0.0;
    for (int q = ptr[t]; q < ptr[t + 1]; q++) v += (src[q] < 0) ? mine : rbuf[src[q]];
    raw[c] = v;
}

int pmlx_exchange(svlgpu_model *m, cudaStream_t st) {
    PmlDev &P = m->pml;
    HaloDev &h = m->halo;
    if (!P.n_xe) return 0;
    k_pmlx_pack<<<(P.n_xe + 255) / 256, 256, 0, st>>>(P.n_xe, P.d_x_send, P.d_raw, P.d_x_sbuf);
    NCCL_OK(g_nccl.GroupStart());
    for (auto &hp : m->halo_peers) {
        if (hp.unk.empty()) continue;
        NCCL_OK(g_nccl.Send(P.d_x_sbuf + hp.unk_offset, hp.unk.size(), kNcclDouble, hp.peer, h.comm, st));
        NCCL_OK(g_nccl.Recv(P.d_x_rbuf + hp.unk_offset, hp.unk.size(), kNcclDouble, hp.peer, h.comm, st));
    }
    NCCL_OK(g_nccl.GroupEnd());
    k_pmlx_sum<<<(P.n_xu + 255) / 256, 256, 0, st>>>(P.n_xu, P.d_x_unk, P.d_x_ptr, P.d_x_src, P.d_raw, P.d_x_rbuf);
    m->total_launches += 2;
    CUDA_OK(cudaGetLastError());
    return 0;
}

int pmlx_allreduce(svlgpu_model *m, double *a, size_t na, double *b, size_t nb, cudaStream_t st) {
    HaloDev &h = m->halo;
    NCCL_OK(g_nccl.AllReduce(a, a, na, kNcclDouble, kNcclSum, h.comm, st));
    if (b && nb) NCCL_OK(g_nccl.AllReduce(b, b, nb, kNcclDouble, kNcclSum, h.comm, st));
    return 0;
}

// Both sides of every pair of ranks must have derived lists of the same length, or the first exchange would wait for data that
// never comes: swap the counts (one double per peer) and stop loudly on a mismatch.  which: 0 soil halo nodes, 1 PML unknowns
static int verify_peer_counts(svlgpu_model *m, int which) {
    HaloDev &h = m->halo;
    const size_t np = m->halo_peers.size();
    if (!np) return 0;
    std::vector<double> mine(np), theirs(np, -1.0);
    for (size_t i = 0; i < np; i++) mine[i] = which ? (double)m->halo_peers[i].unk.size() : (double)m->halo_peers[i].nodes.size();
    double *d = nullptr;
    CUDA_OK(cudaMalloc(&d, sizeof(double) * 2 * np));
    CUDA_OK(cudaMemcpy(d, mine.data(), sizeof(double) * np, cudaMemcpyHostToDevice));
    NCCL_OK(g_nccl.GroupStart());
    for (size_t i = 0; i < np; i++) {
        NCCL_OK(g_nccl.Send(d + i, 1, kNcclDouble, m->halo_peers[i].peer, h.comm, m->stream));
        NCCL_OK(g_nccl.Recv(d + np + i, 1, kNcclDouble, m->halo_peers[i].peer, h.comm, m->stream));
    }
    NCCL_OK(g_nccl.GroupEnd());
    CUDA_OK(cudaMemcpyAsync(theirs.data(), d + np, sizeof(double) * np, cudaMemcpyDeviceToHost, m->stream));
    CUDA_OK(cudaStreamSynchronize(m->stream));
    cudaFree(d);
    for (size_t i = 0; i < np; i++)
        if (mine[i] != theirs[i]) {
            set_error(std::string("comm_init: rank ") + std::to_string(h.rank) + " and rank " + std::to_string(m->halo_peers[i].peer) +
                      " disagree on the number of shared " + (which ? "PML unknowns (" : "interface nodes (") +
                      std::to_string((long long)mine[i]) + " vs " + std::to_string((long long)theirs[i]) + "): the halo lists are not mirror images");
            return 1;
        }
    return 0;
}

// after the communicator exists: rank-ordered source lists of the shared unknowns, the ownership mask of the dot products,
// and the row / column scaling from the diagonal of Keff summed over the ranks
int pmlx_setup(svlgpu_model *m) {
    PmlDev &P = m->pml;
    HaloDev &h = m->halo;
    if (!P.present || !P.d_raw) return 0;                 // no PML block, or one that was planned for a single rank
    if (verify_peer_counts(m, 1)) return 1;
    const int rank = h.rank;
    std::vector<std::vector<std::pair<int, int>>> srcs(P.nc);      // per unknown: (rank, index into rbuf | -1)
    std::vector<double> own(std::max(1, P.nc), 1.0);
    for (auto &hp : m->halo_peers)
        for (size_t j = 0; j < hp.unk.size(); j++) {
            const int c = hp.unk[j];
            if (c < 0 || c >= P.nc) { set_error("comm_init: PML exchange list refers to an unknown that does not exist"); return 1; }
            srcs[c].push_back({hp.peer, hp.unk_offset + (int)j});
            if (hp.peer < rank) own[c] = 0.0;
        }
    std::vector<int32_t> unk, ptr(1, 0), src;
    for (int c = 0; c < P.nc; c++) {
        if (srcs[c].empty()) continue;
        srcs[c].push_back({rank, -1});
        std::sort(srcs[c].begin(), srcs[c].end());
        unk.push_back(c);
        for (auto &pr : srcs[c]) src.push_back(pr.second);
        ptr.push_back((int32_t)src.size());
    }
    P.n_xu = (int)unk.size();
    int32_t *d_unk = nullptr, *d_ptr = nullptr, *d_src = nullptr;
    CUDA_OK(cudaMalloc(&d_unk, sizeof(int32_t) * std::max<size_t>(1, unk.size())));
    CUDA_OK(cudaMalloc(&d_ptr, sizeof(int32_t) * ptr.size()));
    CUDA_OK(cudaMalloc(&d_src, sizeof(int32_t) * std::max<size_t>(1, src.size())));
    m->allocs.push_back(d_unk); m->allocs.push_back(d_ptr); m->allocs.push_back(d_src);
    CUDA_OK(cudaMemcpy(d_unk, unk.data(), sizeof(int32_t) * unk.size(), cudaMemcpyHostToDevice));
    CUDA_OK(cudaMemcpy(d_ptr, ptr.data(), sizeof(int32_t) * ptr.size(), cudaMemcpyHostToDevice));
    CUDA_OK(cudaMemcpy(d_src, src.data(), sizeof(int32_t) * src.size(), cudaMemcpyHostToDevice));
    P.d_x_unk = d_unk; P.d_x_ptr = d_ptr; P.d_x_src = d_src;
    CUDA_OK(cudaMemcpy(P.d_own, own.data(), sizeof(double) * P.nc, cudaMemcpyHostToDevice));
    // global diagonal -> scaling
    if ((int)P.h_diag.size() != P.nc) { set_error("comm_init: PML diagonal missing"); return 1; }
    CUDA_OK(cudaMemcpy(P.d_raw, P.h_diag.data(), sizeof(double) * P.nc, cudaMemcpyHostToDevice));
    if (pmlx_exchange(m, m->stream)) return 1;
    std::vector<double> tot(std::max(1, P.nc));
    CUDA_OK(cudaMemcpyAsync(tot.data(), P.d_raw, sizeof(double) * P.nc, cudaMemcpyDeviceToHost, m->stream));
    CUDA_OK(cudaStreamSynchronize(m->stream));
    for (int c = 0; c < P.nc; c++)
        if (tot[c] == 0.0 || !(tot[c] == tot[c])) { set_error("PML block: zero diagonal in Keff (EigenSolver.cpp:52-55 would fail too)"); return 1; }
    if (pml_rescale(m)) return 1;
    CUDA_OK(cudaStreamSynchronize(m->stream));
    P.multi = true;
    return 0;
}

void halo_destroy(svlgpu_model *m) {
    HaloDev &h = m->halo;
    if (h.comm && g_nccl.CommDestroy) g_nccl.CommDestroy(h.comm);
    if (h.comm_stream) cudaStreamDestroy(h.comm_stream);
    if (h.e_ready) cudaEventDestroy(h.e_ready);
    if (h.e_done) cudaEventDestroy(h.e_done);
    h.comm = nullptr; h.comm_stream = nullptr; h.e_ready = h.e_done = nullptr;
}

}  // namespace svl
