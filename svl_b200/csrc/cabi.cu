// cabi.cu -- extern "C" surface of libsvlgpu.so (include/svlgpu.h).  Thin: argument checks,
// copies into the host model, and forwarding to the planner / kernels.  No exceptions cross
// the boundary; failures return non-zero ("stop") and set svlgpu_last_error().
#include <cmath>
#include <cstring>
#include <new>
#include <string>
#include "model.h"

namespace svl {
static thread_local std::string g_err;
void set_error(const std::string &s) { g_err = s; }
}  // namespace svl
using namespace svl;

#define REQUIRE(c, msg)                    \
    do {                                   \
        if (!(c)) { set_error(msg); return 1; } \
    } while (0)
#define GUARD_BEGIN try {
#define GUARD_END                                                        \
    }                                                                    \
    catch (const std::bad_alloc &) { set_error("host out of memory"); return 1; } \
    catch (const std::exception &e) { set_error(e.what()); return 1; }

extern "C" {

const char *svlgpu_last_error(void) { return g_err.c_str(); }

svlgpu_model *svlgpu_create(int ndim, int lumped) {
    if (ndim != 2 && ndim != 3) { set_error("ndim must be 2 or 3"); return nullptr; }
    svlgpu_model *m = new (std::nothrow) svlgpu_model();
    if (!m) { set_error("host out of memory"); return nullptr; }
    m->ndim = ndim; m->lumped = lumped ? 1 : 0;
    return m;
}

void svlgpu_destroy(svlgpu_model *m) {
    if (!m) return;
    forget_const_owner(m);
    if (m->finalized || m->stream) {
        cudaSetDevice(m->device);
        if (m->stream) cudaStreamSynchronize(m->stream);
        graph_destroy(m);
        halo_destroy(m);
        pml_destroy(m);
        newmark_destroy(m);
        for (void *p : m->allocs) cudaFree(p);
        if (m->h_pl_amp) cudaFreeHost(m->h_pl_amp);
        if (m->h_row) cudaFreeHost(m->h_row);
        for (auto &t : m->timers) { if (t.e0) cudaEventDestroy(t.e0); if (t.e1) cudaEventDestroy(t.e1); }
        if (m->ev0) cudaEventDestroy(m->ev0);
        if (m->ev1) cudaEventDestroy(m->ev1);
        for (auto &d : m->drm_dev) for (auto &e : d.ev_ready) if (e) cudaEventDestroy(e);
        for (auto &st : m->side) if (st) { cudaStreamSynchronize(st); cudaStreamDestroy(st); }
        for (cudaEvent_t e : {m->ev_fork, m->ev_fork2, m->ev_join}) if (e) cudaEventDestroy(e);
        if (m->stream) cudaStreamDestroy(m->stream);
    }
    delete m;
}

int svlgpu_set_nodes(svlgpu_model *m, int n, const int32_t *ndof, const double *coords,
                     const int32_t *totaldof, const int32_t *freedof, int ntotal, int nfree) {
    GUARD_BEGIN
    REQUIRE(m && !m->finalized, "set_nodes: model missing or already finalized");
    REQUIRE(n > 0 && ndof && coords && totaldof && freedof, "set_nodes: null argument");
    m->n_nodes = n; m->n_total = ntotal; m->n_free = nfree;
    m->node_ndof.assign(ndof, ndof + n);
    m->node_ptr.assign(n + 1, 0);
    for (int i = 0; i < n; i++) {
        REQUIRE(ndof[i] > 0 && ndof[i] <= 9, "set_nodes: ndof out of range");
        m->node_ptr[i + 1] = m->node_ptr[i] + ndof[i];
    }
    const int nd = m->node_ptr[n];
    REQUIRE(nd == ntotal, "set_nodes: sum of ndof differs from ntotal");
    m->coords.assign(coords, coords + (size_t)n * m->ndim);
    m->totaldof.assign(totaldof, totaldof + nd);
    m->freedof.assign(freedof, freedof + nd);
    return 0;
    GUARD_END
}

int svlgpu_add_nodal_mass(svlgpu_model *m, int n, const int32_t *node, const double *mass) {
    GUARD_BEGIN
    REQUIRE(m && !m->finalized && m->n_nodes, "add_nodal_mass: set nodes first");
    size_t off = 0;
    for (int i = 0; i < n; i++) {
        REQUIRE(node[i] >= 0 && node[i] < m->n_nodes, "add_nodal_mass: node out of range");
        const int nd = m->node_ndof[node[i]];
        m->masses.push_back({node[i], std::vector<double>(mass + off, mass + off + nd)});
        off += nd;
    }
    return 0;
    GUARD_END
}

int svlgpu_add_constraint(svlgpu_model *m, int tag, int slave_total_dof, int nmaster,
                          const int32_t *master_free_dof, const double *factor) {
    GUARD_BEGIN
    REQUIRE(m && !m->finalized, "add_constraint: model missing or finalized");
    svlgpu_model::Constraint c;
    c.tag = tag; c.slave = slave_total_dof;
    c.master.assign(master_free_dof, master_free_dof + nmaster);
    c.factor.assign(factor, factor + nmaster);
    m->constraints.push_back(c);
    return 0;
    GUARD_END
}

int svlgpu_add_material(svlgpu_model *m, int kind, const double *params, int nparams) {
    try {
        if (!m || m->finalized || nparams > 8 || nparams < (kind == SVLGPU_VISCOUS1DLINEAR ? 1 : 3)) { set_error("add_material: bad arguments"); return -1; }
        Material mat;
        mat.kind = kind;
        for (int i = 0; i < 8; i++) mat.p[i] = (i < nparams) ? params[i] : 0.0;
        m->materials.push_back(mat);
        return (int)m->materials.size() - 1;
    } catch (...) { set_error("add_material: host out of memory"); return -1; }
}

int svlgpu_add_elements(svlgpu_model *m, int kind, int n, const int32_t *conn, const int32_t *material,
                        const double *attrs, int nattr) {
    try {
        if (!m || m->finalized || !m->n_nodes || n <= 0 || !conn || !material) { set_error("add_elements: bad arguments"); return -1; }
        if (kind < SVLGPU_LIN3DHEXA8 || kind > SVLGPU_ZEROLENGTH1D) { set_error("add_elements: unknown element kind"); return -1; }
        const int npe = (kind == SVLGPU_LIN3DHEXA8 || kind == SVLGPU_PML3DHEXA8) ? 8 : (kind == SVLGPU_ZEROLENGTH1D) ? 2 : 4;
        if (kind != SVLGPU_ZEROLENGTH1D && (npe == 8) != (m->ndim == 3)) { set_error("add_elements: element kind does not match ndim"); return -1; }
        if (kind == SVLGPU_ZEROLENGTH1D && (nattr < 1 || !attrs)) { set_error("add_elements: ZeroLength1D needs its direction"); return -1; }
        if (nattr > 10 || (nattr > 0 && !attrs)) { set_error("add_elements: bad attrs"); return -1; }
        const int first = (int)m->elem_kind.size();
        m->elem_kind.resize(first + n, kind);
        m->elem_conn.resize(8ull * (first + n), 0);
        m->elem_mat.resize(first + n);
        if (nattr > 0 || kind == SVLGPU_LIN2DQUAD4 || !m->elem_attr.empty()) m->elem_attr.resize(10ull * (first + n), 0.0);
        if (!m->elem_am.empty()) { m->elem_am.resize(first + n, 0.0); m->elem_ak.resize(first + n, 0.0); }
        for (int e = 0; e < n; e++) {
            for (int l = 0; l < npe; l++) {
                const int nd = conn[(size_t)e * npe + l];
                if (nd < 0 || nd >= m->n_nodes) { set_error("add_elements: node index out of range"); return -1; }
                m->elem_conn[8ull * (first + e) + l] = nd;
            }
            if (material[e] < 0 || material[e] >= (int)m->materials.size()) { set_error("add_elements: material index out of range"); return -1; }
            m->elem_mat[first + e] = material[e];
            for (int a = 0; a < nattr; a++) m->elem_attr[10ull * (first + e) + a] = attrs[(size_t)e * nattr + a];
            if (kind == SVLGPU_LIN2DQUAD4 && nattr == 0) m->elem_attr[10ull * (first + e)] = 1.0;
        }
        return first;
    } catch (...) { set_error("add_elements: host out of memory"); return -1; }
}

int svlgpu_set_rayleigh(svlgpu_model *m, int n, const int32_t *elems, double am, double ak) {
    GUARD_BEGIN
    REQUIRE(m && !m->finalized, "set_rayleigh: model missing or finalized");
    // ak != 0 makes the CentralDifference Keff non-diagonal (refused at finalize); NewmarkBeta takes it (newmark.cu)
    m->elem_am.resize(m->elem_kind.size(), 0.0);
    m->elem_ak.resize(m->elem_kind.size(), 0.0);
    for (int i = 0; i < n; i++) {
        REQUIRE(elems[i] >= 0 && elems[i] < (int)m->elem_kind.size(), "set_rayleigh: element out of range");
        m->elem_am[elems[i]] = am;
        m->elem_ak[elems[i]] = ak;
    }
    return 0;
    GUARD_END
}

int svlgpu_set_option(svlgpu_model *m, const char *name, double value) {
    GUARD_BEGIN
    REQUIRE(m && name && !m->finalized, "set_option: model missing or already finalized");
    const std::string n(name);
    if (n == "lattice_guess") m->opt_lattice_guess = value != 0.0;
    else if (n == "keep_gauss") m->opt_keep_gauss = value != 0.0;
    else if (n == "nbr_classes") m->opt_nbr_classes = value != 0.0;
    else if (n == "renumber") m->opt_renumber = value != 0.0;
    else if (n == "reaction_collective") m->opt_reaction_collective = value != 0.0;
    else if (n == "cuda_graph") m->opt_graph = value != 0.0 ? 1 : 0;
    else if (n == "integrator") { REQUIRE(value == 0.0 || value == 1.0, "set_option: integrator must be 0 (CentralDifference) or 1 (NewmarkBeta)"); m->opt_integrator = (int)value; }
    else if (n == "newmark_rtol") { REQUIRE(value > 0.0 && value < 1.0, "set_option: newmark_rtol out of range"); m->nm.rtol = value; }
    else if (n == "pml_rtol") { REQUIRE(value > 0.0 && value < 1.0, "set_option: pml_rtol out of range"); m->pml.rtol = value; }
    else if (n == "pml_collective") m->pml.collective = value != 0.0;
    else if (n == "ftol") { REQUIRE(value >= 0.0, "set_option: ftol must be >= 0"); m->pml.ftol = value; }
    else { set_error("set_option: unknown option " + n); return 1; }
    return 0;
    GUARD_END
}

int svlgpu_hint_structured_block(svlgpu_model *m, int node0, int nx, int ny, int nz) {
    GUARD_BEGIN
    REQUIRE(m && !m->finalized, "hint: model missing or finalized");
    m->hints.push_back({node0, nx, ny, nz});
    return 0;
    GUARD_END
}

int svlgpu_add_point_load(svlgpu_model *m, int nnodes, const int32_t *nodes, int ndir, const double *dir,
                          int nt, const double *series, double factor) {
    GUARD_BEGIN
    REQUIRE(m && !m->finalized && nnodes > 0 && nt > 0 && ndir >= m->ndim, "add_point_load: bad arguments");
    PointLoad pl;
    pl.nodes.assign(nodes, nodes + nnodes);
    for (int n : pl.nodes) REQUIRE(n >= 0 && n < m->n_nodes, "add_point_load: node out of range");
    for (int c = 0; c < 3; c++) pl.dir[c] = (c < ndir) ? dir[c] : 0.0;
    pl.series.assign(series, series + nt);
    pl.factor = factor;
    m->ploads.push_back(std::move(pl));
    return 0;
    GUARD_END
}

int svlgpu_add_drm_load(svlgpu_model *m, int nelems, const int32_t *elems, int nnodes, const int32_t *nodes,
                        const uint8_t *exterior, int nt, const double *field, double factor) {
    GUARD_BEGIN
    REQUIRE(m && !m->finalized && nelems > 0 && nnodes > 0 && nt > 0 && field, "add_drm_load: bad arguments");
    DrmLoad d;
    d.elems.assign(elems, elems + nelems);
    d.nodes.assign(nodes, nodes + nnodes);
    d.ext.assign(exterior, exterior + nnodes);
    d.nt = nt;
    d.field.assign(field, field + (size_t)nnodes * nt * 3 * m->ndim);
    d.factor = factor;
    m->drms.push_back(std::move(d));
    return 0;
    GUARD_END
}

int svlgpu_add_drm_planewave(svlgpu_model *m, int nelems, const int32_t *elems, int nnodes, const int32_t *nodes,
                             const uint8_t *exterior, const double *dir, const double *pol, const double *xref,
                             double c, double f0, double t0, double amp, double factor) {
    GUARD_BEGIN
    REQUIRE(m && !m->finalized && nelems > 0 && nnodes > 0 && c > 0, "add_drm_planewave: bad arguments");
    DrmLoad d;
    d.elems.assign(elems, elems + nelems);
    d.nodes.assign(nodes, nodes + nnodes);
    d.ext.assign(exterior, exterior + nnodes);
    d.analytic = true;
    for (int i = 0; i < 3; i++) { d.dir[i] = i < m->ndim ? dir[i] : 0; d.pol[i] = i < m->ndim ? pol[i] : 0; d.xref[i] = i < m->ndim ? xref[i] : 0; }
    d.c = c; d.f0 = f0; d.t0 = t0; d.amp = amp; d.factor = factor;
    m->drms.push_back(std::move(d));
    return 0;
    GUARD_END
}

int svlgpu_add_node_recorder(svlgpu_model *m, int field, int nnodes, const int32_t *nodes, int max_rows) {
    try {
        if (!m || m->finalized || nnodes <= 0 || field < 0 || field > 3) { set_error("add_node_recorder: bad arguments (disp / vel / accel / reaction)"); return -1; }
        Recorder r;
        r.field = field; r.nodes.assign(nodes, nodes + nnodes); r.max_rows = max_rows;
        for (int n : r.nodes) if (n < 0 || n >= m->n_nodes) { set_error("add_node_recorder: node out of range"); return -1; }
        m->recorders.push_back(std::move(r));
        return (int)m->recorders.size() - 1;
    } catch (...) { set_error("add_node_recorder: host out of memory"); return -1; }
}

int svlgpu_add_support_motion(svlgpu_model *m, int node, int dof, int nt, const double *series, double factor) {
    GUARD_BEGIN
    REQUIRE(m && !m->finalized && m->n_nodes && nt > 0 && series, "add_support_motion: bad arguments (set nodes first, before finalize)");
    REQUIRE(node >= 0 && node < m->n_nodes && dof >= 0 && dof < m->node_ndof[node], "add_support_motion: node / dof out of range");
    SupportMotion sm;
    sm.node = node; sm.dof = dof; sm.series.assign(series, series + nt); sm.factor = factor;
    m->supports.push_back(std::move(sm));
    return 0;
    GUARD_END
}

int svlgpu_set_initial_state(svlgpu_model *m, const double *U, const double *V, const double *A) {
    GUARD_BEGIN
    REQUIRE(m && !m->finalized && m->n_total, "set_initial_state: set nodes first / before finalize");
    if (U) m->U0.assign(U, U + m->n_total);
    if (V) m->V0.assign(V, V + m->n_total);
    if (A) m->A0.assign(A, A + m->n_total);
    return 0;
    GUARD_END
}

int svlgpu_finalize(svlgpu_model *m, double dt, int device) {
    GUARD_BEGIN
    REQUIRE(m && !m->finalized, "finalize: model missing or already finalized");
    REQUIRE(m->n_nodes > 0 && !m->elem_kind.empty(), "finalize: empty model");
    REQUIRE(dt > 0.0, "finalize: dt must be positive");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { set_error("no CUDA device: libsvlgpu has no CPU fallback"); return 1; }
    REQUIRE(device >= 0 && device < ndev, "finalize: device ordinal out of range");
    m->dt = dt; m->device = device;
    return plan_and_upload(m);
    GUARD_END
}

int svlgpu_step(svlgpu_model *m, int k_begin, int k_end, int sync) {
    GUARD_BEGIN
    REQUIRE(m && m->finalized, "step: model not finalized");
    cudaSetDevice(m->device);
    cudaEventRecord(m->ev0, m->stream);
    if (run_steps(m, k_begin, k_end, nullptr)) return 1;
    cudaEventRecord(m->ev1, m->stream);
    m->steps_since_check += k_end - k_begin;
    if (sync) return svlgpu_sync(m);
    return 0;
    GUARD_END
}

int svlgpu_sync(svlgpu_model *m) {
    REQUIRE(m && m->finalized, "sync: model not finalized");
    cudaError_t e = cudaStreamSynchronize(m->stream);
    if (e != cudaSuccess) { set_error(std::string("device failure: ") + cudaGetErrorString(e)); return 1; }
    float ms = 0;
    if (cudaEventElapsedTime(&ms, m->ev0, m->ev1) == cudaSuccess) m->last_step_ms = ms;
    timer_flush(m);
    // NaN / Inf in U -> stop; one read of U_n per synchronised call, outside the event-timed region
    if (m->steps_since_check > 0) { m->steps_since_check = 0; if (state_is_finite(m)) return 1; }
    return 0;
}

int svlgpu_step_host(svlgpu_model *m, int k, const double *amplitudes, int nloads, int rec, double *row_out,
                     int row_len) {
    GUARD_BEGIN
    REQUIRE(m && m->finalized, "step_host: model not finalized");
    REQUIRE(nloads == m->n_ploads, "step_host: amplitude count differs from the number of point loads");
    cudaSetDevice(m->device);
    // host buffers, zero-copy: the step's load amplitudes are read by k_nodal_loads straight from mapped pinned host
    // memory and the recorder row is written by k_record straight into mapped pinned host memory, so the bytes
    // cross the bus inside the step without two extra copy-engine submissions
    const double *damp = nullptr;
    if (nloads > 0) {
        std::memcpy(m->h_pl_amp, amplitudes, sizeof(double) * nloads);
        damp = m->h_pl_amp;
    }
    if (rec >= 0) {
        REQUIRE(rec < (int)m->recorders.size(), "step_host: recorder out of range");
        Recorder &r = m->recorders[rec];
        REQUIRE(row_len == r.width && r.rows < r.max_rows, "step_host: row length differs from the recorder width");
        m->mirror_rec = rec;
    }
    m->host_step = true;
    const int rc = run_steps(m, k, k + 1, damp);
    m->host_step = false;
    m->mirror_rec = -1;
    if (rc) return 1;
    cudaError_t e = cudaStreamSynchronize(m->stream);
    if (e != cudaSuccess) { set_error(std::string("device failure: ") + cudaGetErrorString(e)); return 1; }
    if (rec >= 0) {
        std::memcpy(row_out, m->h_row, sizeof(double) * row_len);
        for (int i = 0; i < row_len; i++)
            if (!(std::fabs(row_out[i]) <= 1.7976931348623157e308)) { set_error("NaN / Inf in recorded response"); return 1; }
    }
    if (++m->steps_since_check >= 256) { m->steps_since_check = 0; if (state_is_finite(m)) return 1; }   // whole state: every 256 calls
    return 0;
    GUARD_END
}

int svlgpu_get_state(svlgpu_model *m, int field, const int32_t *dofs, int n, double *out) {
    GUARD_BEGIN
    REQUIRE(m && m->finalized && out, "get_state: model not finalized");
    REQUIRE(field >= 0 && field <= 2, "get_state: field must be disp/vel/accel");
    if (!dofs) n = m->n_total;
    else for (int i = 0; i < n; i++) REQUIRE(dofs[i] >= 0 && dofs[i] < m->n_total, "get_state: dof out of range");
    cudaSetDevice(m->device);
    return gather_state(m, field, dofs, n, out);
    GUARD_END
}

int svlgpu_internal_force(svlgpu_model *m, double *F) {
    GUARD_BEGIN
    REQUIRE(m && m->finalized && F, "internal_force: model not finalized");
    cudaSetDevice(m->device);
    return compute_internal_force(m, F);
    GUARD_END
}

int svlgpu_get_mass_diagonal(svlgpu_model *m, double *Md) {
    GUARD_BEGIN
    REQUIRE(m && m->finalized && Md, "get_mass_diagonal: model not finalized");
    for (int t = 0; t < m->n_total; t++) Md[t] = m->h_mass[m->int_of_total[t]];
    return 0;
    GUARD_END
}

int svlgpu_get_gauss(svlgpu_model *m, int field, int nelem, const int32_t *elems, double *out) {
    GUARD_BEGIN
    REQUIRE(m && m->finalized && out, "get_gauss: model not finalized");
    REQUIRE(field == SVLGPU_STRAIN || field == SVLGPU_STRESS, "get_gauss: strain or stress only");
    cudaSetDevice(m->device);
    cudaStreamSynchronize(m->stream);
    for (int i = 0; i < nelem; i++) {
        const int e = elems[i];
        bool found = false;
        for (auto &g : m->gsets) {
            auto it = std::lower_bound(g.elems.begin(), g.elems.end(), e);
            if (it == g.elems.end() || *it != e) continue;
            REQUIRE(g.d_gp, "get_gauss: Gauss-point output not kept (set SVLGPU_KEEP_GAUSS=1 before finalize)");
            const int ncomp = (g.kind == SVLGPU_LIN3DHEXA8) ? 6 : 3;
            const long long ngp = (long long)g.n * g.ngp, q0 = (long long)(it - g.elems.begin()) * g.ngp;
            for (int c = 0; c < ncomp; c++) {
                std::vector<double> tmp(g.ngp);
                cudaMemcpy(tmp.data(), g.d_gp + ((field == SVLGPU_STRESS ? ncomp : 0) + c) * ngp + q0,
                           sizeof(double) * g.ngp, cudaMemcpyDeviceToHost);
                for (int gp = 0; gp < g.ngp; gp++) out[((size_t)i * g.ngp + gp) * ncomp + c] = tmp[gp];
            }
            found = true;
        }
        REQUIRE(found, "get_gauss: element is advanced by the block-stencil kernel (no Gauss-point data)");
    }
    return 0;
    GUARD_END
}

int svlgpu_read_recorder(svlgpu_model *m, int rec, int r0, int r1, double *out) {
    GUARD_BEGIN
    REQUIRE(m && m->finalized && rec >= 0 && rec < (int)m->recorders.size(), "read_recorder: bad recorder");
    Recorder &r = m->recorders[rec];
    REQUIRE(r0 >= 0 && r1 <= r.rows && r0 <= r1, "read_recorder: row range");
    cudaSetDevice(m->device);
    cudaStreamSynchronize(m->stream);
    cudaError_t e = cudaMemcpy(out, r.d_rows + (size_t)r0 * r.width, sizeof(double) * (size_t)(r1 - r0) * r.width,
                               cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { set_error(cudaGetErrorString(e)); return 1; }
    const size_t cnt = (size_t)(r1 - r0) * r.width;
    for (size_t i = 0; i < cnt; i++)
        if (!(std::fabs(out[i]) <= 1.7976931348623157e308)) { set_error("NaN / Inf in recorded response"); return 1; }
    return 0;
    GUARD_END
}
int svlgpu_recorder_rows(svlgpu_model *m, int rec) {
    if (!m || rec < 0 || rec >= (int)m->recorders.size()) return -1;
    return m->recorders[rec].rows;
}
int svlgpu_recorder_width(svlgpu_model *m, int rec) {
    if (!m || rec < 0 || rec >= (int)m->recorders.size()) return -1;
    return m->recorders[rec].width;
}

int svlgpu_get_counters(svlgpu_model *m, svlgpu_counters *o) {
    REQUIRE(m && o, "get_counters: null argument");
    std::memset(o, 0, sizeof(*o));
    o->n_elements = (int64_t)m->elem_kind.size(); o->n_nodes = m->n_nodes; o->n_total_dofs = m->n_total;
    o->n_block_nodes = m->n_block_nodes; o->n_generic_nodes = m->n_gnodes;
    o->n_generic_elements = m->n_generic_elements; o->n_elem_classes = m->n_elem_classes;
    o->n_node_classes = m->n_node_classes; o->launches_per_step = m->launches_per_step;
    o->total_launches = m->total_launches; o->device_bytes = m->device_bytes;
    o->last_step_ms = m->last_step_ms;
    o->stencil_ms = m->timers[0].launches ? m->timers[0].total_ms / m->timers[0].launches : 0.0;
    o->n_pml_elements = m->pml.n_elem; o->n_pml_unknowns = m->pml.nc;
    o->pml_solves = m->pml.solves + m->nm.solves; o->pml_iterations = m->pml.total_iters + m->nm.total_iters;
    o->n_nbr_nodes = m->nbr.n_nodes; o->n_nbr_classes = m->nbr.n_cls;
    return 0;
}

int svlgpu_set_kernel_timing(svlgpu_model *m, int on) {
    REQUIRE(m, "set_kernel_timing: null model");
    m->kernel_timing = on != 0;
    return 0;
}
int svlgpu_kernel_time(svlgpu_model *m, int which, double *avg_ms, int64_t *launches, int reset) {
    REQUIRE(m && which >= 0 && which < kNumTimers, "kernel_time: bad arguments");
    timer_flush(m);
    KernelTimer &t = m->timers[which];
    if (avg_ms) *avg_ms = t.launches ? t.total_ms / t.launches : 0.0;
    if (launches) *launches = t.launches;
    if (reset) { t.total_ms = 0.0; t.launches = 0; }
    return 0;
}

int svlgpu_device_ptr(svlgpu_model *m, int which, void **ptr, int64_t *len) {
    REQUIRE(m && m->finalized && ptr && which >= 0 && which < 3, "device_ptr: bad arguments");
    const int idx = which == 0 ? m->cur : which == 1 ? m->prev : m->next;
    *ptr = m->d_U[idx];
    if (len) *len = m->n_int;
    return 0;
}

int svlgpu_add_halo(svlgpu_model *m, int peer, int nnodes, const int32_t *nodes) {
    GUARD_BEGIN
    REQUIRE(m && !m->finalized && peer >= 0 && nnodes > 0 && nodes, "add_halo: bad arguments (call before finalize)");
    HaloPeer hp;
    hp.peer = peer;
    hp.nodes.assign(nodes, nodes + nnodes);
    for (int n : hp.nodes) REQUIRE(n >= 0 && n < m->n_nodes, "add_halo: node out of range");
    for (auto &o : m->halo_peers) REQUIRE(o.peer != peer, "add_halo: one list per peer");
    m->halo_peers.push_back(std::move(hp));
    return 0;
    GUARD_END
}
int svlgpu_nccl_unique_id(void *out128) {
    GUARD_BEGIN
    REQUIRE(out128, "nccl_unique_id: null argument");
    return halo_unique_id(out128);
    GUARD_END
}
int svlgpu_comm_init(svlgpu_model *m, const void *id128, int rank, int nranks) {
    GUARD_BEGIN
    REQUIRE(m && id128 && rank >= 0 && rank < nranks, "comm_init: bad arguments");
    REQUIRE(!m->halo.active, "comm_init: already initialised");
    return halo_comm_init(m, id128, rank, nranks);
    GUARD_END
}

}  // extern "C"
