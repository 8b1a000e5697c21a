/* TEST INFRASTRUCTURE -- a recording stand-in for libsvlgpu.so (no GPU, no physics).
 *
 * Exports every entry point of include/svlgpu.h.  Builder calls append one line per call to the file named by
 * $SVLGPU_TRACE (".<RANK>" appended when RANK is set): the call name, its scalar arguments and an FNV-1a digest of
 * every array argument over exactly the extent the ABI defines.  Run-time calls succeed and return zeros, so that the
 * C++ host driver (LD_PRELOAD) and the ctypes binding (library path patched by the test) run to completion on a CPU
 * box.  The CPU tests compare traces: the same model reaching the C ABI through different front ends (JSON vs binary
 * tables, C++ driver on the reference's per-rank files vs the Python partitioner) must produce the same calls.
 * Only tests/ builds and loads this file; the product never does (svl_b200/capi.py loads svl_b200/libsvlgpu.so). */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "../../include/svlgpu.h"

struct svlgpu_model {
    int ndim, n_nodes, n_total, keep_gauss, n_force_passes;
    int32_t *ndof;
    int nrec;
    int rec_width[64], rec_rows[64];
};

static uint64_t fnv(const void *p, size_t n) {
    const unsigned char *c = (const unsigned char *)p;
    uint64_t h = 1469598103934665603ull;
    for (size_t i = 0; i < n; i++) { h ^= c[i]; h *= 1099511628211ull; }
    return h;
}
static FILE *out(void) {
    static FILE *f = NULL;
    if (!f) {
        const char *p = getenv("SVLGPU_TRACE"), *r = getenv("RANK");
        char path[4096];
        if (!p) return stderr;
        if (r) snprintf(path, sizeof path, "%s.%s", p, r); else snprintf(path, sizeof path, "%s", p);
        f = fopen(path, "w");
        if (!f) return stderr;
    }
    return f;
}
#define H(ptr, count) ((ptr) ? (unsigned long long)fnv((ptr), sizeof(*(ptr)) * (size_t)(count)) : 0ull)

svlgpu_model *svlgpu_create(int ndim, int lumped) {
    svlgpu_model *m = (svlgpu_model *)calloc(1, sizeof *m);
    m->ndim = ndim;
    fprintf(out(), "create ndim=%d lumped=%d\n", ndim, lumped);
    return m;
}
void svlgpu_destroy(svlgpu_model *m) { if (m) { free(m->ndof); free(m); } fflush(out()); }
const char *svlgpu_last_error(void) { return "trace shim"; }
int svlgpu_set_nodes(svlgpu_model *m, int n, const int32_t *ndof, const double *coords, const int32_t *totaldof,
                     const int32_t *freedof, int ntotal, int nfree) {
    size_t S = 0;
    for (int i = 0; i < n; i++) S += (size_t)ndof[i];
    m->n_nodes = n; m->n_total = ntotal;
    m->ndof = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n > 0 ? n : 1));
    memcpy(m->ndof, ndof, sizeof(int32_t) * (size_t)n);
    fprintf(out(), "set_nodes n=%d ntotal=%d nfree=%d ndof=%llx coords=%llx total=%llx free=%llx\n", n, ntotal, nfree, H(ndof, n),
            H(coords, (size_t)n * m->ndim), H(totaldof, S), H(freedof, S));
    return 0;
}
int svlgpu_add_nodal_mass(svlgpu_model *m, int n, const int32_t *node, const double *mass) {
    size_t S = 0;
    for (int i = 0; i < n; i++) S += (size_t)m->ndof[node[i]];
    fprintf(out(), "add_nodal_mass n=%d node=%llx mass=%llx\n", n, H(node, n), H(mass, S));
    return 0;
}
int svlgpu_add_constraint(svlgpu_model *m, int tag, int slave, int nmaster, const int32_t *master, const double *factor) {
    (void)m;
    fprintf(out(), "add_constraint tag=%d slave=%d nmaster=%d master=%llx factor=%llx\n", tag, slave, nmaster, H(master, nmaster), H(factor, nmaster));
    return 0;
}
int svlgpu_add_material(svlgpu_model *m, int kind, const double *params, int nparams) {
    static int count = 0;
    (void)m;
    fprintf(out(), "add_material kind=%d n=%d params=%llx\n", kind, nparams, H(params, nparams));
    return count++;
}
int svlgpu_add_elements(svlgpu_model *m, int kind, int n, const int32_t *conn, const int32_t *material, const double *attrs, int nattr) {
    static int count = 0;
    const int npe = (kind == SVLGPU_LIN3DHEXA8 || kind == SVLGPU_PML3DHEXA8) ? 8 : (kind == SVLGPU_ZEROLENGTH1D ? 2 : 4);
    (void)m;
    fprintf(out(), "add_elements kind=%d n=%d nattr=%d conn=%llx material=%llx attrs=%llx\n", kind, n, nattr, H(conn, (size_t)n * npe),
            H(material, n), nattr ? H(attrs, (size_t)n * nattr) : 0ull);
    const int first = count;
    count += n;
    return first;
}
int svlgpu_set_rayleigh(svlgpu_model *m, int n, const int32_t *elems, double am, double ak) {
    (void)m;
    fprintf(out(), "set_rayleigh n=%d am=%.17g ak=%.17g elems=%llx\n", n, am, ak, H(elems, n));
    return 0;
}
int svlgpu_hint_structured_block(svlgpu_model *m, int node0, int nx, int ny, int nz) {
    (void)m;
    fprintf(out(), "hint node0=%d nx=%d ny=%d nz=%d\n", node0, nx, ny, nz);
    return 0;
}
int svlgpu_set_option(svlgpu_model *m, const char *name, double value) {
    fprintf(out(), "set_option %s=%.17g\n", name, value);
    if (strcmp(name, "keep_gauss") == 0) m->keep_gauss = value != 0.0;
    return 0;
}
int svlgpu_add_point_load(svlgpu_model *m, int nnodes, const int32_t *nodes, int ndir, const double *dir, int nt, const double *series, double factor) {
    (void)m;
    fprintf(out(), "add_point_load nnodes=%d ndir=%d nt=%d factor=%.17g nodes=%llx dir=%llx series=%llx\n", nnodes, ndir, nt, factor,
            H(nodes, nnodes), H(dir, ndir), H(series, nt));
    return 0;
}
int svlgpu_add_drm_load(svlgpu_model *m, int nelems, const int32_t *elems, int nnodes, const int32_t *nodes, const uint8_t *exterior,
                        int nt, const double *field, double factor) {
    fprintf(out(), "add_drm_load nelems=%d nnodes=%d nt=%d factor=%.17g elems=%llx nodes=%llx ext=%llx field=%llx\n", nelems, nnodes, nt, factor,
            H(elems, nelems), H(nodes, nnodes), H(exterior, nnodes), H(field, (size_t)nnodes * nt * 3 * m->ndim));
    return 0;
}
int svlgpu_add_drm_planewave(svlgpu_model *m, int nelems, const int32_t *elems, int nnodes, const int32_t *nodes, const uint8_t *exterior,
                             const double *dir, const double *pol, const double *xref, double c, double f0, double t0, double amp, double factor) {
    (void)m;
    fprintf(out(), "add_drm_planewave nelems=%d nnodes=%d c=%.17g f0=%.17g t0=%.17g amp=%.17g factor=%.17g elems=%llx nodes=%llx ext=%llx dir=%llx pol=%llx xref=%llx\n",
            nelems, nnodes, c, f0, t0, amp, factor, H(elems, nelems), H(nodes, nnodes), H(exterior, nnodes), H(dir, 3), H(pol, 3), H(xref, 3));
    return 0;
}
int svlgpu_add_node_recorder(svlgpu_model *m, int field, int nnodes, const int32_t *nodes, int max_rows) {
    int w = 0;
    for (int i = 0; i < nnodes; i++) w += m->ndof[nodes[i]];
    if (m->nrec >= 64) return -1;
    m->rec_width[m->nrec] = w; m->rec_rows[m->nrec] = 0;
    fprintf(out(), "add_node_recorder field=%d nnodes=%d nodes=%llx\n", field, nnodes, H(nodes, nnodes));
    (void)max_rows;
    return m->nrec++;
}
int svlgpu_finalize(svlgpu_model *m, double dt, int device) {
    (void)m; (void)device;
    fprintf(out(), "finalize dt=%.17g\n", dt);
    fflush(out());
    return 0;
}
int svlgpu_set_initial_state(svlgpu_model *m, const double *U, const double *V, const double *A) {
    fprintf(out(), "set_initial_state U=%llx V=%llx A=%llx\n", H(U, m->n_total), H(V, m->n_total), H(A, m->n_total));
    return 0;
}
int svlgpu_step(svlgpu_model *m, int k_begin, int k_end, int sync) {
    (void)sync;
    for (int r = 0; r < m->nrec; r++) m->rec_rows[r] += k_end - k_begin;
    return 0;
}
int svlgpu_sync(svlgpu_model *m) { (void)m; return 0; }
int svlgpu_step_host(svlgpu_model *m, int k, const double *amplitudes, int nloads, int rec, double *row, int row_len) {
    (void)k; (void)amplitudes; (void)nloads; (void)rec;
    for (int r = 0; r < m->nrec; r++) m->rec_rows[r] += 1;
    if (row) memset(row, 0, sizeof(double) * (size_t)row_len);
    return 0;
}
int svlgpu_get_state(svlgpu_model *m, int field, const int32_t *dofs, int n, double *outv) {
    (void)field;
    memset(outv, 0, sizeof(double) * (size_t)(dofs ? n : m->n_total));
    return 0;
}
int svlgpu_internal_force(svlgpu_model *m, double *F) { memset(F, 0, sizeof(double) * (size_t)m->n_total); return 0; }
int svlgpu_get_mass_diagonal(svlgpu_model *m, double *M) { memset(M, 0, sizeof(double) * (size_t)m->n_total); return 0; }
int svlgpu_get_gauss(svlgpu_model *m, int field, int nelem, const int32_t *elems, double *outv) {
    (void)field; (void)elems;
    if (!m->keep_gauss) return 1;                      /* as the library: not kept unless asked for before finalize */
    memset(outv, 0, sizeof(double) * (size_t)nelem * (m->ndim == 3 ? 48 : 12));
    return 0;
}
int svlgpu_read_recorder(svlgpu_model *m, int rec, int r0, int r1, double *outv) {
    memset(outv, 0, sizeof(double) * (size_t)(r1 - r0) * (size_t)m->rec_width[rec]);
    return 0;
}
int svlgpu_recorder_rows(svlgpu_model *m, int rec) { return m->rec_rows[rec]; }
int svlgpu_recorder_width(svlgpu_model *m, int rec) { return m->rec_width[rec]; }
int svlgpu_get_counters(svlgpu_model *m, svlgpu_counters *c) { (void)m; memset(c, 0, sizeof *c); return 0; }
int svlgpu_set_kernel_timing(svlgpu_model *m, int on) { (void)m; (void)on; return 0; }
int svlgpu_kernel_time(svlgpu_model *m, int which, double *avg_ms, int64_t *launches, int reset) {
    (void)m; (void)which; (void)reset;
    if (avg_ms) *avg_ms = 0.0;
    if (launches) *launches = 0;
    return 0;
}
int svlgpu_device_ptr(svlgpu_model *m, int which, void **ptr, int64_t *len) { (void)m; (void)which; *ptr = NULL; *len = 0; return 1; }
int svlgpu_add_halo(svlgpu_model *m, int peer, int nnodes, const int32_t *nodes) {
    (void)m;
    fprintf(out(), "add_halo peer=%d nnodes=%d nodes=%llx\n", peer, nnodes, H(nodes, nnodes));
    return 0;
}
int svlgpu_nccl_unique_id(void *out128) { memset(out128, 0x5a, 128); return 0; }
int svlgpu_comm_init(svlgpu_model *m, const void *id128, int rank, int nranks) {
    (void)m;
    fprintf(out(), "comm_init rank=%d nranks=%d id=%llx\n", rank, nranks, (unsigned long long)fnv(id128, 128));
    fflush(out());
    return 0;
}
int svlgpu_add_support_motion(svlgpu_model *m, int node, int dof, int nt, const double *series, double factor) {
    (void)m;
    fprintf(out(), "add_support_motion node=%d dof=%d nt=%d factor=%.17g series=%llx\n", node, dof, nt, factor, H(series, nt));
    return 0;
}
int svlgpu_measure_peaks(int device, double *fp64_tflops, double *copy_gbs) { (void)device; (void)fp64_tflops; (void)copy_gbs; return 1; }
