// svl_host.cpp -- the reference's run-time, re-hosted on libsvlgpu (C++17, no third-party dependencies).
//
//   SeismoVLAB_gpu.exe -dir <Partition dir> -file '<name>.$.json'                    one partition, one GPU
//   SeismoVLAB_gpu.exe -np N -dir <Partition dir> -file '<name>.$.json'              N partitions, one process per GPU
//     (what `mpirun -np N SeismoVLAB.exe ...` is for the reference; -np forks the ranks itself, or start the N processes
//      yourself with RANK / LOCAL_RANK / WORLD_SIZE set, e.g. under torchrun; `-plan` stops after the partition plan)
//
// keeps the outer boundary of SeismoVLAB.exe verbatim: the same command line
// (12-Utilities/Utilities.hpp:97-161; '$' -> rank, Driver.hpp:250-276), the same per-rank JSON partition files
// written by 01-Pre_Process (schema: Driver.hpp:1981-2046), the same time-series / .drm text files
// (Driver.hpp:1514-1527, 1689-1721) and the same NODE recorder files under <dir>/../Solution/<combo>/
// (12-Utilities/Recorder.cpp:73-105, 239-269).  Inside, the reference's class surface is kept as thin host
// classes (Mesh, Node, Element, Material, Load, LoadCombo, Recorder, Assembler, CentralDifference, Linear,
// DynamicAnalysis) whose hot methods forward to the C ABI of include/svlgpu.h; no physics is computed here.
// Error convention as in the reference: methods return `true` to stop (Integrator::ComputeNewStep etc.).
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <initializer_list>
#include <string>
#include <strings.h>
#include <signal.h>
#include <sys/wait.h>
#include <time.h>
#include <tuple>
#include <unistd.h>
#include <vector>

#include "../../include/svlgpu.h"
#include "json.hpp"

using svlhost::JValue;

namespace {

bool ieq(const std::string &a, const char *b) { return strcasecmp(a.c_str(), b) == 0; }

// numeric-tag view of a JSON object, ascending tag (the reference's std::map<unsigned int, ...> order)
std::vector<std::pair<long, const JValue *>> by_tag(const JValue &o) {
    std::vector<std::pair<long, const JValue *>> v;
    for (auto &kv : o.obj) v.push_back({std::strtol(kv.first.c_str(), nullptr, 10), &kv.second});
    std::sort(v.begin(), v.end(), [](auto &a, auto &b) { return a.first < b.first; });
    return v;
}

// ---- Geometry module (01-Node ... 06-Mesh): plain data, the device owns the arithmetic ----------------------
// Per-node / per-element lists are short and bounded (<= 9 dofs, <= 8 nodes, <= 10 attributes): kept inline instead of one
// heap block each -- five std::vector per node were most of the 2.5 GB / 6 s the object graph cost at 4 M elements.
template <class T, int N> struct Small {
    T v[N]; unsigned char n = 0;
    size_t size() const { return n; }
    bool empty() const { return n == 0; }
    void clear() { n = 0; }
    void push_back(const T &x) { if (n >= N) throw std::length_error("entity list longer than the path supports"); v[n++] = x; }
    T &back() { return v[n - 1]; }
    T &operator[](size_t i) { return v[i]; }
    const T &operator[](size_t i) const { return v[i]; }
    T *begin() { return v; }
    T *end() { return v + n; }
    const T *begin() const { return v; }
    const T *end() const { return v + n; }
    template <class It> void assign(It a, It b) { n = 0; for (; a != b; ++a) push_back(*a); }
    Small &operator=(std::initializer_list<T> l) { n = 0; for (const T &x : l) push_back(x); return *this; }
};
struct Node {            // 01-Node/Node.cpp:3-21
    unsigned tag; int index; int ndof;
    Small<int, 9> total, free;           // as in the file: GLOBAL numbers when the model is partitioned
    Small<int, 9> ltotal, lfree;         // this partition's own compact numbering (several ranks only; else empty)
    Small<double, 3> coords;
};
struct Material { unsigned tag; int index; int kind; std::vector<double> par; };      // 02-Materials/Material.hpp
struct Element {         // 04-Elements/Element.hpp:51-249
    unsigned tag; int index; int kind; Small<unsigned, 8> conn; unsigned material; Small<double, 10> attr;
};
struct Load {            // 05-Loads/Load.cpp
    unsigned tag; bool drm = false, support = false;      // support: POINTLOAD_SUPPORT_MOTION (Driver.hpp:1725-1735)
    std::vector<unsigned> nodes, elements;
    std::vector<double> dir, series;
    std::string drm_pattern;
};
struct LoadCombo { unsigned tag; std::string name, folder; std::vector<unsigned> loads; std::vector<double> factors; };
struct RecorderSpec { std::string name, file, resp; int precision = 6, nsample = 1; std::vector<unsigned> ids; };

class Mesh {             // 06-Mesh/Mesh.cpp
  public:
    int ndim = 3, ntotal = 0, nfree = 0;
    int ntotal_dev = 0, nfree_dev = 0;   // what the device model is built with: the partition's own counts (== ntotal / nfree on one rank)
    bool lumped = true;
    bool pml_collective = false;         // several ranks and PML nodes anywhere in the model: every rank joins the block solve
    bool reaction_collective = false;    // several ranks and a REACTION recorder on any of them: every rank joins the reaction pass
    std::map<int, std::vector<int32_t>> Halos;                                             // peer rank -> shared nodes (device indices)
    std::map<unsigned, Node> Nodes;
    std::map<unsigned, Material> Materials;
    std::map<unsigned, Element> Elements;
    std::map<unsigned, Load> Loads;
    std::map<long, std::tuple<int, std::vector<int>, std::vector<double>>> Constraints;   // tag -> slave total, master free, factor
    std::map<unsigned, std::vector<double>> Masses;
    std::map<unsigned, std::pair<double, double>> Rayleigh;                               // element tag -> (am, ak)
    std::map<unsigned, std::vector<std::pair<int, std::vector<double>>>> Supports;         // node tag -> (dof, Xo) (Driver.hpp:509-563)
};

// ---- Driver (12-Utilities/Driver.hpp UpdateMesh :1981-2046) ------------------------------------------------------------
// binary sidecars of a partition file (svl_b200/model.py::pack_partition_tables): flat little-endian arrays
struct BinFile {
    std::ifstream f;
    bool open(const std::string &path, const char magic[4]) {
        f.open(path, std::ios::binary);
        char m[4]; uint32_t ver = 0;
        if (!f.is_open() || !f.read(m, 4) || std::memcmp(m, magic, 4) != 0 || !f.read((char *)&ver, 4) || ver != 1) {
            std::cout << "\x1B[31m ERROR: \x1B[0mcannot read binary table " << path << "\n";
            return false;
        }
        return true;
    }
    template <typename T> bool read(std::vector<T> &v, size_t n) {
        v.resize(n);
        return n == 0 || (bool)f.read((char *)v.data(), (std::streamsize)(n * sizeof(T)));
    }
    template <typename T> bool read1(T &v) { return (bool)f.read((char *)&v, sizeof(T)); }
};

bool UpdateMesh(Mesh &mesh, const JValue &J, const std::string &dir, bool geometry_only = false) {
    const JValue &G = J["Global"];
    mesh.ndim = G["ndim"].as_int(3);
    mesh.ntotal = G["ntotal"].as_int();
    mesh.nfree = G["nfree"].as_int();
    mesh.ntotal_dev = mesh.ntotal; mesh.nfree_dev = mesh.nfree;
    mesh.lumped = ieq(G["massform"].as_string("LUMPED"), "LUMPED");
    int idx = 0;
    if (J["Nodes"].has("binary")) {
        BinFile b;
        uint64_t n = 0; uint32_t nc = 0;
        std::vector<uint32_t> tags; std::vector<int32_t> ndof, tot, fre; std::vector<double> xyz;
        if (!b.open(dir + "/" + J["Nodes"]["binary"].as_string(), "SVLN") || !b.read1(n) || !b.read1(nc) || !b.read(tags, n) ||
            !b.read(ndof, n) || !b.read(xyz, n * nc)) return true;
        size_t S = 0;
        for (int32_t d : ndof) S += (size_t)d;
        if (!b.read(tot, S) || !b.read(fre, S)) { std::cout << "\x1B[31m ERROR: \x1B[0mtruncated node table\n"; return true; }
        size_t at = 0;
        for (uint64_t i = 0; i < n; i++) {
            Node nd;
            nd.tag = tags[i]; nd.index = 0; nd.ndof = ndof[i];
            nd.total.assign(tot.begin() + at, tot.begin() + at + ndof[i]);
            nd.free.assign(fre.begin() + at, fre.begin() + at + ndof[i]);
            nd.coords.assign(xyz.begin() + i * nc, xyz.begin() + (i + 1) * nc);
            at += (size_t)ndof[i];
            mesh.Nodes.emplace_hint(mesh.Nodes.end(), nd.tag, std::move(nd));      // tags ascend in the table: O(1) inserts
        }
        for (auto &kv : mesh.Nodes) kv.second.index = idx++;      // ascending tag
    } else
    for (auto &kv : by_tag(J["Nodes"])) {
        Node n;
        n.tag = (unsigned)kv.first; n.index = idx++;
        n.ndof = (*kv.second)["ndof"].as_int();
        for (auto &v : (*kv.second)["totaldof"].arr) n.total.push_back(v.as_int());
        for (auto &v : (*kv.second)["freedof"].arr) n.free.push_back(v.as_int());
        for (auto &v : (*kv.second)["coords"].arr) n.coords.push_back(v.as_double());
        mesh.Nodes[n.tag] = n;
    }
    if (J["Constraints"].has("binary")) {
        BinFile b;
        uint64_t n = 0;
        std::vector<int64_t> tg; std::vector<int32_t> st, mt; std::vector<double> fc;
        if (!b.open(dir + "/" + J["Constraints"]["binary"].as_string(), "SVLC") || !b.read1(n) || !b.read(tg, n) || !b.read(st, n) ||
            !b.read(mt, n) || !b.read(fc, n)) return true;
        for (uint64_t i = 0; i < n; i++) mesh.Constraints[(long)tg[i]] = std::make_tuple((int)st[i], std::vector<int>{mt[i]}, std::vector<double>{fc[i]});
    } else
    for (auto &kv : J["Constraints"].obj) {
        const long tag = std::strtol(kv.first.c_str(), nullptr, 10);
        std::vector<int> mt; std::vector<double> f;
        for (auto &x : kv.second["mtag"].arr) mt.push_back(x.as_int());
        for (auto &x : kv.second["factor"].arr) f.push_back(x.as_double());
        mesh.Constraints[tag] = std::make_tuple(kv.second["stag"].as_int(), mt, f);
    }
    if (geometry_only) return false;     // a peer's file: only its nodes and ties matter here
    idx = 0;
    for (auto &kv : by_tag(J["Materials"])) {
        Material m;
        m.tag = (unsigned)kv.first; m.index = idx++;
        const std::string name = (*kv.second)["name"].as_string();
        const JValue &a = (*kv.second)["attributes"];
        if (ieq(name, "ELASTIC3DLINEAR")) { m.kind = SVLGPU_ELASTIC3DLINEAR; m.par = {a["E"].as_double(), a["nu"].as_double(), a["rho"].as_double()}; }
        else if (ieq(name, "ELASTIC2DPLANESTRAIN")) { m.kind = SVLGPU_ELASTIC2DPLANESTRAIN; m.par = {a["E"].as_double(), a["nu"].as_double(), a["rho"].as_double()}; }
        else if (ieq(name, "PLASTIC3DJ2")) { m.kind = SVLGPU_PLASTIC3DJ2; m.par = {a["K"].as_double(), a["G"].as_double(), a["rho"].as_double(), a["h"].as_double(), a["beta"].as_double(), a["Sy"].as_double()}; }
        else if (ieq(name, "PLASTICPLANESTRAINJ2")) { m.kind = SVLGPU_PLASTICPLANESTRAINJ2; m.par = {a["K"].as_double(), a["G"].as_double(), a["rho"].as_double(), a["h"].as_double(), a["beta"].as_double(), a["Sy"].as_double()}; }
        else if (ieq(name, "VISCOUS1DLINEAR")) { m.kind = SVLGPU_VISCOUS1DLINEAR; m.par = {a["eta"].as_double()}; }   // Driver.hpp:602-607
        else { std::cout << "\x1B[31m ERROR: \x1B[0mmaterial " << name << " is not on the GPU explicit path\n"; return true; }
        mesh.Materials[m.tag] = m;
    }
    for (auto &kv : by_tag(J["Masses"])) {
        std::vector<double> v;
        for (auto &x : (*kv.second)["mass"].arr) v.push_back(x.as_double());
        mesh.Masses[(unsigned)kv.first] = v;
    }
    idx = 0;
    if (J["Elements"].has("binary")) {
        BinFile b;
        uint64_t n = 0;
        std::vector<uint32_t> tags, mat, conn; std::vector<int32_t> kind, nconn; std::vector<double> attr, am, ak; std::vector<uint8_t> ray;
        if (!b.open(dir + "/" + J["Elements"]["binary"].as_string(), "SVLE") || !b.read1(n) || !b.read(tags, n) || !b.read(kind, n) ||
            !b.read(mat, n) || !b.read(nconn, n) || !b.read(conn, n * 8) || !b.read(attr, n * 10) || !b.read(am, n) || !b.read(ak, n) ||
            !b.read(ray, n)) { std::cout << "\x1B[31m ERROR: \x1B[0mtruncated element table\n"; return true; }
        static const int nattr_of[6] = {0, 0, 1, 9, 8, 1};       // attributes per kind, the order of the JSON branch below
        for (uint64_t i = 0; i < n; i++) {
            Element e;
            e.tag = tags[i]; e.index = 0; e.kind = kind[i]; e.material = mat[i];
            if (e.kind < SVLGPU_LIN3DHEXA8 || e.kind > SVLGPU_ZEROLENGTH1D || nconn[i] < 0 || nconn[i] > 8) { std::cout << "\x1B[31m ERROR: \x1B[0mbad element record\n"; return true; }
            e.conn.assign(conn.begin() + i * 8, conn.begin() + i * 8 + nconn[i]);
            e.attr.assign(attr.begin() + i * 10, attr.begin() + i * 10 + nattr_of[e.kind]);
            if (ray[i]) mesh.Rayleigh[e.tag] = {am[i], ak[i]};
            mesh.Elements.emplace_hint(mesh.Elements.end(), e.tag, std::move(e));
        }
        for (auto &kv : mesh.Elements) kv.second.index = idx++;
    } else
    for (auto &kv : by_tag(J["Elements"])) {
        Element e;
        e.tag = (unsigned)kv.first; e.index = idx++;
        const std::string name = (*kv.second)["name"].as_string();
        const JValue &a = (*kv.second)["attributes"];
        for (auto &v : (*kv.second)["conn"].arr) e.conn.push_back((unsigned)v.as_int());
        e.material = (unsigned)a["material"].as_int();
        auto vec = [&](const char *k) { for (auto &x : a[k].arr) e.attr.push_back(x.as_double()); };
        if (ieq(name, "LIN3DHEXA8")) e.kind = SVLGPU_LIN3DHEXA8;
        else if (ieq(name, "LIN2DQUAD4")) { e.kind = SVLGPU_LIN2DQUAD4; e.attr = {a["th"].as_double(1.0)}; }
        else if (ieq(name, "PML3DHEXA8")) { e.kind = SVLGPU_PML3DHEXA8; e.attr = {a["n"].as_double(), a["L"].as_double(), a["R"].as_double()}; vec("x0"); vec("npml"); }
        else if (ieq(name, "PML2DQUAD4")) { e.kind = SVLGPU_PML2DQUAD4; e.attr = {a["th"].as_double(1.0), a["n"].as_double(), a["L"].as_double(), a["R"].as_double()}; vec("x0"); vec("npml"); }
        else if (ieq(name, "ZEROLENGTH1D")) { e.kind = SVLGPU_ZEROLENGTH1D; e.attr = {(double)a["dir"].as_int()}; }   // Driver.hpp:1072-1078
        else { std::cout << "\x1B[31m ERROR: \x1B[0melement " << name << " is not on the GPU explicit path\n"; return true; }
        mesh.Elements[e.tag] = e;
    }
    if (!J["Dampings"].has("binary"))                              // binary form: the Rayleigh columns of the element table
    for (auto &kv : by_tag(J["Dampings"])) {
        const std::string name = (*kv.second)["name"].as_string();
        const JValue &a = (*kv.second)["attributes"];
        if (ieq(name, "RAYLEIGH"))
            for (auto &x : a["list"].arr) mesh.Rayleigh[(unsigned)x.as_int()] = {a["am"].as_double(), a["ak"].as_double()};
    }
    for (auto &kv : by_tag(J["Loads"])) {
        Load l;
        l.tag = (unsigned)kv.first;
        const std::string name = (*kv.second)["name"].as_string();
        const JValue &a = (*kv.second)["attributes"];
        if (ieq(name, "POINTLOAD")) {
            if (!ieq(a["type"].as_string(), "CONCENTRATED")) { std::cout << "\x1B[31m ERROR: \x1B[0monly CONCENTRATED point loads are on the GPU explicit path\n"; return true; }
            for (auto &x : a["list"].arr) l.nodes.push_back((unsigned)x.as_int());
            for (auto &x : a["dir"].arr) l.dir.push_back(x.as_double());
            if (ieq(a["name"].as_string(), "CONSTANT")) l.series = {a["mag"].as_double()};
            else {
                std::ifstream f(a["file"].as_string());                     // Driver.hpp:1514-1527
                unsigned nt = 0;
                if (f.is_open()) { f >> nt; l.series.resize(nt); for (unsigned j = 0; j < nt; j++) f >> l.series[j]; }
                if (l.series.empty()) { std::cout << "\x1B[31m ERROR: \x1B[0mcannot read load file " << a["file"].as_string() << "\n"; return true; }
            }
        } else if (ieq(name, "ELEMENTLOAD") && ieq(a["type"].as_string(), "GENERALWAVE")) {
            l.drm = true;
            for (auto &x : a["list"].arr) l.elements.push_back((unsigned)x.as_int());
            l.drm_pattern = a["file"].as_string();
        } else if (ieq(name, "SUPPORTMOTION")) {
            l.support = true;
            for (auto &x : a["list"].arr) l.nodes.push_back((unsigned)x.as_int());
        } else { std::cout << "\x1B[31m ERROR: \x1B[0mload " << name << " is not on the GPU explicit path\n"; return true; }
        mesh.Loads[l.tag] = l;
    }
    if (J.has("Supports"))
    for (auto &kv : by_tag(J["Supports"])) {                       // UpdateSupportMotion, Driver.hpp:509-563
        const JValue &S = *kv.second;
        const bool constant = ieq(S["type"].as_string(), "CONSTANT");
        for (size_t n = 0; n < S["dof"].arr.size(); n++) {
            std::vector<double> xo;
            if (constant) xo = {S["value"].arr[n].as_double()};
            else {
                std::ifstream f(S["file"].arr[n].as_string());
                unsigned nt = 0;
                if (f.is_open()) { f >> nt; xo.resize(nt); for (unsigned j = 0; j < nt; j++) f >> xo[j]; }
                if (xo.empty()) { std::cout << "\x1B[31m ERROR: \x1B[0mcannot read support motion file " << S["file"].arr[n].as_string() << "\n"; return true; }
            }
            mesh.Supports[(unsigned)kv.first].push_back({S["dof"].arr[n].as_int(), xo});
        }
    }
    return false;
}

// ---- several ranks -------------------------------------------------------------------------------------------------------
// The reference's ranks meet in one globally numbered system (Global.ntotal / nfree, MumpsSolver.cpp:56,161).  Here every
// rank drives its own GPU with its own compact numbering and exchanges interface values with the ranks that share nodes
// with it, so three things are derived from the per-rank files (all ranks read all files and derive the same answer):
//   (i)  tie closure: the pre-processor gives a partition the master nodes of the constraints whose slave it holds
//        (SeismoVLAB.py:381-392); the device also wants the slave wherever the master is, so that every replica of a tied
//        dof resolves to the same unknown -- missing nodes / constraints are copied from the rank that has them;
//   (ii) local total / free numbers (ascending node tag, then dof);
//   (iii) per peer: the shared node tags, ascending -- svlgpu_add_halo wants mirror-image lists.
bool PlanPartitions(std::vector<Mesh> &all, int rank) {
    const int world = (int)all.size();
    Mesh &me = all[rank];
    std::map<unsigned, const Node *> catalog;                       // any copy of a node (replicas are identical)
    std::map<long, std::tuple<int, std::vector<int>, std::vector<double>>> ties;
    for (auto &m : all) {
        for (auto &kv : m.Nodes) catalog.emplace(kv.first, &kv.second);
        for (auto &kv : m.Constraints) ties.emplace(kv.first, kv.second);
    }
    std::map<int, unsigned> node_of_total, node_of_free;
    for (auto &kv : catalog)
        for (size_t k = 0; k < kv.second->total.size(); k++) {
            node_of_total[kv.second->total[k]] = kv.first;
            if (kv.second->free[k] >= 0) node_of_free[kv.second->free[k]] = kv.first;
        }
    struct Tie { long tag; std::vector<unsigned> nodes; };
    std::vector<Tie> tl;
    for (auto &kv : ties) {
        Tie t; t.tag = kv.first;
        auto &[stag, mt, f] = kv.second;
        if (!node_of_total.count(stag)) { std::cout << "\x1B[31m ERROR: \x1B[0mconstraint " << kv.first << ": slave dof not found in any partition\n"; return true; }
        t.nodes.push_back(node_of_total[stag]);
        for (int m : mt) {
            if (!node_of_free.count(m)) { std::cout << "\x1B[31m ERROR: \x1B[0mconstraint " << kv.first << ": master dof not found in any partition\n"; return true; }
            t.nodes.push_back(node_of_free[m]);
        }
        tl.push_back(t);
    }
    for (auto &m : all) {
        bool changed = true;
        while (changed) {
            changed = false;
            for (auto &t : tl) {
                bool any = false, allp = true;
                for (unsigned n : t.nodes) { const bool h = m.Nodes.count(n) != 0; any |= h; allp &= h; }
                if (!any) continue;
                if (!allp) { for (unsigned n : t.nodes) if (!m.Nodes.count(n)) m.Nodes[n] = *catalog.at(n); changed = true; }
                if (!m.Constraints.count(t.tag)) m.Constraints[t.tag] = ties.at(t.tag);
            }
        }
        int idx = 0;
        for (auto &kv : m.Nodes) kv.second.index = idx++;         // ascending tag, as UpdateMesh numbers them
    }
    // local numbering of this rank
    std::map<int, int> ltot, lfre;
    for (auto &kv : me.Nodes) {
        Node &n = kv.second;
        n.ltotal.clear(); n.lfree.clear();
        for (size_t k = 0; k < n.total.size(); k++) {
            n.ltotal.push_back((int)ltot.size()); ltot[n.total[k]] = n.ltotal.back();
            if (n.free[k] >= 0) { n.lfree.push_back((int)lfre.size()); lfre[n.free[k]] = n.lfree.back(); }
            else n.lfree.push_back(n.free[k]);
        }
    }
    me.ntotal_dev = (int)ltot.size(); me.nfree_dev = (int)lfre.size();
    for (auto &kv : me.Constraints) {
        auto &[stag, mt, f] = kv.second;
        stag = ltot.at(stag);
        for (int &m : mt) m = lfre.at(m);
    }
    const int pml_ndof = me.ndim == 3 ? 9 : 5;
    for (auto &kv : catalog) if (kv.second->ndof == pml_ndof) me.pml_collective = true;
    for (int q = 0; q < world; q++) {
        if (q == rank) continue;
        std::vector<int32_t> shared;
        auto a = me.Nodes.begin();
        auto b = all[q].Nodes.begin();
        while (a != me.Nodes.end() && b != all[q].Nodes.end()) {
            if (a->first < b->first) ++a;
            else if (b->first < a->first) ++b;
            else { shared.push_back(a->second.index); ++a; ++b; }
        }
        if (!shared.empty()) me.Halos[q] = shared;
    }
    return false;
}
void PrintPlan(const Mesh &me, int rank) {
    std::cout << "PLAN rank " << rank << " nodes " << me.Nodes.size() << " ntotal " << me.ntotal_dev << " nfree " << me.nfree_dev
              << " constraints " << me.Constraints.size() << " pml_collective " << (me.pml_collective ? 1 : 0) << "\n";
    {   // digest of everything UpdateMesh built: equal for a JSON file and its binary-sidecar twin
        unsigned long long h = 1469598103934665603ull;
        auto mixb = [&](const void *p, size_t n) { const unsigned char *c = (const unsigned char *)p; for (size_t i = 0; i < n; i++) { h ^= c[i]; h *= 1099511628211ull; } };
        auto mixi = [&](long long v) { mixb(&v, sizeof v); };
        auto mixd = [&](double v) { mixb(&v, sizeof v); };
        for (auto &kv : me.Nodes) {
            const Node &n = kv.second;
            mixi(n.tag); mixi(n.index); mixi(n.ndof);
            for (int v : n.total) mixi(v);
            for (int v : n.free) mixi(v);
            for (double v : n.coords) mixd(v);
        }
        for (auto &kv : me.Elements) {
            const Element &e = kv.second;
            mixi(e.tag); mixi(e.index); mixi(e.kind); mixi(e.material);
            for (unsigned v : e.conn) mixi(v);
            for (double v : e.attr) mixd(v);
        }
        for (auto &kv : me.Constraints) { mixi(kv.first); mixi(std::get<0>(kv.second)); for (int v : std::get<1>(kv.second)) mixi(v); for (double v : std::get<2>(kv.second)) mixd(v); }
        for (auto &kv : me.Rayleigh) { mixi(kv.first); mixd(kv.second.first); mixd(kv.second.second); }
        for (auto &kv : me.Masses) { mixi(kv.first); for (double v : kv.second) mixd(v); }
        std::cout << "PLAN rank " << rank << " mesh digest " << h << " elements " << me.Elements.size() << " rayleigh " << me.Rayleigh.size() << "\n";
    }
    std::vector<unsigned> tag_of;
    for (auto &kv : me.Nodes) tag_of.push_back(kv.first);
    for (auto &kv : me.Halos) {
        unsigned long long h = 1469598103934665603ull;             // FNV-1a over the shared tags
        for (int32_t i : kv.second) { h ^= tag_of[i]; h *= 1099511628211ull; }
        std::cout << "PLAN rank " << rank << " peer " << kv.first << " shared " << kv.second.size() << " hash " << h << "\n";
    }
}

// NCCL bootstrap without MPI: rank 0 drops the unique id into the partition directory, the others pick it up
// The file name carries the job: the -np launcher's own id, else torchrun's run id + port (the same on every rank of a job).
// A run that died after rank 0 wrote the file must not poison the next one with the same name: (i) rank 0 removes stale
// id / ack files as the very first thing it does (RemoveStaleNcclFiles, before it parses anything), (ii) the file carries the
// wall-clock second it was written and a reader refuses one written long before the reader itself started.
std::string NcclIdPath(const std::string &dir) {
    const char *job = getenv("SVLGPU_JOB_ID");
    std::string tag;
    if (job) tag = job;
    else {
        const char *run = getenv("TORCHELASTIC_RUN_ID"), *port = getenv("MASTER_PORT");
        tag = std::string(run ? run : "run") + "." + (port ? port : "0");
    }
    for (char &c : tag) if (c == '/' || c == ' ') c = '_';
    return dir + "/.svlgpu_nccl_id." + tag;
}
static const long long g_process_start = (long long)time(nullptr);
void RemoveStaleNcclFiles(const std::string &dir, int world) {
    const std::string path = NcclIdPath(dir);
    std::remove(path.c_str());
    std::remove((path + ".tmp").c_str());
    for (int q = 1; q < world; q++) std::remove((path + ".ack." + std::to_string(q)).c_str());
}
bool ExchangeNcclId(const std::string &dir, int rank, char id[128]) {
    const std::string path = NcclIdPath(dir);
    if (rank == 0) {
        if (svlgpu_nccl_unique_id(id)) return true;
        const std::string tmp = path + ".tmp";
        const long long stamp = (long long)time(nullptr);
        { std::ofstream f(tmp, std::ios::binary); f.write(id, 128); f.write((const char *)&stamp, sizeof stamp); }
        return std::rename(tmp.c_str(), path.c_str()) != 0;
    }
    for (int tries = 0; tries < 3000; tries++) {                     // up to 5 minutes: rank 0 may still be parsing
        std::ifstream f(path, std::ios::binary);
        long long stamp = 0;
        if (f.is_open() && f.read(id, 128) && f.gcount() == 128 && f.read((char *)&stamp, sizeof stamp) &&
            stamp >= g_process_start - 120) {                        // older: left behind by a run that died (ranks start together)
            std::ofstream ack(path + ".ack." + std::to_string(rank));  // rank 0 keeps the file until every rank has it
            return false;
        }
        struct timespec ts = {0, 100000000};
        nanosleep(&ts, nullptr);
    }
    std::cout << "\x1B[31m ERROR: \x1B[0mrank " << rank << " did not find the NCCL id file " << path << "\n";
    return true;
}

// ---- Assembler (07-Assembler/Assembler.cpp): host-vector access to the device passes -------------------------------
class Assembler {
  public:
    explicit Assembler(svlgpu_model *h, int ntotal) : h(h), n(ntotal) {}
    bool ComputeInternalForceVector(std::vector<double> &F) { F.assign(n, 0.0); return svlgpu_internal_force(h, F.data()) != 0; }   // :239-269
    bool ComputeMassMatrix(std::vector<double> &Mdiag) { Mdiag.assign(n, 0.0); return svlgpu_get_mass_diagonal(h, Mdiag.data()) != 0; }   // :47-67 (lumped)
  private:
    svlgpu_model *h; int n;
};

// ---- Integrator (10-Integrators/02-CentralDifference/CentralDifference.cpp) -----------------------------------------
class CentralDifference {
  public:
    // newmark: the same device handle advanced by NewmarkBeta + Linear (10-Integrators/03-Newmark/NewmarkBeta.cpp); the class
    // keeps the CentralDifference name because it is the plug point INTEGRATION.md describes
    CentralDifference(Mesh &mesh, double dt, bool newmark = false) : mesh(mesh), dt(dt), newmark(newmark) {}
    ~CentralDifference() { if (h) svlgpu_destroy(h); }
    svlgpu_model *handle() { return h; }

    // CentralDifference::Initialize (:35-71) + Mesh::Initialize: hands the object graph to the device
    // keep_gauss: ELEMENT recorders (Gauss-point strain / stress) are asked for -- the listed elements must then run through the
    // Gauss-point kernels, so the lattice fast path is switched off for this run
    bool Initialize(const LoadCombo &combo, const std::vector<RecorderSpec> &recs, int nt, int device, bool keep_gauss = false) {
        h = svlgpu_create(mesh.ndim, mesh.lumped ? 1 : 0);
        if (!h) return fail();
        std::vector<int32_t> ndof, total, freed;
        std::vector<double> xyz;
        for (auto &kv : mesh.Nodes) {
            const Node &n = kv.second;
            ndof.push_back(n.ndof);
            const Small<int, 9> &tt = n.ltotal.empty() ? n.total : n.ltotal, &ff = n.lfree.empty() ? n.free : n.lfree;
            total.insert(total.end(), tt.begin(), tt.end());
            freed.insert(freed.end(), ff.begin(), ff.end());
            for (int c = 0; c < mesh.ndim; c++) xyz.push_back(c < (int)n.coords.size() ? n.coords[c] : 0.0);
        }
        if (svlgpu_set_nodes(h, (int)ndof.size(), ndof.data(), xyz.data(), total.data(), freed.data(), mesh.ntotal_dev, mesh.nfree_dev)) return fail();
        for (auto &kv : mesh.Masses) {
            const int32_t node = mesh.Nodes.at(kv.first).index;
            if (svlgpu_add_nodal_mass(h, 1, &node, kv.second.data())) return fail();
        }
        for (auto &kv : mesh.Constraints) {
            auto &[stag, mt, f] = kv.second;
            std::vector<int32_t> m32(mt.begin(), mt.end());
            if (svlgpu_add_constraint(h, (int)kv.first, stag, (int)m32.size(), m32.data(), f.data())) return fail();
        }
        for (auto &kv : mesh.Materials)
            if (svlgpu_add_material(h, kv.second.kind, kv.second.par.data(), (int)kv.second.par.size()) < 0) return fail();
        // elements in ascending tag order (Assembler.cpp:251), one call per run of equal kind
        auto it = mesh.Elements.begin();
        while (it != mesh.Elements.end()) {
            const int kind = it->second.kind;
            const int nattr = (int)it->second.attr.size();
            std::vector<int32_t> conn, mat;
            std::vector<double> attr;
            auto jt = it;
            for (; jt != mesh.Elements.end() && jt->second.kind == kind; ++jt) {
                for (unsigned n : jt->second.conn) conn.push_back(mesh.Nodes.at(n).index);
                mat.push_back(mesh.Materials.at(jt->second.material).index);
                attr.insert(attr.end(), jt->second.attr.begin(), jt->second.attr.end());
            }
            if (svlgpu_add_elements(h, kind, (int)mat.size(), conn.data(), mat.data(), nattr ? attr.data() : nullptr, nattr) < 0) return fail();
            it = jt;
        }
        {   // Rayleigh damping groups (lin3DHexa8.cpp:354-366)
            std::map<std::pair<double, double>, std::vector<int32_t>> groups;
            for (auto &kv : mesh.Rayleigh) {
                const Element &el = mesh.Elements.at(kv.first);
                if (el.kind == SVLGPU_ZEROLENGTH1D) continue;     // ZeroLength1D::SetDamping does nothing (ZeroLength1D.cpp)
                groups[kv.second].push_back(el.index);
            }
            for (auto &g : groups)
                if (svlgpu_set_rayleigh(h, (int)g.second.size(), g.second.data(), g.first.first, g.first.second)) return fail();
        }
        // loads of the active combination (Assembler::ComputeExternalForceVector :290-489)
        for (size_t q = 0; q < combo.loads.size(); q++) {
            const Load &l = mesh.Loads.at(combo.loads[q]);
            const double factor = q < combo.factors.size() ? combo.factors[q] : 1.0;
            if (l.support) {                                       // Assembler::ComputeSupportMotionIncrement, Assembler.cpp:493-533
                for (unsigned n : l.nodes) {
                    auto it = mesh.Supports.find(n);
                    if (it == mesh.Supports.end() || !mesh.Nodes.count(n)) continue;
                    for (auto &ds : it->second)
                        if (svlgpu_add_support_motion(h, mesh.Nodes.at(n).index, ds.first, (int)ds.second.size(), ds.second.data(), factor)) return fail();
                }
            } else if (!l.drm) {
                std::vector<int32_t> nodes;
                for (unsigned n : l.nodes) nodes.push_back(mesh.Nodes.at(n).index);
                std::vector<double> dir = l.dir;
                dir.resize(3, 0.0);
                if (svlgpu_add_point_load(h, (int)nodes.size(), nodes.data(), 3, dir.data(), (int)l.series.size(), l.series.data(), factor)) return fail();
            } else if (AddDomainReduction(l, factor)) return true;
        }
        for (auto &r : recs) {
            std::vector<int32_t> nodes;
            for (unsigned id : r.ids) nodes.push_back(mesh.Nodes.at(id).index);
            const int field = ieq(r.resp, "DISP") ? SVLGPU_DISP : ieq(r.resp, "VEL") ? SVLGPU_VEL : ieq(r.resp, "REACTION") ? SVLGPU_REACTION : SVLGPU_ACCEL;
            if (svlgpu_add_node_recorder(h, field, (int)nodes.size(), nodes.data(), nt) < 0) return fail();
        }
        if (newmark && svlgpu_set_option(h, "integrator", 1.0)) return fail();
        if (keep_gauss && (svlgpu_set_option(h, "keep_gauss", 1.0) || svlgpu_set_option(h, "lattice_guess", 0.0))) return fail();
        for (auto &kv : mesh.Halos)
            if (svlgpu_add_halo(h, kv.first, (int)kv.second.size(), kv.second.data())) return fail();
        if (mesh.reaction_collective && svlgpu_set_option(h, "reaction_collective", 1.0)) return fail();
        if (mesh.pml_collective && svlgpu_set_option(h, "pml_collective", 1.0)) return fail();   // also on a rank without shared nodes: it still joins the all-reduces
        if (svlgpu_finalize(h, dt, device)) return fail();
        return false;
    }
    // several ranks: joins the NCCL communicator (svlgpu_comm_init sums the interface mass / damping diagonals)
    bool JoinRanks(const std::string &dir, int rank, int world) {
        char id[128];
        if (ExchangeNcclId(dir, rank, id)) return true;
        if (svlgpu_comm_init(h, id, rank, world)) return fail();
        if (rank == 0) {                                            // clean up once every rank has acknowledged the id
            const std::string path = NcclIdPath(dir);
            for (int q = 1; q < world; q++) {
                const std::string ack = path + ".ack." + std::to_string(q);
                for (int tries = 0; tries < 3000 && !std::ifstream(ack).is_open(); tries++) {
                    struct timespec ts = {0, 100000000};
                    nanosleep(&ts, nullptr);
                }
                std::remove(ack.c_str());
            }
            std::remove(path.c_str());
        }
        return false;
    }

    // Integrator::ComputeNewStep (:123-152): the whole step (effective force, diagonal / block solve, state update,
    // CommitState, recorder row) happens on the device
    bool ComputeNewStep(unsigned k) { return svlgpu_step(h, (int)k, (int)k + 1, 0) != 0 && fail(); }
    bool ComputeSteps(unsigned k0, unsigned k1) { return svlgpu_step(h, (int)k0, (int)k1, 1) != 0 && fail(); }
    bool GetDisplacements(std::vector<double> &U) { U.assign(mesh.ntotal_dev, 0.0); return svlgpu_get_state(h, SVLGPU_DISP, nullptr, 0, U.data()) != 0; }
    bool GetVelocities(std::vector<double> &V) { V.assign(mesh.ntotal_dev, 0.0); return svlgpu_get_state(h, SVLGPU_VEL, nullptr, 0, V.data()) != 0; }
    bool GetAccelerations(std::vector<double> &A) { A.assign(mesh.ntotal_dev, 0.0); return svlgpu_get_state(h, SVLGPU_ACCEL, nullptr, 0, A.data()) != 0; }

  private:
    bool fail() { std::cout << "\x1B[31m ERROR: \x1B[0m" << svlgpu_last_error() << "\n"; return true; }

    // ELEMENTLOAD GENERALWAVE: one .drm text file per node, `nt nFields cond` then nt rows (Driver.hpp:1689-1721)
    bool AddDomainReduction(const Load &l, double factor) {
        std::map<unsigned, bool> nodes;
        std::vector<int32_t> elems;
        for (unsigned e : l.elements) {
            const Element &el = mesh.Elements.at(e);
            elems.push_back(el.index);
            for (unsigned n : el.conn) nodes[n] = true;
        }
        std::vector<int32_t> nidx;
        std::vector<uint8_t> ext;
        std::vector<double> field;
        unsigned nt0 = 0;
        const unsigned nf = 3 * (unsigned)mesh.ndim;
        for (auto &kv : nodes) {
            std::string file = l.drm_pattern;
            const size_t pos = file.find('$');
            if (pos != std::string::npos) file.replace(pos, 1, std::to_string(kv.first));
            std::ifstream f(file);
            if (!f.is_open()) { std::cout << "\x1B[31m ERROR: \x1B[0mcannot read DRM file " << file << "\n"; return true; }
            unsigned nt = 0, nFields = 0; int cond = 0;
            f >> nt >> nFields >> cond;
            if (nFields != nf || (nt0 && nt != nt0)) { std::cout << "\x1B[31m ERROR: \x1B[0minconsistent DRM file " << file << "\n"; return true; }
            nt0 = nt;
            const size_t base = field.size();
            field.resize(base + (size_t)nt * nf);
            for (size_t i = 0; i < (size_t)nt * nf; i++) f >> field[base + i];
            nidx.push_back(mesh.Nodes.at(kv.first).index);
            ext.push_back(cond ? 1 : 0);
        }
        if (svlgpu_add_drm_load(h, (int)elems.size(), elems.data(), (int)nidx.size(), nidx.data(), ext.data(), (int)nt0, field.data(), factor)) return fail();
        return false;
    }

    Mesh &mesh;
    double dt;
    bool newmark = false;
    svlgpu_model *h = nullptr;
};

// ---- Recorder (12-Utilities/Recorder.cpp): NODE files in the reference's text layout -----------------------------------
class Recorder {
  public:
    Recorder(const RecorderSpec &s, int id) : spec(s), id(id) {}
    void Initialize(const Mesh &mesh, const std::string &dir, const std::string &combo, unsigned nsteps) {     // :57-105
        out.open(dir + "/../Solution/" + combo + "/" + spec.file);
        out.precision(spec.precision);
        out.setf(std::ios::scientific);
        unsigned ndofs = 0;
        for (unsigned idn : spec.ids) ndofs += mesh.Nodes.at(idn).ndof;
        out << spec.ids.size() << " " << ndofs << " " << mesh.ntotal << " " << nsteps << "\n";
        for (unsigned idn : spec.ids) {
            const Node &n = mesh.Nodes.at(idn);
            out << idn << " " << n.ndof;
            for (int d : n.total) out << " " << d;
            out << "\n";
        }
    }
    bool WriteResponse(svlgpu_model *h) {                                                                   // :239-269
        const int rows = svlgpu_recorder_rows(h, id), w = svlgpu_recorder_width(h, id);
        std::vector<double> buf((size_t)std::max(rows, 0) * std::max(w, 0));
        if (rows > 0 && svlgpu_read_recorder(h, id, 0, rows, buf.data())) return true;
        for (int r = 0; r < rows; r++) {
            for (int c = 0; c < w; c++) out << buf[(size_t)r * w + c] << " ";
            out << "\n";
        }
        return false;
    }
    void Finalize() { out.close(); }
  private:
    RecorderSpec spec; int id; std::ofstream out;
};

// ELEMENT recorder, responses STRAIN / STRESS (Recorder.cpp:106-160 header, :270-300 rows): Element::GetStrain / GetStress at the
// Gauss points of the listed elements, one row per step, element by element, Gauss point by Gauss point
class ElementRecorder {
  public:
    ElementRecorder(const RecorderSpec &s) : spec(s) {}
    void Initialize(const Mesh &mesh, const std::string &dir, const std::string &combo, unsigned nsteps) {
        out.open(dir + "/../Solution/" + combo + "/" + spec.file);
        out.precision(spec.precision);
        out.setf(std::ios::scientific);
        out << spec.ids.size() << " " << mesh.ntotal << " " << nsteps << "\n";
        width = 0;
        for (unsigned id : spec.ids) {
            const Element &e = mesh.Elements.at(id);
            const bool hex = e.kind == SVLGPU_LIN3DHEXA8;
            out << id << " " << (hex ? 8 : 4) << " " << (hex ? 6 : 3) << "\n";
            idx.push_back(e.index);
            width += hex ? 48 : 12;
        }
    }
    // after a step: Gauss-point values of the current state (the caller has just run an internal-force pass on it)
    bool WriteResponse(svlgpu_model *h) {
        std::vector<double> buf(width);
        const int field = ieq(spec.resp, "STRAIN") ? SVLGPU_STRAIN : SVLGPU_STRESS;
        if (svlgpu_get_gauss(h, field, (int)idx.size(), idx.data(), buf.data())) return true;
        for (double v : buf) out << v << " ";
        out << "\n";
        return false;
    }
    void Finalize() { out.close(); }
  private:
    RecorderSpec spec; std::vector<int32_t> idx; size_t width = 0; std::ofstream out;
};

// ---- Analysis (08-Analysis/02-Dynamic/DynamicAnalysis.cpp:25-64) ----------------------------------------------------------
class DynamicAnalysis {
  public:
    DynamicAnalysis(Mesh &mesh, CentralDifference &integ, std::vector<Recorder> &recs, unsigned nt) : mesh(mesh), integ(integ), recs(recs), nt(nt) {}
    std::vector<ElementRecorder> erecs;
    bool Analyze(const std::string &dir, const LoadCombo &combo) {
        for (auto &r : recs) r.Initialize(mesh, dir, combo.folder, nt);
        if (!erecs.empty()) {
            // Gauss-point histories: step by step, with a force pass on the new state before every row (the step itself
            // evaluates the elements at its starting state); DynamicAnalysis.cpp:36-57 order: step, then WriteRecorders
            for (auto &r : erecs) r.Initialize(mesh, dir, combo.folder, nt);
            bool stop = false;
            std::vector<double> scratch(mesh.ntotal_dev);
            for (unsigned k = 1; k < nt && !stop; k++) {
                stop = integ.ComputeSteps(k, k + 1) || svlgpu_internal_force(integ.handle(), scratch.data()) != 0;
                for (auto &r : erecs) stop = stop || r.WriteResponse(integ.handle());
            }
            std::cout << " RUNNING (" << combo.name << ") : 100%\n";
            for (auto &r : recs) { stop = r.WriteResponse(integ.handle()) || stop; r.Finalize(); }
            for (auto &r : erecs) r.Finalize();
            return stop;
        }
        // k = 1 .. nt-1 (DynamicAnalysis.cpp:36): recorder rows are kept on the device and written at the end
        bool stop = false;
        const unsigned chunk = 256;
        for (unsigned k = 1; k < nt && !stop; k += chunk) {
            stop = integ.ComputeSteps(k, std::min(nt, k + chunk));
            std::cout << "\r RUNNING (" << combo.name << ") : " << std::min(100u, 100 * std::min(nt, k + chunk) / std::max(1u, nt - 1)) << "%" << std::flush;
        }
        std::cout << "\n";
        for (auto &r : recs) { stop = r.WriteResponse(integ.handle()) || stop; r.Finalize(); }
        return stop;
    }
  private:
    Mesh &mesh; CentralDifference &integ; std::vector<Recorder> &recs; unsigned nt;
};

std::string read_file(const std::string &path) {
    std::ifstream f(path);
    if (!f.is_open()) throw std::runtime_error("cannot open " + path);
    std::stringstream ss;
    ss << f.rdbuf();
    return ss.str();
}

}  // namespace

int main(int argc, char **argv) {
    std::string dir = ".";
    std::vector<std::string> files;
    int np = 0;
    bool plan_only = false;
    for (int i = 1; i < argc; i++) {                                     // Utilities.hpp:97-161
        if (ieq(argv[i], "-dir") && i + 1 < argc) dir = argv[++i];
        else if (ieq(argv[i], "-file")) { while (i + 1 < argc && argv[i + 1][0] != '-') files.push_back(argv[++i]); }
        else if (ieq(argv[i], "-np") && i + 1 < argc) np = atoi(argv[++i]);
        else if (ieq(argv[i], "-plan")) plan_only = true;
    }
    if (files.empty()) { std::cout << " usage: SeismoVLAB_gpu.exe [-np N] -dir <Partition dir> -file '<name>.$.json'\n"; return 1; }
    if (np > 1) {
        // the launcher `mpirun -np N` is for the reference: one child per rank / GPU, forked before any CUDA call
        const std::string job = std::to_string((long)getpid()) + "." + std::to_string((long)time(nullptr));
        std::vector<pid_t> kids;
        for (int r = 0; r < np; r++) {
            const pid_t pid = fork();
            if (pid < 0) { std::cout << "\x1B[31m ERROR: \x1B[0mfork failed\n"; return 1; }
            if (pid == 0) {
                setenv("RANK", std::to_string(r).c_str(), 1);
                setenv("LOCAL_RANK", std::to_string(r).c_str(), 1);
                setenv("WORLD_SIZE", std::to_string(np).c_str(), 1);
                setenv("SVLGPU_JOB_ID", job.c_str(), 1);
                np = 0;
                kids.clear();
                break;
            }
            kids.push_back(pid);
        }
        if (!kids.empty()) {
            // a rank that stops (bad input, solver stop) would leave the others waiting in a collective: take them down with it
            int worst = 0;
            size_t left = kids.size();
            while (left) {
                int st = 0;
                const pid_t done = waitpid(-1, &st, 0);
                if (done < 0) break;
                left--;
                const int rc = WIFEXITED(st) ? WEXITSTATUS(st) : 1;
                kids.erase(std::remove(kids.begin(), kids.end(), done), kids.end());   // a reaped pid may be recycled: never signal it
                if (rc != 0 && worst == 0)
                    for (pid_t k : kids) kill(k, SIGTERM);
                worst = std::max(worst, rc);
            }
            return worst;
        }
    }
    const char *rk = getenv("RANK"), *lr = getenv("LOCAL_RANK"), *ws = getenv("WORLD_SIZE");
    const int rank = rk ? atoi(rk) : 0, device = lr ? atoi(lr) : 0, world = ws ? std::max(1, atoi(ws)) : 1;
    if (world > 1 && rank == 0 && !plan_only) RemoveStaleNcclFiles(dir, world);
    struct IdFileGuard {                                                 // rank 0 leaves no id file behind on a failure path either
        std::string dir; int world; bool armed;
        ~IdFileGuard() { if (armed) RemoveStaleNcclFiles(dir, world); }
    } guard{dir, world, world > 1 && rank == 0 && !plan_only};
    try {
        for (std::string pattern : files) {                              // staged analyses run in sequence (Driver.hpp:2058-2100)
            auto file_of = [&](int r) {
                std::string f = pattern;
                const size_t pos = f.find('$');
                if (pos != std::string::npos) f.replace(pos, 1, std::to_string(r));
                return f;
            };
            const std::string file = file_of(rank);
            const std::string text = read_file(dir + "/" + file);
            const JValue J = svlhost::JParser(text).parse();
            std::vector<Mesh> all(world);
            Mesh &mesh = all[rank];
            if (UpdateMesh(mesh, J, dir)) return 1;
            if (world > 1) {
                if (pattern.find('$') == std::string::npos) { std::cout << "\x1B[31m ERROR: \x1B[0mseveral ranks need a '$' in -file\n"; return 1; }
                for (int q = 0; q < world; q++) {
                    if (q == rank) continue;
                    const JValue Jq = svlhost::JParser(read_file(dir + "/" + file_of(q))).parse();
                    if (UpdateMesh(all[q], Jq, dir, true)) return 1;
                    if (Jq.has("Recorders"))                        // a REACTION recorder on any rank makes the reaction pass collective
                        for (auto &kv : by_tag(Jq["Recorders"])) if (ieq((*kv.second)["resp"].as_string("disp"), "REACTION")) mesh.reaction_collective = true;
                }
                if (J.has("Recorders"))
                    for (auto &kv : by_tag(J["Recorders"])) if (ieq((*kv.second)["resp"].as_string("disp"), "REACTION")) mesh.reaction_collective = true;
                if (PlanPartitions(all, rank)) return 1;
                for (int q = 0; q < world; q++) if (q != rank) all[q] = Mesh();     // the peers' tables were only needed for the plan
            }
            if (plan_only) { PrintPlan(mesh, rank); continue; }
            // combination / recorders / simulation (Driver.hpp:1930-1975, 1859-1925, 1748-1856)
            const JValue &S = J["Simulations"];
            const unsigned comboTag = (unsigned)S["combo"].as_int(1);
            const JValue &C = J["Combinations"][std::to_string(comboTag)];
            LoadCombo combo;
            combo.tag = comboTag; combo.name = C["name"].as_string("Combo");
            combo.folder = C["attributes"]["folder"].as_string(combo.name);
            for (auto &x : C["attributes"]["load"].arr) combo.loads.push_back((unsigned)x.as_int());
            for (auto &x : C["attributes"]["factor"].arr) combo.factors.push_back(x.as_double());
            const JValue &A = S["attributes"];
            const bool newmark = ieq(A["integrator"]["name"].as_string(), "NEWMARK");          // Driver.hpp:1811-1813
            if (!ieq(A["analysis"]["name"].as_string(), "DYNAMIC") || !ieq(A["algorithm"]["name"].as_string(), "LINEAR") ||
                !(newmark || ieq(A["integrator"]["name"].as_string(), "CENTRALDIFFERENCE"))) {
                std::cout << "\x1B[31m ERROR: \x1B[0monly DYNAMIC + LINEAR + CENTRALDIFFERENCE | NEWMARK is on the GPU path\n";
                return 1;
            }
            const unsigned nt = (unsigned)A["analysis"]["nt"].as_int();
            const double dt = A["integrator"]["dt"].as_double();
            std::vector<RecorderSpec> specs, especs;
            for (auto &kv : by_tag(J["Recorders"])) {
                RecorderSpec r;
                r.name = (*kv.second)["name"].as_string();
                r.file = (*kv.second)["file"].as_string();
                r.resp = (*kv.second)["resp"].as_string("disp");
                // ELEMENT recorders were written after the last GPU pass of the round: opt-in until hardware has seen them
                if (getenv("SVLGPU_ELEMENT_RECORDERS") && ieq(r.name, "ELEMENT") && (ieq(r.resp, "STRAIN") || ieq(r.resp, "STRESS"))) {
                    r.precision = (*kv.second)["ndps"].as_int(6);
                    for (auto &x : (*kv.second)["list"].arr) r.ids.push_back((unsigned)x.as_int());
                    bool solid = true;
                    for (unsigned id : r.ids) { const int kd = mesh.Elements.at(id).kind; solid = solid && (kd == SVLGPU_LIN3DHEXA8 || kd == SVLGPU_LIN2DQUAD4); }
                    if (solid) { especs.push_back(r); continue; }
                }
                if (!ieq(r.name, "NODE")) { std::cout << " WARNING: recorder " << r.name << " (" << r.resp << ") is not written by the GPU path\n"; continue; }
                r.precision = (*kv.second)["ndps"].as_int(6);
                r.nsample = (*kv.second)["nsamp"].as_int(1);
                for (auto &x : (*kv.second)["list"].arr) r.ids.push_back((unsigned)x.as_int());
                specs.push_back(r);
            }
            CentralDifference integrator(mesh, dt, newmark);
            if (!especs.empty()) std::cout << " NOTE: ELEMENT recorders: all elements run through the Gauss-point kernels, one force pass per recorded step\n";
            if (integrator.Initialize(combo, specs, (int)nt, device, !especs.empty())) return 1;
            if (world > 1 && integrator.JoinRanks(dir, rank, world)) return 1;
            std::vector<Recorder> recorders;
            for (size_t i = 0; i < specs.size(); i++) recorders.emplace_back(specs[i], (int)i);
            DynamicAnalysis analysis(mesh, integrator, recorders, nt);
            for (auto &e : especs) analysis.erecs.emplace_back(e);
            if (analysis.Analyze(dir, combo)) { std::cout << "\x1B[31m ERROR: \x1B[0mthe analysis stopped: " << svlgpu_last_error() << "\n"; return 1; }
            svlgpu_counters c;
            svlgpu_get_counters(integrator.handle(), &c);
            std::cout << " elements " << c.n_elements << ", lattice nodes " << c.n_block_nodes << ", Gauss-point elements "
                      << c.n_generic_elements << ", PML unknowns " << c.n_pml_unknowns << ", kernel launches " << c.total_launches
                      << ", last svlgpu_step call " << c.last_step_ms << " ms on the device\n";
        }
    } catch (const std::exception &e) {
        std::cout << "\x1B[31m ERROR: \x1B[0m" << e.what() << "\n";
        return 1;
    }
    return 0;
}
