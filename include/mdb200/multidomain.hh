// mdb200/multidomain.hh — C++ facade over the C ABI (mdb200.h) with the reference's spellings.
//
// The reference's boundary is a C++ template concept, Dune::PDELab::MultiDomain::GridOperator and the participants it
// is built from (gridoperator.hh:38-214, subproblem.hh:44-148, coupling.hh:30-91, subdomainset.hh:82-133).  This
// header keeps those names, argument orders and call sequences so that a driver written against dune-multidomain
// (test/testpoisson.cc:119-298 is the model) reads the same, and lowers every object to a handle of libmdb200.so.
// Arbitrary user local operators cannot run on the device: the operator classes below are the closed set the backend
// implements; anything else fails at compile time (no matching class) or throws mdb200::Exception (DUNE_THROW analogue).
// Header-only, C++11, no dependency besides mdb200.h; link with -lmdb200.
#ifndef MDB200_MULTIDOMAIN_HH
#define MDB200_MULTIDOMAIN_HH

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <fstream>
#include <initializer_list>
#include <map>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../mdb200.h"

namespace mdb200 {

struct Exception : std::runtime_error {
  int code;
  Exception(int c, const std::string& what) : std::runtime_error(what), code(c) {}
};

// One device context (one per host thread and GPU; gridoperator.hh:209-210: the reference is not re-entrant either).
class Context {
  mdb_ctx* c_;
  Context(const Context&);
  Context& operator=(const Context&);
public:
  explicit Context(int device = 0) : c_(mdb_create(device)) {
    if (!c_) throw Exception(MDB_ECUDA, std::string("mdb_create: ") + mdb_last_error(0));
  }
  ~Context() { mdb_destroy(c_); }
  mdb_ctx* get() const { return c_; }
  int check(int rc, const char* what) const {
    if (rc < 0) throw Exception(rc, std::string(what) + ": " + mdb_last_error(c_));
    return rc;
  }
};

// ---- unstructured grids: Dune::GmshReader<UGGrid<2>> (test/stokesdarcy2.cc:575-583) ------------------------------------------
// The mesh a GmshReader hands over: vertices, triangles, and the two index -> physical-group maps the reference's drivers use to
// mark subdomains (`elementIndexToPhysicalGroup[index(e)] > 0 ? 1 : 0`, test/stokesdarcy2.cc:596-601).
struct GmshMesh {
  int dim;
  std::vector<double> vertices;                    // dim doubles per vertex
  std::vector<int64_t> elements;                   // 3 vertex indices per triangle
  std::vector<int> elementIndexToPhysicalGroup;
  std::vector<int64_t> boundarySegments;           // 2 vertex indices per boundary line
  std::vector<int> boundaryIndexToPhysicalGroup;
  GmshMesh() : dim(2) {}
  int64_t numVertices() const { return (int64_t)vertices.size() / dim; }
  int64_t numElements() const { return (int64_t)elements.size() / 3; }
  // baseGrid.globalRefine(levels): regular ("red") refinement, every triangle into four; children inherit the physical group.
  // New vertices (edge midpoints) are appended in order of creation, children replace their father in place:
  // (v0,m01,m02), (m01,v1,m12), (m02,m12,v2), (m01,m12,m02).  [UPSTREAM] UG's own leaf index order after refinement is an
  // implementation detail of UG; this order is ours.
  void globalRefine(int levels) {
    for (int l = 0; l < levels; ++l) {
      std::map<std::pair<int64_t, int64_t>, int64_t> mid;
      auto midpoint = [&](int64_t a, int64_t b) -> int64_t {
        const std::pair<int64_t, int64_t> key(a < b ? a : b, a < b ? b : a);
        std::map<std::pair<int64_t, int64_t>, int64_t>::const_iterator it = mid.find(key);
        if (it != mid.end()) return it->second;
        const int64_t v = numVertices();
        for (int d = 0; d < dim; ++d) vertices.push_back(0.5 * (vertices[(size_t)(a * dim + d)] + vertices[(size_t)(b * dim + d)]));
        mid[key] = v; return v;
      };
      std::vector<int64_t> ne; std::vector<int> ng;
      for (int64_t e = 0; e < numElements(); ++e) {
        const int64_t v0 = elements[(size_t)(3 * e)], v1 = elements[(size_t)(3 * e + 1)], v2 = elements[(size_t)(3 * e + 2)];
        const int64_t m01 = midpoint(v0, v1), m02 = midpoint(v0, v2), m12 = midpoint(v1, v2);
        const int64_t c[12] = {v0, m01, m02, m01, v1, m12, m02, m12, v2, m01, m12, m02};
        ne.insert(ne.end(), c, c + 12);
        for (int k = 0; k < 4; ++k) ng.push_back(elementIndexToPhysicalGroup[(size_t)e]);
      }
      std::vector<int64_t> nb; std::vector<int> nbg;
      for (size_t b = 0; b < boundarySegments.size() / 2; ++b) {
        const int64_t a0 = boundarySegments[2 * b], a1 = boundarySegments[2 * b + 1], m = midpoint(a0, a1);
        const int64_t c[4] = {a0, m, m, a1};
        nb.insert(nb.end(), c, c + 4);
        nbg.push_back(boundaryIndexToPhysicalGroup[b]); nbg.push_back(boundaryIndexToPhysicalGroup[b]);
      }
      elements.swap(ne); elementIndexToPhysicalGroup.swap(ng); boundarySegments.swap(nb); boundaryIndexToPhysicalGroup.swap(nbg);
    }
  }
};

// Dune::GmshReader<Grid>::read(filename, boundaryIndexToPhysicalGroup, elementIndexToPhysicalGroup, verbose, insertBoundarySegments)
// for the Gmsh 2.x ASCII format of the reference's fixtures (test/stokesdarcy*.msh: "$MeshFormat 2.1 0 8", $Nodes, $Elements with
// 2-node lines = boundary segments and 3-node triangles = elements; the first tag is the physical group).  [UPSTREAM dune-grid
// gmshreader.hh] only nodes used by a line or a triangle become vertices, numbered in order of first appearance in $Elements;
// elements and boundary segments keep the file order.  Other element types (points, type 15) are skipped; quadrilaterals and
// second-order elements are rejected.
class GmshReader {
public:
  static GmshMesh read(const std::string& filename, bool verbose = false, bool insertBoundarySegments = true) {
    std::ifstream in(filename.c_str());
    if (!in) throw Exception(MDB_EINVAL, "GmshReader: cannot open " + filename);
    GmshMesh m;
    std::string tok;
    std::map<int64_t, int64_t> node_line;             // gmsh node id -> position in `xyz`
    std::vector<double> xyz;
    std::map<int64_t, int64_t> renumber;
    bool have_format = false;
    while (in >> tok) {
      if (tok == "$MeshFormat") {
        double version; int type, size;
        if (!(in >> version >> type >> size)) throw Exception(MDB_EINVAL, "GmshReader: bad $MeshFormat");
        if (version < 2.0 || version >= 3.0 || type != 0) throw Exception(MDB_EINVAL, "GmshReader: only the ASCII 2.x format is supported");
        have_format = true;
      } else if (tok == "$Nodes") {
        int64_t n; in >> n;
        xyz.resize((size_t)(3 * n));
        for (int64_t i = 0; i < n; ++i) {
          int64_t id; in >> id >> xyz[(size_t)(3 * i)] >> xyz[(size_t)(3 * i + 1)] >> xyz[(size_t)(3 * i + 2)];
          node_line[id] = i;
        }
        if (!in) throw Exception(MDB_EINVAL, "GmshReader: truncated $Nodes section");
      } else if (tok == "$Elements") {
        int64_t n; in >> n;
        for (int64_t i = 0; i < n; ++i) {
          int64_t id; int type, ntags;
          if (!(in >> id >> type >> ntags)) throw Exception(MDB_EINVAL, "GmshReader: truncated $Elements section");
          int physical = 0;
          for (int t = 0; t < ntags; ++t) { int tag; in >> tag; if (t == 0) physical = tag; }
          const int nn = type == 1 ? 2 : type == 2 ? 3 : type == 15 ? 1 : -1;
          if (nn < 0) throw Exception(MDB_EINVAL, "GmshReader: element type " + std::to_string(type) + " is not supported (lines, triangles and points only)");
          int64_t v[3];
          for (int k = 0; k < nn; ++k) {
            in >> v[k];
            if (node_line.find(v[k]) == node_line.end()) throw Exception(MDB_EINVAL, "GmshReader: element refers to an unknown node");
          }
          if (type == 15) continue;
          for (int k = 0; k < nn; ++k)
            if (renumber.find(v[k]) == renumber.end()) {
              const int64_t nv = (int64_t)renumber.size();
              renumber[v[k]] = nv;
              const int64_t line = node_line[v[k]];
              m.vertices.push_back(xyz[(size_t)(3 * line)]); m.vertices.push_back(xyz[(size_t)(3 * line + 1)]);
            }
          if (type == 2) { for (int k = 0; k < 3; ++k) m.elements.push_back(renumber[v[k]]); m.elementIndexToPhysicalGroup.push_back(physical); }
          else if (insertBoundarySegments) { for (int k = 0; k < 2; ++k) m.boundarySegments.push_back(renumber[v[k]]); m.boundaryIndexToPhysicalGroup.push_back(physical); }
        }
      }
    }
    if (!have_format || m.elements.empty()) throw Exception(MDB_EINVAL, "GmshReader: " + filename + " holds no 2.x triangle mesh");
    if (verbose) std::printf("Reading 2d Gmsh grid...\nnumber of real vertices = %lld\nnumber of boundary elements = %lld\nnumber of elements = %lld\n",
                             (long long)m.numVertices(), (long long)m.boundaryIndexToPhysicalGroup.size(), (long long)m.numElements());
    return m;
  }
};

// ---- grid: YaspGrid + MultiDomainGrid<..., FewSubDomainsTraits<dim,8>> (test/testpoisson.cc:134-161) ---------------
class SubDomainGridView { public: int subdomain; explicit SubDomainGridView(int s) : subdomain(s) {} };

class MultiDomainGrid {
  Context& ctx_;
  int dim_;
  std::vector<int64_t> n_;
  std::vector<double> lower_, upper_;
  std::vector<uint8_t> marks_;
  bool marking_;
  bool unstructured_ = false;
  std::vector<double> uvertices_; std::vector<int64_t> uelements_;
public:
  const std::vector<double>& lower() const { return lower_; }
  const std::vector<double>& upper() const { return upper_; }
  MultiDomainGrid(Context& ctx, int dim, const int64_t* cells, const double* lower, const double* upper)
    : ctx_(ctx), dim_(dim), n_(cells, cells + dim), lower_(lower, lower + dim), upper_(upper, upper + dim), marking_(false) {
    ctx_.check(mdb_mesh_structured(ctx_.get(), dim, cells, lower, upper), "mdb_mesh_structured");
  }
  // MultiDomainGrid<UGGrid<2>, ...> grid(baseGrid, false) over a mesh read by GmshReader (test/stokesdarcy2.cc:585-586)
  MultiDomainGrid(Context& ctx, const GmshMesh& mesh)
    : ctx_(ctx), dim_(mesh.dim), n_(1, mesh.numElements()), lower_(mesh.dim, 0.0), upper_(mesh.dim, 1.0), marking_(false), unstructured_(true),
      uvertices_(mesh.vertices), uelements_(mesh.elements) {
    ctx_.check(mdb_mesh_unstructured(ctx_.get(), mesh.dim, mesh.numVertices(), mesh.vertices.data(), mesh.numElements(), MDB_ETYPE_TRIANGLE,
                                     mesh.elements.data()), "mdb_mesh_unstructured");
  }
  // elementIndexToPhysicalGroup of the reader, for parameter classes that look a cell's group up (DarcyParameters::A, test/stokesdarcy2.cc:285-290)
  void setPhysicalGroups(const std::vector<int>& groups) {
    std::vector<int32_t> g(groups.begin(), groups.end());
    if ((int64_t)g.size() != size0()) throw Exception(MDB_EINVAL, "setPhysicalGroups: one group per cell");
    ctx_.check(mdb_cell_groups(ctx_.get(), g.data()), "mdb_cell_groups");
  }
  bool unstructured() const { return unstructured_; }
  const std::vector<double>& vertices() const { return uvertices_; }        // simplex meshes: dim doubles per vertex
  const std::vector<int64_t>& elements() const { return uelements_; }       // simplex meshes: dim + 1 vertex indices per cell
  Context& context() const { return ctx_; }
  int dimension() const { return dim_; }
  int64_t size0() const { if (unstructured_) return n_[0]; int64_t t = 1; for (int d = 0; d < dim_; ++d) t *= n_[d]; return t; }
  const std::vector<int64_t>& cells() const { return n_; }
  SubDomainGridView subDomain(int s) const { return SubDomainGridView(s); }
  SubDomainGridView leafGridView() const { return SubDomainGridView(-1); }      // the MultiDomainGrid view itself
  // grid.startSubDomainMarking(); grid.addToSubDomain(s, e); grid.preUpdate/update/postUpdateSubDomains();
  void startSubDomainMarking() { marks_.assign((size_t)size0(), 0); marking_ = true; }
  void addToSubDomain(int s, int64_t cell) {
    if (!marking_ || s < 0 || s > 7 || cell < 0 || cell >= size0()) throw Exception(MDB_EINVAL, "addToSubDomain: not marking / bad argument");
    marks_[(size_t)cell] |= (uint8_t)(1u << s);
  }
  void preUpdateSubDomains() {}
  void updateSubDomains() { ctx_.check(mdb_subdomains(ctx_.get(), marks_.data()), "mdb_subdomains"); marking_ = false; }
  void postUpdateSubDomains() {}
  // device-side marking for large meshes: centre[axis] > threshold ? above : below (test/testpoisson.cc:152-158)
  void markHalfspace(int axis, double threshold, int below, int above) {
    ctx_.check(mdb_subdomains_halfspace(ctx_.get(), axis, threshold, below, above), "mdb_subdomains_halfspace");
  }
};

// ---- function spaces (multidomaingridfunctionspace.hh:152,215) -----------------------------------------------------------
class GridFunctionSpace {            // GridFunctionSpace<SDGV, QkLocalFiniteElementMap<...,k>, ConformingDirichletConstraints, VBE>
public:
  SubDomainGridView gv; int k; int child;
  bool coupling; int lkind, rkind; uint32_t lmask, rmask;      // set by CouplingGridFunctionSpace only
  int ncomp;                                                   // > 1: PowerCouplingGridFunctionSpace
  bool dg;                                                     // QkDGLocalFiniteElementMap<DF,RF,k,dim> instead of QkLocalFiniteElementMap
  bool simplex;                                                // PkLocalFiniteElementMap (simplex meshes)
  bool opb;                                                    // OPBLocalFiniteElementMap (orthonormal basis, simplex meshes)
  std::vector<GridFunctionSpace*> leaves;                      // non-empty: a composite / vector space (its leaf children, in order)
  GridFunctionSpace(SubDomainGridView gv_, int order, bool dg_ = false)
    : gv(gv_), k(order), child(-1), coupling(false), lkind(0), rkind(0), lmask(0), rmask(0), ncomp(1), dg(dg_), simplex(false), opb(false) {}
  virtual ~GridFunctionSpace() {}
  // the leaf spaces in lexicographic order (LexicographicOrderingTag of the composite spaces, test/stokesdarcy2.cc:657-664)
  void collect(std::vector<GridFunctionSpace*>& out) { if (leaves.empty()) out.push_back(this); else for (GridFunctionSpace* g : leaves) g->collect(out); }
  void collect_children(std::vector<int>& out) const { if (leaves.empty()) out.push_back(child); else for (const GridFunctionSpace* g : leaves) g->collect_children(out); }
  int fe() const {
    if (opb) { if (k < 1 || k > 3) throw Exception(MDB_EINVAL, "OPBGridFunctionSpace: k = 1, 2, 3"); return MDB_FE_OPB1 + k - 1; }
    if (simplex) { if (k != 1 && k != 2) throw Exception(MDB_EINVAL, "PkGridFunctionSpace: k = 1, 2 have device kernels in this build"); return k == 1 ? MDB_FE_P1 : MDB_FE_P2; }
    return dg ? (k == 1 ? MDB_FE_DGQ1 : MDB_FE_DGQ2) : (k == 1 ? MDB_FE_Q1 : MDB_FE_Q2);
  }
};
// GridFunctionSpace<SDGV, PkLocalFiniteElementMap<SDGV,DF,RF,k>, CON, VBE> on a simplex mesh (test/stokesdarcy2.cc:613-614, :646-655)
class PkGridFunctionSpace : public GridFunctionSpace {
public:
  PkGridFunctionSpace(SubDomainGridView gv_, int order) : GridFunctionSpace(gv_, order) { simplex = true; }
};
// GridFunctionSpace<SDGV, OPBLocalFiniteElementMap<DF,RF,k,dim,simplex,GMPField<512>>, NoConstraints, VBE> (test/stokesdarcy2.cc:615, :666-668)
class OPBGridFunctionSpace : public GridFunctionSpace {
public:
  OPBGridFunctionSpace(SubDomainGridView gv_, int order) : GridFunctionSpace(gv_, order) { simplex = true; opb = true; }
};
// VectorGridFunctionSpace<SDGV, PkLocalFiniteElementMap<..,k>, dim, VBE, VBE, CON> (test/stokesdarcy2.cc:646-655): dim copies of the
// scalar space, component block after component block
class VectorGridFunctionSpace : public GridFunctionSpace {
  std::vector<PkGridFunctionSpace> comps_;
public:
  VectorGridFunctionSpace(SubDomainGridView gv_, int order, int dim) : GridFunctionSpace(gv_, order), comps_((size_t)dim, PkGridFunctionSpace(gv_, order)) {
    simplex = true;
    for (PkGridFunctionSpace& c : comps_) leaves.push_back(&c);
  }
  VectorGridFunctionSpace(const VectorGridFunctionSpace&) = delete;
};
// CompositeGridFunctionSpace<VBE, LexicographicOrderingTag, Children...> (test/stokesdarcy2.cc:657-664)
class CompositeGridFunctionSpace : public GridFunctionSpace {
public:
  CompositeGridFunctionSpace(std::initializer_list<GridFunctionSpace*> children) : GridFunctionSpace(SubDomainGridView(-1), 0) { leaves.assign(children.begin(), children.end()); }
};
// GridFunctionSpace<SDGV, QkDGLocalFiniteElementMap<DF,double,k,dim>, CON, VBE> (test/testpoisson-dirichlet-neumann.cc:925-926, :968-971)
class QkDGGridFunctionSpace : public GridFunctionSpace {
public:
  QkDGGridFunctionSpace(SubDomainGridView gv_, int order) : GridFunctionSpace(gv_, order, true) {}
};

// CouplingGridFunctionSpace<MDGV, QkLocalFiniteElementMap<(dim-1)-cube,1>, SubProblemSubProblemInterface<MDGV,Left,Right>, NoConstraints, VBE>
// (couplinggridfunctionspace.hh:14-249; test/testmortarpoisson.cc:365-375): a Q1 space on the intersections whose inside cell
// satisfies `left` and whose outside cell satisfies `right`.  Pass it to MultiDomainGridFunctionSpace like any other child.
class CouplingGridFunctionSpace : public GridFunctionSpace {
public:
  template <class Left, class Right>
  CouplingGridFunctionSpace(int order, const Left& left, const Right& right) : GridFunctionSpace(SubDomainGridView(-1), order) {
    coupling = true; lkind = Left::kind(); lmask = left.mask; rkind = Right::kind(); rmask = right.mask;
  }
};

// PowerCouplingGridFunctionSpace<CouplingGFS, k, VBE>(couplinggfs)   (powercouplinggridfunctionspace.hh:13-75;
// test/testpowermortarpoisson.cc:384-386): k copies of a coupling space, block after block.  Pass IT (not the CouplingGFS it was
// built from) to MultiDomainGridFunctionSpace; `child` is then the first component block.
class PowerCouplingGridFunctionSpace : public GridFunctionSpace {
public:
  PowerCouplingGridFunctionSpace(const CouplingGridFunctionSpace& c, int k_) : GridFunctionSpace(c) { ncomp = k_; child = -1; }
};

class MultiDomainGridFunctionSpace { // children in argument order, LexicographicOrderingTag
  MultiDomainGrid& grid_;
  bool finalized_;
  int64_t ndof_;
public:
  MultiDomainGridFunctionSpace(MultiDomainGrid& grid, std::initializer_list<GridFunctionSpace*> children)
    : grid_(grid), finalized_(false), ndof_(0) {
    std::vector<GridFunctionSpace*> flat;
    for (GridFunctionSpace* g : children) g->collect(flat);
    for (GridFunctionSpace* g : flat)
      g->child = g->coupling && g->ncomp > 1
        ? context().check(mdb_add_power_coupling_space(context().get(), g->k == 1 ? MDB_FE_Q1 : MDB_FE_Q2, g->lkind, g->lmask, g->rkind, g->rmask, g->ncomp), "mdb_add_power_coupling_space")
        : g->coupling
        ? context().check(mdb_add_coupling_space(context().get(), g->k == 1 ? MDB_FE_Q1 : MDB_FE_Q2, g->lkind, g->lmask, g->rkind, g->rmask), "mdb_add_coupling_space")
        : context().check(mdb_add_space(context().get(), g->fe(), g->gv.subdomain), "mdb_add_space");
  }
  Context& context() const { return grid_.context(); }
  MultiDomainGrid& grid() const { return grid_; }
  // ordering(): computed lazily on first use (multidomaingridfunctionspace.hh:308-369); participants must exist by then
  void update() { if (!finalized_) { context().check(mdb_finalize(context().get(), &ndof_), "mdb_finalize"); finalized_ = true; } }
  bool finalized() const { return finalized_; }
  int64_t size() { update(); return ndof_; }
};

// ---- conditions (subdomainset.hh:82-133) -----------------------------------------------------------------------------------
struct SubDomainEqualityCondition {
  uint32_t mask; SubDomainEqualityCondition(std::initializer_list<int> s) : mask(0) { for (int v : s) mask |= 1u << v; }
  static int kind() { return MDB_COND_EQUALITY; }
};
struct SubDomainSubsetCondition {
  uint32_t mask; SubDomainSubsetCondition(std::initializer_list<int> s) : mask(0) { for (int v : s) mask |= 1u << v; }
  static int kind() { return MDB_COND_SUBSET; }
};

// ---- problem data and local operators (closed set) ----------------------------------------------------------------------------
inline mdb_function function(int kind, double p0 = 0, double p1 = 0, double p2 = 0, double p3 = 0, double p4 = 0, double p5 = 0) {
  mdb_function f; f.kind = kind; f.reserved = 0; f.p[0] = p0; f.p[1] = p1; f.p[2] = p2; f.p[3] = p3; f.p[4] = p4; f.p[5] = p5; return f;
}
inline mdb_bctype boundary(int kind) { mdb_bctype b; b.kind = kind; b.reserved = 0; for (int i = 0; i < 4; ++i) b.p[i] = 0; return b; }

struct Poisson {                     // Dune::PDELab::Poisson<F,B,J>(f, b, j, quadOrder), test/testpoisson.cc:218-219
  mdb_poisson_params p;
  Poisson(const mdb_function& f, const mdb_bctype& b, const mdb_function& j, int quadOrder) { p.f = f; p.bc = b; p.j = j; p.quad_order = quadOrder; p.reserved = 0; }
  static int id() { return MDB_OP_POISSON; }
  const mdb_bctype& bctype() const { return p.bc; }
};
struct L2 {                          // Dune::PDELab::L2(intorder) / test/simpletimeoperator.hh
  mdb_l2_params p;
  explicit L2(int intorder = 2) { p.quad_order = intorder; p.reserved = 0; }
  static int id() { return MDB_OP_L2; }
};
// [UPSTREAM] Dune::PDELab::ConvectionDiffusionDG<Params,FEM>(params, method, weights, alpha) as built by dglop()
// (test/testpoisson-dirichlet-neumann.cc:864-881) over the parameter class ConvectionDiffusionModelProblem (:169-285): A = I, b = 0,
// c = 0, source f, boundary type b (None on interior faces), Dirichlet value g, Neumann flux j.  QkDG spaces only.
namespace ConvectionDiffusionDGMethod { enum Type { NIPG = MDB_DG_METHOD_NIPG, SIPG = MDB_DG_METHOD_SIPG }; }
namespace ConvectionDiffusionDGWeights { enum Type { weightsOff = MDB_DG_WEIGHTS_OFF, weightsOn = MDB_DG_WEIGHTS_ON }; }
struct ConvectionDiffusionModelProblem {
  mdb_function f, g, j; mdb_bctype b;
  ConvectionDiffusionModelProblem(const mdb_function& f_, const mdb_bctype& b_, const mdb_function& g_, const mdb_function& j_) : f(f_), g(g_), j(j_), b(b_) {}
};
struct ConvectionDiffusionDG {
  mdb_convdiffdg_params p;
  ConvectionDiffusionDG(const ConvectionDiffusionModelProblem& param, ConvectionDiffusionDGMethod::Type method = ConvectionDiffusionDGMethod::NIPG,
                        ConvectionDiffusionDGWeights::Type weights = ConvectionDiffusionDGWeights::weightsOff, double alpha = 0.0, int intorderadd = 0) {
    p.f = param.f; p.bc = param.b; p.g = param.g; p.j = param.j; p.method = method; p.weights = weights; p.intorderadd = intorderadd; p.reserved = 0; p.alpha = alpha;
  }
  static int id() { return MDB_OP_CONVDIFF_DG; }
  const mdb_bctype& bctype() const { return p.bc; }
};
// ---- the Stokes-Darcy operators of test/stokesdarcy2.cc ------------------------------------------------------------------------------
// [UPSTREAM] TaylorHoodNavierStokes<NavierStokesParameters>(params, stokes_q) (test/stokesdarcy2.cc:226-251, :719-720): Stokes, full tensor
struct NavierStokesParameters {
  mdb_taylorhood_params p;
  // params.sub("parameters"): fluid.mu, fluid.rho, gravity; b = StokesBoundaryType (MDB_BC_ALL_NEUMANN = StressNeumann as shipped); j = PressureDropFlux
  NavierStokesParameters(double mu, double rho, double gravity, const mdb_bctype& b, const mdb_function& j) {
    p.mu = mu; p.rho = rho; p.f[0] = 0.0; p.f[1] = -gravity; p.bc = b; p.j = j; p.quad_order = 4; p.full_tensor = 1;
  }
};
struct TaylorHoodNavierStokes {
  mdb_taylorhood_params p;
  TaylorHoodNavierStokes(const NavierStokesParameters& params, int superintegration_order) : p(params.p) { p.quad_order = superintegration_order; }
  static int id() { return MDB_OP_TAYLORHOOD_STOKES; }
  const mdb_bctype& bctype() const { return p.bc; }
};
// DarcyParameters (test/stokesdarcy2.cc:254-407) + [UPSTREAM] ConvectionDiffusionDG<DarcyParameters, OPB>(params, method, weights, alpha) (:733-739)
struct DarcyParameters {
  mdb_darcy_params p;
  DarcyParameters(double kabs, double kabs_low, const mdb_bctype& b, const mdb_function& g, const mdb_function& j) {
    p.kabs = kabs; p.kabs_low = kabs_low; p.bc = b; p.g = g; p.j = j; p.high_group = 1; p.method = MDB_DG_METHOD_SIPG; p.weights = MDB_DG_WEIGHTS_ON; p.intorderadd = 0; p.alpha = 1.0;
  }
};
struct DarcyConvectionDiffusionDG {
  mdb_darcy_params p;
  DarcyConvectionDiffusionDG(const DarcyParameters& params, ConvectionDiffusionDGMethod::Type method, ConvectionDiffusionDGWeights::Type weights, double alpha)
    : p(params.p) { p.method = method; p.weights = weights; p.alpha = alpha; }
  static int id() { return MDB_OP_DARCY_DG; }
};
// CouplingParameters + StokesDarcyCouplingOperator (test/stokesdarcycouplingoperator.hh:11-252); Jacobian = NumericalJacobianCoupling(epsilon)
struct CouplingParameters {
  mdb_stokesdarcy_params p;
  CouplingParameters(double alpha, double gamma, double porosity, double gravity, double density, double viscosity, double epsilon, double kabs) {
    p.alpha = alpha; p.gamma = gamma; p.porosity = porosity; p.gravity = gravity; p.density = density; p.viscosity = viscosity; p.epsilon = epsilon; p.kabs = kabs;
    p.quirk = 1; p.reserved = 0;
  }
};
struct StokesDarcyCouplingOperator {
  mdb_stokesdarcy_params p; int jac_mode;
  explicit StokesDarcyCouplingOperator(const CouplingParameters& params, int jac = MDB_JAC_FD_BITMATCH) : p(params.p), jac_mode(jac) {}
  static int id() { return MDB_OP_STOKESDARCY_COUPLING; }
};

struct ContinuousValueContinuousFlowCoupling {   // test/proportionalflowcoupling.hh:94-239
  mdb_cvcf_params p; int jac_mode;
  ContinuousValueContinuousFlowCoupling(int intorder, double intensity, int jacobian_mode = MDB_JAC_FD_BITMATCH) : jac_mode(jacobian_mode) {
    p.quad_order = intorder; p.reserved = 0; p.intensity = intensity; }
  static int id() { return MDB_OP_CVCF_COUPLING; }
};
struct ProportionalFlowCoupling {                // test/proportionalflowcoupling.hh:10-91
  mdb_propflow_params p; int jac_mode;
  explicit ProportionalFlowCoupling(double intensity, int jacobian_mode = MDB_JAC_FD_BITMATCH) : jac_mode(jacobian_mode) { p.intensity = intensity; }
  static int id() { return MDB_OP_PROPFLOW_COUPLING; }
};

// Dune::PDELab::ConvectionDiffusionDGNeumannDirichletCoupling<CouplingParameters>(coupling_mode, param, method, weights, alpha, intorderadd)
// (test/neumanndirichletcoupling.hh:56-74) with the parameter class of test/testpoisson-dirichlet-neumann.cc:36-69 (A = I, b = 0, c = 0)
namespace CouplingMode { enum Type { Neumann = MDB_ND_MODE_NEUMANN, Dirichlet = MDB_ND_MODE_DIRICHLET, Both = MDB_ND_MODE_BOTH }; }
struct ConvectionDiffusionDGNeumannDirichletCoupling {
  mdb_ndcoupling_params p; int jac_mode;
  ConvectionDiffusionDGNeumannDirichletCoupling(CouplingMode::Type mode, ConvectionDiffusionDGMethod::Type method = ConvectionDiffusionDGMethod::NIPG,
                                                ConvectionDiffusionDGWeights::Type weights = ConvectionDiffusionDGWeights::weightsOff, double alpha = 0.0,
                                                int intorderadd = 0) : jac_mode(MDB_JAC_ANALYTIC) {
    p.mode = mode; p.method = method; p.weights = weights; p.intorderadd = intorderadd; p.alpha = alpha;
  }
  static int id() { return MDB_OP_ND_COUPLING; }
};

struct MortarPoissonCoupling {                   // test/testmortarpoisson.cc:117-250, ctor arg = "scaling factor for coupling" (:402)
  mdb_mortar_params p; int jac_mode;
  explicit MortarPoissonCoupling(double gamma_0, int jacobian_mode = MDB_JAC_FD_BITMATCH) : jac_mode(jacobian_mode) { p.gamma_0 = gamma_0; }
  static int id() { return MDB_OP_MORTAR_POISSON_COUPLING; }
};

// ---- participants --------------------------------------------------------------------------------------------------------------
// TypeBasedSubProblem<MultiGFS,MultiGFS,LOP,Condition,GFS...>(lop, condition): the child space is named by object here
template <class LOP, class Condition>
class SubProblem {
public:
  MultiDomainGridFunctionSpace& gfs; LOP lop; int handle;
  // `child` may be a composite space: the subproblem then spans all its leaves (subproblemlocalfunctionspace.hh:144-261)
  SubProblem(MultiDomainGridFunctionSpace& gfs_, const LOP& lop_, const Condition& cond, const GridFunctionSpace& child) : gfs(gfs_), lop(lop_) {
    std::vector<int> ch; child.collect_children(ch);
    handle = gfs.context().check(mdb_add_subproblem(gfs.context().get(), LOP::id(), &lop.p, sizeof(lop.p), Condition::kind(), cond.mask, ch.data(), (int)ch.size()),
                                 "mdb_add_subproblem");
  }
};
// Coupling<LocalSubProblem,RemoteSubProblem,CouplingOperator>(local, remote, op)   (coupling.hh:30-91)
template <class Local, class Remote, class Op>
class Coupling {
public:
  int handle;
  Coupling(const Local& local, const Remote& remote, const Op& op) {
    handle = local.gfs.context().check(mdb_add_coupling(local.gfs.context().get(), Op::id(), &op.p, sizeof(op.p), local.handle, remote.handle,
                                                        op.jac_mode), "mdb_add_coupling");
  }
};

// EnrichedCoupling<LocalSubProblem,RemoteSubProblem,CouplingOperator,couplingLFSIndex>(local, remote, op)   (coupling.hh:123-192); the
// coupling space is named by object instead of by its child index
template <class Local, class Remote, class Op>
class EnrichedCoupling {
public:
  int handle;
  EnrichedCoupling(const Local& local, const Remote& remote, const Op& op, const CouplingGridFunctionSpace& couplinggfs) {
    handle = local.gfs.context().check(mdb_add_enriched_coupling(local.gfs.context().get(), Op::id(), &op.p, sizeof(op.p), local.handle, remote.handle,
                                                                 couplinggfs.child, op.jac_mode), "mdb_add_enriched_coupling");
  }
  // the operator of test/testpowermortarpoisson.cc:117-262 on a power coupling space (one pass of alpha_generic per component block)
  EnrichedCoupling(const Local& local, const Remote& remote, const Op& op, const PowerCouplingGridFunctionSpace& couplinggfs) {
    handle = local.gfs.context().check(mdb_add_power_enriched_coupling(local.gfs.context().get(), Op::id(), &op.p, sizeof(op.p), local.handle, remote.handle,
                                                                       couplinggfs.child, couplinggfs.ncomp, op.jac_mode), "mdb_add_power_enriched_coupling");
  }
};

// ---- vectors and matrices --------------------------------------------------------------------------------------------------------
typedef std::vector<double> Vector;   // flat ISTL BlockVector<FieldVector<double,1>>; device-resident callers pass raw pointers

// constraints<R>(gfs, constrainSubProblem(sp, b)...).assemble(cg)   (constraints.hh:539-558)
class ConstraintsContainer { public: int64_t constrained; ConstraintsContainer() : constrained(0) {} int64_t size() const { return constrained; } };
class ConstraintsAssembler {
  MultiDomainGridFunctionSpace& gfs_; std::vector<int> sps_; std::vector<mdb_bctype> bcs_;
public:
  explicit ConstraintsAssembler(MultiDomainGridFunctionSpace& gfs) : gfs_(gfs) {}
  template <class SP> ConstraintsAssembler& constrainSubProblem(const SP& sp, const mdb_bctype& b) { sps_.push_back(sp.handle); bcs_.push_back(b); return *this; }
  void assemble(ConstraintsContainer& cg) {
    gfs_.update();
    gfs_.context().check(mdb_constraints_assemble(gfs_.context().get(), sps_.data(), bcs_.data(), (int)sps_.size(), &cg.constrained),
                         "mdb_constraints_assemble");
  }
};
inline ConstraintsAssembler constraints(MultiDomainGridFunctionSpace& gfs) { return ConstraintsAssembler(gfs); }

// interpolateOnTrialSpace(gfs, x, f0, sp0, f1, sp1, ...)   (interpolate.hh:76-81)
class Interpolator {
  MultiDomainGridFunctionSpace& gfs_; std::vector<int> sps_; std::vector<mdb_function> fns_;
public:
  explicit Interpolator(MultiDomainGridFunctionSpace& gfs) : gfs_(gfs) {}
  template <class SP> Interpolator& on(const mdb_function& f, const SP& sp) { fns_.push_back(f); sps_.push_back(sp.handle); return *this; }
  void apply(double* x, double time = 0.0) {
    gfs_.update();
    gfs_.context().check(mdb_interpolate(gfs_.context().get(), sps_.data(), fns_.data(), (int)sps_.size(), time, x), "mdb_interpolate");
  }
};
template <class SP0>
inline void interpolateOnTrialSpace(MultiDomainGridFunctionSpace& gfs, Vector& x, const mdb_function& f0, const SP0& sp0) {
  x.resize((size_t)gfs.size()); Interpolator(gfs).on(f0, sp0).apply(x.data());
}
template <class SP0, class SP1>
inline void interpolateOnTrialSpace(MultiDomainGridFunctionSpace& gfs, Vector& x, const mdb_function& f0, const SP0& sp0,
                                    const mdb_function& f1, const SP1& sp1) {
  x.resize((size_t)gfs.size()); Interpolator(gfs).on(f0, sp0).on(f1, sp1).apply(x.data());
}

// ---- GridOperator (gridoperator.hh:38-214) -----------------------------------------------------------------------------------------
class Jacobian;
class GridOperator {
  MultiDomainGridFunctionSpace& gfs_;
  int handle_;
  double weight_;
public:
  // GridOperator(gfsu, gfsv, cu, cv, mb, participants...): participants in argument order, subproblems then couplings
  GridOperator(MultiDomainGridFunctionSpace& gfsu, MultiDomainGridFunctionSpace&, const ConstraintsContainer&, const ConstraintsContainer&,
               const std::vector<int>& subproblems, const std::vector<int>& couplings) : gfs_(gfsu), weight_(1.0) {
    gfs_.update();
    handle_ = gfs_.context().check(mdb_gridoperator_create(gfs_.context().get(), subproblems.data(), (int)subproblems.size(),
                                                           couplings.data(), (int)couplings.size()), "mdb_gridoperator_create");
  }
  int handle() const { return handle_; }
  MultiDomainGridFunctionSpace& trialGridFunctionSpace() const { return gfs_; }
  Context& context() const { return gfs_.context(); }
  void setWeight(double w) { weight_ = w; }                         // localassembler.hh:220-223
  void setTime(double t) { context().check(mdb_set_time(context().get(), t), "mdb_set_time"); }
  // r += weight * R(x), constrained entries zeroed   (gridoperator.hh:89-92)
  void residual(const double* x, double* r) const { context().check(mdb_residual(context().get(), handle_, weight_, x, r), "mdb_residual"); }
  void residual(const Vector& x, Vector& r) const { r.resize(x.size()); residual(x.data(), r.data()); }
  // a += weight * J(x), constrained rows <- unit rows   (gridoperator.hh:94-97); the caller zeroes `a` first (a = 0.0)
  void jacobian(const double* x, Jacobian& a) const;
  void jacobian(const Vector& x, Jacobian& a) const { jacobian(x.data(), a); }
};

// Jacobian container: `Jacobian a(go)` runs fill_pattern (gridoperator.hh:83-87); `a = 0.0` zeroes the entries
class Jacobian {
  Context& ctx_; int handle_; int64_t nnz_;
public:
  explicit Jacobian(const GridOperator& go) : ctx_(go.context()), nnz_(0) {
    const int h = go.handle();
    handle_ = ctx_.check(mdb_matrix_create(ctx_.get(), &h, 1, &nnz_), "mdb_matrix_create");
  }
  Jacobian(const GridOperator& go0, const GridOperator& go1) : ctx_(go0.context()), nnz_(0) {     // OneStepGridOperator: dt0 + dt1 pattern
    const int h[2] = {go0.handle(), go1.handle()};
    handle_ = ctx_.check(mdb_matrix_create(ctx_.get(), h, 2, &nnz_), "mdb_matrix_create");
  }
  int handle() const { return handle_; }
  int64_t nonzeroes() const { return nnz_; }
  Jacobian& operator=(double v) {
    if (v != 0.0) throw Exception(MDB_EINVAL, "Jacobian = v: only 0.0 is supported");
    ctx_.check(mdb_matrix_zero(ctx_.get(), handle_), "mdb_matrix_zero"); return *this;
  }
  void pattern(std::vector<int64_t>& rowptr, std::vector<int64_t>& colidx, int64_t nrows) const {
    rowptr.resize((size_t)nrows + 1); colidx.resize((size_t)nnz_);
    ctx_.check(mdb_matrix_get_pattern(ctx_.get(), handle_, rowptr.data(), colidx.data()), "mdb_matrix_get_pattern");
  }
  void values(std::vector<double>& v) const { v.resize((size_t)nnz_); ctx_.check(mdb_matrix_get_values(ctx_.get(), handle_, v.data()), "mdb_matrix_get_values"); }
};
inline void GridOperator::jacobian(const double* x, Jacobian& a) const {
  context().check(mdb_jacobian(context().get(), handle_, a.handle(), weight_, x, 0), "mdb_jacobian");
}

// ---- linear solver backends + StationaryLinearProblemSolver ([UPSTREAM] PDELab; test/testpoisson.cc:291-298) --------------------------
struct LinearSolverResult { int iterations; bool converged; double reduction; double elapsed; };
class ISTLBackend_SEQ_Base {
protected:
  int method_, precond_; unsigned maxiter_; int sweeps_;
public:
  LinearSolverResult res;
  ISTLBackend_SEQ_Base(int method, int precond, unsigned maxiter, int sweeps = 0) : method_(method), precond_(precond), maxiter_(maxiter), sweeps_(sweeps) { res.iterations = 0; res.converged = false; res.reduction = 0; res.elapsed = 0; }
  void apply(Jacobian& a, Context& ctx, double* z, const double* r, double reduction) {
    mdb_solve_stats st;
    if (sweeps_ > 0) ctx.check(mdb_set_ssor_sweeps(ctx.get(), sweeps_), "mdb_set_ssor_sweeps");
    int rc = mdb_solve(ctx.get(), a.handle(), method_, precond_, reduction, (int)maxiter_, r, z, &st);
    res.iterations = st.iterations; res.converged = st.converged != 0; res.elapsed = st.seconds;
    res.reduction = st.defect0 > 0 ? st.defect / st.defect0 : 0.0;
    if (rc != MDB_ENOTCONVERGED) ctx.check(rc, "mdb_solve");
  }
  int method() const { return method_; }
  int precond() const { return precond_; }
  unsigned maxiter() const { return maxiter_; }
  void configure(Context& ctx) const { if (sweeps_ > 0) ctx.check(mdb_set_ssor_sweeps(ctx.get(), sweeps_), "mdb_set_ssor_sweeps"); }
  // StationaryLinearProblemSolver::apply in one device call (mdb_stationary_solve): x <- x - J(x)^-1 R(x)
  bool applyStationary(Context& ctx, int go, int mat, double* x, double reduction, bool keep_matrix) {
    mdb_solve_stats st;
    configure(ctx);
    int rc = mdb_stationary_solve(ctx.get(), go, mat, method_, precond_, reduction, (int)maxiter_, x, keep_matrix ? 1 : 0, &st);
    res.iterations = st.iterations; res.converged = st.converged != 0; res.elapsed = st.seconds;
    res.reduction = st.defect0 > 0 ? st.defect / st.defect0 : 0.0;
    if (rc != MDB_ENOTCONVERGED) ctx.check(rc, "mdb_stationary_solve");
    return true;
  }
};
// Jacobi is the fast device path.  The SSOR / ILU0 backends run the reference's sequential ISTL preconditioners operation
// for operation (sync-free sweeps, csrc/precond.cu): [UPSTREAM] ISTLBackend_SEQ_BCGS_SSOR = BiCGSTAB + SeqSSOR(A, 3, 1.0)
// (test/testpoisson.cc:291-292), ISTLBackend_SEQ_CG_SSOR = CG + SeqSSOR(A, 1, 1.0) (test/explicitlycoupledpoisson.cc:297-313).
struct ISTLBackend_SEQ_BCGS_SSOR : ISTLBackend_SEQ_Base { explicit ISTLBackend_SEQ_BCGS_SSOR(unsigned maxiter = 5000) : ISTLBackend_SEQ_Base(MDB_SOLVER_BICGSTAB, MDB_PRECOND_SSOR, maxiter, 3) {} };
struct ISTLBackend_SEQ_CG_SSOR : ISTLBackend_SEQ_Base { explicit ISTLBackend_SEQ_CG_SSOR(unsigned maxiter = 5000) : ISTLBackend_SEQ_Base(MDB_SOLVER_CG, MDB_PRECOND_SSOR, maxiter, 1) {} };
struct ISTLBackend_SEQ_BCGS_ILU0 : ISTLBackend_SEQ_Base { explicit ISTLBackend_SEQ_BCGS_ILU0(unsigned maxiter = 5000) : ISTLBackend_SEQ_Base(MDB_SOLVER_BICGSTAB, MDB_PRECOND_ILU0, maxiter) {} };
struct ISTLBackend_SEQ_CG_ILU0 : ISTLBackend_SEQ_Base { explicit ISTLBackend_SEQ_CG_ILU0(unsigned maxiter = 5000) : ISTLBackend_SEQ_Base(MDB_SOLVER_CG, MDB_PRECOND_ILU0, maxiter) {} };
struct ISTLBackend_SEQ_CG_Jac : ISTLBackend_SEQ_Base { explicit ISTLBackend_SEQ_CG_Jac(unsigned maxiter = 5000) : ISTLBackend_SEQ_Base(MDB_SOLVER_CG, MDB_PRECOND_JACOBI, maxiter) {} };
struct ISTLBackend_SEQ_BCGS_Jac : ISTLBackend_SEQ_Base { explicit ISTLBackend_SEQ_BCGS_Jac(unsigned maxiter = 5000) : ISTLBackend_SEQ_Base(MDB_SOLVER_BICGSTAB, MDB_PRECOND_JACOBI, maxiter) {} };
// [UPSTREAM] PDELab's multigrid-preconditioned backends ISTLBackend_SEQ_CG_AMG_SSOR / ISTLBackend_SEQ_BCGS_AMG_SSOR: here the Krylov method
// with one GEOMETRIC multigrid V-cycle per application (MDB_PRECOND_MG, csrc/mg.cu) — structured meshes, Q1 spaces; not used by the
// reference's drivers, offered because it turns ~1000 Jacobi-CG iterations at 256^3 into ~11
struct ISTLBackend_SEQ_CG_AMG_SSOR : ISTLBackend_SEQ_Base { explicit ISTLBackend_SEQ_CG_AMG_SSOR(unsigned maxiter = 5000) : ISTLBackend_SEQ_Base(MDB_SOLVER_CG, MDB_PRECOND_MG, maxiter) {} };
struct ISTLBackend_SEQ_BCGS_AMG_SSOR : ISTLBackend_SEQ_Base { explicit ISTLBackend_SEQ_BCGS_AMG_SSOR(unsigned maxiter = 5000) : ISTLBackend_SEQ_Base(MDB_SOLVER_BICGSTAB, MDB_PRECOND_MG, maxiter) {} };
// [UPSTREAM] ISTLBackend_SEQ_SuperLU(verbose) (test/stokesdarcy2.cc:784-785): a direct solve — banded LU with partial pivoting on the device
struct ISTLBackend_SEQ_SuperLU : ISTLBackend_SEQ_Base { explicit ISTLBackend_SEQ_SuperLU(bool = true) : ISTLBackend_SEQ_Base(MDB_SOLVER_DIRECT, MDB_PRECOND_NONE, 1) {} };

// ISTL_SEQ_Subblock_Backend<Dune::SeqSSOR, Dune::BiCGSTABSolver>(block, maxiter) (test/testpoisson-dirichlet-neumann.cc:264-370):
// BiCGSTAB + SeqSSOR(3) on the diagonal block of one child space; z outside the block is left untouched
class ISTL_SEQ_Subblock_Backend {
  int child_; unsigned maxiter_;
public:
  LinearSolverResult res;
  explicit ISTL_SEQ_Subblock_Backend(const GridFunctionSpace& block, unsigned maxiter = 5000) : child_(block.child), maxiter_(maxiter) {
    res.iterations = 0; res.converged = false; res.reduction = 0; res.elapsed = 0; }
  bool applyStationary(Context&, int, int, double*, double, bool) { return false; }      // block solves keep the three-call path
  void apply(Jacobian& a, Context& ctx, double* z, const double* r, double reduction) {
    mdb_solve_stats st;
    ctx.check(mdb_set_ssor_sweeps(ctx.get(), 3), "mdb_set_ssor_sweeps");
    int rc = mdb_solve_block(ctx.get(), a.handle(), child_, MDB_SOLVER_BICGSTAB, MDB_PRECOND_SSOR, reduction, (int)maxiter_, r, z, &st);
    res.iterations = st.iterations; res.converged = st.converged != 0; res.elapsed = st.seconds;
    res.reduction = st.defect0 > 0 ? st.defect / st.defect0 : 0.0;
    if (rc != MDB_ENOTCONVERGED) ctx.check(rc, "mdb_solve_block");
  }
};

// StationaryLinearProblemSolver(go, ls, reduction).apply(x [, keep_matrix]): J(x0), r = R(x0), solve J z = r, x = x0 - z
// (SURVEY §8c vii); keep_matrix = true reuses the Jacobian of the previous call (test/testpoisson-dirichlet-neumann.cc:806, :813)
template <class LS>
class StationaryLinearProblemSolver {
  GridOperator& go_; LS& ls_; double reduction_; Jacobian* a_;
  StationaryLinearProblemSolver(const StationaryLinearProblemSolver&);
  StationaryLinearProblemSolver& operator=(const StationaryLinearProblemSolver&);
public:
  StationaryLinearProblemSolver(GridOperator& go, LS& ls, double reduction) : go_(go), ls_(ls), reduction_(reduction), a_(0) {}
  ~StationaryLinearProblemSolver() { delete a_; }
  void setReduction(double reduction) { reduction_ = reduction; }
  double reduction() const { return reduction_; }
  void apply(Vector& x, bool keep_matrix = false) {
    if (!a_) { a_ = new Jacobian(go_); keep_matrix = false; }
    if (ls_.applyStationary(go_.context(), go_.handle(), a_->handle(), x.data(), reduction_, keep_matrix)) {
      if (!ls_.res.converged) throw Exception(MDB_ENOTCONVERGED, "StationaryLinearProblemSolver: linear solver did not converge");
      return;
    }
    if (!keep_matrix) { *a_ = 0.0; go_.jacobian(x, *a_); }
    Vector r(x.size(), 0.0), z(x.size(), 0.0);
    go_.residual(x, r);
    ls_.apply(*a_, go_.context(), z.data(), r.data(), reduction_);
    if (!ls_.res.converged) throw Exception(MDB_ENOTCONVERGED, "StationaryLinearProblemSolver: linear solver did not converge");
    for (size_t i = 0; i < x.size(); ++i) x[i] -= z[i];
  }
};

// ---- instationary problems ([UPSTREAM] PDELab OneStepGridOperator / OneStepMethod; test/testinstationarypoisson.cc:285-403) ------------
// interpolateOnSubProblems(g, sp0, g, sp1): the boundary-value functor handed to OneStepMethod::apply (:373)
struct SubProblemFunctions { std::vector<int> sps; std::vector<mdb_function> fns; };
template <class SP0, class SP1>
inline SubProblemFunctions interpolateOnSubProblems(const mdb_function& f0, const SP0& sp0, const mdb_function& f1, const SP1& sp1) {
  SubProblemFunctions f; f.sps.push_back(sp0.handle); f.fns.push_back(f0); f.sps.push_back(sp1.handle); f.fns.push_back(f1); return f;
}
// time-stepping parameter sets (the diagonally implicit ones)
struct ImplicitEulerParameter { static int scheme() { return MDB_ONESTEP_IMPLICIT_EULER; } double theta() const { return 1.0; } };
struct OneStepThetaParameter { double th; explicit OneStepThetaParameter(double t) : th(t) {} static int scheme() { return MDB_ONESTEP_THETA; } double theta() const { return th; } };
struct Alexander2Parameter { static int scheme() { return MDB_ONESTEP_ALEXANDER2; } double theta() const { return 0.0; } };
// OneStepGridOperator(go_dt0, go_dt1): one matrix whose pattern is the union of both operators' patterns
class OneStepGridOperator {
  GridOperator& go0_; GridOperator& go1_; Jacobian a_;
public:
  OneStepGridOperator(GridOperator& go0, GridOperator& go1) : go0_(go0), go1_(go1), a_(go0, go1) {}
  GridOperator& spatial() { return go0_; }
  GridOperator& temporal() { return go1_; }
  Jacobian& matrix() { return a_; }
  Context& context() const { return go0_.context(); }
};
// OneStepMethod(method, gridoperator, linear solver backend, reduction).apply(time, dt, xold, f, xnew): one mdb_onestep_apply per step;
// xold / xnew may be host vectors or device pointers (DeviceVector keeps the state in HBM between steps)
template <class Parameter, class LS>
class OneStepMethod {
  Parameter method_; OneStepGridOperator& go_; LS& ls_; double reduction_;
public:
  LinearSolverResult res;
  OneStepMethod(const Parameter& method, OneStepGridOperator& go, LS& ls, double reduction) : method_(method), go_(go), ls_(ls), reduction_(reduction) {}
  void apply(double time, double dt, const double* xold, const SubProblemFunctions& f, double* xnew) {
    Context& ctx = go_.context();
    mdb_solve_stats st;
    ls_.configure(ctx);
    int rc = mdb_onestep_apply(ctx.get(), go_.spatial().handle(), go_.temporal().handle(), go_.matrix().handle(), f.sps.data(), f.fns.data(),
                               (int)f.sps.size(), time, dt, Parameter::scheme(), method_.theta(), ls_.method(), ls_.precond(), reduction_,
                               (int)ls_.maxiter(), xold, xnew, &st);
    res.iterations = st.iterations; res.converged = st.converged != 0; res.elapsed = st.seconds;
    res.reduction = st.defect0 > 0 ? st.defect / st.defect0 : 0.0;
    ctx.check(rc, "mdb_onestep_apply");
  }
  void apply(double time, double dt, const Vector& xold, const SubProblemFunctions& f, Vector& xnew) {
    xnew.resize(xold.size()); apply(time, dt, xold.data(), f, xnew.data());
  }
};

// ---- output ([UPSTREAM] Dune::VTKWriter + multidomain/vtk.hh:83-98 addSolutionToVTKWriter; dune-istl writeMatrixToMatlab) -----------
// VTKWriter<SDGV> vtkwriter(sdgv, VTK::conforming); addSolutionToVTKWriter(vtkwriter, multigfs, u, subdomain_predicate(sd));
// vtkwriter.write(name, VTK::ascii)  — as every reference driver ends (e.g. test/testpoisson.cc:301-323).  One unstructured-grid
// piece per subdomain: its cells (VTK_QUAD / VTK_HEXAHEDRON, VTK_TRIANGLE on simplex meshes), their vertices, and per child space living on that subdomain one
// point-data array with the solution at the vertices (conforming output: for Q2 / P2 the corner values, like VTK::conforming; an
// orthonormal (modal DG) space is evaluated cell by cell at the corners and averaged per vertex).  Composite and vector spaces count with their leaves.
class VTKWriter {
  MultiDomainGrid& grid_; int subdomain_;
  struct Field { std::string name; std::vector<double> point_values; };
  std::vector<Field> fields_;
  std::vector<int64_t> cells_;                 // MultiDomainGrid indices of the subdomain's cells, host order
  std::vector<int64_t> vertex_of_point_;       // MultiDomainGrid vertex index of every output point (ascending = subdomain vertex order)
  std::vector<int64_t> point_of_vertex_;
  int corners() const { return grid_.unstructured() ? grid_.dimension() + 1 : (1 << grid_.dimension()); }
  void collect() {
    if (!cells_.empty()) return;
    std::vector<uint8_t> sd((size_t)grid_.size0());
    grid_.context().check(mdb_get_subdomains(grid_.context().get(), sd.data()), "mdb_get_subdomains");
    const int dim = grid_.dimension(); const std::vector<int64_t>& n = grid_.cells();
    int64_t nv = 1; for (int d = 0; d < dim; ++d) nv *= n[d] + 1;
    if (grid_.unstructured()) nv = (int64_t)grid_.vertices().size() / dim;
    std::vector<char> used((size_t)nv, 0);
    for (int64_t e = 0; e < grid_.size0(); ++e) {
      if (subdomain_ >= 0 && !((sd[(size_t)e] >> subdomain_) & 1)) continue;
      cells_.push_back(e);
      for (int c = 0; c < corners(); ++c) used[(size_t)vertex(e, c)] = 1;
    }
    point_of_vertex_.assign((size_t)nv, -1);
    for (int64_t v = 0; v < nv; ++v) if (used[(size_t)v]) { point_of_vertex_[(size_t)v] = (int64_t)vertex_of_point_.size(); vertex_of_point_.push_back(v); }
  }
  int64_t vertex(int64_t e, int corner) const {          // [UPSTREAM] Yasp: corner bit d = offset along axis d; vertices lexicographic
    const int dim = grid_.dimension(); const std::vector<int64_t>& n = grid_.cells();
    if (grid_.unstructured()) return grid_.elements()[(size_t)(e * (dim + 1) + corner)];
    int64_t idx = 0, stride = 1, rem = e;
    for (int d = 0; d < dim; ++d) { const int64_t i = rem % n[d] + ((corner >> d) & 1); rem /= n[d]; idx += i * stride; stride *= n[d] + 1; }
    return idx;
  }
  // ---- local spaces as the cell DOF maps list them ----------------------------------------------------------------------------
  static void flatten(const GridFunctionSpace* g, std::vector<const GridFunctionSpace*>& out) {          // composite / vector spaces: their leaves in order
    if (g->leaves.empty()) out.push_back(g); else for (size_t i = 0; i < g->leaves.size(); ++i) flatten(g->leaves[i], out);
  }
public:
  // coefficients of a leaf space on one cell: (k+1)^dim for Qk / QkDG, binom(k + dim, dim) for Pk and the orthonormal basis of degree k
  static int local_size(const GridFunctionSpace& h, int dim) {
    int n = 1;
    if (h.simplex) { for (int d = 1; d <= dim; ++d) n = n * (h.k + d) / d; return n; }
    for (int d = 0; d < dim; ++d) n *= h.k + 1;
    return n;
  }
  // local index of the shape function that sits at corner c.  Qk: lattice index with a_d = k * bit_d.  Pk on a triangle ([UPSTREAM]
  // Pk2DLocalBasis: lattice order, i fastest): corner 0 -> 0, corner 1 -> k, corner 2 -> the last one; P1 on any simplex: c.
  static int corner_index(const GridFunctionSpace& g, int dim, int c) {
    if (g.simplex) return (g.k == 1 || c == 0) ? c : (c == 1 ? g.k : local_size(g, dim) - 1);
    int li = 0, stride = 1;
    for (int d = 0; d < dim; ++d) { li += g.k * ((c >> d) & 1) * stride; stride *= g.k + 1; }
    return li;
  }
  // values of the orthonormal basis of degree k (OPBLocalFiniteElementMap, test/stokesdarcy2.cc:615) at the corners (0,0), (1,0), (0,1) of
  // the reference triangle: out[c * n + i].  Gram-Schmidt of the monomials in Dune::MonomialBasis order (1 | x, y | x^2, xy, y^2 | ...) with
  // respect to the L2 product of the reference triangle, i.e. L = C^-1 for the Cholesky factor C of G_ij = a! b! / (a + b + 2)!; here in
  // long double (output only: the device and the oracle evaluate the 60-digit tables of csrc/opb_tables.h; a test compares the two)
  static void opb_matrix(int k, std::vector<long double>& L, std::vector<int>& ea, std::vector<int>& eb) {
    std::vector<int> pa(1, 0), pb(1, 0);
    ea.assign(1, 0); eb.assign(1, 0);
    for (int d = 1; d <= k; ++d) {
      std::vector<int> ra(1, d), rb(1, 0);
      for (size_t i = 0; i < pa.size(); ++i) { ra.push_back(pa[i]); rb.push_back(pb[i] + 1); }
      ea.insert(ea.end(), ra.begin(), ra.end()); eb.insert(eb.end(), rb.begin(), rb.end());
      pa = ra; pb = rb;
    }
    const int n = (int)ea.size();
    std::vector<long double> fact(2 * k + 3, 1.0L);
    for (size_t i = 1; i < fact.size(); ++i) fact[i] = fact[i - 1] * (long double)i;
    std::vector<long double> C((size_t)n * n, 0.0L);
    L.assign((size_t)n * n, 0.0L);
    for (int i = 0; i < n; ++i)
      for (int j = 0; j <= i; ++j) {
        const int a = ea[i] + ea[j], b = eb[i] + eb[j];
        long double t = fact[a] * fact[b] / fact[a + b + 2];
        for (int m = 0; m < j; ++m) t -= C[(size_t)i * n + m] * C[(size_t)j * n + m];
        C[(size_t)i * n + j] = i == j ? std::sqrt(t) : t / C[(size_t)j * n + j];
      }
    for (int col = 0; col < n; ++col)                                                    // C L = I by forward substitution
      for (int i = col; i < n; ++i) {
        long double t = i == col ? 1.0L : 0.0L;
        for (int m = col; m < i; ++m) t -= C[(size_t)i * n + m] * L[(size_t)m * n + col];
        L[(size_t)i * n + col] = t / C[(size_t)i * n + i];
      }
  }
  // out[c * n + i] = phi_i(corner c); grad (optional) [(c * n + i) * 2 + d] = d phi_i / d xhat_d at corner c (REFERENCE gradients)
  static void opb_corner_values(int k, std::vector<double>& out, std::vector<double>* grad = 0) {
    std::vector<long double> L; std::vector<int> ea, eb;
    opb_matrix(k, L, ea, eb);
    const int n = (int)ea.size();
    static const long double cx[3] = {0.0L, 1.0L, 0.0L}, cy[3] = {0.0L, 0.0L, 1.0L};
    out.assign((size_t)3 * n, 0.0);
    if (grad) grad->assign((size_t)6 * n, 0.0);
    for (int c = 0; c < 3; ++c) {
      std::vector<long double> px(k + 1, 1.0L), py(k + 1, 1.0L);
      for (int e = 1; e <= k; ++e) { px[e] = px[e - 1] * cx[c]; py[e] = py[e - 1] * cy[c]; }
      for (int i = 0; i < n; ++i) {
        long double v = 0.0L, gx = 0.0L, gy = 0.0L;
        for (int j = 0; j <= i; ++j) {
          const long double l = L[(size_t)i * n + j];
          v += l * px[ea[j]] * py[eb[j]];
          if (ea[j] > 0) gx += l * (long double)ea[j] * px[ea[j] - 1] * py[eb[j]];
          if (eb[j] > 0) gy += l * (long double)eb[j] * px[ea[j]] * py[eb[j] - 1];
        }
        out[(size_t)c * n + i] = (double)v;
        if (grad) { (*grad)[((size_t)c * n + i) * 2] = (double)gx; (*grad)[((size_t)c * n + i) * 2 + 1] = (double)gy; }
      }
    }
  }
  // The derived fields the reference adds to the Darcy piece (test/stokesdarcy2.cc:876-884), on plain arrays like point_values:
  //   VTKPressureFromPotential (:443-494):   p = (phi - x_last) * gravity * density
  //   VTKDarcyFlowFromPotential (:497-560):  v = -A grad(phi) / porosity,  A = the cell's (isotropic) tensor value (DarcyParameters::A :285-290)
  // evaluated per cell at its corners from the orthonormal coefficients (grad = J^-T grad_ref on the affine triangle) and averaged per
  // vertex.  corner_xy[(ci * 3 + c) * 2 + d]: coordinates of corner c of listed cell ci; cell_A[ci]: its tensor value.
  static void darcy_point_fields(const std::vector<const GridFunctionSpace*>& leaves, size_t gi, const std::vector<int64_t>& cells, const std::vector<uint8_t>& sd,
                                 const std::vector<int64_t>& ptr, const std::vector<int64_t>& dofs, const double* x, const std::vector<int64_t>& cell_points,
                                 const std::vector<double>& corner_xy, const std::vector<double>& cell_A, double porosity, double gravity, double density,
                                 std::vector<double>& p, std::vector<double>& v0, std::vector<double>& v1 /* all sized to the number of points */) {
    const GridFunctionSpace* g = leaves[gi];
    if (!g->opb) throw Exception(MDB_EINVAL, "VTKWriter: the Darcy fields are derived from an orthonormal (OPB) potential space");
    const int nloc = local_size(*g, 2);
    std::vector<double> phi, gphi; opb_corner_values(g->k, phi, &gphi);
    std::vector<int> count(p.size(), 0);
    std::fill(p.begin(), p.end(), 0.0); std::fill(v0.begin(), v0.end(), 0.0); std::fill(v1.begin(), v1.end(), 0.0);
    for (size_t ci = 0; ci < cells.size(); ++ci) {
      const int64_t e = cells[ci];
      int64_t off = ptr[(size_t)e];
      for (size_t hi = 0; hi < gi; ++hi) {
        const GridFunctionSpace* h = leaves[hi];
        if (h->coupling) continue;
        if (h->gv.subdomain < 0 || ((sd[(size_t)e] >> h->gv.subdomain) & 1)) off += local_size(*h, 2);
      }
      if (off + nloc > ptr[(size_t)e + 1]) throw Exception(MDB_EINVAL, "VTKWriter: the cell DOF map is shorter than the listed spaces");
      const double* P = &corner_xy[ci * 6];
      const double ax = P[2] - P[0], ay = P[3] - P[1], bx = P[4] - P[0], by = P[5] - P[1];      // columns of J: x = P0 + a xhat + b yhat
      const double det = ax * by - ay * bx;
      for (int c = 0; c < 3; ++c) {
        double val = 0.0, gx = 0.0, gy = 0.0;
        for (int i = 0; i < nloc; ++i) {
          const double u = x[(size_t)dofs[(size_t)(off + i)]];
          val += u * phi[(size_t)c * nloc + i];
          gx += u * gphi[((size_t)c * nloc + i) * 2]; gy += u * gphi[((size_t)c * nloc + i) * 2 + 1];
        }
        const double dx = (by * gx - ay * gy) / det, dy = (-bx * gx + ax * gy) / det;           // J^-T (gx, gy)
        const size_t pt = (size_t)cell_points[ci * 3 + (size_t)c];
        p[pt] += (val - P[2 * c + 1]) * gravity * density;
        v0[pt] += -cell_A[ci] * dx / porosity; v1[pt] += -cell_A[ci] * dy / porosity;
        count[pt] += 1;
      }
    }
    for (size_t pt = 0; pt < p.size(); ++pt) if (count[pt]) { p[pt] /= count[pt]; v0[pt] /= count[pt]; v1[pt] /= count[pt]; }
  }
  // The evaluation behind addSolution on plain arrays (unit-tested on the CPU, tests/test_vtk_spaces.py): point values of leaf `gi` of
  // `leaves` on the listed cells.  ptr / dofs: the cell DOF maps of mdb_get_dofmap (children in order, a child present iff its subdomain
  // holds the cell); cell_points[ci * corners + c]: output point of corner c of listed cell ci.  Conforming spaces: the coefficient of the
  // shape function at the corner.  Orthonormal (modal, discontinuous) spaces: the polynomial of each cell evaluated at its corners,
  // averaged over the cells that meet in the point (the reference writes such spaces through a SubsamplingVTKWriter, one point per cell
  // corner; a conforming piece holds one value per vertex).
  static void point_values(const std::vector<const GridFunctionSpace*>& leaves, size_t gi, int dim, int corners, const std::vector<int64_t>& cells,
                           const std::vector<uint8_t>& sd, const std::vector<int64_t>& ptr, const std::vector<int64_t>& dofs, const double* x,
                           const std::vector<int64_t>& cell_points, std::vector<double>& values /* sized to the number of points */) {
    const GridFunctionSpace* g = leaves[gi];
    std::vector<double> corner_phi; std::vector<int> count;
    const int nloc = local_size(*g, dim);
    if (g->opb) {
      if (dim != 2) throw Exception(MDB_EINVAL, "VTKWriter: orthonormal spaces are built on triangles only");
      opb_corner_values(g->k, corner_phi); count.assign(values.size(), 0);
      std::fill(values.begin(), values.end(), 0.0);
    }
    for (size_t ci = 0; ci < cells.size(); ++ci) {
      const int64_t e = cells[ci];
      int64_t off = ptr[(size_t)e];                    // position of this leaf's block inside the cell's local space
      for (size_t hi = 0; hi < gi; ++hi) {
        const GridFunctionSpace* h = leaves[hi];
        if (h->coupling) continue;
        if (h->gv.subdomain < 0 || ((sd[(size_t)e] >> h->gv.subdomain) & 1)) off += local_size(*h, dim);
      }
      if (off + nloc > ptr[(size_t)e + 1]) throw Exception(MDB_EINVAL, "VTKWriter: the cell DOF map is shorter than the listed spaces");
      for (int c = 0; c < corners; ++c) {
        const size_t pt = (size_t)cell_points[ci * (size_t)corners + (size_t)c];
        if (g->opb) {
          double v = 0.0;
          for (int i = 0; i < nloc; ++i) v += x[(size_t)dofs[(size_t)(off + i)]] * corner_phi[(size_t)c * nloc + i];
          values[pt] += v; count[pt] += 1;
        } else {
          values[pt] = x[(size_t)dofs[(size_t)(off + corner_index(*g, dim, c))]];
        }
      }
    }
    if (g->opb) for (size_t pt = 0; pt < values.size(); ++pt) if (count[pt]) values[pt] /= (double)count[pt];
  }
  VTKWriter(MultiDomainGrid& grid, const SubDomainGridView& gv) : grid_(grid), subdomain_(gv.subdomain) {}
  // addSolutionToVTKWriter(vtkwriter, multigfs, x, subdomain_predicate): every leaf space whose grid view is this writer's subdomain
  // (composite and vector spaces count with their leaves, test/stokesdarcy2.cc:836-846, :871-876)
  void addSolution(MultiDomainGridFunctionSpace& gfs, const Vector& x, std::initializer_list<const GridFunctionSpace*> children, const std::string& name = "u") {
    collect();
    gfs.update();
    Context& ctx = gfs.context();
    const int dim = grid_.dimension();
    std::vector<int64_t> ptr((size_t)grid_.size0() + 1);
    ctx.check(mdb_get_dofmap(ctx.get(), ptr.data(), 0), "mdb_get_dofmap");
    std::vector<int64_t> dofs((size_t)ptr.back());
    ctx.check(mdb_get_dofmap(ctx.get(), ptr.data(), dofs.data()), "mdb_get_dofmap");
    std::vector<uint8_t> sd((size_t)grid_.size0());
    ctx.check(mdb_get_subdomains(ctx.get(), sd.data()), "mdb_get_subdomains");
    std::vector<const GridFunctionSpace*> leaves;
    for (const GridFunctionSpace* g : children) flatten(g, leaves);
    std::vector<int64_t> cell_points(cells_.size() * (size_t)corners());
    for (size_t ci = 0; ci < cells_.size(); ++ci)
      for (int c = 0; c < corners(); ++c) cell_points[ci * (size_t)corners() + (size_t)c] = point_of_vertex_[(size_t)vertex(cells_[ci], c)];
    int nout = 0;
    for (size_t gi = 0; gi < leaves.size(); ++gi) {
      const GridFunctionSpace* g = leaves[gi];
      if (g->coupling || g->gv.subdomain != subdomain_) continue;              // the predicate
      Field f; f.name = (nout++ == 0) ? name : name + "_" + std::to_string(nout - 1);
      f.point_values.assign(vertex_of_point_.size(), 0.0);
      point_values(leaves, gi, dim, corners(), cells_, sd, ptr, dofs, x.data(), cell_points, f.point_values);
      fields_.push_back(f);
    }
  }
  // .addVertexFunction<VTKPressureFromPotential>(TreePath<1>(), "p", darcyParams).addVertexFunction(VTKDarcyFlowFromPotential(darcyParams),
  // TreePath<1>(), "v")  (test/stokesdarcy2.cc:876-884): fields "p", "v_0", "v_1" derived from the potential space `potential` (a leaf of
  // `children`).  groups: elementIndexToPhysicalGroup of the mesh; the tensor is kabs on cells of `high_group`, kabs_low elsewhere.
  void addDarcyFields(MultiDomainGridFunctionSpace& gfs, const Vector& x, std::initializer_list<const GridFunctionSpace*> children, const GridFunctionSpace& potential,
                      const std::vector<int>& groups, double kabs, double kabs_low, int high_group, double porosity, double gravity, double density) {
    collect();
    gfs.update();
    if (!grid_.unstructured() || grid_.dimension() != 2) throw Exception(MDB_EINVAL, "addDarcyFields: triangle meshes");
    Context& ctx = gfs.context();
    std::vector<int64_t> ptr((size_t)grid_.size0() + 1);
    ctx.check(mdb_get_dofmap(ctx.get(), ptr.data(), 0), "mdb_get_dofmap");
    std::vector<int64_t> dofs((size_t)ptr.back());
    ctx.check(mdb_get_dofmap(ctx.get(), ptr.data(), dofs.data()), "mdb_get_dofmap");
    std::vector<uint8_t> sd((size_t)grid_.size0());
    ctx.check(mdb_get_subdomains(ctx.get(), sd.data()), "mdb_get_subdomains");
    std::vector<const GridFunctionSpace*> leaves;
    for (const GridFunctionSpace* g : children) flatten(g, leaves);
    size_t gi = leaves.size();
    for (size_t i = 0; i < leaves.size(); ++i) if (leaves[i] == &potential) gi = i;
    if (gi == leaves.size() || potential.gv.subdomain != subdomain_) throw Exception(MDB_EINVAL, "addDarcyFields: the potential space is not a leaf on this writer's subdomain");
    std::vector<int64_t> cell_points(cells_.size() * 3);
    std::vector<double> corner_xy(cells_.size() * 6), cell_A(cells_.size());
    for (size_t ci = 0; ci < cells_.size(); ++ci) {
      for (int c = 0; c < 3; ++c) {
        const int64_t v = vertex(cells_[ci], c);
        cell_points[ci * 3 + (size_t)c] = point_of_vertex_[(size_t)v];
        corner_xy[(ci * 3 + (size_t)c) * 2] = grid_.vertices()[(size_t)(v * 2)]; corner_xy[(ci * 3 + (size_t)c) * 2 + 1] = grid_.vertices()[(size_t)(v * 2 + 1)];
      }
      cell_A[ci] = groups[(size_t)cells_[ci]] == high_group ? kabs : kabs_low;
    }
    Field p, v0, v1; p.name = "p"; v0.name = "v_0"; v1.name = "v_1";
    p.point_values.assign(vertex_of_point_.size(), 0.0); v0.point_values = p.point_values; v1.point_values = p.point_values;
    darcy_point_fields(leaves, gi, cells_, sd, ptr, dofs, x.data(), cell_points, corner_xy, cell_A, porosity, gravity, density, p.point_values, v0.point_values, v1.point_values);
    fields_.push_back(p); fields_.push_back(v0); fields_.push_back(v1);
  }
  // write(name, VTK::ascii): name.vtu
  void write(const std::string& name) {
    collect();
    const int dim = grid_.dimension(); const std::vector<int64_t>& n = grid_.cells();
    std::FILE* fp = std::fopen((name + ".vtu").c_str(), "w");
    if (!fp) throw Exception(MDB_EINVAL, "VTKWriter: cannot open " + name + ".vtu");
    std::fprintf(fp, "<?xml version=\"1.0\"?>\n<VTKFile type=\"UnstructuredGrid\" version=\"0.1\" byte_order=\"LittleEndian\">\n<UnstructuredGrid>\n");
    std::fprintf(fp, "<Piece NumberOfPoints=\"%lld\" NumberOfCells=\"%lld\">\n", (long long)vertex_of_point_.size(), (long long)cells_.size());
    std::fprintf(fp, "<PointData%s>\n", fields_.empty() ? "" : (" Scalars=\"" + fields_[0].name + "\"").c_str());
    for (size_t f = 0; f < fields_.size(); ++f) {
      std::fprintf(fp, "<DataArray type=\"Float64\" Name=\"%s\" NumberOfComponents=\"1\" format=\"ascii\">\n", fields_[f].name.c_str());
      for (size_t i = 0; i < fields_[f].point_values.size(); ++i) std::fprintf(fp, "%.17g\n", fields_[f].point_values[i]);
      std::fprintf(fp, "</DataArray>\n");
    }
    std::fprintf(fp, "</PointData>\n<Points>\n<DataArray type=\"Float64\" NumberOfComponents=\"3\" format=\"ascii\">\n");
    for (size_t p = 0; p < vertex_of_point_.size(); ++p) {
      int64_t rem = vertex_of_point_[p]; double xyz[3] = {0, 0, 0};
      if (grid_.unstructured()) for (int d = 0; d < dim; ++d) xyz[d] = grid_.vertices()[(size_t)(rem * dim + d)];
      else for (int d = 0; d < dim; ++d) { xyz[d] = grid_.lower()[d] + (double)(rem % (n[d] + 1)) * ((grid_.upper()[d] - grid_.lower()[d]) / (double)n[d]); rem /= n[d] + 1; }
      std::fprintf(fp, "%.17g %.17g %.17g\n", xyz[0], xyz[1], xyz[2]);
    }
    std::fprintf(fp, "</DataArray>\n</Points>\n<Cells>\n<DataArray type=\"Int64\" Name=\"connectivity\" format=\"ascii\">\n");
    static const int vtk_quad[4] = {0, 1, 3, 2}, vtk_hex[8] = {0, 1, 3, 2, 4, 5, 7, 6};      // DUNE corner order -> VTK order
    for (size_t ci = 0; ci < cells_.size(); ++ci) {
      for (int c = 0; c < corners(); ++c) {
        const int dc = grid_.unstructured() ? c : dim == 3 ? vtk_hex[c] : dim == 2 ? vtk_quad[c] : c;
        std::fprintf(fp, "%lld ", (long long)point_of_vertex_[(size_t)vertex(cells_[ci], dc)]);
      }
      std::fprintf(fp, "\n");
    }
    std::fprintf(fp, "</DataArray>\n<DataArray type=\"Int64\" Name=\"offsets\" format=\"ascii\">\n");
    for (size_t ci = 0; ci < cells_.size(); ++ci) std::fprintf(fp, "%lld\n", (long long)((ci + 1) * (size_t)corners()));
    std::fprintf(fp, "</DataArray>\n<DataArray type=\"UInt8\" Name=\"types\" format=\"ascii\">\n");
    for (size_t ci = 0; ci < cells_.size(); ++ci) std::fprintf(fp, "%d\n", grid_.unstructured() ? (dim == 3 ? 10 : 5) : dim == 3 ? 12 : dim == 2 ? 9 : 3);
    std::fprintf(fp, "</DataArray>\n</Cells>\n</Piece>\n</UnstructuredGrid>\n</VTKFile>\n");
    std::fclose(fp);
  }
};

// Dune::writeMatrixToMatlab(J.base(), "jacobian.mat") (test/testmortarpoisson.cc:453; [UPSTREAM] dune-istl io.hh): one line
// "row col value" per stored entry, 1-based indices, rows ascending, entries in BCRS (ascending column) order, 18 digits
inline void writeMatrixToMatlab(const Jacobian& a, int64_t nrows, const std::string& filename, int outputPrecision = 18) {
  std::vector<int64_t> rowptr, colidx; std::vector<double> v;
  a.pattern(rowptr, colidx, nrows); a.values(v);
  std::FILE* fp = std::fopen(filename.c_str(), "w");
  if (!fp) throw Exception(MDB_EINVAL, "writeMatrixToMatlab: cannot open " + filename);
  for (int64_t r = 0; r < nrows; ++r)
    for (int64_t p = rowptr[(size_t)r]; p < rowptr[(size_t)r + 1]; ++p)
      std::fprintf(fp, "%lld %lld %.*g\n", (long long)(r + 1), (long long)(colidx[(size_t)p] + 1), outputPrecision, v[(size_t)p]);
  std::fclose(fp);
}

} // namespace mdb200
#endif
