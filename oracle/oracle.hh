// oracle/oracle.hh — TEST INFRASTRUCTURE ONLY (parity oracle + CPU baseline).
//
// PARITY UNPINNED: the reference (smuething/dune-multidomain) cannot be built in
// this image (it needs dune-common/geometry/grid/istl/localfunctions/typetree/
// pdelab/multidomaingrid, none of which are present) and its own tests contain
// no expected values.  This file is therefore a from-scratch CPU
// restatement of the reference's algorithm for the hot path, pinned only by
// self-derived known-answer tests (tests/test_oracle*.py; the table in DESIGN.md
// section 2 lists them per assumption): closed-form element matrices, child blocks
// against Kronecker products of 1-D matrices, the couplings and the lambda terms in
// closed form, the Dirichlet set from the lattice, the solver stack against SciPy /
// textbook recurrences, patch tests, FD-vs-analytic, SciPy direct solves.
//
// Nothing under the product tree (dune-multidomain_b200/, include/) may include,
// link or call this; only tests/, __graft_entry__.smoke() and bench.py's CPU
// baseline do.
//
// What is restated, and from where (paths relative to /root/reference):
//   traversal                globalassembler.hh:43-298
//   dispatch                 residualassemblerfunctors.hh:16-257, jacobianassemblerfunctors.hh:16-158,
//                            patternassemblerfunctors.hh:16-150
//   engine flag logic        jacobianassemblerengine.hh:149-192, residualassemblerengine.hh:117-188,
//                            patternassemblerengine.hh:96-139
//   scatter / Dirichlet      jacobianassemblerengine.hh:195-364,480-504, residualassemblerengine.hh:190-343,553-567
//   participants             subproblem.hh:84-87, coupling.hh:78-83, subdomainset.hh:102-131
//   local spaces             multidomainlocalfunctionspace.hh:79-124,191-237,348-358
//   coupling pattern + FD    couplingutilities.hh:29-49, 59-141
//   operators                test/proportionalflowcoupling.hh, test/instationarypoissonoperator.hh,
//                            test/simpletimeoperator.hh, problem data test/testpoisson.cc:32-116
//   constraints              constraints.hh:312-558, constraintsfunctors.hh:41-131
//   interpolation            interpolate.hh:18-81
// [UPSTREAM] semantics assumed (SURVEY §8c i-ix) are marked in the code.
#ifndef ORACLE_HH
#define ORACLE_HH

#include <cstdint>
#include <cmath>
#include <vector>
#include <array>
#include <string>
#include <algorithm>
#include <memory>
#include <stdexcept>
#include <cassert>
#include <cstring>

#include "../include/mdb200.h"   // descriptor structs / enums only (the interface specification)

namespace orc {

constexpr int MAXLOC = 27;   // Q2 hex
constexpr int MAXQ1D = 5;

// ---------------------------------------------------------------------------------------------
// Quadrature: [UPSTREAM] (viii) Gauss-Legendre tensor rules on [0,1]^d, order p -> ceil((p+1)/2) points.
// Point order: axis 0 fastest.
// ---------------------------------------------------------------------------------------------
struct Rule1D { int m; double x[MAXQ1D]; double w[MAXQ1D]; };

inline Rule1D gauss1d(int order)
{
  Rule1D r;
  int m = (order + 2) / 2;           // ceil((order+1)/2)
  if (m < 1) m = 1;
  if (m > 5) throw std::runtime_error("quadrature order too high");
  r.m = m;
  switch (m) {
  case 1: r.x[0] = 0.5; r.w[0] = 1.0; break;
  case 2: {
    double a = 0.5 / std::sqrt(3.0);
    r.x[0] = 0.5 - a; r.x[1] = 0.5 + a; r.w[0] = 0.5; r.w[1] = 0.5; break; }
  case 3: {
    double a = 0.5 * std::sqrt(3.0 / 5.0);
    r.x[0] = 0.5 - a; r.x[1] = 0.5; r.x[2] = 0.5 + a;
    r.w[0] = 5.0 / 18.0; r.w[1] = 8.0 / 18.0; r.w[2] = 5.0 / 18.0; break; }
  case 4: {
    double s = std::sqrt(6.0 / 5.0);
    double a = 0.5 * std::sqrt(3.0 / 7.0 - 2.0 / 7.0 * s);
    double b = 0.5 * std::sqrt(3.0 / 7.0 + 2.0 / 7.0 * s);
    double wa = (18.0 + std::sqrt(30.0)) / 72.0;
    double wb = (18.0 - std::sqrt(30.0)) / 72.0;
    r.x[0] = 0.5 - b; r.x[1] = 0.5 - a; r.x[2] = 0.5 + a; r.x[3] = 0.5 + b;
    r.w[0] = wb; r.w[1] = wa; r.w[2] = wa; r.w[3] = wb; break; }
  case 5: {
    double s = 2.0 * std::sqrt(10.0 / 7.0);
    double a = 0.5 * (1.0 / 3.0) * std::sqrt(5.0 - s);
    double b = 0.5 * (1.0 / 3.0) * std::sqrt(5.0 + s);
    double wa = (322.0 + 13.0 * std::sqrt(70.0)) / 1800.0;
    double wb = (322.0 - 13.0 * std::sqrt(70.0)) / 1800.0;
    r.x[0] = 0.5 - b; r.x[1] = 0.5 - a; r.x[2] = 0.5; r.x[3] = 0.5 + a; r.x[4] = 0.5 + b;
    r.w[0] = wb; r.w[1] = wa; r.w[2] = 128.0 / 450.0; r.w[3] = wa; r.w[4] = wb; break; }
  }
  return r;
}

struct QuadPoint { double x[3]; double w; };

inline std::vector<QuadPoint> make_tensor_rule(int dim, int order)
{
  Rule1D r = gauss1d(order);
  std::vector<QuadPoint> q;
  if (dim == 0) { QuadPoint p{{0, 0, 0}, 1.0}; q.push_back(p); return q; }
  int m = r.m;
  int n = 1; for (int d = 0; d < dim; ++d) n *= m;
  for (int idx = 0; idx < n; ++idx) {
    int rem = idx; QuadPoint p{{0, 0, 0}, 1.0};
    for (int d = 0; d < dim; ++d) {
      int i = rem % m; rem /= m;
      p.x[d] = r.x[i];
      p.w = (d == 0) ? r.w[i] : p.w * r.w[i];
    }
    q.push_back(p);
  }
  return q;
}

// [UPSTREAM] Dune::QuadratureRules<DF,dim>::rule(gt, order) hands out cached rules
inline const std::vector<QuadPoint>& tensor_rule(int dim, int order)
{
  static thread_local std::vector<QuadPoint> cache[4][12];
  if (order < 0 || order > 11 || dim < 0 || dim > 3) throw std::runtime_error("quadrature order/dim out of range");
  std::vector<QuadPoint>& r = cache[dim][order];
  if (r.empty()) r = make_tensor_rule(dim, order);
  return r;
}

// ---------------------------------------------------------------------------------------------
// Qk local basis, [UPSTREAM] dune-localfunctions QkLocalBasis<D,R,k,d>: Lagrange on the equidistant
// (k+1)^d lattice, shape function i <-> multi-index alpha, alpha_j = (i / (k+1)^j) % (k+1).
// ---------------------------------------------------------------------------------------------
struct QkBasis {
  int k, dim, n;
  QkBasis(int k_, int dim_) : k(k_), dim(dim_) { n = 1; for (int d = 0; d < dim; ++d) n *= (k + 1); }
  void alpha(int i, int a[3]) const { for (int d = 0; d < 3; ++d) { a[d] = (d < dim) ? i % (k + 1) : 0; i /= (k + 1); } }
  double p(int i, double x) const {
    double result = 1.0;
    for (int j = 0; j <= k; ++j) if (j != i) result *= (k * x - j) / (i - j);
    return result;
  }
  double dp(int i, double x) const {
    double result = 0.0;
    for (int j = 0; j <= k; ++j) if (j != i) {
      double prod = (k * 1.0) / (i - j);
      for (int l = 0; l <= k; ++l) if (l != i && l != j) prod *= (k * x - l) / (i - l);
      result += prod;
    }
    return result;
  }
  void evaluateFunction(const double* x, double* out) const {
    for (int i = 0; i < n; ++i) {
      int a[3]; alpha(i, a);
      out[i] = 1.0;
      for (int j = 0; j < dim; ++j) out[i] *= p(a[j], x[j]);
    }
  }
  // out[i][j] = d phi_i / d x_j
  void evaluateJacobian(const double* x, double (*out)[3]) const {
    for (int i = 0; i < n; ++i) {
      int a[3]; alpha(i, a);
      for (int j = 0; j < dim; ++j) {
        out[i][j] = dp(a[j], x[j]);
        for (int l = 0; l < dim; ++l) if (l != j) out[i][j] *= p(a[l], x[l]);
      }
    }
  }
};

// ---------------------------------------------------------------------------------------------
// Structured grid ([UPSTREAM] (i) YaspGrid: cells/vertices lexicographic, axis 0 fastest; faces -x,+x,-y,+y,-z,+z)
// ---------------------------------------------------------------------------------------------
struct Grid {
  int dim = 0;
  int64_t n[3] = {1, 1, 1};
  double lower[3] = {0, 0, 0}, upper[3] = {1, 1, 1}, h[3] = {1, 1, 1};
  // slab of a larger mesh (multi-rank tests; include/mdb200.h mdb_mesh_partition): global cell counts, offset of the local
  // cells, global lower corner.  Unpartitioned: gn = n, off = 0, glower = lower.
  int64_t gn[3] = {1, 1, 1}, off[3] = {0, 0, 0};
  double glower[3] = {0, 0, 0};
  bool partitioned = false;
  int64_t ncells() const { return n[0] * n[1] * n[2]; }
  double coord(int d, int64_t i) const { return partitioned ? glower[d] + double(i + off[d]) * h[d] : lower[d] + i * h[d]; }
  // a face without local neighbour that is not on the global boundary: processor boundary (globalassembler.hh:275-283)
  bool processor_face(const int64_t ijk[3], int f) const {
    if (!partitioned) return false;
    int a = f / 2; int64_t j = ijk[a] + (f % 2 ? 1 : -1);
    return (j < 0 || j >= n[a]) && (j + off[a] >= 0 && j + off[a] < gn[a]);
  }
  void cell_ijk(int64_t c, int64_t ijk[3]) const { ijk[0] = c % n[0]; c /= n[0]; ijk[1] = c % n[1]; ijk[2] = c / n[1]; }
  int64_t cell_index(const int64_t ijk[3]) const { return ijk[0] + n[0] * (ijk[1] + n[1] * ijk[2]); }
};

// element geometry wrapper ([UPSTREAM] AxisAlignedCubeGeometry)
struct EG {
  const Grid* g; int64_t ijk[3]; int64_t index; uint8_t sd;
  double lo(int d) const { return g->coord(d, ijk[d]); }
  double up(int d) const { return g->coord(d, ijk[d] + 1); }
  double jacInvT(int d) const { return 1.0 / (up(d) - lo(d)); }
  double integrationElement() const { double v = up(0) - lo(0); for (int d = 1; d < g->dim; ++d) v *= (up(d) - lo(d)); return v; }
  void global(const double* local, double* x) const { for (int d = 0; d < g->dim; ++d) x[d] = lo(d) + local[d] * (up(d) - lo(d)); }
};

// intersection geometry wrapper
struct IG {
  EG inside, outside; int axis, side; bool boundary; bool oneSidedDirection; int face;
  int dim() const { return inside.g->dim; }
  // geometryInInside().global(pos)
  void localInside(const double* pos, double* local) const {
    int t = 0; for (int d = 0; d < dim(); ++d) local[d] = (d == axis) ? double(side) : pos[t++];
  }
  void localOutside(const double* pos, double* local) const {
    int t = 0; for (int d = 0; d < dim(); ++d) local[d] = (d == axis) ? double(1 - side) : pos[t++];
  }
  double integrationElement() const {
    double v = 1.0; bool first = true;
    for (int d = 0; d < dim(); ++d) if (d != axis) { double e = inside.up(d) - inside.lo(d); v = first ? e : v * e; first = false; }
    return v;
  }
  void unitOuterNormal(double* nrm) const { for (int d = 0; d < dim(); ++d) nrm[d] = 0.0; nrm[axis] = side ? 1.0 : -1.0; }
  void global(const double* pos, double* x) const {
    int t = 0;
    for (int d = 0; d < dim(); ++d) {
      if (d == axis) x[d] = side ? inside.up(d) : inside.lo(d);
      else { x[d] = inside.lo(d) + pos[t] * (inside.up(d) - inside.lo(d)); ++t; }
    }
  }
  // |corner(0) - corner(1)| of the face geometry
  double cornerDistance01() const {
    if (dim() == 1) return 0.0;
    int t0 = (axis == 0) ? 1 : 0;
    double d = inside.lo(t0) - inside.up(t0);
    return std::sqrt(d * d);
  }
};

// ---------------------------------------------------------------------------------------------
// analytic data (closed set, see include/mdb200.h)
// ---------------------------------------------------------------------------------------------
double eval_function(const mdb_function& f, int dim, const double* x, double t);
// returns 0 = dirichlet, 1 = neumann, 2 = none  (test/testpoisson.cc:47 enum order)
int eval_bctype(const mdb_bctype& b, int dim, const double* xg, bool is_boundary);

// ---------------------------------------------------------------------------------------------
// local function space view handed to a local operator (subproblemlocalfunctionspace.hh:462-501 single-child)
// ---------------------------------------------------------------------------------------------
struct LFS {
  int n = 0;          // size()
  int idx[MAXLOC];    // localIndex(i): position in the cell-local coefficient vector
  int k = 1;          // Qk order
  bool dg = false;    // QkDG space: every coefficient belongs to the cell (local key codim 0)
};

struct LocalVectorView {   // [UPSTREAM] LocalVector::WeightedAccumulationView
  double* data; double weight;
  void accumulate(const LFS& l, int i, double v) { data[l.idx[i]] += weight * v; }
};
struct LocalMatrixView {   // [UPSTREAM] LocalMatrix::WeightedAccumulationView, row-major rows x cols
  double* data; int cols; double weight;
  void accumulate(const LFS& lv, int i, const LFS& lu, int j, double v) { data[lv.idx[i] * cols + lu.idx[j]] += weight * v; }
};

struct LocalOperator {
  bool doPatternVolume = false, doAlphaVolume = false, doLambdaVolume = false;
  bool doAlphaBoundary = false, doLambdaBoundary = false;
  bool doAlphaSkeleton = false, doLambdaSkeleton = false, doPatternSkeleton = false, doSkeletonTwoSided = false;
  double time = 0.0;
  virtual ~LocalOperator() {}
  virtual void alpha_volume(const EG&, const LFS&, const double*, const LFS&, LocalVectorView&) const {}
  virtual void jacobian_volume(const EG&, const LFS&, const double*, const LFS&, LocalMatrixView&) const {}
  virtual void lambda_volume(const EG&, const LFS&, LocalVectorView&) const {}
  virtual void lambda_boundary(const IG&, const LFS&, LocalVectorView&) const {}
  virtual void alpha_skeleton(const IG&, const LFS& /*lfsu_s*/, const double* /*x_s*/, const LFS& /*lfsv_s*/, const LFS& /*lfsu_n*/,
                              const double* /*x_n*/, const LFS& /*lfsv_n*/, LocalVectorView& /*r_s*/, LocalVectorView& /*r_n*/) const {}
  virtual void jacobian_skeleton(const IG&, const LFS&, const double*, const LFS&, const LFS&, const double*, const LFS&,
                                 LocalMatrixView& /*ss*/, LocalMatrixView& /*sn*/, LocalMatrixView& /*ns*/, LocalMatrixView& /*nn*/) const {}
  virtual void alpha_boundary(const IG&, const LFS&, const double*, const LFS&, LocalVectorView&) const {}
  virtual void jacobian_boundary(const IG&, const LFS&, const double*, const LFS&, LocalMatrixView&) const {}
};

struct CouplingOperator {
  bool doAlphaCoupling = false, doPatternCoupling = false;
  double epsilon = 1e-8;       // NumericalJacobianCoupling default, couplingutilities.hh:63-65
  int jac_mode = MDB_JAC_FD_BITMATCH;
  bool own_jacobian = false;   // the operator has its own jacobian_coupling (test/neumanndirichletcoupling.hh:291-432) writing mat_ss only
  virtual void jacobian_own(const IG&, const LFS& /*lfsu_s*/, const LFS& /*lfsu_r*/, LocalMatrixView& /*mat_ss*/) const {}
  virtual ~CouplingOperator() {}
  virtual void alpha_coupling(const IG&, const LFS& lfsu_s, const double* x_s, const LFS& lfsv_s,
                              const LFS& lfsu_n, const double* x_n, const LFS& lfsv_n,
                              LocalVectorView& r_s, LocalVectorView& r_n) const = 0;
  // couplingutilities.hh:75-135
  void jacobian_coupling(const IG&, const LFS& lfsu_s, const double* x_s, int nx_s, const LFS& lfsv_s,
                         const LFS& lfsu_n, const double* x_n, int nx_n, const LFS& lfsv_n,
                         LocalMatrixView& mat_ss, LocalMatrixView& mat_sn,
                         LocalMatrixView& mat_ns, LocalMatrixView& mat_nn) const;
};

// enriched (mortar) coupling operator: test/testmortarpoisson.cc:117-250 + couplingutilities.hh:288-484
struct EnrichedOperator {
  double gamma_0 = 1.0;
  double epsilon = 1e-8;       // NumericalJacobianEnrichedCouplingTo{First,Second}SubProblem default
  int jac_mode = MDB_JAC_FD_BITMATCH;
  int ncomp = 1;               // component blocks of a PowerCouplingGridFunctionSpace (test/testpowermortarpoisson.cc:186-189)
  // the operator on the whole (power) coupling space: loop over the component blocks, lfsu_c / x_c / r_c hold ncomp blocks of equal size
  void alpha_power(int sign, const IG& ig, const LFS& lfsu, const double* x, const LFS& lfsv, const LFS& lfsu_c, const double* x_c,
                   const LFS& lfsv_c, LocalVectorView& r, LocalVectorView& r_c) const;
  // alpha_generic<sign>: sign = +1 first (inside element), -1 second (outside element)
  void alpha_generic(int sign, const IG& ig, const LFS& lfsu, const double* x, const LFS& lfsv, const LFS& lfsu_c, const double* x_c,
                     const LFS& lfsv_c, LocalVectorView& r, LocalVectorView& r_c) const;
  // jacobian_enriched_coupling_first / _second (identical algorithm, couplingutilities.hh:305-388 / :404-484)
  void jacobian_generic(int sign, const IG& ig, const LFS& lfsu, const double* x, int nx, const LFS& lfsv, const LFS& lfsu_c,
                        const double* x_c, int nxc, const LFS& lfsv_c, LocalMatrixView& mat_ss, LocalMatrixView& mat_sc,
                        LocalMatrixView& mat_cs, LocalMatrixView& mat_cc) const;
};

std::unique_ptr<LocalOperator> make_local_operator(int op_id, const void* params, size_t nbytes);
std::unique_ptr<CouplingOperator> make_coupling_operator(int op_id, const void* params, size_t nbytes, int jac_mode);

// ---------------------------------------------------------------------------------------------
// function space: children + DOF map
// ---------------------------------------------------------------------------------------------
struct ChildSpace {
  int fe_kind; int k; int subdomain;           // subdomain -1: MultiDomainGrid view
  // coupling (mortar) space on the interface faces inside in L / outside in R (couplinggridfunctionspace.hh:14-44)
  bool iface = false; int lkind = 0, rkind = 0; uint32_t lmask = 0, rmask = 0;
  bool dg = false;                              // QkDG: DOF = cellblk[cell] * nloc + i
  std::vector<int64_t> cellblk;                 // dg: index of the cell in the child's grid view (-1 outside)
  int64_t lat_n[3];                             // lattice points per axis = k*n+1
  std::vector<int64_t> lat2dof;                 // lattice point -> index within the child block, -1 absent
  std::vector<int64_t> dof2lat;
  int64_t size = 0, offset = 0;
  int nloc(int dim) const { int r = 1; for (int d = 0; d < dim; ++d) r *= (k + 1); return r; }
};

struct SubProblem {
  std::unique_ptr<LocalOperator> lop;
  int cond_kind; uint32_t mask;
  std::vector<int> children;
  bool appliesTo(uint8_t sd) const {
    if (cond_kind == MDB_COND_EQUALITY) return uint32_t(sd) == mask;      // subdomainset.hh:129-131
    return (uint32_t(sd) & mask) == mask;                                  // subdomainset.hh:102-104
  }
};

struct Coupling {
  std::unique_ptr<CouplingOperator> op;
  int local_sp, remote_sp;
  std::unique_ptr<EnrichedOperator> eop;       // EnrichedCoupling (coupling.hh:123-192): op is null
  int coupling_child = -1, ncomp = 1;          // ncomp consecutive child blocks from coupling_child on (power coupling space)
};

struct GridOperatorDesc { std::vector<int> sps, cps; };

struct CSR {
  int64_t nrows = 0;
  std::vector<int64_t> rowptr, colidx;
  std::vector<double> val;
  int64_t find(int64_t r, int64_t c) const {
    auto b = colidx.begin() + rowptr[r], e = colidx.begin() + rowptr[r + 1];
    auto it = std::lower_bound(b, e, c);
    if (it == e || *it != c) return -1;
    return it - colidx.begin();
  }
};

struct SolveStats { int iterations = 0; bool converged = false; double defect0 = 0, defect = 0; };

struct Context {
  Grid grid;
  std::vector<uint8_t> cell_sd;
  std::vector<ChildSpace> children;
  std::vector<SubProblem> sps;
  std::vector<Coupling> cps;
  std::vector<GridOperatorDesc> gos;
  std::vector<CSR> mats;
  std::vector<uint8_t> constrained;
  int64_t ndof = 0;
  bool finalized = false;
  double time = 0.0;
  std::string err;
  int nthreads = 1;

  void finalize();
  // MultiDomainLocalFunctionSpace::bind(e): children offsets in the cell-local vector + container indices
  int bind(int64_t cell, int* child_off /*nchildren+1*/, int64_t* dofs) const;
  LFS subproblem_lfs(const SubProblem& sp, const int* child_off) const;
  // CouplingLocalFunctionSpace::bind(intersection): the mortar DOFs of face f of `cell` in face-local order; returns their number
  int bind_coupling(int child, int64_t cell, int f, int64_t* dofs) const;
  int bind_coupling_power(const Coupling& cp, int64_t cell, int f, int64_t* dofs) const;     // all component blocks, block after block
  bool iface_applies(const ChildSpace& ch, uint8_t sd_in, uint8_t sd_out) const;

  void constraints_assemble(const int* sps_, const mdb_bctype* bcs, int n);
  void interpolate(const int* sps_, const mdb_function* fns, int n, double t, double* x) const;

  int matrix_create(const int* gos_, int ngos);
  void residual(int go, double weight, const double* x, double* r) const;
  void jacobian(int go, int mat, double weight, const double* x);
  // threaded variant used only as the CPU baseline (z-slab two-colouring); same result up to summation order
  void jacobian_parallel(int go, int mat, double weight, const double* x, int nthreads);
};

// linear algebra ([UPSTREAM] dune-istl restated)
void spmv(const CSR& A, const double* x, double* y);
SolveStats solve(const CSR& A, int method, int precond, int precond_sweeps, double reduction, int maxit, const double* b, double* x);
enum { ORC_PRECOND_SSOR = 100 };

} // namespace orc
#endif
