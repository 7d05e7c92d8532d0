// oracle/oracle.cpp — TEST INFRASTRUCTURE ONLY.  See oracle.hh for the scope statement ("parity unpinned").
// CPU restatement of dune-multidomain's single-traversal assembly + Krylov solve.
#include "oracle.hh"
#include <set>
#include <thread>
#include <functional>

namespace orc {

// =============================================================================================
// problem data (test/testpoisson.cc:32-116)
// =============================================================================================
double eval_function(const mdb_function& f, int dim, const double* x, double t)
{
  switch (f.kind) {
  case MDB_FN_ZERO: return 0.0;
  case MDB_FN_CONST: return f.p[0];
  case MDB_FN_GAUSS: {                     // testpoisson.cc:92-99: center -= x; y = exp(-center.two_norm2())
    double n2 = 0.0;
    for (int d = 0; d < dim; ++d) { double c = f.p[1 + d] - x[d]; n2 += c * c; }
    return std::exp(-(f.p[0] * n2));
  }
  case MDB_FN_BOX: {                       // testpoisson.cc:34-37
    for (int d = 0; d < dim; ++d) if (!(x[d] > f.p[1] && x[d] < f.p[2])) return 0.0;
    return f.p[0];
  }
  case MDB_FN_TESTPOISSON_J: {             // testpoisson.cc:103-116
    if (x[1] < 1E-6 || x[1] > 1.0 - 1E-6) return 0.0;
    if (x[0] > 1.0 - 1E-6 && x[1] > 0.5 + 1E-6) return -5.0;
    return 0.0;
  }
  case MDB_FN_INSTAT_G: {
    double n2 = 0.0;
    for (int d = 0; d < dim; ++d) { double c = 0.5 - x[d]; n2 += c * c; }
    return f.p[0] * std::sin(f.p[1] * t) * std::exp(-n2);
  }
  case MDB_FN_POLY2: {
    double v = f.p[0], n2 = 0.0;
    for (int d = 0; d < dim; ++d) { v += f.p[1 + d] * x[d]; n2 += x[d] * x[d]; }
    return v + f.p[4] * n2;
  }
  case MDB_FN_DN_SOURCE: {               // testpoisson-dirichlet-neumann.cc:73-87
    bool box = true;
    for (int d = 0; d < dim; ++d) box = box && x[d] > 0.25 && x[d] < 0.375;
    if (box) return 50.0;
    double m2 = 0.0, n2 = 0.0;
    for (int d = 0; d < dim; ++d) { double c = 0.5 - x[d]; m2 += c * c; n2 += x[d] * x[d]; }
    if (std::sqrt(m2) < 0.2) return 60 * std::exp(-10 * n2);
    return 0.0;
  }
  }
  throw std::runtime_error("unknown function kind");
}

int eval_bctype(const mdb_bctype& b, int dim, const double* xg, bool is_boundary)
{
  (void)dim;
  switch (b.kind) {
  case MDB_BC_ALL_DIRICHLET: return is_boundary ? 0 : 2;
  case MDB_BC_ALL_NEUMANN: return is_boundary ? 1 : 2;
  case MDB_BC_TESTPOISSON:                 // testpoisson.cc:49-71
    if (!is_boundary) return 2;
    if (xg[1] < 1E-6 || xg[1] > 1.0 - 1E-6) return 1;
    if (xg[0] > 1.0 - 1E-6 && xg[1] > 0.5 + 1E-6) return 1;
    return 0;
  }
  throw std::runtime_error("unknown bctype kind");
}

// =============================================================================================
// local operators
// =============================================================================================
namespace {

// [UPSTREAM] Dune::PDELab::Poisson<F,B,J> = Laplace stiffness + source + Neumann flux; the in-tree twin is
// test/instationarypoissonoperator.hh:30-168 (alpha_volume :52-77, lambda_volume :103-117, lambda_boundary :142-167)
struct PoissonOp : LocalOperator {
  mdb_poisson_params P;
  explicit PoissonOp(const mdb_poisson_params& p) : P(p) {
    doPatternVolume = true; doAlphaVolume = true; doLambdaVolume = true; doLambdaBoundary = true;
  }
  void alpha_volume(const EG& eg, const LFS& lfsu, const double* x, const LFS& lfsv, LocalVectorView& r) const override {
    const int dim = eg.g->dim;
    QkBasis B(lfsu.k, dim);
    const auto& rule = tensor_rule(dim, P.quad_order);
    double js[MAXLOC][3], gradphi[MAXLOC][3];
    for (const auto& q : rule) {
      B.evaluateJacobian(q.x, js);
      for (int i = 0; i < lfsu.n; ++i)
        for (int d = 0; d < dim; ++d) { gradphi[i][d] = 0.0; gradphi[i][d] += eg.jacInvT(d) * js[i][d]; }
      double gradu[3] = {0, 0, 0};
      for (int i = 0; i < lfsu.n; ++i)
        for (int d = 0; d < dim; ++d) gradu[d] += x[lfsu.idx[i]] * gradphi[i][d];
      double factor = q.w * eg.integrationElement();
      for (int i = 0; i < lfsu.n; ++i) {
        double dot = 0.0;
        for (int d = 0; d < dim; ++d) dot += gradu[d] * gradphi[i][d];
        r.accumulate(lfsv, i, dot * factor);
      }
    }
  }
  void jacobian_volume(const EG& eg, const LFS& lfsu, const double*, const LFS& lfsv, LocalMatrixView& mat) const override {
    const int dim = eg.g->dim;
    QkBasis B(lfsu.k, dim);
    const auto& rule = tensor_rule(dim, P.quad_order);
    double js[MAXLOC][3], gradphi[MAXLOC][3];
    for (const auto& q : rule) {
      B.evaluateJacobian(q.x, js);
      for (int i = 0; i < lfsu.n; ++i)
        for (int d = 0; d < dim; ++d) { gradphi[i][d] = 0.0; gradphi[i][d] += eg.jacInvT(d) * js[i][d]; }
      double factor = q.w * eg.integrationElement();
      for (int i = 0; i < lfsu.n; ++i)
        for (int j = 0; j < lfsu.n; ++j) {
          double dot = 0.0;
          for (int d = 0; d < dim; ++d) dot += gradphi[j][d] * gradphi[i][d];
          mat.accumulate(lfsv, i, lfsu, j, dot * factor);
        }
    }
  }
  void lambda_volume(const EG& eg, const LFS& lfsv, LocalVectorView& r) const override {
    if (P.f.kind == MDB_FN_ZERO) return;   // adds exact zeros otherwise
    const int dim = eg.g->dim;
    QkBasis B(lfsv.k, dim);
    const auto& rule = tensor_rule(dim, P.quad_order);
    double phi[MAXLOC], xg[3];
    for (const auto& q : rule) {
      B.evaluateFunction(q.x, phi);
      eg.global(q.x, xg);
      double y = eval_function(P.f, dim, xg, time);
      double factor = q.w * eg.integrationElement();
      for (int i = 0; i < lfsv.n; ++i) r.accumulate(lfsv, i, -y * phi[i] * factor);
    }
  }
  void lambda_boundary(const IG& ig, const LFS& lfsv, LocalVectorView& r) const override {
    const int dim = ig.dim();
    QkBasis B(lfsv.k, dim);
    const auto& rule = tensor_rule(dim - 1, P.quad_order);
    double phi[MAXLOC], local[3], xg[3];
    for (const auto& q : rule) {
      ig.global(q.x, xg);
      if (eval_bctype(P.bc, dim, xg, ig.boundary) != 1) continue;     // !isNeumann
      ig.localInside(q.x, local);
      B.evaluateFunction(local, phi);
      double xe[3]; ig.inside.global(local, xe);
      double y = eval_function(P.j, dim, xe, time);
      double factor = q.w * ig.integrationElement();
      for (int i = 0; i < lfsv.n; ++i) r.accumulate(lfsv, i, y * phi[i] * factor);
    }
  }
};

// [UPSTREAM] Dune::PDELab::L2 / test/simpletimeoperator.hh:29-101
struct L2Op : LocalOperator {
  mdb_l2_params P;
  explicit L2Op(const mdb_l2_params& p) : P(p) { doPatternVolume = true; doAlphaVolume = true; }
  void alpha_volume(const EG& eg, const LFS& lfsu, const double* x, const LFS& lfsv, LocalVectorView& r) const override {
    const int dim = eg.g->dim;
    QkBasis B(lfsu.k, dim);
    const auto& rule = tensor_rule(dim, P.quad_order);
    double phi[MAXLOC];
    for (const auto& q : rule) {
      B.evaluateFunction(q.x, phi);
      double u = 0.0;
      for (int i = 0; i < lfsu.n; ++i) u += x[lfsu.idx[i]] * phi[i];
      double factor = q.w * eg.integrationElement();
      for (int i = 0; i < lfsu.n; ++i) r.accumulate(lfsv, i, u * phi[i] * factor);
    }
  }
  void jacobian_volume(const EG& eg, const LFS& lfsu, const double*, const LFS& lfsv, LocalMatrixView& mat) const override {
    const int dim = eg.g->dim;
    QkBasis B(lfsu.k, dim);
    const auto& rule = tensor_rule(dim, P.quad_order);
    double phi[MAXLOC];
    for (const auto& q : rule) {
      B.evaluateFunction(q.x, phi);
      double factor = q.w * eg.integrationElement();
      for (int i = 0; i < lfsu.n; ++i)
        for (int j = 0; j < lfsu.n; ++j)
          mat.accumulate(lfsv, i, lfsu, j, phi[j] * phi[i] * factor);
    }
  }
};

// [UPSTREAM] Dune::PDELab::ConvectionDiffusionDG<T,FEM> (dune-pdelab convectiondiffusiondg.hh, the 2.0/2.4-dev version with the
// "Houston" face diameter — the same text test/neumanndirichletcoupling.hh:85-432 was derived from and can be read against),
// instantiated by dglop() (test/testpoisson-dirichlet-neumann.cc:864-881) with the parameter class ConvectionDiffusionModelProblem
// (:169-285): A = I, b = 0, c = 0.  Restated from the published algorithm; the velocity terms are kept (with b = 0) so that the
// operation count per entry is the upstream one.
struct ConvDiffDGOp : LocalOperator {
  mdb_convdiffdg_params P;
  double theta;
  explicit ConvDiffDGOp(const mdb_convdiffdg_params& p) : P(p) {
    doPatternVolume = true; doPatternSkeleton = true;
    doAlphaVolume = true; doAlphaSkeleton = true; doAlphaBoundary = true; doLambdaVolume = true;
    doSkeletonTwoSided = false;
    theta = 1.0; if (P.method == MDB_DG_METHOD_SIPG) theta = -1.0;
  }
  static void tgrad(const QkBasis& B, const EG& e, const double* local, int n, double (*js)[3], double (*out)[3]) {
    B.evaluateJacobian(local, js);
    for (int i = 0; i < n; ++i)
      for (int d = 0; d < B.dim; ++d) { out[i][d] = 0.0; out[i][d] += e.jacInvT(d) * js[i][d]; }     // jac.mv(js[i][0], gradphi[i])
  }
  // ---- volume ------------------------------------------------------------------------------------------------------
  void alpha_volume(const EG& eg, const LFS& lfsu, const double* x, const LFS& lfsv, LocalVectorView& r) const override {
    const int dim = eg.g->dim;
    QkBasis B(lfsu.k, dim);
    const auto& rule = tensor_rule(dim, P.intorderadd + 2 * lfsu.k);
    double phi[MAXLOC], js[MAXLOC][3], gradphi[MAXLOC][3];
    for (const auto& q : rule) {
      B.evaluateFunction(q.x, phi);
      double u = 0.0;
      for (int i = 0; i < lfsu.n; ++i) u += x[lfsu.idx[i]] * phi[i];
      tgrad(B, eg, q.x, lfsu.n, js, gradphi);
      double gradu[3] = {0, 0, 0};
      for (int i = 0; i < lfsu.n; ++i) for (int d = 0; d < dim; ++d) gradu[d] += x[lfsu.idx[i]] * gradphi[i][d];
      double Agradu[3];
      for (int i = 0; i < dim; ++i) { Agradu[i] = 0.0; for (int j = 0; j < dim; ++j) Agradu[i] += ((i == j) ? 1.0 : 0.0) * gradu[j]; }
      const double c = 0.0;
      const double factor = q.w * eg.integrationElement();
      for (int i = 0; i < lfsv.n; ++i) {
        double Agg = 0.0, bg = 0.0;
        for (int d = 0; d < dim; ++d) { Agg += Agradu[d] * gradphi[i][d]; bg += 0.0 * gradphi[i][d]; }
        r.accumulate(lfsv, i, (Agg - u * bg + c * u * phi[i]) * factor);
      }
    }
  }
  void jacobian_volume(const EG& eg, const LFS& lfsu, const double*, const LFS& lfsv, LocalMatrixView& mat) const override {
    const int dim = eg.g->dim;
    QkBasis B(lfsu.k, dim);
    const auto& rule = tensor_rule(dim, P.intorderadd + 2 * lfsu.k);
    double phi[MAXLOC], js[MAXLOC][3], gradphi[MAXLOC][3];
    for (const auto& q : rule) {
      B.evaluateFunction(q.x, phi);
      tgrad(B, eg, q.x, lfsu.n, js, gradphi);
      const double c = 0.0;
      const double factor = q.w * eg.integrationElement();
      for (int j = 0; j < lfsu.n; ++j)
        for (int i = 0; i < lfsu.n; ++i) {
          double Agg = 0.0, bg = 0.0;
          for (int d = 0; d < dim; ++d) { Agg += gradphi[j][d] * gradphi[i][d]; bg += 0.0 * gradphi[i][d]; }     // Agradphi[j] = gradphi[j]
          mat.accumulate(lfsv, i, lfsu, j, (Agg - phi[j] * bg + c * phi[j] * phi[i]) * factor);
        }
    }
  }
  void lambda_volume(const EG& eg, const LFS& lfsv, LocalVectorView& r) const override {
    const int dim = eg.g->dim;
    QkBasis B(lfsv.k, dim);
    const auto& rule = tensor_rule(dim, P.intorderadd + 2 * lfsv.k);
    double phi[MAXLOC], xg[3];
    for (const auto& q : rule) {
      B.evaluateFunction(q.x, phi);
      eg.global(q.x, xg);
      const double f = eval_function(P.f, dim, xg, time);
      const double factor = q.w * eg.integrationElement();
      for (int i = 0; i < lfsv.n; ++i) r.accumulate(lfsv, i, -f * phi[i] * factor);
    }
  }
  // ---- interior faces ----------------------------------------------------------------------------------------------
  struct Face { double n_F[3], omega_s, omega_n, penalty_factor; int intorder; };
  Face face(const IG& ig, int k_s, int k_n, bool boundary) const {
    Face F; const int dim = ig.dim();
    const int order = std::max(k_s, k_n);
    F.intorder = P.intorderadd + 2 * order;
    const double h_F = (boundary ? ig.inside.integrationElement()
                                 : std::min(ig.inside.integrationElement(), ig.outside.integrationElement())) / ig.integrationElement();
    ig.unitOuterNormal(F.n_F);                                           // An_F_s = An_F_n = n_F (A = I)
    double harmonic_average;
    if (P.weights == MDB_DG_WEIGHTS_ON) {
      double delta_s = 0.0, delta_n = 0.0;
      for (int d = 0; d < dim; ++d) { delta_s += F.n_F[d] * F.n_F[d]; delta_n += F.n_F[d] * F.n_F[d]; }
      if (boundary) { F.omega_s = 1.0; F.omega_n = 0.0; harmonic_average = delta_s; }
      else {
        F.omega_s = delta_n / (delta_s + delta_n + 1e-20); F.omega_n = delta_s / (delta_s + delta_n + 1e-20);
        harmonic_average = 2.0 * delta_s * delta_n / (delta_s + delta_n + 1e-20);
      }
    } else { F.omega_s = F.omega_n = 0.5; harmonic_average = 1.0; if (boundary) { F.omega_s = 1.0; F.omega_n = 0.0; } }
    const int degree = order;
    F.penalty_factor = (P.alpha / h_F) * harmonic_average * degree * (degree + dim - 1);
    return F;
  }
  void alpha_skeleton(const IG& ig, const LFS& lfsu_s, const double* x_s, const LFS& lfsv_s, const LFS& lfsu_n, const double* x_n,
                      const LFS& lfsv_n, LocalVectorView& r_s, LocalVectorView& r_n) const override {
    const int dim = ig.dim();
    const Face F = face(ig, lfsu_s.k, lfsu_n.k, false);
    QkBasis Bs(lfsu_s.k, dim), Bn(lfsu_n.k, dim);
    const auto& rule = tensor_rule(dim - 1, F.intorder);
    double phi_s[MAXLOC], phi_n[MAXLOC], js[MAXLOC][3], tg_s[MAXLOC][3], tg_n[MAXLOC][3], ip_s[3], ip_n[3];
    for (const auto& q : rule) {
      ig.localInside(q.x, ip_s); ig.localOutside(q.x, ip_n);
      Bs.evaluateFunction(ip_s, phi_s); Bn.evaluateFunction(ip_n, phi_n);
      double u_s = 0.0, u_n = 0.0;
      for (int i = 0; i < lfsu_s.n; ++i) u_s += x_s[lfsu_s.idx[i]] * phi_s[i];
      for (int i = 0; i < lfsu_n.n; ++i) u_n += x_n[lfsu_n.idx[i]] * phi_n[i];
      tgrad(Bs, ig.inside, ip_s, lfsu_s.n, js, tg_s); tgrad(Bn, ig.outside, ip_n, lfsu_n.n, js, tg_n);
      double gradu_s[3] = {0, 0, 0}, gradu_n[3] = {0, 0, 0};
      for (int i = 0; i < lfsu_s.n; ++i) for (int d = 0; d < dim; ++d) gradu_s[d] += x_s[lfsu_s.idx[i]] * tg_s[i][d];
      for (int i = 0; i < lfsu_n.n; ++i) for (int d = 0; d < dim; ++d) gradu_n[d] += x_n[lfsu_n.idx[i]] * tg_n[i][d];
      double normalflux = 0.0; for (int d = 0; d < dim; ++d) normalflux += 0.0 * F.n_F[d];             // b = 0
      const double omegaup_s = normalflux >= 0.0 ? 1.0 : 0.0, omegaup_n = normalflux >= 0.0 ? 0.0 : 1.0;
      const double factor = q.w * ig.integrationElement();
      const double term1 = (omegaup_s * u_s + omegaup_n * u_n) * normalflux * factor;                   // convection
      for (int i = 0; i < lfsv_s.n; ++i) r_s.accumulate(lfsv_s, i, term1 * phi_s[i]);
      for (int i = 0; i < lfsv_n.n; ++i) r_n.accumulate(lfsv_n, i, -term1 * phi_n[i]);
      double Ags = 0.0, Agn = 0.0; for (int d = 0; d < dim; ++d) { Ags += F.n_F[d] * gradu_s[d]; Agn += F.n_F[d] * gradu_n[d]; }
      const double term2 = -(F.omega_s * Ags + F.omega_n * Agn) * factor;                               // diffusion
      for (int i = 0; i < lfsv_s.n; ++i) r_s.accumulate(lfsv_s, i, term2 * phi_s[i]);
      for (int i = 0; i < lfsv_n.n; ++i) r_n.accumulate(lfsv_n, i, -term2 * phi_n[i]);
      const double term3 = (u_s - u_n) * factor;                                                        // (non-)symmetric IP term
      for (int i = 0; i < lfsv_s.n; ++i) { double a = 0.0; for (int d = 0; d < dim; ++d) a += F.n_F[d] * tg_s[i][d]; r_s.accumulate(lfsv_s, i, term3 * theta * F.omega_s * a); }
      for (int i = 0; i < lfsv_n.n; ++i) { double a = 0.0; for (int d = 0; d < dim; ++d) a += F.n_F[d] * tg_n[i][d]; r_n.accumulate(lfsv_n, i, term3 * theta * F.omega_n * a); }
      const double term4 = F.penalty_factor * (u_s - u_n) * factor;                                     // standard IP term
      for (int i = 0; i < lfsv_s.n; ++i) r_s.accumulate(lfsv_s, i, term4 * phi_s[i]);
      for (int i = 0; i < lfsv_n.n; ++i) r_n.accumulate(lfsv_n, i, -term4 * phi_n[i]);
    }
  }
  void jacobian_skeleton(const IG& ig, const LFS& lfsu_s, const double*, const LFS& lfsv_s, const LFS& lfsu_n, const double*, const LFS& lfsv_n,
                         LocalMatrixView& mat_ss, LocalMatrixView& mat_sn, LocalMatrixView& mat_ns, LocalMatrixView& mat_nn) const override {
    const int dim = ig.dim();
    const Face F = face(ig, lfsu_s.k, lfsu_n.k, false);
    QkBasis Bs(lfsu_s.k, dim), Bn(lfsu_n.k, dim);
    const auto& rule = tensor_rule(dim - 1, F.intorder);
    double phi_s[MAXLOC], phi_n[MAXLOC], js[MAXLOC][3], tg_s[MAXLOC][3], tg_n[MAXLOC][3], ip_s[3], ip_n[3], an_s[MAXLOC], an_n[MAXLOC];
    for (const auto& q : rule) {
      ig.localInside(q.x, ip_s); ig.localOutside(q.x, ip_n);
      Bs.evaluateFunction(ip_s, phi_s); Bn.evaluateFunction(ip_n, phi_n);
      tgrad(Bs, ig.inside, ip_s, lfsu_s.n, js, tg_s); tgrad(Bn, ig.outside, ip_n, lfsu_n.n, js, tg_n);
      for (int i = 0; i < lfsu_s.n; ++i) { an_s[i] = 0.0; for (int d = 0; d < dim; ++d) an_s[i] += F.n_F[d] * tg_s[i][d]; }       // An_F_s * tgradphi_s[i]
      for (int i = 0; i < lfsu_n.n; ++i) { an_n[i] = 0.0; for (int d = 0; d < dim; ++d) an_n[i] += F.n_F[d] * tg_n[i][d]; }
      double normalflux = 0.0; for (int d = 0; d < dim; ++d) normalflux += 0.0 * F.n_F[d];
      const double omegaup_s = normalflux >= 0.0 ? 1.0 : 0.0, omegaup_n = normalflux >= 0.0 ? 0.0 : 1.0;
      const double factor = q.w * ig.integrationElement();
      const double ipfactor = F.penalty_factor * factor;
      // all terms in the order: I convection, II diffusion, III consistency, IV ip
      for (int j = 0; j < lfsu_s.n; ++j) {
        const double temp1 = -an_s[j] * F.omega_s * factor;
        for (int i = 0; i < lfsv_s.n; ++i) {
          mat_ss.accumulate(lfsv_s, i, lfsu_s, j, omegaup_s * phi_s[j] * normalflux * factor * phi_s[i]);
          mat_ss.accumulate(lfsv_s, i, lfsu_s, j, temp1 * phi_s[i]);
          mat_ss.accumulate(lfsv_s, i, lfsu_s, j, phi_s[j] * factor * theta * F.omega_s * an_s[i]);
          mat_ss.accumulate(lfsv_s, i, lfsu_s, j, phi_s[j] * ipfactor * phi_s[i]);
        }
      }
      for (int j = 0; j < lfsu_n.n; ++j) {
        const double temp1 = -an_n[j] * F.omega_n * factor;
        for (int i = 0; i < lfsv_s.n; ++i) {
          mat_sn.accumulate(lfsv_s, i, lfsu_n, j, omegaup_n * phi_n[j] * normalflux * factor * phi_s[i]);
          mat_sn.accumulate(lfsv_s, i, lfsu_n, j, temp1 * phi_s[i]);
          mat_sn.accumulate(lfsv_s, i, lfsu_n, j, -phi_n[j] * factor * theta * F.omega_s * an_s[i]);
          mat_sn.accumulate(lfsv_s, i, lfsu_n, j, -phi_n[j] * ipfactor * phi_s[i]);
        }
      }
      for (int j = 0; j < lfsu_s.n; ++j) {
        const double temp1 = -an_s[j] * F.omega_s * factor;
        for (int i = 0; i < lfsv_n.n; ++i) {
          mat_ns.accumulate(lfsv_n, i, lfsu_s, j, -omegaup_s * phi_s[j] * normalflux * factor * phi_n[i]);
          mat_ns.accumulate(lfsv_n, i, lfsu_s, j, -temp1 * phi_n[i]);
          mat_ns.accumulate(lfsv_n, i, lfsu_s, j, phi_s[j] * factor * theta * F.omega_n * an_n[i]);
          mat_ns.accumulate(lfsv_n, i, lfsu_s, j, -phi_s[j] * ipfactor * phi_n[i]);
        }
      }
      for (int j = 0; j < lfsu_n.n; ++j) {
        const double temp1 = -an_n[j] * F.omega_n * factor;
        for (int i = 0; i < lfsv_n.n; ++i) {
          mat_nn.accumulate(lfsv_n, i, lfsu_n, j, -omegaup_n * phi_n[j] * normalflux * factor * phi_n[i]);
          mat_nn.accumulate(lfsv_n, i, lfsu_n, j, -temp1 * phi_n[i]);
          mat_nn.accumulate(lfsv_n, i, lfsu_n, j, -phi_n[j] * factor * theta * F.omega_n * an_n[i]);
          mat_nn.accumulate(lfsv_n, i, lfsu_n, j, phi_n[j] * ipfactor * phi_n[i]);
        }
      }
    }
  }
  // ---- boundary faces (and subdomain interfaces, where the parameter class answers None) -----------------------------------
  int bctype(const IG& ig) const {          // evaluated at the face centre
    double centre[2] = {0.5, 0.5}, xg[3];
    ig.global(centre, xg);
    return eval_bctype(P.bc, ig.dim(), xg, ig.boundary);
  }
  void alpha_boundary(const IG& ig, const LFS& lfsu_s, const double* x_s, const LFS& lfsv_s, LocalVectorView& r_s) const override {
    const int bc = bctype(ig);
    if (bc == 2) return;                                                // ConvectionDiffusionBoundaryConditions::None
    const int dim = ig.dim();
    const Face F = face(ig, lfsu_s.k, lfsu_s.k, true);
    QkBasis Bs(lfsu_s.k, dim);
    const auto& rule = tensor_rule(dim - 1, F.intorder);
    double phi_s[MAXLOC], js[MAXLOC][3], tg_s[MAXLOC][3], ip_s[3], xg[3];
    for (const auto& q : rule) {
      ig.localInside(q.x, ip_s);
      Bs.evaluateFunction(ip_s, phi_s);
      const double factor = q.w * ig.integrationElement();
      if (bc == 1) {                                                    // Neumann: j at the face point
        ig.global(q.x, xg);
        const double j = eval_function(P.j, dim, xg, time);
        for (int i = 0; i < lfsv_s.n; ++i) r_s.accumulate(lfsv_s, i, j * phi_s[i] * factor);
        continue;
      }
      double u_s = 0.0;
      for (int i = 0; i < lfsu_s.n; ++i) u_s += x_s[lfsu_s.idx[i]] * phi_s[i];
      double normalflux = 0.0; for (int d = 0; d < dim; ++d) normalflux += 0.0 * F.n_F[d];
      tgrad(Bs, ig.inside, ip_s, lfsu_s.n, js, tg_s);
      double gradu_s[3] = {0, 0, 0};
      for (int i = 0; i < lfsu_s.n; ++i) for (int d = 0; d < dim; ++d) gradu_s[d] += x_s[lfsu_s.idx[i]] * tg_s[i][d];
      ig.inside.global(ip_s, xg);
      const double g = eval_function(P.g, dim, xg, time);               // param.g(inside, iplocal_s)
      const double omegaup_s = normalflux >= 0.0 ? 1.0 : 0.0, omegaup_n = normalflux >= 0.0 ? 0.0 : 1.0;
      const double term1 = (omegaup_s * u_s + omegaup_n * g) * normalflux * factor;
      for (int i = 0; i < lfsv_s.n; ++i) r_s.accumulate(lfsv_s, i, term1 * phi_s[i]);
      double Ag = 0.0; for (int d = 0; d < dim; ++d) Ag += F.n_F[d] * gradu_s[d];
      const double term2 = Ag * factor;
      for (int i = 0; i < lfsv_s.n; ++i) r_s.accumulate(lfsv_s, i, -term2 * phi_s[i]);
      const double term3 = (u_s - g) * factor;
      for (int i = 0; i < lfsv_s.n; ++i) { double a = 0.0; for (int d = 0; d < dim; ++d) a += F.n_F[d] * tg_s[i][d]; r_s.accumulate(lfsv_s, i, term3 * theta * a); }
      const double term4 = F.penalty_factor * (u_s - g) * factor;
      for (int i = 0; i < lfsv_s.n; ++i) r_s.accumulate(lfsv_s, i, term4 * phi_s[i]);
    }
  }
  void jacobian_boundary(const IG& ig, const LFS& lfsu_s, const double*, const LFS& lfsv_s, LocalMatrixView& mat_ss) const override {
    const int bc = bctype(ig);
    if (bc == 2 || bc == 1) return;                                     // None / Neumann: no dependence on u
    const int dim = ig.dim();
    const Face F = face(ig, lfsu_s.k, lfsu_s.k, true);
    QkBasis Bs(lfsu_s.k, dim);
    const auto& rule = tensor_rule(dim - 1, F.intorder);
    double phi_s[MAXLOC], js[MAXLOC][3], tg_s[MAXLOC][3], ip_s[3], an_s[MAXLOC];
    for (const auto& q : rule) {
      ig.localInside(q.x, ip_s);
      Bs.evaluateFunction(ip_s, phi_s);
      tgrad(Bs, ig.inside, ip_s, lfsu_s.n, js, tg_s);
      for (int i = 0; i < lfsu_s.n; ++i) { an_s[i] = 0.0; for (int d = 0; d < dim; ++d) an_s[i] += F.n_F[d] * tg_s[i][d]; }
      double normalflux = 0.0; for (int d = 0; d < dim; ++d) normalflux += 0.0 * F.n_F[d];
      const double omegaup_s = normalflux >= 0.0 ? 1.0 : 0.0;
      const double factor = q.w * ig.integrationElement();
      const double ipfactor = F.penalty_factor * factor;
      for (int j = 0; j < lfsu_s.n; ++j) {
        const double temp1 = -an_s[j] * factor;
        for (int i = 0; i < lfsv_s.n; ++i) {
          mat_ss.accumulate(lfsv_s, i, lfsu_s, j, omegaup_s * phi_s[j] * normalflux * factor * phi_s[i]);
          mat_ss.accumulate(lfsv_s, i, lfsu_s, j, temp1 * phi_s[i]);
          mat_ss.accumulate(lfsv_s, i, lfsu_s, j, phi_s[j] * factor * theta * an_s[i]);
          mat_ss.accumulate(lfsv_s, i, lfsu_s, j, phi_s[j] * ipfactor * phi_s[i]);
        }
      }
    }
  }
};

// test/proportionalflowcoupling.hh:10-91
struct PropFlowOp : CouplingOperator {
  mdb_propflow_params P;
  PropFlowOp(const mdb_propflow_params& p, int jm) : P(p) { doAlphaCoupling = true; doPatternCoupling = true; jac_mode = jm; }
  void alpha_coupling(const IG& ig, const LFS& lfsu_s, const double* x_s, const LFS& lfsv_s,
                      const LFS& lfsu_n, const double* x_n, const LFS& lfsv_n,
                      LocalVectorView& r_s, LocalVectorView& r_n) const override {
    const int dim = ig.dim();
    const int intorder = 4;                                            // :43
    QkBasis B1(lfsv_s.k, dim), B2(lfsv_n.k, dim);
    const auto& rule = tensor_rule(dim - 1, intorder);
    double phi1[MAXLOC], phi2[MAXLOC], local1[3], local2[3];
    for (const auto& q : rule) {
      ig.localInside(q.x, local1); ig.localOutside(q.x, local2);
      B1.evaluateFunction(local1, phi1);
      B2.evaluateFunction(local2, phi2);
      double u_s = 0.0; for (int i = 0; i < lfsu_s.n; ++i) u_s += x_s[lfsu_s.idx[i]] * phi1[i];
      double u_n = 0.0; for (int i = 0; i < lfsu_n.n; ++i) u_n += x_n[lfsu_n.idx[i]] * phi2[i];
      double u_diff = P.intensity * (u_s - u_n);
      double factor = q.w * ig.integrationElement();
      for (int i = 0; i < lfsv_s.n; ++i) r_s.accumulate(lfsv_s, i, u_diff * phi1[i] * factor);
      for (int i = 0; i < lfsv_n.n; ++i) r_n.accumulate(lfsv_n, i, -u_diff * phi2[i] * factor);
    }
  }
};

// test/proportionalflowcoupling.hh:94-239
struct CVCFOp : CouplingOperator {
  mdb_cvcf_params P;
  CVCFOp(const mdb_cvcf_params& p, int jm) : P(p) { doAlphaCoupling = true; doPatternCoupling = true; jac_mode = jm; }
  void alpha_coupling(const IG& ig, const LFS& lfsu1, const double* x1, const LFS& lfsv1,
                      const LFS& lfsu2, const double* x2, const LFS& lfsv2,
                      LocalVectorView& r1, LocalVectorView& r2) const override {
    const int dim = ig.dim();
    const double h_F = ig.cornerDistance01();                           // :135
    QkBasis B1(lfsv1.k, dim), B2(lfsv2.k, dim);
    const auto& rule = tensor_rule(dim - 1, P.quad_order);
    double phi1[MAXLOC], phi2[MAXLOC], js1[MAXLOC][3], js2[MAXLOC][3], gradphi1[MAXLOC][3], gradphi2[MAXLOC][3];
    double local1[3], local2[3];
    for (const auto& q : rule) {
      ig.localInside(q.x, local1); ig.localOutside(q.x, local2);
      B1.evaluateFunction(local1, phi1);
      B2.evaluateFunction(local2, phi2);
      B1.evaluateJacobian(local1, js1);
      B2.evaluateJacobian(local2, js2);
      for (int i = 0; i < lfsu1.n; ++i)
        for (int d = 0; d < dim; ++d) { gradphi1[i][d] = 0.0; gradphi1[i][d] += ig.inside.jacInvT(d) * js1[i][d]; }
      for (int i = 0; i < lfsu2.n; ++i)
        for (int d = 0; d < dim; ++d) { gradphi2[i][d] = 0.0; gradphi2[i][d] += ig.outside.jacInvT(d) * js2[i][d]; }
      double gradu1[3] = {0, 0, 0}, gradu2[3] = {0, 0, 0};
      for (int i = 0; i < lfsu1.n; ++i) for (int d = 0; d < dim; ++d) gradu1[d] += x1[lfsu1.idx[i]] * gradphi1[i][d];
      for (int i = 0; i < lfsu2.n; ++i) for (int d = 0; d < dim; ++d) gradu2[d] += x2[lfsu2.idx[i]] * gradphi2[i][d];
      double u1 = 0.0; for (int i = 0; i < lfsu1.n; ++i) u1 += x1[lfsu1.idx[i]] * phi1[i];
      double u2 = 0.0; for (int i = 0; i < lfsu2.n; ++i) u2 += x2[lfsu2.idx[i]] * phi2[i];
      double normal1[3], normal2[3];
      ig.unitOuterNormal(normal1);
      ig.unitOuterNormal(normal2);
      for (int d = 0; d < dim; ++d) normal2[d] *= -1;
      const double jump_u1 = u1 - u2;
      const double jump_u2 = u2 - u1;
      double mean_gradu[3];
      for (int d = 0; d < dim; ++d) { mean_gradu[d] = gradu1[d] + gradu2[d]; mean_gradu[d] *= 0.5; }
      const double theta = 1;
      const double alpha = P.intensity;
      const double factor = q.w * ig.integrationElement();
      for (int i = 0; i < lfsv1.n; ++i) {
        double mgn = 0.0, gpn = 0.0;
        for (int d = 0; d < dim; ++d) mgn += mean_gradu[d] * normal1[d];
        for (int d = 0; d < dim; ++d) gpn += gradphi1[i][d] * normal1[d];
        double value = 0.0;
        value -= mgn * phi1[i];
        value -= theta * 0.5 * gpn * jump_u1;
        value += alpha / h_F * jump_u1 * phi1[i];
        r1.accumulate(lfsv1, i, factor * value);
      }
      for (int i = 0; i < lfsv2.n; ++i) {
        double mgn = 0.0, gpn = 0.0;
        for (int d = 0; d < dim; ++d) mgn += mean_gradu[d] * normal2[d];
        for (int d = 0; d < dim; ++d) gpn += gradphi2[i][d] * normal2[d];
        double value = 0.0;
        value -= mgn * phi2[i];
        value -= theta * 0.5 * gpn * jump_u2;
        value += alpha / h_F * jump_u2 * phi2[i];
        r2.accumulate(lfsv2, i, factor * value);
      }
    }
  }
};

// test/neumanndirichletcoupling.hh:41-441 with the parameter class CouplingParameters of
// test/testpoisson-dirichlet-neumann.cc:36-69 (A = identity, b = 0, c = 0).  One-directional: only r_s / mat_ss are written
// (r_r, mat_sr, mat_rs, mat_rr are const references there); the Jacobian is the operator's own analytic jacobian_coupling
// (:291-432), which deliberately has no derivative with respect to the remote coefficients (explicit coupling).
struct NDCouplingOp : CouplingOperator {
  mdb_ndcoupling_params P;
  double theta;
  explicit NDCouplingOp(const mdb_ndcoupling_params& p) : P(p) {
    doAlphaCoupling = true; doPatternCoupling = false;                  // :51-55 "no additional pattern relative to the subproblems"
    own_jacobian = true;
    theta = 1.0; if (P.method == MDB_DG_METHOD_SIPG) theta = -1.0;      // :72-73
  }
  struct Setup { double An_F_s[3], n_F[3], penalty_factor; int intorder; };
  Setup setup(const IG& ig, const LFS& lfsu_s, const LFS& lfsu_r) const {
    Setup S; const int dim = ig.dim();
    const int order = std::max(lfsu_s.k, lfsu_r.k);                     // :100-106 (lfsv = lfsu)
    S.intorder = P.intorderadd + 2 * order;                             // quadrature_factor(2) :69, :107
    // A_s = A_n = I (param.A), h_F = min(vol_s, vol_n) / vol_face (:128, the element_size value is overwritten)
    const double h_F = std::min(ig.inside.integrationElement(), ig.outside.integrationElement()) / ig.integrationElement();
    ig.unitOuterNormal(S.n_F);
    double An_F_n[3];
    for (int i = 0; i < dim; ++i) {                                     // A.mv(n_F, An_F): y_i = sum_j A_ij x_j
      S.An_F_s[i] = 0.0; An_F_n[i] = 0.0;
      for (int j = 0; j < dim; ++j) { S.An_F_s[i] += ((i == j) ? 1.0 : 0.0) * S.n_F[j]; An_F_n[i] += ((i == j) ? 1.0 : 0.0) * S.n_F[j]; }
    }
    double harmonic_average = 0.0;
    if (P.weights == MDB_DG_WEIGHTS_ON) {                               // :149-155
      double delta_s = 0.0, delta_n = 0.0;
      for (int d = 0; d < dim; ++d) { delta_s += S.An_F_s[d] * S.n_F[d]; delta_n += An_F_n[d] * S.n_F[d]; }
      harmonic_average = 2.0 * delta_s * delta_n / (delta_s + delta_n + 1e-20);
    } else harmonic_average = 1.0;
    const int degree = std::max(lfsu_s.k, lfsu_r.k);                    // :163-165
    S.penalty_factor = (P.alpha / h_F) * harmonic_average * degree * (degree + dim - 1);     // :168
    return S;
  }
  void alpha_coupling(const IG& ig, const LFS& lfsu_s, const double* x_s, const LFS& lfsv_s,
                      const LFS& lfsu_r, const double* x_r, const LFS&, LocalVectorView& r_s, LocalVectorView&) const override {
    const int dim = ig.dim();
    const Setup S = setup(ig, lfsu_s, lfsu_r);
    QkBasis Bs(lfsu_s.k, dim), Br(lfsu_r.k, dim);
    const auto& rule = tensor_rule(dim - 1, S.intorder);
    double psi_s[MAXLOC], phi_r[MAXLOC], js[MAXLOC][3], tgrad_r[MAXLOC][3], tgrad_s[MAXLOC][3], iplocal_s[3], iplocal_n[3];
    for (const auto& q : rule) {
      const double* n_F_local = S.n_F;                                  // ig.unitOuterNormal(pos): constant on an axis-aligned face
      ig.localInside(q.x, iplocal_s); ig.localOutside(q.x, iplocal_n);
      Bs.evaluateFunction(iplocal_s, psi_s);
      const double factor = q.w * ig.integrationElement();
      double normalflux = 0.0;                                          // b = 0 (param.b)
      for (int d = 0; d < dim; ++d) normalflux += 0.0 * n_F_local[d];
      Br.evaluateFunction(iplocal_n, phi_r);
      double u_r = 0.0;
      for (int i = 0; i < lfsu_r.n; ++i) u_r += x_r[lfsu_r.idx[i]] * phi_r[i];
      if (P.mode == MDB_ND_MODE_NEUMANN || P.mode == MDB_ND_MODE_BOTH) {      // :194-217: flux of the neighbour
        Br.evaluateJacobian(iplocal_n, js);
        for (int i = 0; i < lfsu_r.n; ++i)
          for (int d = 0; d < dim; ++d) { tgrad_r[i][d] = 0.0; tgrad_r[i][d] += ig.outside.jacInvT(d) * js[i][d]; }
        double gradu_r[3] = {0, 0, 0};
        for (int i = 0; i < lfsu_r.n; ++i) for (int d = 0; d < dim; ++d) gradu_r[d] += x_r[lfsu_r.idx[i]] * tgrad_r[i][d];
        double gn = 0.0; for (int d = 0; d < dim; ++d) gn += gradu_r[d] * n_F_local[d];
        const double j = normalflux * u_r - gn;
        for (int i = 0; i < lfsv_s.n; ++i) r_s.accumulate(lfsv_s, i, j * psi_s[i] * factor);
        if (P.mode == MDB_ND_MODE_NEUMANN) continue;
      }
      const double* phi_s = psi_s;                                      // lfsu_s = lfsv_s: the same basis at the same point
      double u_s = 0.0;
      for (int i = 0; i < lfsu_s.n; ++i) u_s += x_s[lfsu_s.idx[i]] * phi_s[i];
      Bs.evaluateJacobian(iplocal_s, js);
      for (int i = 0; i < lfsu_s.n; ++i)
        for (int d = 0; d < dim; ++d) { tgrad_s[i][d] = 0.0; tgrad_s[i][d] += ig.inside.jacInvT(d) * js[i][d]; }
      double gradu_s[3] = {0, 0, 0};
      for (int i = 0; i < lfsu_s.n; ++i) for (int d = 0; d < dim; ++d) gradu_s[d] += x_s[lfsu_s.idx[i]] * tgrad_s[i][d];
      const double omegaup_s = normalflux >= 0.0 ? 1.0 : 0.0, omegaup_r = normalflux >= 0.0 ? 0.0 : 1.0;
      const double term1 = (omegaup_s * u_s + omegaup_r * u_r) * normalflux * factor;               // convection
      for (int i = 0; i < lfsv_s.n; ++i) r_s.accumulate(lfsu_s, i, term1 * psi_s[i]);
      double Agu = 0.0; for (int d = 0; d < dim; ++d) Agu += S.An_F_s[d] * gradu_s[d];
      const double term2 = -factor * Agu;                                                            // diffusion
      for (int i = 0; i < lfsv_s.n; ++i) r_s.accumulate(lfsv_s, i, term2 * psi_s[i]);
      const double term3 = (u_s - u_r) * factor;                                                     // (non-)symmetric IP term
      for (int i = 0; i < lfsv_s.n; ++i) {
        double Agp = 0.0; for (int d = 0; d < dim; ++d) Agp += S.An_F_s[d] * tgrad_s[i][d];
        r_s.accumulate(lfsv_s, i, term3 * theta * Agp);
      }
      const double term4 = S.penalty_factor * (u_s - u_r) * factor;                                  // standard IP term
      for (int i = 0; i < lfsv_s.n; ++i) r_s.accumulate(lfsv_s, i, term4 * psi_s[i]);
    }
  }
  void jacobian_own(const IG& ig, const LFS& lfsu_s, const LFS& lfsu_r, LocalMatrixView& mat_ss) const override {
    if (P.mode == MDB_ND_MODE_NEUMANN) return;                          // :299-300
    const int dim = ig.dim();
    const Setup S = setup(ig, lfsu_s, lfsu_r);
    QkBasis Bs(lfsu_s.k, dim);
    const auto& rule = tensor_rule(dim - 1, S.intorder);
    double phi_s[MAXLOC], js[MAXLOC][3], tgrad_s[MAXLOC][3], iplocal_s[3];
    for (const auto& q : rule) {
      const double* n_F_local = S.n_F;
      ig.localInside(q.x, iplocal_s);
      Bs.evaluateFunction(iplocal_s, phi_s);
      Bs.evaluateJacobian(iplocal_s, js);
      for (int i = 0; i < lfsu_s.n; ++i)
        for (int d = 0; d < dim; ++d) { tgrad_s[i][d] = 0.0; tgrad_s[i][d] += ig.inside.jacInvT(d) * js[i][d]; }
      double normalflux = 0.0;
      for (int d = 0; d < dim; ++d) normalflux += 0.0 * n_F_local[d];
      const double omegaup_s = normalflux >= 0.0 ? 1.0 : 0.0;
      const double factor = q.w * ig.integrationElement();
      const double ipfactor = S.penalty_factor * factor;
      for (int j = 0; j < lfsu_s.n; ++j) {                              // :421-429: I convection, II diffusion, III consistency, IV ip
        double Agj = 0.0; for (int d = 0; d < dim; ++d) Agj += S.An_F_s[d] * tgrad_s[j][d];
        const double temp1 = -Agj * factor;
        for (int i = 0; i < lfsu_s.n; ++i) {
          double Agi = 0.0; for (int d = 0; d < dim; ++d) Agi += S.An_F_s[d] * tgrad_s[i][d];
          mat_ss.accumulate(lfsu_s, i, lfsu_s, j, omegaup_s * phi_s[j] * normalflux * factor * phi_s[i]);
          mat_ss.accumulate(lfsu_s, i, lfsu_s, j, temp1 * phi_s[i]);
          mat_ss.accumulate(lfsu_s, i, lfsu_s, j, phi_s[j] * factor * theta * Agi);
          mat_ss.accumulate(lfsu_s, i, lfsu_s, j, phi_s[j] * ipfactor * phi_s[i]);
        }
      }
    }
  }
};

} // anonymous namespace

// couplingutilities.hh:75-135 (NumericalJacobianCoupling).  MDB_JAC_ANALYTIC exploits linearity of the shipped
// coupling operators: column j = alpha_coupling(e_j) - alpha_coupling(0) evaluated exactly (no division).
void CouplingOperator::jacobian_coupling(const IG& ig, const LFS& lfsu_s, const double* x_s, int nx_s, const LFS& lfsv_s,
                                         const LFS& lfsu_n, const double* x_n, int nx_n, const LFS& lfsv_n,
                                         LocalMatrixView& mat_ss, LocalMatrixView& mat_sn,
                                         LocalMatrixView& mat_ns, LocalMatrixView& mat_nn) const
{
  if (own_jacobian) { jacobian_own(ig, lfsu_s, lfsu_n, mat_ss); return; }      // the operator implements jacobian_coupling itself
  const int m_s = lfsv_s.n, m_n = lfsv_n.n, n_s = lfsu_s.n, n_n = lfsu_n.n;
  std::vector<double> u_s(x_s, x_s + nx_s), u_n(x_n, x_n + nx_n);
  std::vector<double> down_s(nx_s, 0.0), up_s(nx_s, 0.0), down_n(nx_n, 0.0), up_n(nx_n, 0.0);
  LocalVectorView dv_s{down_s.data(), 1.0}, uv_s{up_s.data(), 1.0}, dv_n{down_n.data(), 1.0}, uv_n{up_n.data(), 1.0};
  if (jac_mode == MDB_JAC_ANALYTIC) {
    std::fill(u_s.begin(), u_s.end(), 0.0); std::fill(u_n.begin(), u_n.end(), 0.0);
    for (int j = 0; j < n_s; ++j) {
      std::fill(up_s.begin(), up_s.end(), 0.0); std::fill(up_n.begin(), up_n.end(), 0.0);
      u_s[lfsu_s.idx[j]] = 1.0;
      alpha_coupling(ig, lfsu_s, u_s.data(), lfsv_s, lfsu_n, u_n.data(), lfsv_n, uv_s, uv_n);
      for (int i = 0; i < m_s; ++i) mat_ss.accumulate(lfsv_s, i, lfsu_s, j, up_s[lfsv_s.idx[i]]);
      for (int i = 0; i < m_n; ++i) mat_ns.accumulate(lfsv_n, i, lfsu_s, j, up_n[lfsv_n.idx[i]]);
      u_s[lfsu_s.idx[j]] = 0.0;
    }
    for (int j = 0; j < n_n; ++j) {
      std::fill(up_s.begin(), up_s.end(), 0.0); std::fill(up_n.begin(), up_n.end(), 0.0);
      u_n[lfsu_n.idx[j]] = 1.0;
      alpha_coupling(ig, lfsu_s, u_s.data(), lfsv_s, lfsu_n, u_n.data(), lfsv_n, uv_s, uv_n);
      for (int i = 0; i < m_s; ++i) mat_sn.accumulate(lfsv_s, i, lfsu_n, j, up_s[lfsv_s.idx[i]]);
      for (int i = 0; i < m_n; ++i) mat_nn.accumulate(lfsv_n, i, lfsu_n, j, up_n[lfsv_n.idx[i]]);
      u_n[lfsu_n.idx[j]] = 0.0;
    }
    return;
  }
  // base line
  alpha_coupling(ig, lfsu_s, u_s.data(), lfsv_s, lfsu_n, u_n.data(), lfsv_n, dv_s, dv_n);
  // jiggle in self
  for (int j = 0; j < n_s; ++j) {
    std::fill(up_s.begin(), up_s.end(), 0.0); std::fill(up_n.begin(), up_n.end(), 0.0);
    double delta = epsilon * (1.0 + std::abs(u_s[lfsu_s.idx[j]]));
    u_s[lfsu_s.idx[j]] += delta;
    alpha_coupling(ig, lfsu_s, u_s.data(), lfsv_s, lfsu_n, u_n.data(), lfsv_n, uv_s, uv_n);
    for (int i = 0; i < m_s; ++i)
      mat_ss.accumulate(lfsv_s, i, lfsu_s, j, (up_s[lfsv_s.idx[i]] - down_s[lfsv_s.idx[i]]) / delta);
    for (int i = 0; i < m_n; ++i)
      mat_ns.accumulate(lfsv_n, i, lfsu_s, j, (up_n[lfsv_n.idx[i]] - down_n[lfsv_n.idx[i]]) / delta);
    u_s[lfsu_s.idx[j]] = x_s[lfsu_s.idx[j]];
  }
  // jiggle in neighbor
  for (int j = 0; j < n_n; ++j) {
    std::fill(up_s.begin(), up_s.end(), 0.0); std::fill(up_n.begin(), up_n.end(), 0.0);
    double delta = epsilon * (1.0 + std::abs(u_n[lfsu_n.idx[j]]));
    u_n[lfsu_n.idx[j]] += delta;
    alpha_coupling(ig, lfsu_s, u_s.data(), lfsv_s, lfsu_n, u_n.data(), lfsv_n, uv_s, uv_n);
    for (int i = 0; i < m_s; ++i)
      mat_sn.accumulate(lfsv_s, i, lfsu_n, j, (up_s[lfsv_s.idx[i]] - down_s[lfsv_s.idx[i]]) / delta);
    for (int i = 0; i < m_n; ++i)
      mat_nn.accumulate(lfsv_n, i, lfsu_n, j, (up_n[lfsv_n.idx[i]] - down_n[lfsv_n.idx[i]]) / delta);
    u_n[lfsu_n.idx[j]] = x_n[lfsu_n.idx[j]];
  }
}

// test/testmortarpoisson.cc:156-244 (MortarPoissonCoupling::alpha_generic<sign>): sign = +1 works on the inside element
// (alpha_enriched_coupling_first :131-139), sign = -1 on the outside element (_second :143-151).  Both use the unit outer
// normal of the INSIDE element (:198) and the mortar basis at the face-local quadrature point (:204).
void EnrichedOperator::alpha_generic(int sign, const IG& ig, const LFS& lfsu, const double* x, const LFS& lfsv, const LFS& lfsu_c,
                                     const double* x_c, const LFS& lfsv_c, LocalVectorView& r, LocalVectorView& r_c) const
{
  const int dim = ig.dim();
  const double gamma = gamma_0 / ig.integrationElement();                 // :191, [UPSTREAM] volume() of an axis-aligned face
  const int qorder = 2 * std::max(lfsu.k, lfsu_c.k);                      // :193-195, [UPSTREAM] QkLocalBasis::order() = k
  QkBasis B(lfsu.k, dim), Bc(lfsu_c.k, dim - 1);
  const auto& rule = tensor_rule(dim - 1, qorder);
  const EG& elem = sign > 0 ? ig.inside : ig.outside;
  double v[MAXLOC], mu[MAXLOC], js[MAXLOC][3], gradv[MAXLOC][3], element_pos[3], normal[3];
  for (const auto& q : rule) {
    if (sign > 0) ig.localInside(q.x, element_pos); else ig.localOutside(q.x, element_pos);
    ig.unitOuterNormal(normal);
    B.evaluateFunction(element_pos, v);
    Bc.evaluateFunction(q.x, mu);
    B.evaluateJacobian(element_pos, js);
    for (int i = 0; i < lfsu.n; ++i)
      for (int d = 0; d < dim; ++d) { gradv[i][d] = 0.0; gradv[i][d] += elem.jacInvT(d) * js[i][d]; }
    double u = 0.0, gradu[3] = {0, 0, 0};
    for (int i = 0; i < lfsu.n; ++i) {
      u += x[lfsu.idx[i]] * v[i];
      for (int d = 0; d < dim; ++d) gradu[d] += x[lfsu.idx[i]] * gradv[i][d];
    }
    double lambda = 0.0;
    for (int i = 0; i < lfsu_c.n; ++i) lambda += x_c[lfsu_c.idx[i]] * mu[i];
    const double factor = q.w * ig.integrationElement();
    double ngu = 0.0; for (int d = 0; d < dim; ++d) ngu += normal[d] * gradu[d];
    for (int i = 0; i < lfsu.n; ++i) {
      double ngv = 0.0; for (int d = 0; d < dim; ++d) ngv += normal[d] * gradv[i][d];
      r.accumulate(lfsv, i, (-ngu * v[i] - (u - lambda) * ngv + 2 * gamma * (u - lambda) * v[i]) * factor);
    }
    for (int i = 0; i < lfsu_c.n; ++i)
      r_c.accumulate(lfsv_c, i, (ngu - 2 * gamma * (u - lambda)) * mu[i] * factor);
  }
}

// test/testpowermortarpoisson.cc:186-262: `for (c_block < PowerLFSU_C::CHILDREN)` around the body of alpha_generic, each block
// with its own multiplier coefficients and its own rows of r_c; the element residual r collects every block's terms
void EnrichedOperator::alpha_power(int sign, const IG& ig, const LFS& lfsu, const double* x, const LFS& lfsv, const LFS& lfsu_c,
                                   const double* x_c, const LFS& lfsv_c, LocalVectorView& r, LocalVectorView& r_c) const
{
  if (ncomp == 1) { alpha_generic(sign, ig, lfsu, x, lfsv, lfsu_c, x_c, lfsv_c, r, r_c); return; }
  const int nb = lfsu_c.n / ncomp;
  for (int c_block = 0; c_block < ncomp; ++c_block) {
    LFS child; child.k = lfsu_c.k; child.n = nb;
    for (int i = 0; i < nb; ++i) child.idx[i] = lfsu_c.idx[c_block * nb + i];
    alpha_generic(sign, ig, lfsu, x, lfsv, child, x_c, child, r, r_c);
  }
}

// couplingutilities.hh:305-388 / :404-484 (NumericalJacobianEnrichedCouplingTo{First,Second}SubProblem; the two differ only
// in which alpha_enriched_coupling_* they call).  MDB_JAC_ANALYTIC: the operator is linear in (u, lambda), column j = alpha(e_j).
void EnrichedOperator::jacobian_generic(int sign, const IG& ig, const LFS& lfsu, const double* x, int nx, const LFS& lfsv,
                                        const LFS& lfsu_c, const double* x_c, int nxc, const LFS& lfsv_c, LocalMatrixView& mat_ss,
                                        LocalMatrixView& mat_sc, LocalMatrixView& mat_cs, LocalMatrixView& mat_cc) const
{
  const int m_s = lfsv.n, m_c = lfsv_c.n, n_s = lfsu.n, n_c = lfsu_c.n;
  std::vector<double> u_s(x, x + nx), u_c(x_c, x_c + nxc);
  std::vector<double> down_s(nx, 0.0), up_s(nx, 0.0), down_c(std::max(nxc, 1), 0.0), up_c(std::max(nxc, 1), 0.0);
  LocalVectorView dv_s{down_s.data(), 1.0}, uv_s{up_s.data(), 1.0}, dv_c{down_c.data(), 1.0}, uv_c{up_c.data(), 1.0};
  const bool analytic = jac_mode == MDB_JAC_ANALYTIC;
  if (analytic) { std::fill(u_s.begin(), u_s.end(), 0.0); std::fill(u_c.begin(), u_c.end(), 0.0); }
  else alpha_power(sign, ig, lfsu, u_s.data(), lfsv, lfsu_c, u_c.data(), lfsv_c, dv_s, dv_c);     // base line
  for (int j = 0; j < n_s; ++j) {                                                                      // jiggle in self
    std::fill(up_s.begin(), up_s.end(), 0.0); std::fill(up_c.begin(), up_c.end(), 0.0);
    double& uj = u_s[lfsu.idx[j]];
    const double keep = uj;
    const double delta = analytic ? 1.0 : epsilon * (1.0 + std::abs(uj));
    uj += delta;
    alpha_power(sign, ig, lfsu, u_s.data(), lfsv, lfsu_c, u_c.data(), lfsv_c, uv_s, uv_c);
    for (int i = 0; i < m_s; ++i) mat_ss.accumulate(lfsv, i, lfsu, j, (up_s[lfsv.idx[i]] - down_s[lfsv.idx[i]]) / delta);
    for (int i = 0; i < m_c; ++i) mat_cs.accumulate(lfsv_c, i, lfsu, j, (up_c[lfsv_c.idx[i]] - down_c[lfsv_c.idx[i]]) / delta);
    uj = keep;
  }
  for (int j = 0; j < n_c; ++j) {                                                                      // jiggle in coupling
    std::fill(up_s.begin(), up_s.end(), 0.0); std::fill(up_c.begin(), up_c.end(), 0.0);
    double& uj = u_c[lfsu_c.idx[j]];
    const double keep = uj;
    const double delta = analytic ? 1.0 : epsilon * (1.0 + std::abs(uj));
    uj += delta;
    alpha_power(sign, ig, lfsu, u_s.data(), lfsv, lfsu_c, u_c.data(), lfsv_c, uv_s, uv_c);
    for (int i = 0; i < m_s; ++i) mat_sc.accumulate(lfsv, i, lfsu_c, j, (up_s[lfsv.idx[i]] - down_s[lfsv.idx[i]]) / delta);
    for (int i = 0; i < m_c; ++i) mat_cc.accumulate(lfsv_c, i, lfsu_c, j, (up_c[lfsv_c.idx[i]] - down_c[lfsv_c.idx[i]]) / delta);
    uj = keep;
  }
}

std::unique_ptr<LocalOperator> make_local_operator(int op_id, const void* params, size_t nbytes)
{
  switch (op_id) {
  case MDB_OP_POISSON:
    if (nbytes != sizeof(mdb_poisson_params)) throw std::runtime_error("bad poisson params size");
    return std::unique_ptr<LocalOperator>(new PoissonOp(*static_cast<const mdb_poisson_params*>(params)));
  case MDB_OP_L2:
    if (nbytes != sizeof(mdb_l2_params)) throw std::runtime_error("bad l2 params size");
    return std::unique_ptr<LocalOperator>(new L2Op(*static_cast<const mdb_l2_params*>(params)));
  case MDB_OP_CONVDIFF_DG:
    if (nbytes != sizeof(mdb_convdiffdg_params)) throw std::runtime_error("bad convdiffdg params size");
    return std::unique_ptr<LocalOperator>(new ConvDiffDGOp(*static_cast<const mdb_convdiffdg_params*>(params)));
  }
  throw std::runtime_error("unknown local operator id");
}

std::unique_ptr<CouplingOperator> make_coupling_operator(int op_id, const void* params, size_t nbytes, int jac_mode)
{
  switch (op_id) {
  case MDB_OP_CVCF_COUPLING:
    if (nbytes != sizeof(mdb_cvcf_params)) throw std::runtime_error("bad cvcf params size");
    return std::unique_ptr<CouplingOperator>(new CVCFOp(*static_cast<const mdb_cvcf_params*>(params), jac_mode));
  case MDB_OP_PROPFLOW_COUPLING:
    if (nbytes != sizeof(mdb_propflow_params)) throw std::runtime_error("bad propflow params size");
    return std::unique_ptr<CouplingOperator>(new PropFlowOp(*static_cast<const mdb_propflow_params*>(params), jac_mode));
  case MDB_OP_ND_COUPLING:
    if (nbytes != sizeof(mdb_ndcoupling_params)) throw std::runtime_error("bad ndcoupling params size");
    return std::unique_ptr<CouplingOperator>(new NDCouplingOp(*static_cast<const mdb_ndcoupling_params*>(params)));
  }
  throw std::runtime_error("unknown coupling operator id");
}

// =============================================================================================
// DOF map  ([UPSTREAM] (ii)-(iv))
// =============================================================================================
void Context::finalize()
{
  const Grid& g = grid;
  const int dim = g.dim;
  if ((int64_t)cell_sd.size() != g.ncells()) throw std::runtime_error("subdomains not set");
  int64_t off = 0;
  for (auto& ch : children) {
    const int k = ch.k;
    int64_t total = 1;
    for (int d = 0; d < 3; ++d) { ch.lat_n[d] = (d < dim) ? k * g.n[d] + 1 : 1; total *= ch.lat_n[d]; }
    if (ch.dg) {
      // [UPSTREAM] (ii),(iii): all (k+1)^dim coefficients sit on the cell (codim 0): container index = subdomain cell index * nloc + i,
      // subdomain cell index = rank among the subdomain's cells in host order
      ch.cellblk.assign(g.ncells(), -1);
      int64_t cnt = 0;
      for (int64_t c = 0; c < g.ncells(); ++c)
        if (ch.subdomain < 0 || ((cell_sd[c] >> ch.subdomain) & 1)) ch.cellblk[c] = cnt++;
      ch.size = cnt * ch.nloc(dim); ch.offset = off; off += ch.size;
      ch.lat2dof.clear(); ch.dof2lat.clear();
      continue;
    }
    std::vector<uint8_t> present(total, 0);
    // coupling space (couplinggfsordering.hh:113-201): the used entities are the vertices of the intersections the
    // predicate accepts; entity_dof_offsets is indexed by the MultiDomainGrid vertex index, so after the partial sum the
    // DOFs are numbered by ascending vertex index — the same rule as the k = 1 branch below.
    if (ch.iface) {
      if (k != 1) throw std::runtime_error("oracle: coupling spaces are restated for Q1 only");
      for (int64_t c = 0; c < g.ncells(); ++c) {
        int64_t ijk[3]; g.cell_ijk(c, ijk);
        for (int f = 0; f < 2 * dim; ++f) {
          int a = f / 2, sgn = f % 2; int64_t nijk[3] = {ijk[0], ijk[1], ijk[2]};
          nijk[a] += sgn ? 1 : -1;
          if (nijk[a] < 0 || nijk[a] >= g.n[a]) continue;                          // !is.neighbor()
          if (!iface_applies(ch, cell_sd[c], cell_sd[g.cell_index(nijk)])) continue;
          for (int i = 0; i < (1 << (dim - 1)); ++i) {
            int64_t v[3] = {ijk[0], ijk[1], ijk[2]}; int t = 0;
            for (int d = 0; d < dim; ++d) { if (d == a) v[d] += sgn; else { v[d] += (i >> t) & 1; ++t; } }
            present[v[0] + ch.lat_n[0] * (v[1] + ch.lat_n[1] * v[2])] = 1;
          }
        }
      }
    }
    // closure of the cells of the child's grid view
    for (int64_t c = 0; c < g.ncells() && !ch.iface; ++c) {
      if (ch.subdomain >= 0 && !((cell_sd[c] >> ch.subdomain) & 1)) continue;
      int64_t ijk[3]; g.cell_ijk(c, ijk);
      int a[3];
      int nl = ch.nloc(dim);
      for (int i = 0; i < nl; ++i) {
        int rem = i; for (int d = 0; d < 3; ++d) { a[d] = (d < dim) ? rem % (k + 1) : 0; rem /= (k + 1); }
        int64_t p = (k * ijk[0] + a[0]) + ch.lat_n[0] * ((k * ijk[1] + a[1]) + ch.lat_n[1] * (k * ijk[2] + a[2]));
        present[p] = 1;
      }
    }
    // [UPSTREAM] (ii),(iii): number per geometry type (entity dimension ascending: vertex < line < quad < hex);
    // within a dimension entities are grouped by the set of axes they extend along (ascending bit pattern) and
    // lexicographic (axis 0 fastest) inside a group; subdomain entity index = rank in that host order.
    ch.lat2dof.assign(total, -1);
    ch.dof2lat.clear();
    int64_t cnt = 0;
    for (int m = 0; m <= dim; ++m)
      for (int ext = 0; ext < (1 << dim); ++ext) {
        if (__builtin_popcount(ext) != m) continue;
        if (k == 1 && ext != 0) continue;
        for (int64_t p = 0; p < total; ++p) {
          if (!present[p]) continue;
          int64_t rem = p; int e = 0;
          for (int d = 0; d < dim; ++d) { int64_t pd = rem % ch.lat_n[d]; rem /= ch.lat_n[d]; if (pd % k != 0) e |= (1 << d); }
          if (e != ext) continue;
          ch.lat2dof[p] = cnt++; ch.dof2lat.push_back(p);
        }
      }
    ch.size = cnt;
    ch.offset = off;            // [UPSTREAM] (iv) LexicographicOrderingTag: prefix sums of child sizes
    off += cnt;
  }
  ndof = off;
  constrained.assign(ndof, 0);
  finalized = true;
}

int Context::bind(int64_t cell, int* child_off, int64_t* dofs) const
{
  // multidomainlocalfunctionspace.hh:104-122: a child contributes iff the cell is in the child's subdomain
  const int dim = grid.dim;
  int64_t ijk[3]; grid.cell_ijk(cell, ijk);
  int n = 0;
  for (size_t c = 0; c < children.size(); ++c) {
    const ChildSpace& ch = children[c];
    child_off[c] = n;
    if (ch.iface) continue;      // bound per intersection (globalassembler.hh:212-231), not per cell
    if (ch.subdomain >= 0 && !((cell_sd[cell] >> ch.subdomain) & 1)) continue;
    const int k = ch.k, nl = ch.nloc(dim);
    if (ch.dg) { for (int i = 0; i < nl; ++i) dofs[n++] = ch.offset + ch.cellblk[cell] * nl + i; continue; }
    for (int i = 0; i < nl; ++i) {
      int a[3]; int rem = i; for (int d = 0; d < 3; ++d) { a[d] = (d < dim) ? rem % (k + 1) : 0; rem /= (k + 1); }
      int64_t p = (k * ijk[0] + a[0]) + ch.lat_n[0] * ((k * ijk[1] + a[1]) + ch.lat_n[1] * (k * ijk[2] + a[2]));
      dofs[n++] = ch.offset + ch.lat2dof[p];
    }
  }
  child_off[children.size()] = n;
  return n;
}

bool Context::iface_applies(const ChildSpace& ch, uint8_t sd_in, uint8_t sd_out) const
{
  // SubProblemSubProblemInterface::operator() couplinggridfunctionspace.hh:29-35
  auto cond = [](int kind, uint32_t mask, uint8_t sd) {
    return kind == MDB_COND_EQUALITY ? uint32_t(sd) == mask : (uint32_t(sd) & mask) == mask;
  };
  return cond(ch.lkind, ch.lmask, sd_in) && cond(ch.rkind, ch.rmask, sd_out);
}

int Context::bind_coupling(int child, int64_t cell, int f, int64_t* dofs) const
{
  // couplinglocalfunctionspace.hh:208-287: local DOF i <-> corner i of the intersection (Q1 on the face), mapped to the
  // inside element's vertex by DOFMapper::mapSubIndex (dofmapper.hh:37-50).  [UPSTREAM] Yasp: corner i of face (axis, side)
  // has bit t of i = coordinate along the t-th axis other than `axis`.
  const ChildSpace& ch = children[child];
  const int dim = grid.dim;
  int64_t ijk[3]; grid.cell_ijk(cell, ijk);
  const int a = f / 2, sgn = f % 2;
  int64_t nijk[3] = {ijk[0], ijk[1], ijk[2]};
  nijk[a] += sgn ? 1 : -1;
  if (nijk[a] < 0 || nijk[a] >= grid.n[a]) return 0;
  if (!iface_applies(ch, cell_sd[cell], cell_sd[grid.cell_index(nijk)])) return 0;     // !gfs.contains(is)
  const int nl = 1 << (dim - 1);
  for (int i = 0; i < nl; ++i) {
    int64_t v[3] = {ijk[0], ijk[1], ijk[2]}; int t = 0;
    for (int d = 0; d < dim; ++d) { if (d == a) v[d] += sgn; else { v[d] += (i >> t) & 1; ++t; } }
    dofs[i] = ch.offset + ch.lat2dof[v[0] + ch.lat_n[0] * (v[1] + ch.lat_n[1] * v[2])];
  }
  return nl;
}

int Context::bind_coupling_power(const Coupling& cp, int64_t cell, int f, int64_t* dofs) const
{
  int n = 0;
  for (int c = 0; c < cp.ncomp; ++c) n += bind_coupling(cp.coupling_child + c, cell, f, dofs + n);
  return n;
}

LFS Context::subproblem_lfs(const SubProblem& sp, const int* child_off) const
{
  // subproblemlocalfunctionspace.hh:462-501: single-child subproblem = proxy of that child
  if (sp.children.size() != 1) throw std::runtime_error("oracle: only single-child subproblems are restated");
  LFS l; int c = sp.children[0];
  l.k = children[c].k; l.dg = children[c].dg;
  l.n = child_off[c + 1] - child_off[c];
  for (int i = 0; i < l.n; ++i) l.idx[i] = child_off[c] + i;
  return l;
}

// =============================================================================================
// the traversal (globalassembler.hh:43-298) with three engines
// =============================================================================================
namespace {

constexpr int MAXCELL = 64;  // max size of a cell-local space (sum over children)

struct FaceIter {
  // faces of a cell in grid order -x,+x,-y,+y,-z,+z
  static bool neighbor(const Grid& g, const int64_t ijk[3], int f, int64_t nijk[3]) {
    int a = f / 2, s = f % 2;
    for (int d = 0; d < 3; ++d) nijk[d] = ijk[d];
    nijk[a] += s ? 1 : -1;
    return nijk[a] >= 0 && nijk[a] < g.n[a];
  }
};

EG make_eg(const Context& c, int64_t cell) {
  EG e; e.g = &c.grid; e.index = cell; c.grid.cell_ijk(cell, e.ijk); e.sd = c.cell_sd[cell]; return e;
}

enum Mode { PATTERN, RESIDUAL, JACOBIAN };

struct Flags { bool visit_faces, two_sided; };

Flags engine_flags(const Context& c, const GridOperatorDesc& go, Mode mode)
{
  bool a_skel = false, a_bnd = false, l_skel = false, l_bnd = false, two = false, a_coup = false, p_skel = false, p_coup = false;
  for (int s : go.sps) {
    const LocalOperator& l = *c.sps[s].lop;
    a_skel |= l.doAlphaSkeleton; a_bnd |= l.doAlphaBoundary; l_skel |= l.doLambdaSkeleton; l_bnd |= l.doLambdaBoundary;
    two |= l.doSkeletonTwoSided; p_skel |= l.doPatternSkeleton;
  }
  bool enr = false;                // do{Alpha,Pattern}EnrichedCouplingToSubProblems: all true for the restated mortar operator
  for (int k : go.cps) {
    if (c.cps[k].eop) { enr = true; continue; }
    a_coup |= c.cps[k].op->doAlphaCoupling; p_coup |= c.cps[k].op->doPatternCoupling;
  }
  Flags f;
  if (mode == PATTERN) {           // patternassemblerengine.hh:96-139
    f.visit_faces = p_skel || p_coup || enr;
    f.two_sided = p_coup || two || enr;
  } else if (mode == JACOBIAN) {   // jacobianassemblerengine.hh:149-192
    f.visit_faces = a_skel || a_bnd || a_coup || enr;
    f.two_sided = a_bnd || two || a_coup || enr;
  } else {                         // residualassemblerengine.hh:117-188
    f.visit_faces = a_skel || a_bnd || a_coup || l_skel || l_bnd || enr;
    f.two_sided = a_bnd || l_bnd || two || a_coup || enr;
  }
  return f;
}

} // namespace

// ---------------------------------------------------------------------------------------------
// pattern  (patternassemblerengine.hh + patternassemblerfunctors.hh; [UPSTREAM] (v) union of local links, rows ascending)
// ---------------------------------------------------------------------------------------------
int Context::matrix_create(const int* gos_, int ngos)
{
  if (!finalized) throw std::runtime_error("not finalized");
  std::vector<std::vector<int64_t>> rows(ndof);
  const int nch = (int)children.size();
  std::vector<int> off_s(nch + 1), off_n(nch + 1);
  int64_t dofs_s[MAXCELL], dofs_n[MAXCELL];
  for (int gi = 0; gi < ngos; ++gi) {
    const GridOperatorDesc& go = gos.at(gos_[gi]);
    Flags fl = engine_flags(*this, go, PATTERN);
    for (int64_t cell = 0; cell < grid.ncells(); ++cell) {
      EG eg = make_eg(*this, cell);
      bind(cell, off_s.data(), dofs_s);
      // volume: FullVolumePattern [UPSTREAM]
      for (int s : go.sps) {
        const SubProblem& sp = sps[s];
        if (!sp.appliesTo(eg.sd) || !sp.lop->doPatternVolume) continue;
        LFS l = subproblem_lfs(sp, off_s.data());
        for (int i = 0; i < l.n; ++i) for (int j = 0; j < l.n; ++j) rows[dofs_s[l.idx[i]]].push_back(dofs_s[l.idx[j]]);
      }
      if (!fl.visit_faces) continue;
      for (int f = 0; f < 2 * grid.dim; ++f) {
        int64_t nijk[3];
        if (!FaceIter::neighbor(grid, eg.ijk, f, nijk)) continue;     // pattern has no boundary terms for our operators
        int64_t ncell = grid.cell_index(nijk);
        if (!fl.two_sided && cell < ncell) continue;
        EG en = make_eg(*this, ncell);
        bind(ncell, off_n.data(), dofs_n);
        for (int s : go.sps) {                                                 // invoke_pattern_skeleton_or_boundary :36-66
          const SubProblem& sp = sps[s];
          if (!sp.appliesTo(eg.sd) || !sp.appliesTo(en.sd)) continue;          // (pattern_boundary: none of the restated operators)
          if (!sp.lop->doPatternSkeleton || !(sp.lop->doSkeletonTwoSided || cell > ncell)) continue;
          LFS ls = subproblem_lfs(sp, off_s.data()), ln = subproblem_lfs(sp, off_n.data());
          // [UPSTREAM] FullSkeletonPattern: sn and ns
          for (int i = 0; i < ls.n; ++i) for (int j = 0; j < ln.n; ++j) rows[dofs_s[ls.idx[i]]].push_back(dofs_n[ln.idx[j]]);
          for (int i = 0; i < ln.n; ++i) for (int j = 0; j < ls.n; ++j) rows[dofs_n[ln.idx[i]]].push_back(dofs_s[ls.idx[j]]);
        }
        for (int k : go.cps) {
          const Coupling& cp = cps[k];
          if (cp.eop) {                                                        // invoke_pattern_enriched_coupling :116-150
            const SubProblem& lsp = sps[cp.local_sp]; const SubProblem& rsp = sps[cp.remote_sp];
            if (!(lsp.appliesTo(eg.sd) && rsp.appliesTo(en.sd))) continue;
            LFS ls = subproblem_lfs(lsp, off_s.data()), ln = subproblem_lfs(rsp, off_n.data());
            int64_t dofs_c[MAXLOC];
            const int nc = bind_coupling_power(cp, cell, f, dofs_c);
            // FullEnrichedCoupling{First,Second}SubProblemPattern + FullEnrichedCouplingPattern couplingutilities.hh:219-286
            for (int i = 0; i < ls.n; ++i) for (int j = 0; j < nc; ++j) rows[dofs_s[ls.idx[i]]].push_back(dofs_c[j]);
            for (int i = 0; i < nc; ++i) for (int j = 0; j < ls.n; ++j) rows[dofs_c[i]].push_back(dofs_s[ls.idx[j]]);
            for (int i = 0; i < ln.n; ++i) for (int j = 0; j < nc; ++j) rows[dofs_n[ln.idx[i]]].push_back(dofs_c[j]);
            for (int i = 0; i < nc; ++i) for (int j = 0; j < ln.n; ++j) rows[dofs_c[i]].push_back(dofs_n[ln.idx[j]]);
            for (int i = 0; i < nc; ++i) for (int j = 0; j < nc; ++j) rows[dofs_c[i]].push_back(dofs_c[j]);
            continue;
          }
          if (!cp.op->doPatternCoupling) continue;
          const SubProblem& lsp = sps[cp.local_sp]; const SubProblem& rsp = sps[cp.remote_sp];
          if (!(lsp.appliesTo(eg.sd) && rsp.appliesTo(en.sd))) continue;     // coupling.hh:78-83
          LFS ls = subproblem_lfs(lsp, off_s.data()), ln = subproblem_lfs(rsp, off_n.data());
          // FullCouplingPattern couplingutilities.hh:35-47
          for (int i = 0; i < ls.n; ++i) for (int j = 0; j < ln.n; ++j) rows[dofs_s[ls.idx[i]]].push_back(dofs_n[ln.idx[j]]);
          for (int i = 0; i < ln.n; ++i) for (int j = 0; j < ls.n; ++j) rows[dofs_n[ln.idx[i]]].push_back(dofs_s[ls.idx[j]]);
        }
      }
    }
  }
  CSR A; A.nrows = ndof; A.rowptr.assign(ndof + 1, 0);
  for (int64_t r = 0; r < ndof; ++r) {
    auto& v = rows[r]; std::sort(v.begin(), v.end()); v.erase(std::unique(v.begin(), v.end()), v.end());
    A.rowptr[r + 1] = A.rowptr[r] + (int64_t)v.size();
  }
  A.colidx.resize(A.rowptr[ndof]);
  for (int64_t r = 0; r < ndof; ++r) std::copy(rows[r].begin(), rows[r].end(), A.colidx.begin() + A.rowptr[r]);
  A.val.assign(A.colidx.size(), 0.0);
  mats.push_back(std::move(A));
  return (int)mats.size() - 1;
}

// ---------------------------------------------------------------------------------------------
// residual  (residualassemblerengine.hh + residualassemblerfunctors.hh)
// ---------------------------------------------------------------------------------------------
void Context::residual(int goi, double weight, const double* x, double* r) const
{
  const GridOperatorDesc& go = gos.at(goi);
  Flags fl = engine_flags(*this, go, RESIDUAL);
  const int nch = (int)children.size();
  std::vector<int> off_s(nch + 1), off_n(nch + 1);
  int64_t dofs_s[MAXCELL], dofs_n[MAXCELL];
  double x_s[MAXCELL], x_n[MAXCELL], r_s[MAXCELL], r_n[MAXCELL];
  for (int s : go.sps) sps[s].lop->time = time;
  for (int64_t cell = 0; cell < grid.ncells(); ++cell) {
    EG eg = make_eg(*this, cell);
    int ns = bind(cell, off_s.data(), dofs_s);
    for (int i = 0; i < ns; ++i) { x_s[i] = x[dofs_s[i]]; r_s[i] = 0.0; }
    LocalVectorView rv_s{r_s, weight};
    for (int s : go.sps) {                              // alpha_volume functor :16-35, lambda_volume :38-55
      const SubProblem& sp = sps[s];
      if (!sp.appliesTo(eg.sd)) continue;
      LFS l = subproblem_lfs(sp, off_s.data());
      if (sp.lop->doAlphaVolume) sp.lop->alpha_volume(eg, l, x_s, l, rv_s);
    }
    for (int s : go.sps) {
      const SubProblem& sp = sps[s];
      if (!sp.appliesTo(eg.sd)) continue;
      LFS l = subproblem_lfs(sp, off_s.data());
      if (sp.lop->doLambdaVolume) sp.lop->lambda_volume(eg, l, rv_s);
    }
    if (fl.visit_faces) {
      for (int f = 0; f < 2 * grid.dim; ++f) {
        int64_t nijk[3];
        IG ig; ig.inside = eg; ig.axis = f / 2; ig.side = f % 2; ig.face = f;
        if (FaceIter::neighbor(grid, eg.ijk, f, nijk)) {
          int64_t ncell = grid.cell_index(nijk);
          if (!fl.two_sided && cell < ncell) continue;
          EG en = make_eg(*this, ncell);
          ig.outside = en; ig.boundary = false; ig.oneSidedDirection = cell > ncell;
          int nn = bind(ncell, off_n.data(), dofs_n);
          for (int i = 0; i < nn; ++i) { x_n[i] = x[dofs_n[i]]; r_n[i] = 0.0; }
          LocalVectorView rv_n{r_n, weight};
          // assembleUVSkeleton: alpha_skeleton_or_boundary on subproblems (:58-88) then alpha_coupling (:165-191)
          for (int s : go.sps) {
            const SubProblem& sp = sps[s];
            if (!sp.appliesTo(eg.sd)) continue;
            LFS l = subproblem_lfs(sp, off_s.data());
            if (sp.appliesTo(en.sd)) {
              if (!sp.lop->doAlphaSkeleton || !(sp.lop->doSkeletonTwoSided || ig.oneSidedDirection)) continue;
              LFS ln = subproblem_lfs(sp, off_n.data());
              sp.lop->alpha_skeleton(ig, l, x_s, l, ln, x_n, ln, rv_s, rv_n);
            } else if (sp.lop->doAlphaBoundary) sp.lop->alpha_boundary(ig, l, x_s, l, rv_s);     // the interface seen as boundary (:80-86)
          }
          for (int k : go.cps) {
            const Coupling& cp = cps[k];
            if (cp.eop || !cp.op->doAlphaCoupling) continue;
            const SubProblem& lsp = sps[cp.local_sp]; const SubProblem& rsp = sps[cp.remote_sp];
            if (!(lsp.appliesTo(eg.sd) && rsp.appliesTo(en.sd))) continue;
            LFS ls = subproblem_lfs(lsp, off_s.data()), ln = subproblem_lfs(rsp, off_n.data());
            cp.op->alpha_coupling(ig, ls, x_s, ls, ln, x_n, ln, rv_s, rv_n);
          }
          // assembleVSkeleton: lambda_skeleton_or_boundary (:91-120)
          for (int s : go.sps) {
            const SubProblem& sp = sps[s];
            if (!sp.appliesTo(eg.sd)) continue;
            if (sp.appliesTo(en.sd)) { /* lambda_skeleton: none */ }
            else if (sp.lop->doLambdaBoundary) {       // subdomain interface seen as boundary (:113-118)
              LFS l = subproblem_lfs(sp, off_s.data());
              sp.lop->lambda_boundary(ig, l, rv_s);
            }
          }
          // assembleUVEnrichedCoupling (globalassembler.hh:210-238; alpha_enriched_coupling functor :219-257): the mortar
          // space is bound to the intersection, first works on (inside, mortar), second on (outside, mortar)
          for (int k : go.cps) {
            const Coupling& cp = cps[k];
            if (!cp.eop) continue;
            const SubProblem& lsp = sps[cp.local_sp]; const SubProblem& rsp = sps[cp.remote_sp];
            if (!(lsp.appliesTo(eg.sd) && rsp.appliesTo(en.sd))) continue;
            LFS ls = subproblem_lfs(lsp, off_s.data()), ln = subproblem_lfs(rsp, off_n.data());
            int64_t dofs_c[MAXLOC]; double x_c[MAXLOC], r_c[MAXLOC];
            LFS lc; lc.k = children[cp.coupling_child].k;
            lc.n = bind_coupling_power(cp, cell, f, dofs_c);
            for (int i = 0; i < lc.n; ++i) { lc.idx[i] = i; x_c[i] = x[dofs_c[i]]; r_c[i] = 0.0; }
            LocalVectorView rv_c{r_c, weight};
            cp.eop->alpha_power(+1, ig, ls, x_s, ls, lc, x_c, lc, rv_s, rv_c);
            cp.eop->alpha_power(-1, ig, ln, x_n, ln, lc, x_c, lc, rv_n, rv_c);
            for (int i = 0; i < lc.n; ++i) r[dofs_c[i]] += r_c[i];           // onUnbindLFSVCoupling
          }
          for (int i = 0; i < nn; ++i) r[dofs_n[i]] += r_n[i];           // onUnbindLFSVOutside
        } else {
          if (grid.processor_face(eg.ijk, f)) continue;                   // assembleVProcessor: no shipped operator has it
          ig.outside = eg; ig.boundary = true; ig.oneSidedDirection = true;
          for (int s : go.sps) {                                          // alpha_boundary functor :123-143
            const SubProblem& sp = sps[s];
            if (!sp.appliesTo(eg.sd) || !sp.lop->doAlphaBoundary) continue;
            LFS l = subproblem_lfs(sp, off_s.data());
            sp.lop->alpha_boundary(ig, l, x_s, l, rv_s);
          }
          for (int s : go.sps) {                                          // lambda_boundary functor :145-162
            const SubProblem& sp = sps[s];
            if (!sp.appliesTo(eg.sd) || !sp.lop->doLambdaBoundary) continue;
            LFS l = subproblem_lfs(sp, off_s.data());
            sp.lop->lambda_boundary(ig, l, rv_s);
          }
        }
      }
    }
    for (int i = 0; i < ns; ++i) r[dofs_s[i]] += r_s[i];                  // onUnbindLFSV
  }
  // postAssembly: constrain_residual [UPSTREAM] (vi)
  for (int64_t i = 0; i < ndof; ++i) if (constrained[i]) r[i] = 0.0;
}

// ---------------------------------------------------------------------------------------------
// jacobian  (jacobianassemblerengine.hh + jacobianassemblerfunctors.hh)
// ---------------------------------------------------------------------------------------------
namespace {
struct JacScratch {
  std::vector<int> off_s, off_n;
  int64_t dofs_s[MAXCELL], dofs_n[MAXCELL];
  double x_s[MAXCELL], x_n[MAXCELL];
  std::vector<double> a_ss, a_sn, a_ns, a_nn;
  explicit JacScratch(int nch) : off_s(nch + 1), off_n(nch + 1), a_ss(MAXCELL * MAXCELL), a_sn(MAXCELL * MAXCELL),
                                 a_ns(MAXCELL * MAXCELL), a_nn(MAXCELL * MAXCELL) {}
};

void scatter(CSR& A, const int64_t* rd, int nr, const int64_t* cd, int nc, const double* a, bool must_exist)
{
  // [UPSTREAM] (vi) scatter_jacobian = plain add of every local entry whose (row,col) is in the pattern; the local
  // blocks are dense, entries outside the operator's own pattern are exact zeros and are skipped.
  for (int i = 0; i < nr; ++i)
    for (int j = 0; j < nc; ++j) {
      double v = a[i * nc + j];
      int64_t p = A.find(rd[i], cd[j]);
      if (p < 0) { if (v != 0.0 && must_exist) throw std::runtime_error("jacobian entry outside pattern"); continue; }
      A.val[p] += v;
    }
}

void jacobian_cell(const Context& c, const GridOperatorDesc& go, const Flags& fl, CSR& A, double weight, const double* x,
                   int64_t cell, JacScratch& S)
{
  const Grid& grid = c.grid;
  EG eg = make_eg(c, cell);
  int ns = c.bind(cell, S.off_s.data(), S.dofs_s);
  for (int i = 0; i < ns; ++i) S.x_s[i] = x[S.dofs_s[i]];
  std::fill(S.a_ss.begin(), S.a_ss.begin() + ns * ns, 0.0);
  LocalMatrixView m_ss{S.a_ss.data(), ns, weight};
  for (int s : go.sps) {                                 // invoke_jacobian_volume :16-35
    const SubProblem& sp = c.sps[s];
    if (!sp.appliesTo(eg.sd) || !sp.lop->doAlphaVolume) continue;
    LFS l = c.subproblem_lfs(sp, S.off_s.data());
    sp.lop->jacobian_volume(eg, l, S.x_s, l, m_ss);
  }
  if (fl.visit_faces) {
    for (int f = 0; f < 2 * grid.dim; ++f) {
      int64_t nijk[3];
      if (!FaceIter::neighbor(grid, eg.ijk, f, nijk)) {
        if (grid.processor_face(eg.ijk, f)) continue;
        IG ig; ig.inside = eg; ig.outside = eg; ig.axis = f / 2; ig.side = f % 2; ig.face = f; ig.boundary = true; ig.oneSidedDirection = true;
        for (int s : go.sps) {                              // invoke_jacobian_boundary :76-94
          const SubProblem& sp = c.sps[s];
          if (!sp.appliesTo(eg.sd) || !sp.lop->doAlphaBoundary) continue;
          LFS l = c.subproblem_lfs(sp, S.off_s.data());
          sp.lop->jacobian_boundary(ig, l, S.x_s, l, m_ss);
        }
        continue;
      }
      int64_t ncell = grid.cell_index(nijk);
      if (!fl.two_sided && cell < ncell) continue;
      EG en = make_eg(c, ncell);
      bool any = false;
      for (int s : go.sps) {
        const SubProblem& sp = c.sps[s];
        if (sp.appliesTo(eg.sd) && (sp.lop->doAlphaSkeleton || sp.lop->doAlphaBoundary)) any = true;
      }
      for (int k : go.cps) {
        const Coupling& cp = c.cps[k];
        if ((cp.eop || cp.op->doAlphaCoupling) && c.sps[cp.local_sp].appliesTo(eg.sd) && c.sps[cp.remote_sp].appliesTo(en.sd)) any = true;
      }
      if (!any) continue;      // the reference binds lfs_n and scatters all-zero blocks here; no effect on values
      IG ig; ig.inside = eg; ig.outside = en; ig.axis = f / 2; ig.side = f % 2; ig.face = f; ig.boundary = false;
      ig.oneSidedDirection = cell > ncell;
      int nn = c.bind(ncell, S.off_n.data(), S.dofs_n);
      for (int i = 0; i < nn; ++i) S.x_n[i] = x[S.dofs_n[i]];
      std::fill(S.a_sn.begin(), S.a_sn.begin() + ns * nn, 0.0);
      std::fill(S.a_ns.begin(), S.a_ns.begin() + nn * ns, 0.0);
      std::fill(S.a_nn.begin(), S.a_nn.begin() + nn * nn, 0.0);
      LocalMatrixView m_sn{S.a_sn.data(), nn, weight}, m_ns{S.a_ns.data(), ns, weight}, m_nn{S.a_nn.data(), nn, weight};
      for (int s : go.sps) {                              // invoke_jacobian_skeleton_or_boundary :36-72
        const SubProblem& sp = c.sps[s];
        if (!sp.appliesTo(eg.sd)) continue;
        LFS l = c.subproblem_lfs(sp, S.off_s.data());
        if (sp.appliesTo(en.sd)) {
          if (!sp.lop->doAlphaSkeleton || !(sp.lop->doSkeletonTwoSided || ig.oneSidedDirection)) continue;
          LFS ln = c.subproblem_lfs(sp, S.off_n.data());
          sp.lop->jacobian_skeleton(ig, l, S.x_s, l, ln, S.x_n, ln, m_ss, m_sn, m_ns, m_nn);
        } else if (sp.lop->doAlphaBoundary) sp.lop->jacobian_boundary(ig, l, S.x_s, l, m_ss);
      }
      for (int k : go.cps) {                              // invoke_jacobian_coupling :91-117
        const Coupling& cp = c.cps[k];
        if (cp.eop || !cp.op->doAlphaCoupling) continue;
        const SubProblem& lsp = c.sps[cp.local_sp]; const SubProblem& rsp = c.sps[cp.remote_sp];
        if (!(lsp.appliesTo(eg.sd) && rsp.appliesTo(en.sd))) continue;
        LFS ls = c.subproblem_lfs(lsp, S.off_s.data()), ln = c.subproblem_lfs(rsp, S.off_n.data());
        cp.op->jacobian_coupling(ig, ls, S.x_s, ns, ls, ln, S.x_n, nn, ln, m_ss, m_sn, m_ns, m_nn);
      }
      for (int k : go.cps) {                              // invoke_jacobian_enriched_coupling :121-158
        const Coupling& cp = c.cps[k];
        if (!cp.eop) continue;
        const SubProblem& lsp = c.sps[cp.local_sp]; const SubProblem& rsp = c.sps[cp.remote_sp];
        if (!(lsp.appliesTo(eg.sd) && rsp.appliesTo(en.sd))) continue;
        LFS ls = c.subproblem_lfs(lsp, S.off_s.data()), ln = c.subproblem_lfs(rsp, S.off_n.data());
        int64_t dofs_c[MAXLOC]; double x_c[MAXLOC];
        LFS lc; lc.k = c.children[cp.coupling_child].k;
        lc.n = c.bind_coupling_power(cp, cell, f, dofs_c);
        const int ncl = lc.n;
        for (int i = 0; i < ncl; ++i) { lc.idx[i] = i; x_c[i] = x[dofs_c[i]]; }
        std::vector<double> a_cc(ncl * ncl, 0.0), a_sc(ns * ncl, 0.0), a_cs(ncl * ns, 0.0), a_nc(nn * ncl, 0.0), a_cn(ncl * nn, 0.0);
        LocalMatrixView m_cc{a_cc.data(), ncl, weight}, m_sc{a_sc.data(), ncl, weight}, m_cs{a_cs.data(), ns, weight},
                        m_nc{a_nc.data(), ncl, weight}, m_cn{a_cn.data(), nn, weight};
        cp.eop->jacobian_generic(+1, ig, ls, S.x_s, ns, ls, lc, x_c, ncl, lc, m_ss, m_sc, m_cs, m_cc);
        cp.eop->jacobian_generic(-1, ig, ln, S.x_n, nn, ln, lc, x_c, ncl, lc, m_nn, m_nc, m_cn, m_cc);
        // onUnbindLFSUVCoupling: cc, sc, cs, nc, cn (jacobianassemblerengine.hh:332-377)
        scatter(A, dofs_c, ncl, dofs_c, ncl, a_cc.data(), true);
        scatter(A, S.dofs_s, ns, dofs_c, ncl, a_sc.data(), true);
        scatter(A, dofs_c, ncl, S.dofs_s, ns, a_cs.data(), true);
        scatter(A, S.dofs_n, nn, dofs_c, ncl, a_nc.data(), true);
        scatter(A, dofs_c, ncl, S.dofs_n, nn, a_cn.data(), true);
      }
      // onUnbindLFSUVOutside: nn, sn, ns (jacobianassemblerengine.hh:296-330)
      scatter(A, S.dofs_n, nn, S.dofs_n, nn, S.a_nn.data(), true);
      scatter(A, S.dofs_s, ns, S.dofs_n, nn, S.a_sn.data(), true);
      scatter(A, S.dofs_n, nn, S.dofs_s, ns, S.a_ns.data(), true);
    }
  }
  scatter(A, S.dofs_s, ns, S.dofs_s, ns, S.a_ss.data(), true);   // onUnbindLFSUV
}

void dirichlet_rows(const Context& c, CSR& A)
{
  // [UPSTREAM] (vi) handle_dirichlet_constraints / set_trivial_rows: row <- 0, diagonal <- 1
  for (int64_t r = 0; r < c.ndof; ++r) if (c.constrained[r]) {
    for (int64_t p = A.rowptr[r]; p < A.rowptr[r + 1]; ++p) A.val[p] = (A.colidx[p] == r) ? 1.0 : 0.0;
  }
}
} // namespace

void Context::jacobian(int goi, int mat, double weight, const double* x)
{
  const GridOperatorDesc& go = gos.at(goi);
  CSR& A = mats.at(mat);
  Flags fl = engine_flags(*this, go, JACOBIAN);
  JacScratch S((int)children.size());
  for (int64_t cell = 0; cell < grid.ncells(); ++cell) jacobian_cell(*this, go, fl, A, weight, x, cell, S);
  dirichlet_rows(*this, A);
}

void Context::jacobian_parallel(int goi, int mat, double weight, const double* x, int nthreads)
{
  // CPU-baseline variant: slabs along the slowest axis, two-coloured (even slabs, then odd slabs) so that
  // concurrently processed cells never share a matrix row.  Same arithmetic per cell as jacobian().
  const GridOperatorDesc& go = gos.at(goi);
  CSR& A = mats.at(mat);
  Flags fl = engine_flags(*this, go, JACOBIAN);
  const int ax = grid.dim - 1;
  const int64_t nslow = grid.n[ax];
  const int64_t per = grid.ncells() / nslow;
  if (nthreads < 1) nthreads = 1;
  // slab thickness >= 2 layers so that coupling faces (which touch the neighbour cell's rows) stay inside a colour gap
  int64_t nslabs = std::min<int64_t>(nslow / 2, (int64_t)nthreads * 2);
  if (nslabs < 2) { jacobian(goi, mat, weight, x); return; }
  if (nslabs % 2) nslabs -= 1;
  auto slab_begin = [&](int64_t s) { return (nslow * s) / nslabs; };
  for (int colour = 0; colour < 2; ++colour) {
    std::vector<std::thread> th;
    std::vector<std::string> errs(nslabs);
    for (int64_t s = colour; s < nslabs; s += 2) {
      th.emplace_back([&, s]() {
        try {
          JacScratch S((int)children.size());
          for (int64_t cell = slab_begin(s) * per; cell < slab_begin(s + 1) * per; ++cell)
            jacobian_cell(*this, go, fl, A, weight, x, cell, S);
        } catch (std::exception& e) { errs[s] = e.what(); }
      });
    }
    for (auto& t : th) t.join();
    for (auto& e : errs) if (!e.empty()) throw std::runtime_error(e);
  }
  dirichlet_rows(*this, A);
}

// =============================================================================================
// constraints (constraints.hh:312-558, constraintsfunctors.hh:41-131; [UPSTREAM] ConformingDirichletConstraints:
// predicate evaluated at the face centre, all DOFs on the face are constrained — SURVEY Appendix B)
// =============================================================================================
void Context::constraints_assemble(const int* sps_, const mdb_bctype* bcs, int n)
{
  const int dim = grid.dim;
  const int nch = (int)children.size();
  std::vector<int> off_s(nch + 1);
  int64_t dofs_s[MAXCELL];
  std::fill(constrained.begin(), constrained.end(), 0);
  for (int64_t cell = 0; cell < grid.ncells(); ++cell) {
    EG eg = make_eg(*this, cell);
    bind(cell, off_s.data(), dofs_s);
    for (int f = 0; f < 2 * dim; ++f) {
      int64_t nijk[3];
      bool interior = FaceIter::neighbor(grid, eg.ijk, f, nijk);
      if (!interior && grid.processor_face(eg.ijk, f)) continue;
      uint8_t nsd = interior ? cell_sd[grid.cell_index(nijk)] : 0;
      IG ig; ig.inside = eg; ig.outside = eg; ig.axis = f / 2; ig.side = f % 2; ig.face = f; ig.boundary = !interior;
      for (int q = 0; q < n; ++q) {
        const SubProblem& sp = sps.at(sps_[q]);
        if (!sp.appliesTo(eg.sd)) continue;
        // constraintsfunctors.hh:77-103: skeleton constraints if the subproblem also lives outside (none for
        // conforming spaces), otherwise the *boundary* visitor runs, also on subdomain interfaces
        if (interior && sp.appliesTo(nsd)) continue;
        double centre[2] = {0.5, 0.5}, xg[3];
        ig.global(centre, xg);
        if (eval_bctype(bcs[q], dim, xg, ig.boundary) != 0) continue;
        LFS l = subproblem_lfs(sp, off_s.data());
        if (l.dg) continue;          // [UPSTREAM] ConformingDirichletConstraints skips coefficients with localKey().codim() == 0
        const int k = l.k;
        for (int i = 0; i < l.n; ++i) {
          int rem = i, a[3];
          for (int d = 0; d < 3; ++d) { a[d] = (d < dim) ? rem % (k + 1) : 0; rem /= (k + 1); }
          if (a[ig.axis] == (ig.side ? k : 0)) constrained[dofs_s[l.idx[i]]] = 1;
        }
      }
    }
  }
}

// interpolate.hh:35-73: one grid loop; for every subproblem that applies, Lagrange interpolation of f
void Context::interpolate(const int* sps_, const mdb_function* fns, int n, double t, double* x) const
{
  const int dim = grid.dim;
  const int nch = (int)children.size();
  std::vector<int> off_s(nch + 1);
  int64_t dofs_s[MAXCELL];
  for (int64_t cell = 0; cell < grid.ncells(); ++cell) {
    EG eg = make_eg(*this, cell);
    bind(cell, off_s.data(), dofs_s);
    for (int q = 0; q < n; ++q) {
      const SubProblem& sp = sps.at(sps_[q]);
      if (!sp.appliesTo(eg.sd)) continue;
      LFS l = subproblem_lfs(sp, off_s.data());
      const int k = l.k;
      for (int i = 0; i < l.n; ++i) {
        int rem = i; double local[3], xg[3];
        for (int d = 0; d < 3; ++d) { int a = (d < dim) ? rem % (k + 1) : 0; rem /= (k + 1); local[d] = double(a) / k; }
        eg.global(local, xg);
        x[dofs_s[l.idx[i]]] = eval_function(fns[q], dim, xg, t);
      }
    }
  }
}

// =============================================================================================
// linear algebra  ([UPSTREAM] dune-istl: MatrixAdapter, SeqJac, SeqSSOR, SeqILU0, CGSolver, BiCGSTABSolver)
// =============================================================================================
void spmv(const CSR& A, const double* x, double* y)
{
  for (int64_t r = 0; r < A.nrows; ++r) {
    double s = 0.0;
    for (int64_t p = A.rowptr[r]; p < A.rowptr[r + 1]; ++p) s += A.val[p] * x[A.colidx[p]];
    y[r] = s;
  }
}

namespace {
struct Precond {
  const CSR& A; int kind; int sweeps;
  std::vector<double> ilu; std::vector<int64_t> diag;
  Precond(const CSR& A_, int kind_, int sweeps_) : A(A_), kind(kind_), sweeps(sweeps_) {
    diag.resize(A.nrows);
    for (int64_t r = 0; r < A.nrows; ++r) { diag[r] = A.find(r, r); if (diag[r] < 0) throw std::runtime_error("missing diagonal"); }
    if (kind == MDB_PRECOND_ILU0) {          // [UPSTREAM] ilu.hh bilu0_decomposition on the pattern of A: L_ik = a_ik * (a_kk)^-1 with
      ilu = A.val;                           // the pivot of every finished row stored INVERTED ((*ij).rightmultiply(*jj), (*ij).invert())
      for (int64_t i = 0; i < A.nrows; ++i) {
        for (int64_t p = A.rowptr[i]; p < diag[i]; ++p) {
          int64_t k = A.colidx[p];
          ilu[p] = ilu[p] * ilu[diag[k]];
          double lik = ilu[p];
          int64_t q = diag[k] + 1, qe = A.rowptr[k + 1];
          int64_t s = p + 1, se = A.rowptr[i + 1];
          while (q < qe && s < se) {
            if (A.colidx[q] == A.colidx[s]) { ilu[s] -= lik * ilu[q]; ++q; ++s; }
            else if (A.colidx[q] < A.colidx[s]) ++q; else ++s;
          }
        }
        ilu[diag[i]] = 1.0 / ilu[diag[i]];
      }
    }
  }
  // v = M^{-1} d   (v starts at zero, as the ISTL solvers do)
  void apply(double* v, const double* d) const {
    const int64_t n = A.nrows;
    std::fill(v, v + n, 0.0);
    if (kind == MDB_PRECOND_NONE) { std::copy(d, d + n, v); return; }
    if (kind == MDB_PRECOND_JACOBI) { for (int64_t i = 0; i < n; ++i) v[i] = d[i] / A.val[diag[i]]; return; }
    if (kind == ORC_PRECOND_SSOR) {          // SeqSSOR(A,n,1.0): n x (bsorf, bsorb), (ix)
      for (int s = 0; s < sweeps; ++s) {
        for (int64_t i = 0; i < n; ++i) {
          double rhs = d[i];
          for (int64_t p = A.rowptr[i]; p < A.rowptr[i + 1]; ++p) rhs -= A.val[p] * v[A.colidx[p]];
          v[i] += rhs / A.val[diag[i]];
        }
        for (int64_t i = n - 1; i >= 0; --i) {
          double rhs = d[i];
          for (int64_t p = A.rowptr[i]; p < A.rowptr[i + 1]; ++p) rhs -= A.val[p] * v[A.colidx[p]];
          v[i] += rhs / A.val[diag[i]];
        }
      }
      return;
    }
    if (kind == MDB_PRECOND_ILU0) {          // bilu_backsolve
      for (int64_t i = 0; i < n; ++i) {
        double s = d[i];
        for (int64_t p = A.rowptr[i]; p < diag[i]; ++p) s -= ilu[p] * v[A.colidx[p]];
        v[i] = s;
      }
      for (int64_t i = n - 1; i >= 0; --i) {
        double s = v[i];
        for (int64_t p = diag[i] + 1; p < A.rowptr[i + 1]; ++p) s -= ilu[p] * v[A.colidx[p]];
        v[i] = ilu[diag[i]] * s;            // the diagonal holds the inverse pivot
      }
      return;
    }
    throw std::runtime_error("unknown preconditioner");
  }
};
double dot(int64_t n, const double* a, const double* b) { double s = 0; for (int64_t i = 0; i < n; ++i) s += a[i] * b[i]; return s; }
} // namespace

SolveStats solve(const CSR& A, int method, int precond, int sweeps, double reduction, int maxit, const double* b, double* x)
{
  const int64_t n = A.nrows;
  Precond M(A, precond, sweeps);
  SolveStats st;
  std::vector<double> r(b, b + n);
  std::fill(x, x + n, 0.0);
  st.defect0 = std::sqrt(dot(n, r.data(), r.data()));
  st.defect = st.defect0;
  if (st.defect0 == 0.0) { st.converged = true; return st; }
  if (method == MDB_SOLVER_CG) {
    std::vector<double> p(n, 0.0), q(n);
    double rholast = 1.0;
    for (int it = 1; it <= maxit; ++it) {
      M.apply(q.data(), r.data());
      double rho = dot(n, q.data(), r.data());
      double beta = rho / rholast;
      for (int64_t i = 0; i < n; ++i) p[i] = p[i] * beta + q[i];
      rholast = rho;
      spmv(A, p.data(), q.data());
      double alpha = dot(n, p.data(), q.data());
      double lambda = rho / alpha;
      for (int64_t i = 0; i < n; ++i) { x[i] += lambda * p[i]; r[i] -= lambda * q[i]; }
      st.defect = std::sqrt(dot(n, r.data(), r.data()));
      st.iterations = it;
      if (st.defect <= reduction * st.defect0) { st.converged = true; break; }
    }
    return st;
  }
  // BiCGSTAB
  std::vector<double> rt(r), p(n, 0.0), v(n, 0.0), y(n), t(n);
  double rho = 1, alpha = 1, omega = 1;
  for (int it = 1; it <= maxit; ++it) {
    double rho_new = dot(n, rt.data(), r.data());
    if (it == 1) p = r;
    else {
      double beta = (rho_new / rho) * (alpha / omega);
      for (int64_t i = 0; i < n; ++i) p[i] = r[i] + beta * (p[i] - omega * v[i]);
    }
    M.apply(y.data(), p.data());
    spmv(A, y.data(), v.data());
    double hh = dot(n, rt.data(), v.data());
    alpha = rho_new / hh;
    for (int64_t i = 0; i < n; ++i) { x[i] += alpha * y[i]; r[i] -= alpha * v[i]; }
    st.defect = std::sqrt(dot(n, r.data(), r.data()));
    st.iterations = it;
    if (st.defect <= reduction * st.defect0) { st.converged = true; break; }
    M.apply(y.data(), r.data());
    spmv(A, y.data(), t.data());
    omega = dot(n, t.data(), r.data()) / dot(n, t.data(), t.data());
    for (int64_t i = 0; i < n; ++i) { x[i] += omega * y[i]; r[i] -= omega * t[i]; }
    rho = rho_new;
    st.defect = std::sqrt(dot(n, r.data(), r.data()));
    if (st.defect <= reduction * st.defect0) { st.converged = true; break; }
  }
  return st;
}

} // namespace orc

// =============================================================================================
// C interface for ctypes (tests / bench cpu_baseline only); mirrors the shape of include/mdb200.h
// =============================================================================================
using orc::Context;
#define ORC_TRY(ctx, ...) try { __VA_ARGS__; return 0; } catch (std::exception& e) { (ctx)->err = e.what(); return -1; }

extern "C" {

Context* orc_create() { return new Context(); }
void orc_destroy(Context* c) { delete c; }
const char* orc_last_error(Context* c) { return c->err.c_str(); }

int orc_mesh_structured(Context* c, int dim, const int64_t* n, const double* lower, const double* upper)
{
  ORC_TRY(c, {
    if (dim < 1 || dim > 3) throw std::runtime_error("dim");
    c->grid = orc::Grid(); c->grid.dim = dim;
    for (int d = 0; d < dim; ++d) { c->grid.n[d] = n[d]; c->grid.lower[d] = lower[d]; c->grid.upper[d] = upper[d];
                                    c->grid.h[d] = (upper[d] - lower[d]) / n[d]; }
    c->cell_sd.clear(); c->finalized = false;
  })
}
int orc_mesh_partition(Context* c, const int64_t* global_n, const int64_t* cell_offset, const double* global_lower, const double* global_upper)
{
  ORC_TRY(c, {
    orc::Grid& g = c->grid;
    for (int d = 0; d < g.dim; ++d) {
      g.gn[d] = global_n[d]; g.off[d] = cell_offset[d]; g.glower[d] = global_lower[d];
      g.h[d] = (global_upper[d] - global_lower[d]) / global_n[d];
    }
    g.partitioned = true;
  })
}
int orc_subdomains(Context* c, const uint8_t* bits)
{
  ORC_TRY(c, { c->cell_sd.assign(bits, bits + c->grid.ncells()); })
}
int orc_get_subdomains(Context* c, uint8_t* bits) { ORC_TRY(c, { std::copy(c->cell_sd.begin(), c->cell_sd.end(), bits); }) }
int orc_subdomains_halfspace(Context* c, int axis, double threshold, int sd_below, int sd_above)
{
  ORC_TRY(c, {
    c->cell_sd.resize(c->grid.ncells());
    for (int64_t cell = 0; cell < c->grid.ncells(); ++cell) {
      int64_t ijk[3]; c->grid.cell_ijk(cell, ijk);
      // it->geometry().center()[axis] > threshold  (testpoisson.cc:154); [UPSTREAM] centre = lower + 0.5*(upper-lower)
      double lo = c->grid.coord(axis, ijk[axis]), up = c->grid.coord(axis, ijk[axis] + 1);
      double centre = lo + 0.5 * (up - lo);
      c->cell_sd[cell] = uint8_t(1u << (centre > threshold ? sd_above : sd_below));
    }
  })
}
int orc_add_space(Context* c, int fe_kind, int subdomain)
{
  try {
    orc::ChildSpace ch; ch.fe_kind = fe_kind; ch.subdomain = subdomain;
    ch.k = (fe_kind == MDB_FE_Q1 || fe_kind == MDB_FE_DGQ1) ? 1 : (fe_kind == MDB_FE_Q2 || fe_kind == MDB_FE_DGQ2) ? 2 : -1;
    ch.dg = fe_kind == MDB_FE_DGQ1 || fe_kind == MDB_FE_DGQ2;
    if (ch.k < 0) throw std::runtime_error("unknown fe kind");
    c->children.push_back(ch);
    return (int)c->children.size() - 1;
  } catch (std::exception& e) { c->err = e.what(); return -1; }
}
int orc_add_subproblem(Context* c, int op_id, const void* params, size_t nbytes, int cond_kind, uint32_t mask,
                       const int* children, int nchildren)
{
  try {
    orc::SubProblem sp; sp.lop = orc::make_local_operator(op_id, params, nbytes);
    sp.cond_kind = cond_kind; sp.mask = mask; sp.children.assign(children, children + nchildren);
    c->sps.push_back(std::move(sp));
    return (int)c->sps.size() - 1;
  } catch (std::exception& e) { c->err = e.what(); return -1; }
}
int orc_add_coupling(Context* c, int op_id, const void* params, size_t nbytes, int local_sp, int remote_sp, int jac_mode)
{
  try {
    orc::Coupling cp; cp.op = orc::make_coupling_operator(op_id, params, nbytes, jac_mode);
    cp.local_sp = local_sp; cp.remote_sp = remote_sp;
    c->cps.push_back(std::move(cp));
    return (int)c->cps.size() - 1;
  } catch (std::exception& e) { c->err = e.what(); return -1; }
}
int orc_add_coupling_space(Context* c, int fe_kind, int lkind, uint32_t lmask, int rkind, uint32_t rmask)
{
  try {
    if (fe_kind != MDB_FE_Q1) throw std::runtime_error("coupling spaces: Q1 only");
    orc::ChildSpace ch; ch.fe_kind = fe_kind; ch.subdomain = -1; ch.k = 1;
    ch.iface = true; ch.lkind = lkind; ch.lmask = lmask; ch.rkind = rkind; ch.rmask = rmask;
    c->children.push_back(ch);
    return (int)c->children.size() - 1;
  } catch (std::exception& e) { c->err = e.what(); return -1; }
}
int orc_add_enriched_coupling(Context* c, int op_id, const void* params, size_t nbytes, int local_sp, int remote_sp,
                              int coupling_child, int jac_mode)
{
  try {
    if (op_id != MDB_OP_MORTAR_POISSON_COUPLING || nbytes != sizeof(mdb_mortar_params)) throw std::runtime_error("unknown enriched coupling operator");
    if (coupling_child < 0 || coupling_child >= (int)c->children.size() || !c->children[coupling_child].iface)
      throw std::runtime_error("coupling_child is not a coupling space");
    orc::Coupling cp; cp.eop.reset(new orc::EnrichedOperator());
    cp.eop->gamma_0 = static_cast<const mdb_mortar_params*>(params)->gamma_0; cp.eop->jac_mode = jac_mode;
    cp.local_sp = local_sp; cp.remote_sp = remote_sp; cp.coupling_child = coupling_child;
    c->cps.push_back(std::move(cp));
    return (int)c->cps.size() - 1;
  } catch (std::exception& e) { c->err = e.what(); return -1; }
}
int orc_add_coupling_space(Context* c, int fe_kind, int lkind, uint32_t lmask, int rkind, uint32_t rmask);
int orc_add_power_coupling_space(Context* c, int fe_kind, int lkind, uint32_t lmask, int rkind, uint32_t rmask, int ncomp)
{
  if (ncomp < 1 || ncomp > 2) { c->err = "power coupling spaces: 1 or 2 components"; return -1; }
  int first = -1;
  for (int k = 0; k < ncomp; ++k) { int ch = orc_add_coupling_space(c, fe_kind, lkind, lmask, rkind, rmask); if (ch < 0) return ch; if (k == 0) first = ch; }
  return first;
}
int orc_add_enriched_coupling(Context* c, int op_id, const void* params, size_t nbytes, int local_sp, int remote_sp, int coupling_child, int jac_mode);
int orc_add_power_enriched_coupling(Context* c, int op_id, const void* params, size_t nbytes, int local_sp, int remote_sp, int first_child,
                                    int ncomp, int jac_mode)
{
  if (ncomp < 1 || ncomp > 2 || first_child < 0 || first_child + ncomp > (int)c->children.size()) { c->err = "bad power coupling space"; return -1; }
  for (int k = 0; k < ncomp; ++k) if (!c->children[first_child + k].iface) { c->err = "bad power coupling space"; return -1; }
  int h = orc_add_enriched_coupling(c, op_id, params, nbytes, local_sp, remote_sp, first_child, jac_mode);
  if (h < 0) return h;
  c->cps[h].ncomp = ncomp; c->cps[h].eop->ncomp = ncomp;
  return h;
}
int orc_finalize(Context* c, int64_t* ndof) { ORC_TRY(c, { c->finalize(); *ndof = c->ndof; }) }
int orc_get_dofmap(Context* c, int64_t* cell_ptr, int64_t* cell_dofs)
{
  ORC_TRY(c, {
    std::vector<int> off(c->children.size() + 1); int64_t dofs[64]; int64_t pos = 0;
    for (int64_t cell = 0; cell < c->grid.ncells(); ++cell) {
      int n = c->bind(cell, off.data(), dofs);
      cell_ptr[cell] = pos;
      if (cell_dofs) for (int i = 0; i < n; ++i) cell_dofs[pos + i] = dofs[i];
      pos += n;
    }
    cell_ptr[c->grid.ncells()] = pos;
  })
}
int orc_get_child_offsets(Context* c, int64_t* offsets)
{
  ORC_TRY(c, { for (size_t i = 0; i < c->children.size(); ++i) offsets[i] = c->children[i].offset; offsets[c->children.size()] = c->ndof; })
}
int orc_constraints_assemble(Context* c, const int* sps, const mdb_bctype* bcs, int n, int64_t* ncon)
{
  ORC_TRY(c, { c->constraints_assemble(sps, bcs, n); int64_t k = 0; for (auto f : c->constrained) k += f; *ncon = k; })
}
int orc_set_constraints(Context* c, int64_t n, const int64_t* dofs)
{
  ORC_TRY(c, { std::fill(c->constrained.begin(), c->constrained.end(), 0); for (int64_t i = 0; i < n; ++i) c->constrained.at(dofs[i]) = 1; })
}
int orc_get_constraints(Context* c, uint8_t* flags) { ORC_TRY(c, { std::copy(c->constrained.begin(), c->constrained.end(), flags); }) }
int orc_interpolate(Context* c, const int* sps, const mdb_function* fns, int n, double t, double* x) { ORC_TRY(c, { c->interpolate(sps, fns, n, t, x); }) }
int orc_gridoperator_create(Context* c, const int* sps, int nsp, const int* cps, int ncp)
{
  try { orc::GridOperatorDesc g; g.sps.assign(sps, sps + nsp); g.cps.assign(cps, cps + ncp); c->gos.push_back(g); return (int)c->gos.size() - 1; }
  catch (std::exception& e) { c->err = e.what(); return -1; }
}
int orc_matrix_create(Context* c, const int* gos, int ngos, int64_t* nnz)
{
  try { int m = c->matrix_create(gos, ngos); *nnz = (int64_t)c->mats[m].colidx.size(); return m; }
  catch (std::exception& e) { c->err = e.what(); return -1; }
}
int orc_matrix_get_pattern(Context* c, int mat, int64_t* rowptr, int64_t* colidx)
{
  ORC_TRY(c, { auto& A = c->mats.at(mat); std::copy(A.rowptr.begin(), A.rowptr.end(), rowptr); std::copy(A.colidx.begin(), A.colidx.end(), colidx); })
}
int orc_matrix_get_values(Context* c, int mat, double* v) { ORC_TRY(c, { auto& A = c->mats.at(mat); std::copy(A.val.begin(), A.val.end(), v); }) }
int orc_matrix_zero(Context* c, int mat) { ORC_TRY(c, { auto& A = c->mats.at(mat); std::fill(A.val.begin(), A.val.end(), 0.0); }) }
int orc_set_time(Context* c, double t) { c->time = t; return 0; }
// copy_nonconstrained_dofs(cu, xold, xnew)  (gridoperator.hh:146)
int orc_copy_nonconstrained(Context* c, const double* xold, double* xnew)
{
  ORC_TRY(c, { for (int64_t i = 0; i < c->ndof; ++i) if (!c->constrained[i]) xnew[i] = xold[i]; })
}
int orc_residual(Context* c, int go, double weight, const double* x, double* r) { ORC_TRY(c, { c->residual(go, weight, x, r); }) }
int orc_jacobian(Context* c, int go, int mat, double weight, const double* x) { ORC_TRY(c, { c->jacobian(go, mat, weight, x); }) }
int orc_jacobian_parallel(Context* c, int go, int mat, double weight, const double* x, int nthreads)
{ ORC_TRY(c, { c->jacobian_parallel(go, mat, weight, x, nthreads); }) }
int orc_spmv(Context* c, int mat, const double* x, double* y) { ORC_TRY(c, { orc::spmv(c->mats.at(mat), x, y); }) }
// ISTL_SEQ_Subblock_Backend<Preconditioner, Solver>::apply (test/testpoisson-dirichlet-neumann.cc:264-370): the solver runs on
// the diagonal block raw(A)[block][block] and the matching blocks of z and r only; every other entry of z is left untouched.
// The blocks of the flat vectors are the child blocks of the MultiDomainGridFunctionSpace.
int orc_solve_block(Context* c, int mat, int child, int method, int precond, int sweeps, double reduction, int maxit, const double* b, double* x,
                    mdb_solve_stats* stats)
{
  ORC_TRY(c, {
    const orc::CSR& A = c->mats.at(mat);
    const orc::ChildSpace& ch = c->children.at(child);
    const int64_t lo = ch.offset, n = ch.size;
    orc::CSR B; B.nrows = n; B.rowptr.assign(n + 1, 0);
    for (int64_t r = 0; r < n; ++r) {
      for (int64_t p = A.rowptr[lo + r]; p < A.rowptr[lo + r + 1]; ++p)
        if (A.colidx[p] >= lo && A.colidx[p] < lo + n) { B.colidx.push_back(A.colidx[p] - lo); B.val.push_back(A.val[p]); }
      B.rowptr[r + 1] = (int64_t)B.colidx.size();
    }
    orc::SolveStats s = orc::solve(B, method, precond, sweeps, reduction, maxit, b + lo, x + lo);
    if (stats) { stats->iterations = s.iterations; stats->converged = s.converged; stats->defect0 = s.defect0; stats->defect = s.defect; stats->seconds = 0; }
  })
}
int orc_solve(Context* c, int mat, int method, int precond, int sweeps, double reduction, int maxit, const double* b, double* x,
              mdb_solve_stats* stats)
{
  ORC_TRY(c, {
    orc::SolveStats s = orc::solve(c->mats.at(mat), method, precond, sweeps, reduction, maxit, b, x);
    if (stats) { stats->iterations = s.iterations; stats->converged = s.converged; stats->defect0 = s.defect0; stats->defect = s.defect; stats->seconds = 0; }
  })
}
// element-level probes for known-answer tests: local matrix / residual of one operator on one cell
int orc_local_jacobian_volume(Context* c, int sp, int64_t cell, double* a /* nloc*nloc row-major */)
{
  ORC_TRY(c, {
    std::vector<int> off(c->children.size() + 1); int64_t dofs[64];
    int n = c->bind(cell, off.data(), dofs);
    orc::EG eg; eg.g = &c->grid; eg.index = cell; c->grid.cell_ijk(cell, eg.ijk); eg.sd = c->cell_sd[cell];
    orc::LFS l = c->subproblem_lfs(c->sps.at(sp), off.data());
    std::vector<double> full(n * n, 0.0), x(n, 0.0);
    orc::LocalMatrixView m{full.data(), n, 1.0};
    c->sps[sp].lop->jacobian_volume(eg, l, x.data(), l, m);
    for (int i = 0; i < l.n; ++i) for (int j = 0; j < l.n; ++j) a[i * l.n + j] = full[l.idx[i] * n + l.idx[j]];
  })
}

} // extern "C"
