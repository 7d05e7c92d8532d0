// examples/testpoisson-dirichlet-neumann.cc — the reference driver test/testpoisson-dirichlet-neumann.cc (`solve()` :371-862, main
// :883-977) written against the B200 facade.  main() picks the instantiation of solve() like the reference's (commented)
// alternatives :964-975:
//
//   testpoisson-dirichlet-neumann <ini file> [fem left = dg2] [fem right = dg2]
//       dg1 / dg2   QkDGLocalFiniteElementMap<DF,double,k,dim> + dglop() = ConvectionDiffusionDG(params, SIPG, weightsOn, 2.0)  — dg2 dg2 is
//                   what the reference ships (:968-971)
//       1 / 2       QkLocalFiniteElementMap<SDGV,DF,double,k> + cglop = Poisson(f,b,j,6) (:955-956, the commented alternative)
//
// Monolithic solve: left_sp, right_sp, Coupling(ContinuousValueContinuousFlowCoupling(4, scaling)), BiCGSTAB + SSOR(3) (:619-665).
// Dirichlet-Neumann solve (:683-822): LeftOperator {left_sp, Coupling(left,right,ND)} and RightOperator {right_sp,
// Coupling(right,left,ND)} solved alternately on their diagonal blocks (ISTL_SEQ_Subblock_Backend) with damping, stopping on the
// residual of the FullOperator.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <map>
#include <string>

#include "mdb200/multidomain.hh"

using namespace mdb200;

// Dune::ParameterTreeParser::readINITree: "[section]" prefixes, "key = value", '#' comments
struct ParameterTree {
  std::map<std::string, std::string> kv;
  static std::string trim(const std::string& s) {
    size_t a = s.find_first_not_of(" \t\r"), b = s.find_last_not_of(" \t\r");
    return a == std::string::npos ? std::string() : s.substr(a, b - a + 1);
  }
  explicit ParameterTree(const char* file) {
    std::ifstream in(file);
    if (!in) throw Exception(MDB_EINVAL, std::string("cannot open ini file ") + file);
    std::string line, prefix;
    while (std::getline(in, line)) {
      line = trim(line.substr(0, line.find('#')));
      if (line.empty()) continue;
      if (line[0] == '[') { prefix = trim(line.substr(1, line.find(']') - 1)); continue; }
      const size_t eq = line.find('=');
      if (eq == std::string::npos) continue;
      kv[(prefix.empty() ? "" : prefix + ".") + trim(line.substr(0, eq))] = trim(line.substr(eq + 1));
    }
  }
  const std::string& operator[](const std::string& key) const {
    std::map<std::string, std::string>::const_iterator it = kv.find(key);
    if (it == kv.end()) throw Exception(MDB_EINVAL, "missing ini key " + key);
    return it->second;
  }
  double get(const std::string& key) const { return std::atof((*this)[key].c_str()); }
  bool flag(const std::string& key) const { const std::string& v = (*this)[key]; return v == "true" || v == "1" || v == "yes"; }
};

static CouplingMode::Type parse_coupling_type(const std::string& name) {      // :490-497
  if (name == "Dirichlet") return CouplingMode::Dirichlet;
  if (name == "Neumann") return CouplingMode::Neumann;
  throw Exception(MDB_EINVAL, "Unknown coupling type " + name);
}

static double norm2(const Vector& v) { double s = 0.0; for (size_t i = 0; i < v.size(); ++i) s += v[i] * v[i]; return std::sqrt(s); }

template <class LOP>
static int solve(const ParameterTree& parameters, int k0, int k1, bool dg, const LOP& lop)
{
    const int dim = 2;
    const int64_t cells = (int64_t)1 << (int)parameters.get("mesh.refine");
    const int64_t n[3] = {cells, cells, 1};
    const double lower[3] = {0, 0, 0}, upper[3] = {1, 1, 1};

    Context ctx(0);
    MultiDomainGrid grid(ctx, dim, n, lower, upper);
    grid.startSubDomainMarking();
    for (int64_t e = 0; e < grid.size0(); ++e)                        // center()[0] > 0.5 -> 1 else 0 (:917-923)
      grid.addToSubDomain(((e % cells) + 0.5) / cells > 0.5 ? 1 : 0, e);
    grid.preUpdateSubDomains(); grid.updateSubDomains(); grid.postUpdateSubDomains();

    GridFunctionSpace gfs0(grid.subDomain(0), k0, dg), gfs1(grid.subDomain(1), k1, dg);
    MultiDomainGridFunctionSpace multigfs(grid, {&gfs0, &gfs1});

    const mdb_bctype b = boundary(MDB_BC_TESTPOISSON);                // DirichletBoundary :89-135
    const mdb_function g = function(MDB_FN_GAUSS, 1.0, 0.5, 0.5, 0.5);

    SubDomainEqualityCondition c0({0}), c1({1});
    typedef SubProblem<LOP, SubDomainEqualityCondition> SP;
    SP left_sp(multigfs, lop, c0, gfs0), right_sp(multigfs, lop, c1, gfs1);

    ContinuousValueContinuousFlowCoupling proportionalFlowCoupling(4, parameters.get("monolithic.coupling.scaling"));
    Coupling<SP, SP, ContinuousValueContinuousFlowCoupling> coupling(left_sp, right_sp, proportionalFlowCoupling);

    typedef ConvectionDiffusionDGNeumannDirichletCoupling NDCoupling;
    NDCoupling nd_coupling_neumann(parse_coupling_type(parameters["dn.coupling.left"]), ConvectionDiffusionDGMethod::SIPG,
                                   ConvectionDiffusionDGWeights::weightsOn, parameters.get("dn.coupling.alpha"));       // :502-508
    NDCoupling nd_coupling_dirichlet(parse_coupling_type(parameters["dn.coupling.right"]), ConvectionDiffusionDGMethod::SIPG,
                                     ConvectionDiffusionDGWeights::weightsOn, parameters.get("dn.coupling.alpha"));     // :510-516
    Coupling<SP, SP, NDCoupling> dirichlet_coupling(left_sp, right_sp, nd_coupling_dirichlet);                        // :518-528
    Coupling<SP, SP, NDCoupling> neumann_coupling(right_sp, left_sp, nd_coupling_neumann);                            // :530-540

    ConstraintsContainer cg;
    constraints(multigfs).constrainSubProblem(left_sp, b).constrainSubProblem(right_sp, b).assemble(cg);

    Vector u;
    interpolateOnTrialSpace(multigfs, u, g, left_sp, g, right_sp);
    std::printf("%lld dof total, %lld dof constrained\n", (long long)u.size(), (long long)cg.size());

    GridOperator gridoperator(multigfs, multigfs, cg, cg, {left_sp.handle, right_sp.handle}, {coupling.handle});
    GridOperator full_operator(multigfs, multigfs, cg, cg, {left_sp.handle, right_sp.handle}, {dirichlet_coupling.handle, neumann_coupling.handle});
    GridOperator left_operator(multigfs, multigfs, cg, cg, {left_sp.handle}, {dirichlet_coupling.handle});
    GridOperator right_operator(multigfs, multigfs, cg, cg, {right_sp.handle}, {neumann_coupling.handle});

    if (parameters.flag("monolithic.solve")) {
      ISTLBackend_SEQ_BCGS_SSOR ls((unsigned)parameters.get("monolithic.linearsolver.iterations"));
      Vector u2(u);
      StationaryLinearProblemSolver<ISTLBackend_SEQ_BCGS_SSOR> pdesolver(gridoperator, ls, parameters.get("monolithic.solver.reduction"));
      pdesolver.apply(u2);
      double sum = 0.0; for (size_t i = 0; i < u2.size(); ++i) sum += u2[i];
      std::printf("Monolithic solver: %d iterations, reduction %.3e, sum(u) = %.12e |u|_2 = %.12e\n", ls.res.iterations, ls.res.reduction, sum, norm2(u2));
    }

    if (parameters.flag("dn.solve")) {
      ISTL_SEQ_Subblock_Backend linear_solver_dn_0(gfs0, (unsigned)parameters.get("dn.left.linearsolver.iterations"));
      ISTL_SEQ_Subblock_Backend linear_solver_dn_1(gfs1, (unsigned)parameters.get("dn.right.linearsolver.iterations"));
      StationaryLinearProblemSolver<ISTL_SEQ_Subblock_Backend> left_pde_solver(left_operator, linear_solver_dn_0, parameters.get("dn.left.solver.reduction"));
      StationaryLinearProblemSolver<ISTL_SEQ_Subblock_Backend> right_pde_solver(right_operator, linear_solver_dn_1, parameters.get("dn.right.solver.reduction"));

      Vector residual(u.size(), 0.0);
      full_operator.residual(u, residual);
      double r = norm2(residual); const double start_r = r;
      std::printf("Start residual: %.12e\n", r);

      size_t i = 0; const size_t max_iterations = (size_t)parameters.get("dn.coupling.iterations");
      const double alpha = parameters.get("dn.coupling.damping"), rel_reduction = parameters.get("dn.coupling.reduction");
      const double max_error = parameters.get("dn.coupling.maxerror");
      const bool reassemble_left = parameters.flag("dn.left.reassemble"), reassemble_right = parameters.flag("dn.right.reassemble");
      const double l_max = parameters.get("dn.left.solver.reduction"), l_min = parameters.get("dn.left.solver.minreduction");
      const double l_decay = parameters.get("dn.left.solver.reductiondecay"), l_offset = 1.0 / (1.0 - l_max / l_min);
      const double r_max = parameters.get("dn.right.solver.reduction"), r_min = parameters.get("dn.right.solver.minreduction");
      const double r_decay = parameters.get("dn.right.solver.reductiondecay"), r_offset = 1.0 / (1.0 - r_max / r_min);
      Vector u_old(u);
      int inner = 0;
      while (r / start_r > rel_reduction && r > max_error && i < max_iterations) {
        u_old = u;
        left_pde_solver.setReduction(l_min * (1.0 - 1.0 / (l_offset + l_decay * i)));
        left_pde_solver.apply(u, !reassemble_left && i > 0);
        inner += linear_solver_dn_0.res.iterations;
        for (size_t q = 0; q < u.size(); ++q) u[q] = alpha * u[q] + (1.0 - alpha) * u_old[q];
        u_old = u;
        right_pde_solver.setReduction(r_min * (1.0 - 1.0 / (r_offset + r_decay * i)));
        right_pde_solver.apply(u, !reassemble_right && i > 0);
        inner += linear_solver_dn_1.res.iterations;
        for (size_t q = 0; q < u.size(); ++q) u[q] = alpha * u[q] + (1.0 - alpha) * u_old[q];
        residual.assign(u.size(), 0.0);
        full_operator.residual(u, residual);
        r = norm2(residual);
        ++i;
        std::printf("%4zu %.12e\n", i, r);
      }
      double sum = 0.0; for (size_t q = 0; q < u.size(); ++q) sum += u[q];
      std::printf("Dirichlet-Neumann solver: %zu sweeps, %d inner iterations, sum(u) = %.12e |u|_2 = %.12e\n", i, inner, sum, norm2(u));
    }
    return 0;
}

int main(int argc, char** argv)
{
  try {
    if (argc < 2) { std::fprintf(stderr, "Usage: %s <ini file> [fem left = dg2] [fem right = dg2]   (fem: dg1, dg2, 1, 2)\n", argv[0]); return 1; }
    ParameterTree parameters(argv[1]);
    const std::string fem0 = argc > 2 ? argv[2] : "dg2", fem1 = argc > 3 ? argv[3] : "dg2";
    const bool dg = fem0.compare(0, 2, "dg") == 0;
    if (dg != (fem1.compare(0, 2, "dg") == 0)) throw Exception(MDB_EINVAL, "both sides conforming (1, 2) or both sides DG (dg1, dg2)");
    const int k0 = std::atoi(fem0.c_str() + (dg ? 2 : 0)), k1 = std::atoi(fem1.c_str() + (dg ? 2 : 0));
    const mdb_bctype b = boundary(MDB_BC_TESTPOISSON);                // DirichletBoundary :89-135 / bctype :220-242
    const mdb_function f = function(MDB_FN_DN_SOURCE);                // F :73-87 / f :203-218
    const mdb_function g = function(MDB_FN_GAUSS, 1.0, 0.5, 0.5, 0.5);
    const mdb_function j = function(MDB_FN_TESTPOISSON_J);
    if (dg) {
      ConvectionDiffusionModelProblem params(f, b, g, j);             // :958-962
      return solve(parameters, k0, k1, true, ConvectionDiffusionDG(params, ConvectionDiffusionDGMethod::SIPG,
                                                                   ConvectionDiffusionDGWeights::weightsOn, 2.0));   // dglop() :864-881
    }
    return solve(parameters, k0, k1, false, Poisson(f, b, j, 6));      // CGLOP cglop(f,b,j,6) :955-956
  } catch (Exception& e) {
    std::fprintf(stderr, "mdb200 error %d: %s\n", e.code, e.what());
    return 1;
  }
}
