// examples/testinstationarypoisson.cc — the reference driver test/testinstationarypoisson.cc:150-403 written against the
// B200 facade: two subdomains, dt0 operator = Poisson on both sides + ProportionalFlowCoupling, dt1 operator = L2 mass,
// OneStepGridOperator, Alexander2 time stepping (:320), time-dependent Dirichlet data; dim is a run-time argument.
//
//   testinstationarypoisson <refinement level> <coupling intensity> <dt> <tend> [dim=2] [scheme: alexander2 | euler | cn] [solver: jac | mg]
//
// Prints the DOF line, per-step solver statistics and a checksum of the final state.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <string>

#include "mdb200/multidomain.hh"

using namespace mdb200;

template <class LS, class Param>
static void timeloop(const Param& method, OneStepGridOperator& go, const SubProblemFunctions& f, Vector& uold, double dt, double tend)
{
  LS ls(5000);                                                       // the reference uses BCGS_SSOR (:311); Jacobi / multigrid are the device paths
  OneStepMethod<Param, LS> osm(method, go, ls, 1e-10);
  Vector unew(uold);
  double time = 0;
  int steps = 0, its = 0;
  while (time < tend - 1e-8) {                                       // :360-401
    osm.apply(time, dt, uold, f, unew);
    its += osm.res.iterations; ++steps;
    uold = unew;
    time += dt;
  }
  std::printf("time loop: %d steps, %d linear iterations\n", steps, its);
}

int main(int argc, char** argv)
{
  try {
    if (argc < 5) { std::fprintf(stderr, "Usage: %s <refinement level> <coupling intensity> <dt> <tend> [dim] [alexander2|euler|cn] [jac|mg]\n", argv[0]); return 1; }
    const int dim = argc > 5 ? std::atoi(argv[5]) : 2;
    const std::string scheme = argc > 6 ? argv[6] : "alexander2";
    const int64_t cells = (int64_t)1 << std::atoi(argv[1]);
    const int64_t n[3] = {cells, cells, cells};
    const double lower[3] = {0, 0, 0}, upper[3] = {1, 1, 1};

    Context ctx(0);
    MultiDomainGrid grid(ctx, dim, n, lower, upper);
    grid.startSubDomainMarking();
    for (int64_t e = 0; e < grid.size0(); ++e) {
      const int64_t i = e % cells;                                   // center[0] > 0.5
      if ((i + 0.5) / cells > 0.5) grid.addToSubDomain(1, e); else grid.addToSubDomain(0, e);
    }
    grid.preUpdateSubDomains(); grid.updateSubDomains(); grid.postUpdateSubDomains();

    GridFunctionSpace gfs0(grid.subDomain(0), 1), gfs1(grid.subDomain(1), 1);
    MultiDomainGridFunctionSpace multigfs(grid, {&gfs0, &gfs1});

    const mdb_bctype b = boundary(MDB_BC_TESTPOISSON);
    const mdb_function f = function(MDB_FN_DN_SOURCE);
    const mdb_function g = function(MDB_FN_INSTAT_G, 1.0, 3.0);      // time-dependent Dirichlet data
    const mdb_function j = function(MDB_FN_TESTPOISSON_J);
    Poisson lop(f, b, j, 4);
    L2 mass(2);

    SubDomainEqualityCondition c0({0}), c1({1});
    typedef SubProblem<Poisson, SubDomainEqualityCondition> SP;
    typedef SubProblem<L2, SubDomainEqualityCondition> TSP;
    SP left_sp_dt0(multigfs, lop, c0, gfs0), right_sp_dt0(multigfs, lop, c1, gfs1);
    TSP left_sp_dt1(multigfs, mass, c0, gfs0), right_sp_dt1(multigfs, mass, c1, gfs1);

    ProportionalFlowCoupling proportionalFlowCoupling(std::atof(argv[2]));
    Coupling<SP, SP, ProportionalFlowCoupling> coupling(left_sp_dt0, right_sp_dt0, proportionalFlowCoupling);

    ConstraintsContainer cg;
    constraints(multigfs).constrainSubProblem(left_sp_dt0, b).constrainSubProblem(right_sp_dt0, b).assemble(cg);

    Vector uold;
    interpolateOnTrialSpace(multigfs, uold, g, left_sp_dt0, g, right_sp_dt0);
    std::printf("%lld dof total, %lld dof constrained\n", (long long)uold.size(), (long long)cg.size());

    GridOperator go_dt_0(multigfs, multigfs, cg, cg, {left_sp_dt0.handle, right_sp_dt0.handle}, {coupling.handle});
    GridOperator go_dt_1(multigfs, multigfs, cg, cg, {left_sp_dt1.handle, right_sp_dt1.handle}, {});
    OneStepGridOperator gridoperator(go_dt_0, go_dt_1);

    const SubProblemFunctions f_ = interpolateOnSubProblems(g, left_sp_dt0, g, right_sp_dt0);
    const double dt = std::atof(argv[3]), tend = std::atof(argv[4]);
    const bool mg = argc > 7 && std::string(argv[7]) == "mg";      // BiCGSTAB + one multigrid V-cycle instead of BiCGSTAB + Jacobi
    if (mg) {
      if (scheme == "euler") timeloop<ISTLBackend_SEQ_BCGS_AMG_SSOR>(ImplicitEulerParameter(), gridoperator, f_, uold, dt, tend);
      else if (scheme == "cn") timeloop<ISTLBackend_SEQ_BCGS_AMG_SSOR>(OneStepThetaParameter(0.5), gridoperator, f_, uold, dt, tend);
      else timeloop<ISTLBackend_SEQ_BCGS_AMG_SSOR>(Alexander2Parameter(), gridoperator, f_, uold, dt, tend);
    } else {
      if (scheme == "euler") timeloop<ISTLBackend_SEQ_BCGS_Jac>(ImplicitEulerParameter(), gridoperator, f_, uold, dt, tend);
      else if (scheme == "cn") timeloop<ISTLBackend_SEQ_BCGS_Jac>(OneStepThetaParameter(0.5), gridoperator, f_, uold, dt, tend);
      else timeloop<ISTLBackend_SEQ_BCGS_Jac>(Alexander2Parameter(), gridoperator, f_, uold, dt, tend);
    }

    double sum = 0.0, sum2 = 0.0;
    for (size_t i = 0; i < uold.size(); ++i) { sum += uold[i]; sum2 += uold[i] * uold[i]; }
    std::printf("checksum: sum(u) = %.12e  |u|_2 = %.12e\n", sum, std::sqrt(sum2));
    return 0;
  } catch (Exception& e) {
    std::fprintf(stderr, "mdb200 error %d: %s\n", e.code, e.what());
    return 1;
  }
}
