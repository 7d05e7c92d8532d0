// examples/testpoisson.cc — the reference driver test/testpoisson.cc:119-298 written against the B200 facade
// (include/mdb200/multidomain.hh).  Same sequence of objects and calls; dim is a run-time argument so that the same
// program runs the shipped 2-D case ("4 2.0", test/CMakeLists.txt:31-32) and the synthetic 3-D extension.
//
//   testpoisson <refinement level> <coupling intensity> [dim=2] [k left=1] [k right=2] [solver: ssor (default, as shipped) | jac]
//
// Prints the DOF / constrained-DOF line of the reference (testpoisson.cc:269), the solver statistics and a checksum of u.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <string>

#include "mdb200/multidomain.hh"

using namespace mdb200;

int main(int argc, char** argv)
{
  try {
    if (argc < 3) { std::fprintf(stderr, "Usage: %s <refinement level> <coupling intensity> [dim] [k left] [k right] [ssor|jac]\n", argv[0]); return 1; }
    const int dim = argc > 3 ? std::atoi(argv[3]) : 2;
    const int k0 = argc > 4 ? std::atoi(argv[4]) : 1, k1 = argc > 5 ? std::atoi(argv[5]) : 2;   // Q1 left, Q2 right (:168-169)
    const bool jacobi = argc > 6 && std::string(argv[6]) == "jac";
    const int64_t cells = (int64_t)1 << std::atoi(argv[1]);          // YaspGrid 1^d cells, globalRefine(level)
    const int64_t n[3] = {cells, cells, cells};
    const double lower[3] = {0, 0, 0}, upper[3] = {1, 1, 1};

    Context ctx(0);
    MultiDomainGrid grid(ctx, dim, n, lower, upper);
    grid.startSubDomainMarking();
    for (int64_t e = 0; e < grid.size0(); ++e) {
      const int64_t j = (e / cells) % cells;                         // it->geometry().center()[1] > 0.5   (:154)
      if ((j + 0.5) / cells > 0.5) grid.addToSubDomain(1, e); else grid.addToSubDomain(0, e);
    }
    grid.preUpdateSubDomains(); grid.updateSubDomains(); grid.postUpdateSubDomains();

    GridFunctionSpace gfs0(grid.subDomain(0), k0), gfs1(grid.subDomain(1), k1);
    MultiDomainGridFunctionSpace multigfs(grid, {&gfs0, &gfs1});

    const mdb_bctype b = boundary(MDB_BC_TESTPOISSON);               // DirichletBoundary :42-88
    const mdb_function f = function(MDB_FN_ZERO);                    // F: y = 0 (:38)
    const mdb_function g = function(MDB_FN_GAUSS, 1.0, 0.5, 0.5, 0.5);   // G :92-99
    const mdb_function j = function(MDB_FN_TESTPOISSON_J);           // J :103-116
    Poisson lop(f, b, j, 4);

    SubDomainEqualityCondition c0({0}), c1({1});
    typedef SubProblem<Poisson, SubDomainEqualityCondition> SP;
    SP left_sp(multigfs, lop, c0, gfs0), right_sp(multigfs, lop, c1, gfs1);

    ContinuousValueContinuousFlowCoupling proportionalFlowCoupling(4, std::atof(argv[2]));
    Coupling<SP, SP, ContinuousValueContinuousFlowCoupling> coupling(left_sp, right_sp, proportionalFlowCoupling);

    ConstraintsContainer cg;
    constraints(multigfs).constrainSubProblem(left_sp, b).constrainSubProblem(right_sp, b).assemble(cg);

    Vector u;
    interpolateOnTrialSpace(multigfs, u, g, left_sp, g, right_sp);
    std::printf("%lld dof total, %lld dof constrained\n", (long long)u.size(), (long long)cg.size());

    GridOperator gridoperator(multigfs, multigfs, cg, cg, {left_sp.handle, right_sp.handle}, {coupling.handle});

    LinearSolverResult res;
    if (jacobi) {                                                    // the fast device path
      ISTLBackend_SEQ_BCGS_Jac ls(5000);
      StationaryLinearProblemSolver<ISTLBackend_SEQ_BCGS_Jac> pdesolver(gridoperator, ls, 1e-10);
      pdesolver.apply(u); res = ls.res;
    } else {                                                         // ISTLBackend_SEQ_BCGS_SSOR(5000, verb), :291-298
      ISTLBackend_SEQ_BCGS_SSOR ls(5000);
      StationaryLinearProblemSolver<ISTLBackend_SEQ_BCGS_SSOR> pdesolver(gridoperator, ls, 1e-10);
      pdesolver.apply(u); res = ls.res;
    }

    double sum = 0.0, sum2 = 0.0;
    for (size_t i = 0; i < u.size(); ++i) { sum += u[i]; sum2 += u[i] * u[i]; }
    std::printf("solver: %d iterations, reduction %.3e, %.3f ms\n", res.iterations, res.reduction, res.elapsed * 1e3);
    std::printf("checksum: sum(u) = %.12e  |u|_2 = %.12e\n", sum, std::sqrt(sum2));
    return 0;
  } catch (Exception& e) {
    std::fprintf(stderr, "mdb200 error %d: %s\n", e.code, e.what());
    return 1;
  }
}
