// examples/testpowermortarpoisson.cc — the reference driver test/testpowermortarpoisson.cc:287-497 written against the B200 facade
// (include/mdb200/multidomain.hh): as testmortarpoisson, but the last child is PowerCouplingGridFunctionSpace<CouplingGFS,2,VBE>
// (:384-386) — two copies of the Q1 mortar space on the interface, block after block — and MortarPoissonCoupling::alpha_generic
// loops over the component blocks (:186-189).  Poisson on both sides, EnrichedCoupling, residual + Jacobian, BiCGSTAB + SSOR(3) to
// 1e-10 (:456-468).
//
//   testpowermortarpoisson <refinement> <scaling factor for coupling> [dim=2] [output prefix]   (reference defaults: 5 1.0, :285-286)
//
// With an output prefix the program also writes what the reference writes: <prefix>jacobian.mat (Dune::writeMatrixToMatlab, :453)
// and the solution per subdomain as <prefix>testpowermortarpoisson-left.vtu / -right.vtu (:465-486).
//
// Prints multigfs.ordering().size() like the reference (:386), the solver statistics and a checksum of u.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <string>

#include "mdb200/multidomain.hh"

using namespace mdb200;

int main(int argc, char** argv)
{
  try {
    const int refinement = argc >= 3 ? std::atoi(argv[1]) : 5;
    const double scaling = argc >= 3 ? std::atof(argv[2]) : 1.0;
    const int dim = argc > 3 ? std::atoi(argv[3]) : 2;
    const std::string out = argc > 4 ? argv[4] : "";
    const int64_t cells = (int64_t)1 << refinement;                  // YaspGrid 1^d cells, globalRefine(refinement) (:291-296)
    const int64_t n[3] = {cells, cells, cells};
    const double lower[3] = {0, 0, 0}, upper[3] = {1, 1, 1};

    Context ctx(0);
    MultiDomainGrid grid(ctx, dim, n, lower, upper);
    grid.startSubDomainMarking();
    for (int64_t e = 0; e < grid.size0(); ++e) {
      const double centre0 = ((e % cells) + 0.5) / cells;            // center[0] > 0.5 -> 1, < 0.5 -> 0 (:325-330)
      if (centre0 > 0.5) grid.addToSubDomain(1, e);
      if (centre0 < 0.5) grid.addToSubDomain(0, e);
    }
    grid.preUpdateSubDomains(); grid.updateSubDomains(); grid.postUpdateSubDomains();

    SubDomainEqualityCondition c0({0}), c1({1});
    GridFunctionSpace gfs0(grid.subDomain(0), 1), gfs1(grid.subDomain(1), 1);
    CouplingGridFunctionSpace couplinggfs(1, c0, c1);                 // Pred pred(mdgv, c0, c1) (:371-375)
    PowerCouplingGridFunctionSpace powercouplinggfs(couplinggfs, 2);   // :384-386
    MultiDomainGridFunctionSpace multigfs(grid, {&gfs0, &gfs1, &powercouplinggfs});

    const mdb_bctype b = boundary(MDB_BC_TESTPOISSON);
    const mdb_function f = function(MDB_FN_ZERO), g = function(MDB_FN_GAUSS, 1.0, 0.5, 0.5, 0.5), j = function(MDB_FN_TESTPOISSON_J);
    Poisson lop(f, b, j, 4);
    MortarPoissonCoupling coupling_lop(scaling);

    typedef SubProblem<Poisson, SubDomainEqualityCondition> SP;
    SP subproblem0(multigfs, lop, c0, gfs0), subproblem1(multigfs, lop, c1, gfs1);
    EnrichedCoupling<SP, SP, MortarPoissonCoupling> coupling(subproblem0, subproblem1, coupling_lop, powercouplinggfs);

    std::printf("%lld\n", (long long)multigfs.size());               // std::cout << multigfs.ordering().size() (:386)

    Vector u;
    interpolateOnTrialSpace(multigfs, u, g, subproblem0, g, subproblem1);
    ConstraintsContainer cg;
    constraints(multigfs).constrainSubProblem(subproblem0, b).constrainSubProblem(subproblem1, b).assemble(cg);

    GridOperator gridoperator(multigfs, multigfs, cg, cg, {subproblem0.handle, subproblem1.handle}, {coupling.handle});

    Vector r(u.size(), 0.0);
    gridoperator.residual(u, r);                                     // :442-444
    Jacobian J(gridoperator);
    gridoperator.jacobian(u, J);                                     // :446-448
    if (!out.empty()) writeMatrixToMatlab(J, (int64_t)u.size(), out + "jacobian.mat");        // :453
    std::printf("%lld dof total, %lld dof constrained, %lld nonzeroes\n", (long long)u.size(), (long long)cg.size(), (long long)J.nonzeroes());

    ISTLBackend_SEQ_BCGS_SSOR ls(5000);                              // :456-457
    StationaryLinearProblemSolver<ISTLBackend_SEQ_BCGS_SSOR> pdesolver(gridoperator, ls, 1e-10);
    pdesolver.apply(u);

    if (!out.empty()) {
      VTKWriter left(grid, grid.subDomain(0)), right(grid, grid.subDomain(1));
      left.addSolution(multigfs, u, {&gfs0, &gfs1, &powercouplinggfs});   // addSolutionToVTKWriter(..., subdomain_predicate(sdgv0.grid().domain()))
      right.addSolution(multigfs, u, {&gfs0, &gfs1, &powercouplinggfs});
      left.write(out + "testpowermortarpoisson-left"); right.write(out + "testpowermortarpoisson-right");
    }

    double sum = 0.0, sum2 = 0.0, rsum2 = 0.0;
    for (size_t i = 0; i < u.size(); ++i) { sum += u[i]; sum2 += u[i] * u[i]; rsum2 += r[i] * r[i]; }
    std::printf("residual at the interpolated u: |r|_2 = %.12e\n", std::sqrt(rsum2));
    std::printf("solver: %d iterations, reduction %.3e, %.3f ms\n", ls.res.iterations, ls.res.reduction, ls.res.elapsed * 1e3);
    std::printf("checksum: sum(u) = %.12e  |u|_2 = %.12e\n", sum, std::sqrt(sum2));
    return 0;
  } catch (Exception& e) {
    std::fprintf(stderr, "mdb200 error %d: %s\n", e.code, e.what());
    return 1;
  }
}
