// examples/testgmshpoisson.cc — a two-subdomain Poisson problem on a Gmsh triangle mesh, written against the B200 facade
// (include/mdb200/multidomain.hh).  The grid part is the reference's test/stokesdarcy2.cc:575-604 (GmshReader, globalRefine,
// physical group > 0 -> subdomain 1), the problem part is test/testpoisson.cc:163-298 with P1 spaces on both subdomains and the
// reference's ProportionalFlowCoupling (test/proportionalflowcoupling.hh:10-91) across the interface.
//
//   testgmshpoisson <mesh.msh> [refinements=0] [coupling intensity=2.0] [solver: ssor (default) | jac] [vtk output prefix]
//
// Prints the grid sizes, the DOF / constrained-DOF line, the solver statistics and a checksum of u.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <string>

#include "mdb200/multidomain.hh"

using namespace mdb200;

int main(int argc, char** argv)
{
  try {
    if (argc < 2) { std::fprintf(stderr, "Usage: %s <mesh.msh> [refinements] [coupling intensity] [ssor|jac]\n", argv[0]); return 1; }
    const int refine = argc > 2 ? std::atoi(argv[2]) : 0;
    const double intensity = argc > 3 ? std::atof(argv[3]) : 2.0;
    const bool jacobi = argc > 4 && std::string(argv[4]) == "jac";
    const std::string out = argc > 5 ? argv[5] : "";

    GmshMesh mesh = GmshReader::read(argv[1], true, false);          // gmshreader.read(filename, b2p, e2p, true, false)  (:578)
    mesh.globalRefine(refine);
    std::printf("%lld vertices, %lld elements after %d refinements\n", (long long)mesh.numVertices(), (long long)mesh.numElements(), refine);

    Context ctx(0);
    MultiDomainGrid grid(ctx, mesh);
    grid.startSubDomainMarking();
    for (int64_t e = 0; e < grid.size0(); ++e)
      grid.addToSubDomain(mesh.elementIndexToPhysicalGroup[(size_t)e] > 0 ? 1 : 0, e);      // :596-601
    grid.preUpdateSubDomains(); grid.updateSubDomains(); grid.postUpdateSubDomains();

    PkGridFunctionSpace gfs0(grid.subDomain(0), 1), gfs1(grid.subDomain(1), 1);
    MultiDomainGridFunctionSpace multigfs(grid, {&gfs0, &gfs1});

    const mdb_bctype b = boundary(MDB_BC_ALL_DIRICHLET);
    const mdb_function f = function(MDB_FN_CONST, 1.0);
    const mdb_function g = function(MDB_FN_ZERO);
    const mdb_function j = function(MDB_FN_ZERO);
    Poisson lop(f, b, j, 2);

    SubDomainEqualityCondition c0({0}), c1({1});
    typedef SubProblem<Poisson, SubDomainEqualityCondition> SP;
    SP left_sp(multigfs, lop, c0, gfs0), right_sp(multigfs, lop, c1, gfs1);

    ProportionalFlowCoupling proportionalFlowCoupling(intensity);
    Coupling<SP, SP, ProportionalFlowCoupling> coupling(left_sp, right_sp, proportionalFlowCoupling);

    ConstraintsContainer cg;
    constraints(multigfs).constrainSubProblem(left_sp, b).constrainSubProblem(right_sp, b).assemble(cg);

    Vector u;
    interpolateOnTrialSpace(multigfs, u, g, left_sp, g, right_sp);
    std::printf("%lld dof total, %lld dof constrained\n", (long long)u.size(), (long long)cg.size());

    GridOperator gridoperator(multigfs, multigfs, cg, cg, {left_sp.handle, right_sp.handle}, {coupling.handle});

    LinearSolverResult res;
    if (jacobi) {
      ISTLBackend_SEQ_BCGS_Jac ls(20000);
      StationaryLinearProblemSolver<ISTLBackend_SEQ_BCGS_Jac> pdesolver(gridoperator, ls, 1e-10);
      pdesolver.apply(u); res = ls.res;
    } else {
      ISTLBackend_SEQ_BCGS_SSOR ls(5000);
      StationaryLinearProblemSolver<ISTLBackend_SEQ_BCGS_SSOR> pdesolver(gridoperator, ls, 1e-10);
      pdesolver.apply(u); res = ls.res;
    }

    if (!out.empty()) {                                              // one piece per subdomain, as every reference driver ends (vtk.hh:83-98)
      VTKWriter left(grid, grid.subDomain(0)), right(grid, grid.subDomain(1));
      left.addSolution(multigfs, u, {&gfs0, &gfs1});
      right.addSolution(multigfs, u, {&gfs0, &gfs1});
      left.write(out + "testgmshpoisson-left"); right.write(out + "testgmshpoisson-right");
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
