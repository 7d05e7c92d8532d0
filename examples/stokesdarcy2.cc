// examples/stokesdarcy2.cc — the reference's Stokes-Darcy driver test/stokesdarcy2.cc:564-790 written against the B200 facade
// (include/mdb200/multidomain.hh): Gmsh mesh -> MultiDomainGrid (physical group > 0 -> Darcy subdomain), Taylor-Hood P2/P1 velocity /
// pressure on the free-flow subdomain, orthonormal P3 hydraulic head on the soil, TaylorHoodNavierStokes + ConvectionDiffusionDG(SIPG,
// weightsOn) subproblems, StokesDarcyCouplingOperator across the interface (finite-difference Jacobian with the ini's epsilon),
// StationaryLinearProblemSolver with a direct solve (ISTLBackend_SEQ_SuperLU in the reference) from u = 0.
//
//   stokesdarcy2 <ini file> [refinements] [mesh file (overrides mesh.filename, relative to the ini's directory)] [VTK file prefix]
//
// Prints the grid sizes, "<n> DOF, <m> restricted" (:764), the solver statistics, the norm of the residual at the solution (:792-794)
// and checksums of the four solution blocks.
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
  bool has(const std::string& key) const { return kv.count(key) != 0; }
  const std::string& operator[](const std::string& key) const {
    std::map<std::string, std::string>::const_iterator it = kv.find(key);
    if (it == kv.end()) throw Exception(MDB_EINVAL, "missing ini key " + key);
    return it->second;
  }
  double get(const std::string& key) const { return std::atof((*this)[key].c_str()); }
  double get(const std::string& key, double dflt) const { return has(key) ? get(key) : dflt; }
};

static double norm2(const Vector& v, int64_t a, int64_t b) { double s = 0.0; for (int64_t i = a; i < b; ++i) s += v[(size_t)i] * v[(size_t)i]; return std::sqrt(s); }

int main(int argc, char** argv)
{
  try {
    if (argc < 2) { std::fprintf(stderr, "Usage: %s <ini file> [refinements] [mesh file]\n", argv[0]); return 1; }
    ParameterTree parameters(argv[1]);
    std::string dir(argv[1]);
    dir = dir.find_last_of('/') == std::string::npos ? std::string() : dir.substr(0, dir.find_last_of('/') + 1);
    const int refine = argc > 2 ? std::atoi(argv[2]) : (int)parameters.get("mesh.refine", 0.0);
    const std::string meshfile = argc > 3 ? std::string(argv[3]) : dir + parameters["mesh.filename"];

    GmshMesh mesh = GmshReader::read(meshfile, true, false);                        // gmshreader.read(mesh.filename, b2p, e2p, true, false)  :578
    mesh.globalRefine(refine);
    std::printf("%lld vertices, %lld elements after %d refinements\n", (long long)mesh.numVertices(), (long long)mesh.numElements(), refine);

    Context ctx(0);
    MultiDomainGrid grid(ctx, mesh);
    grid.startSubDomainMarking();                                                   // :594-604
    for (int64_t e = 0; e < grid.size0(); ++e)
      grid.addToSubDomain(mesh.elementIndexToPhysicalGroup[(size_t)e] > 0 ? 1 : 0, e);
    grid.preUpdateSubDomains(); grid.updateSubDomains(); grid.postUpdateSubDomains();
    grid.setPhysicalGroups(mesh.elementIndexToPhysicalGroup);                       // DarcyParameters keeps the map (:262, :289)
    SubDomainGridView stokesGV = grid.subDomain(0), darcyGV = grid.subDomain(1);

    const int dim = 2, stokes_k = 2, stokes_q = 2 * stokes_k, darcy_k = 3;          // :609-611
    VectorGridFunctionSpace powervgfs(stokesGV, stokes_k, dim);                     // :646-655
    PkGridFunctionSpace pgfs(stokesGV, stokes_k - 1);                               // :657-659
    CompositeGridFunctionSpace stokesgfs({&powervgfs, &pgfs});                      // :661-669
    OPBGridFunctionSpace darcygfs(darcyGV, darcy_k);                                // :671-673
    MultiDomainGridFunctionSpace multigfs(grid, {&stokesgfs, &darcygfs});           // :675-683

    const double X = 100.0;
    // StokesBoundaryType :117-142 answers StressNeumann on the whole boundary; PressureDropFlux :171-205
    NavierStokesParameters navierStokesParams(parameters.get("parameters.fluid.mu"), parameters.get("parameters.fluid.rho"), parameters.get("parameters.gravity", 9.81),
                                              boundary(MDB_BC_ALL_NEUMANN),
                                              function(MDB_FN_PRESSURE_DROP, parameters.get("boundaries.inflow"), parameters.get("boundaries.outflow"), X, 50.0, 100.0));
    TaylorHoodNavierStokes stokesOperator(navierStokesParams, stokes_q);            // :719-720

    mdb_bctype darcybc = boundary(MDB_BC_DARCY_RIGHT); darcybc.p[0] = X; darcybc.p[1] = 13.0;                                      // :338-353
    DarcyParameters darcyParams(parameters.get("parameters.soil.permeability"), parameters.get("parameters.soil.permeability.low"), darcybc,
                                function(MDB_FN_DARCY_G, parameters.get("boundaries.rightpotential.top"), parameters.get("boundaries.rightpotential.bottom"), 15.0),
                                function(MDB_FN_DARCY_J, parameters.get("boundaries.bottomflux"), X));
    DarcyConvectionDiffusionDG darcyOperator(darcyParams, ConvectionDiffusionDGMethod::SIPG, ConvectionDiffusionDGWeights::weightsOn,
                                             parameters.get("parameters.darcy.alpha", 1.0));                                       // :733-739

    CouplingParameters couplingParams(parameters.get("parameters.coupling.alpha", 1.0), parameters.get("parameters.coupling.gamma", 1.0),
                                      parameters.get("parameters.soil.porosity"), parameters.get("parameters.gravity", 9.81), parameters.get("parameters.fluid.rho"),
                                      parameters.get("parameters.fluid.mu"), parameters.get("parameters.epsilon", 1e-8), parameters.get("parameters.soil.permeability"));
    StokesDarcyCouplingOperator couplingOperator(couplingParams);                   // :741-745

    SubDomainEqualityCondition c0({0}), c1({1});
    typedef SubProblem<TaylorHoodNavierStokes, SubDomainEqualityCondition> StokesSubProblem;
    typedef SubProblem<DarcyConvectionDiffusionDG, SubDomainEqualityCondition> DarcySubProblem;
    StokesSubProblem stokesSubProblem(multigfs, stokesOperator, c0, stokesgfs);     // three children: v_0, v_1, p  :747-752
    DarcySubProblem darcySubProblem(multigfs, darcyOperator, c1, darcygfs);
    Coupling<StokesSubProblem, DarcySubProblem, StokesDarcyCouplingOperator> coupling(stokesSubProblem, darcySubProblem, couplingOperator);   // :755-756

    ConstraintsContainer cg;
    constraints(multigfs).constrainSubProblem(stokesSubProblem, stokesOperator.bctype()).assemble(cg);                             // :759-767
    std::printf("%lld DOF, %lld restricted\n", (long long)multigfs.size(), (long long)cg.size());

    GridOperator gridoperator(multigfs, multigfs, cg, cg, {stokesSubProblem.handle, darcySubProblem.handle}, {coupling.handle});  // :769-778

    Vector u((size_t)multigfs.size(), 0.0);                                         // V u(multigfs, 0.0)  :782
    ISTLBackend_SEQ_SuperLU ls(true);                                               // :786-787
    StationaryLinearProblemSolver<ISTLBackend_SEQ_SuperLU> pdesolver(gridoperator, ls, 1e-15);
    pdesolver.apply(u);                                                             // :801
    std::printf("direct solve: defect reduction %.3e in %.3f s (device)\n", ls.res.reduction, ls.res.elapsed);

    Vector r(u.size(), 0.0);
    gridoperator.residual(u, r);                                                    // :803-805
    std::printf("|R(u)| = %.6e\n", norm2(r, 0, (int64_t)r.size()));

    int64_t off[5];
    ctx.check(mdb_get_child_offsets(ctx.get(), off), "mdb_get_child_offsets");
    const char* names[4] = {"v_0", "v_1", "p", "phi"};
    for (int k = 0; k < 4; ++k) std::printf("|%s| = %.12e  (%lld coefficients)\n", names[k], norm2(u, off[k], off[k + 1]), (long long)(off[k + 1] - off[k]));

    // addSolutionToVTKWriter(vtkwriter, multigfs, u, subdomain_predicate(...)); vtkwriter.write("stokes" / "darcy")  :831-880.
    // One conforming piece per subdomain: v_0, v_1 (P2: the vertex coefficients) and p on the Stokes view, the hydraulic head (orthonormal
    // P3: every cell's polynomial at its corners, averaged per vertex) with the reference's derived fields — pressure and Darcy velocity
    // from the potential (:443-560) — on the Darcy view.  Off unless the deck says `output.vtk = true` (fifth argument: prefix).
    if (parameters.has("output.vtk") && (parameters["output.vtk"] == "true" || parameters["output.vtk"] == "1")) {
      const std::string prefix = argc > 4 ? std::string(argv[4]) : std::string();
      VTKWriter stokes(grid, stokesGV), darcy(grid, darcyGV);
      stokes.addSolution(multigfs, u, {&stokesgfs, &darcygfs});
      darcy.addSolution(multigfs, u, {&stokesgfs, &darcygfs}, "phi");
      darcy.addDarcyFields(multigfs, u, {&stokesgfs, &darcygfs}, darcygfs, mesh.elementIndexToPhysicalGroup,       // "p", "v"  :876-884
                           parameters.get("parameters.soil.permeability"), parameters.get("parameters.soil.permeability.low"), 1,
                           parameters.get("parameters.soil.porosity"), parameters.get("parameters.gravity", 9.81), parameters.get("parameters.fluid.rho"));
      stokes.write(prefix + "stokes"); darcy.write(prefix + "darcy");
      std::printf("wrote %sstokes.vtu, %sdarcy.vtu\n", prefix.c_str(), prefix.c_str());
    }
    return 0;
  } catch (Exception& e) { std::fprintf(stderr, "error %d: %s\n", e.code, e.what()); return 1; }
  catch (std::exception& e) { std::fprintf(stderr, "error: %s\n", e.what()); return 1; }
}
