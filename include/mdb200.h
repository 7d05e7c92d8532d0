/* mdb200.h — C ABI of libmdb200.so: the B200 (sm_100a) backend for the
 * dune-multidomain assembly-and-solve hot path.
 *
 * There is no FFI/plugin boundary in the reference; the boundary is the C++
 * template concept realised by Dune::PDELab::MultiDomain::GridOperator
 * (reference dune/pdelab/multidomain/gridoperator.hh:38-214).  This header is the
 * plain-C layer that a C++ facade with the reference's spellings
 * (include/mdb200/multidomain.hh) lowers to.  Each entry point cites the
 * reference interface it replaces (paths relative to the reference root).
 *
 * Conventions
 *  - every function returns 0 on success, a negative MDB_E* code on failure;
 *    mdb_last_error(ctx) returns the message.  No exception crosses the ABI.
 *  - one context per host thread and per GPU; calls are ordered on the context's
 *    CUDA stream.  Not re-entrant (the reference is not either:
 *    gridoperator.hh:209-210 mutable assembler members).
 *  - pointer arguments documented "host or device" may be either; the library
 *    classifies them with cudaPointerGetAttributes.  Host buffers are copied
 *    in/out inside the call (that is the e2e path); device buffers are used in
 *    place (that is the resident path).
 *  - all floating point is IEEE double; indices exported for parity are int64
 *    (std::size_t in the reference), device-internal indices are int32.
 *  - there is NO CPU fallback: if no CUDA device is usable mdb_create fails.
 */
#ifndef MDB200_H
#define MDB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mdb_ctx mdb_ctx;

/* ---- error codes -------------------------------------------------------- */
enum {
  MDB_OK = 0,
  MDB_EINVAL = -1,      /* bad argument / unsupported combination (DUNE_THROW analogue) */
  MDB_ECUDA = -2,       /* CUDA runtime error */
  MDB_ESTATE = -3,      /* call out of order (e.g. assemble before finalize) */
  MDB_ENOMEM = -4,
  MDB_ENCCL = -5,
  MDB_ENOTCONVERGED = -6
};

/* ---- finite elements (child spaces of the MultiDomainGridFunctionSpace) -- */
enum {
  MDB_FE_Q1 = 1,        /* QkLocalFiniteElementMap<...,1>  test/testpoisson.cc:168 */
  MDB_FE_Q2 = 2,        /* QkLocalFiniteElementMap<...,2>  test/testpoisson.cc:169 */
  /* QkDGLocalFiniteElementMap<DF,double,k,dim> (test/testpoisson-dirichlet-neumann.cc:925-926): the Qk Lagrange basis with all
   * (k+1)^dim coefficients attached to the cell — DOF = (index of the cell in its subdomain) * (k+1)^dim + i [UPSTREAM (iii)];
   * no DOF is shared, ConformingDirichletConstraints constrains nothing (Dirichlet data enter weakly through the operator). */
  MDB_FE_DGQ1 = 11,
  MDB_FE_DGQ2 = 12,     /* the shipped spaces of BASELINE configs[0] (test/testpoisson-dirichlet-neumann.cc:968-971) */
  /* PkLocalFiniteElementMap<GV,DF,RF,1> on simplices (the P1 space of test/stokesdarcy2.cc:614, test/testcouplinggfs.cc:263):
   * one DOF per vertex, shape functions 1-x-y, x, y at the cell's vertices 0, 1, 2; numbered by ascending vertex index of the
   * child's (sub)domain [UPSTREAM (ii),(iii)].  Unstructured meshes only (mdb_mesh_unstructured). */
  MDB_FE_P1 = 21,
  /* PkLocalFiniteElementMap<SDGV,DF,RF,2> (the Taylor-Hood velocity space V_FEM of test/stokesdarcy2.cc:613): one DOF per vertex and
   * per edge; Pk2D lattice order on a cell: vertex 0, edge {0,1}, vertex 1, edge {0,2}, edge {1,2}, vertex 2.  Block numbering
   * [UPSTREAM (iii)]: the subdomain's vertices (ascending vertex index), then its edges (ascending edge index; the edge index of the
   * host grid is the order of first appearance in (cell, local face) order). */
  MDB_FE_P2 = 22,
  /* OPBLocalFiniteElementMap<DF,RF,k,2,simplex,GMPField<512>> (DarcyFEM of test/stokesdarcy2.cc:615): the orthonormal polynomial basis
   * of degree k on the reference triangle (csrc/opb_tables.h), all (k+1)(k+2)/2 coefficients attached to the cell:
   * DOF = (index of the cell in its subdomain) * size + i. */
  MDB_FE_OPB1 = 31, MDB_FE_OPB2 = 32, MDB_FE_OPB3 = 33
};

/* ---- subdomain conditions  (subdomainset.hh:82-133) ---------------------- */
enum {
  MDB_COND_EQUALITY = 0,  /* SubDomainEqualityCondition: set(e) == mask        */
  MDB_COND_SUBSET = 1     /* SubDomainSubsetCondition:   set(e) contains mask  */
};

/* ---- analytic problem data.  Arbitrary user C++ functors cannot run on the
 * device, so the data of the reference's drivers form a closed set. -------- */
enum {
  MDB_FN_ZERO = 0,
  MDB_FN_CONST = 1,          /* p[0]                                                          */
  MDB_FN_GAUSS = 2,          /* exp(-p[0] * |x - (p[1],p[2],p[3])|^2)   G of test/testpoisson.cc:92-99 (p0=1,c=.5) */
  MDB_FN_BOX = 3,            /* p[0] if p[1]<x_d<p[2] for all d<dim else 0   F of test/testpoisson.cc:32-37 before :38 */
  MDB_FN_TESTPOISSON_J = 4,  /* 0 on x1<1e-6 or x1>1-1e-6; -5 on x0>1-1e-6 && x1>.5+1e-6; else 0  test/testpoisson.cc:103-116 */
  MDB_FN_INSTAT_G = 5,       /* time dependent Dirichlet data, test/testinstationarypoisson.cc:94-100 family: p[0]*sin(p[1]*t)*exp(-|x-c|^2), c=.5 */
  MDB_FN_POLY2 = 6,          /* p[0] + sum_d p[1+d]*x_d + p[4]*|x|^2  (manufactured-solution tests)     */
  MDB_FN_DN_SOURCE = 7,      /* F of test/testpoisson-dirichlet-neumann.cc:73-87 (= testpoisson.cc:32-37 before the
                                overwrite): 50 on (0.25,0.375)^dim; else 60 exp(-10 |x|^2) if |0.5 - x| < 0.2; else 0 */
  MDB_FN_DARCY_G = 8,        /* DarcyParameters::g  test/stokesdarcy2.cc:355-361: s = x1 / p[2]; s p[0] + (1 - s) p[1]            */
  MDB_FN_DARCY_J = 9,        /* DarcyParameters::j  :364-371: p[0] x0 (p[1] - x0) on x1 < 1e-6, else 0                           */
  MDB_FN_PRESSURE_DROP = 10  /* PressureDropFlux::evaluateGlobal  :191-203: (x0 < 1e-6 ? (x1 > p[3] ? p[0] : -p[1])
                                : x0 > p[2] - 1e-6 ? (x1 > p[3] ? -p[1] : p[0]) : 0) * (p[4] - x1)                                */
};
typedef struct mdb_function {
  int32_t kind;
  int32_t reserved;
  double p[6];
} mdb_function;

enum {
  MDB_BC_ALL_DIRICHLET = 0,
  MDB_BC_ALL_NEUMANN = 1,
  MDB_BC_TESTPOISSON = 2,    /* test/testpoisson.cc:42-88: none on interior faces; Neumann on x1=0, x1=1 and
                                on x0=1 && x1>.5; Dirichlet elsewhere                                        */
  MDB_BC_DARCY_RIGHT = 3     /* DarcyParameters::bctype test/stokesdarcy2.cc:338-353: none on interior faces; Outflow on x1 < 1e-6 &&
                                x0 < p[0] (with b = 0 it acts like Neumann with o = j); Dirichlet on x0 > p[0] - 1e-6 && x1 < p[1];
                                Neumann elsewhere                                                                */
};
typedef struct mdb_bctype {
  int32_t kind;
  int32_t reserved;
  double p[4];
} mdb_bctype;

/* ---- local operators (closed set) --------------------- */
enum {
  /* volume / boundary operators used by SubProblems */
  MDB_OP_POISSON = 1,        /* [UPSTREAM] Dune::PDELab::Poisson<F,B,J>(f,b,j,quadOrder)  test/testpoisson.cc:218-219;
                                 in-tree twin test/instationarypoissonoperator.hh:30-168                            */
  MDB_OP_L2 = 2,             /* [UPSTREAM] Dune::PDELab::L2(intorder) / test/simpletimeoperator.hh:29-101 (mass)    */
  MDB_OP_CONVDIFF_DG = 3,    /* [UPSTREAM] Dune::PDELab::ConvectionDiffusionDG<ConvectionDiffusionModelProblem,FEM>(param, method,
                                 weights, alpha)  test/testpoisson-dirichlet-neumann.cc:864-881 (dglop), parameter class :169-285
                                 (A = I, b = 0, c = 0): volume + interior-penalty skeleton + weak boundary terms             */
  /* [UPSTREAM] Dune::PDELab::TaylorHoodNavierStokes<NavierStokesParameters>(params, stokes_q) with navier = false, full_tensor = true
   * (test/stokesdarcy2.cc:226-251, :719-720) on THREE children {v_0, v_1 : MDB_FE_P2, p : MDB_FE_P1}: the multi-child subproblem of
   * subproblemlocalfunctionspace.hh:144-261 */
  MDB_OP_TAYLORHOOD_STOKES = 4,
  /* [UPSTREAM] ConvectionDiffusionDG<DarcyParameters, OPB>(SIPG, weightsOn, alpha) (test/stokesdarcy2.cc:254-407, :733-739) on an
   * MDB_FE_OPBk child of a triangle mesh, tensor by physical group (mdb_cell_groups) */
  MDB_OP_DARCY_DG = 5,
  /* coupling operators used by Couplings */
  MDB_OP_STOKESDARCY_COUPLING = 104, /* StokesDarcyCouplingOperator<CouplingParameters>  test/stokesdarcycouplingoperator.hh:79-252 */
  MDB_OP_CVCF_COUPLING = 101,     /* ContinuousValueContinuousFlowCoupling  test/proportionalflowcoupling.hh:94-239 */
  MDB_OP_PROPFLOW_COUPLING = 102, /* ProportionalFlowCoupling               test/proportionalflowcoupling.hh:10-91  */
  MDB_OP_ND_COUPLING = 103,       /* ConvectionDiffusionDGNeumannDirichletCoupling<CouplingParameters>  test/neumanndirichletcoupling.hh:41-441,
                                     parameter class test/testpoisson-dirichlet-neumann.cc:36-69 (A = I, b = 0, c = 0)                  */
  /* enriched (mortar) coupling operators used by EnrichedCouplings */
  MDB_OP_MORTAR_POISSON_COUPLING = 201   /* MortarPoissonCoupling(gamma_0)  test/testmortarpoisson.cc:117-250        */
};

typedef struct mdb_poisson_params {
  mdb_function f;            /* source term            */
  mdb_bctype bc;             /* boundary-type predicate */
  mdb_function j;            /* Neumann flux            */
  int32_t quad_order;        /* 4 in testpoisson.cc:219, 6 in testpoisson-assembly.cc:920 */
  int32_t reserved;
} mdb_poisson_params;

typedef struct mdb_l2_params {
  int32_t quad_order;        /* default 2 */
  int32_t reserved;
} mdb_l2_params;

typedef struct mdb_cvcf_params {
  int32_t quad_order;        /* ctor arg intorder (4 in testpoisson.cc:234) */
  int32_t reserved;
  double intensity;          /* ctor arg intensity = alpha (argv[2])        */
} mdb_cvcf_params;

typedef struct mdb_propflow_params {
  double intensity;
} mdb_propflow_params;

/* ConvectionDiffusionDGNeumannDirichletCoupling(coupling_mode, param, method, weights, alpha, intorderadd)
 * (test/neumanndirichletcoupling.hh:56-74).  One-directional: alpha_coupling writes r_s only, jacobian_coupling mat_ss only
 * (:85-283, :291-432); analytic Jacobian (jac_mode is ignored); no pattern of its own (:51-52). */
enum { MDB_ND_MODE_NEUMANN = 0, MDB_ND_MODE_DIRICHLET = 1, MDB_ND_MODE_BOTH = 2 };       /* CouplingMode :19-24 */
enum { MDB_DG_METHOD_NIPG = 0, MDB_DG_METHOD_SIPG = 1 };                                 /* ConvectionDiffusionDGMethod [UPSTREAM] */
enum { MDB_DG_WEIGHTS_OFF = 0, MDB_DG_WEIGHTS_ON = 1 };                                  /* ConvectionDiffusionDGWeights [UPSTREAM] */
typedef struct mdb_ndcoupling_params {
  int32_t mode;              /* MDB_ND_MODE_*                                                          */
  int32_t method;            /* MDB_DG_METHOD_*: theta = 1 (NIPG) or -1 (SIPG), :72-73                 */
  int32_t weights;           /* MDB_DG_WEIGHTS_*                                                        */
  int32_t intorderadd;       /* quadrature order = intorderadd + 2 * max basis order, :107              */
  double alpha;              /* interior penalty parameter                                              */
} mdb_ndcoupling_params;

/* ConvectionDiffusionDG(param, method, weights, alpha, intorderadd = 0) [UPSTREAM dune-pdelab convectiondiffusiondg.hh; its face
 * terms are the ones test/neumanndirichletcoupling.hh:85-283 was derived from].  The parameter class is the reference's
 * ConvectionDiffusionModelProblem (test/testpoisson-dirichlet-neumann.cc:169-285): A = identity, b = 0, c = 0, and
 *   f        source, evaluated at the quadrature points of the cell (:203-218 = MDB_FN_DN_SOURCE)
 *   bc       boundary type at the FACE CENTRE: Dirichlet / Neumann / None — None on every interior face (:226-229), which is what
 *            makes alpha_boundary a no-op when the assembler calls it on a subdomain interface (residualassemblerfunctors.hh:80-86)
 *   g        Dirichlet value (:244-252 = MDB_FN_GAUSS), imposed weakly (Nitsche terms)
 *   j        Neumann flux (:254-268 = MDB_FN_TESTPOISSON_J)
 * doAlphaVolume, doAlphaSkeleton, doAlphaBoundary, doLambdaVolume, doPatternVolume, doPatternSkeleton; doSkeletonTwoSided = false.
 * Quadrature order = intorderadd + 2 * basis order for every term. */
typedef struct mdb_convdiffdg_params {
  mdb_function f;
  mdb_bctype bc;
  mdb_function g;
  mdb_function j;
  int32_t method;            /* MDB_DG_METHOD_*: SIPG in dglop() (:876)      */
  int32_t weights;           /* MDB_DG_WEIGHTS_*: weightsOn in dglop() (:877) */
  int32_t intorderadd;       /* 0                                             */
  int32_t reserved;
  double alpha;              /* interior penalty parameter: 2.0 in dglop() (:878) */
} mdb_convdiffdg_params;

/* NavierStokesParameters of test/stokesdarcy2.cc:226-251 over [UPSTREAM] NavierStokesDefaultParameters<GV,RF,F,B,V,J,false,true>:
 * Stokes (no convective term), full (symmetric-gradient) viscous tensor.  Residual rows, for velocity component d and test function v_i,
 * pressure test function q_i:   int mu (grad u_d . grad v_i + sum_dd d_d u_dd d_dd v_i) - p d_d v_i - f_d v_i ;   - int (div u) q_i ;
 * on faces of the physical boundary whose type is StressNeumann:  + int j (n_d v_i)  (the scalar flux times the outer normal). */
typedef struct mdb_taylorhood_params {
  double mu, rho;            /* parameters.fluid.mu, fluid.rho (rho only multiplies the disabled convective term)          */
  double f[2];               /* NavierStokesParameters::f = (0, -gravity)  :239-249                                        */
  mdb_bctype bc;             /* StokesBoundaryType :117-142: MDB_BC_ALL_NEUMANN = StressNeumann on the whole boundary (as shipped),
                                MDB_BC_ALL_DIRICHLET = VelocityDirichlet; interior faces: DoNothing                        */
  mdb_function j;            /* PressureDropFlux :171-205 = MDB_FN_PRESSURE_DROP                                           */
  int32_t quad_order;        /* stokes_q = 2 * stokes_k = 4  :610                                                          */
  int32_t full_tensor;       /* 1 as shipped                                                                               */
} mdb_taylorhood_params;

/* DarcyParameters of test/stokesdarcy2.cc:254-407: A = (group == high_group ? kabs : kabs_low) * identity per cell (:285-290),
 * b = 0, c = 0, f = 0; ConvectionDiffusionDG(params, method, weights, alpha) :733-739. */
typedef struct mdb_darcy_params {
  double kabs, kabs_low;     /* parameters.soil.permeability, soil.permeability.low                                        */
  mdb_bctype bc;             /* MDB_BC_DARCY_RIGHT (p[0] = domain length 100, p[1] = 13)                                   */
  mdb_function g;            /* MDB_FN_DARCY_G(rightpotential.top, rightpotential.bottom, 15)                              */
  mdb_function j;            /* MDB_FN_DARCY_J(bottomflux, 100); also the outflow flux o  :374-378                         */
  int32_t high_group;        /* 1                                                                                          */
  int32_t method;            /* MDB_DG_METHOD_SIPG                                                                         */
  int32_t weights;           /* MDB_DG_WEIGHTS_ON                                                                          */
  int32_t intorderadd;       /* 0: quadrature order 2 k                                                                    */
  double alpha;              /* parameters.darcy.alpha                                                                     */
} mdb_darcy_params;

/* CouplingParameters of test/stokesdarcycouplingoperator.hh:11-77.  The operator's Jacobian is NumericalJacobianCoupling(epsilon)
 * (:97-101; epsilon = 1e-2 in test/topbottomfiltration.ini:10).  quirk = 1 keeps the operator as written: the loop that transforms
 * the Darcy gradients runs over the number of VELOCITY functions (:194-195), so the gradients of the last head functions stay zero. */
typedef struct mdb_stokesdarcy_params {
  double gravity, alpha, viscosity, porosity, gamma, density;
  double kabs;               /* soil.permeability (isotropic tensor)                                                       */
  double epsilon;
  int32_t quirk;
  int32_t reserved;
} mdb_stokesdarcy_params;

typedef struct mdb_mortar_params {
  double gamma_0;            /* ctor arg ("scaling factor for coupling", test/testmortarpoisson.cc:402): gamma = gamma_0 / |face| */
} mdb_mortar_params;

/* Jacobian mode for coupling operators. */
enum {
  MDB_JAC_FD_BITMATCH = 0,   /* NumericalJacobianCoupling, eps=1e-8, couplingutilities.hh:59-141, evaluated
                                with the reference's operation order and without FMA contraction           */
  MDB_JAC_ANALYTIC = 1       /* exact derivative of the (linear) coupling residual                          */
};

/* ---- life cycle ---------------------------------------------------------- */
mdb_ctx* mdb_create(int device);
void mdb_destroy(mdb_ctx* ctx);
const char* mdb_last_error(mdb_ctx* ctx);   /* ctx may be NULL: returns the creation error */
const char* mdb_version(void);

/* ---- mesh + subdomains  (replaces YaspGrid + MultiDomainGrid marking,
 *      test/testpoisson.cc:134-161) ---------------------------------------- */
/* Structured axis-aligned grid, n[d] cells per axis on [lower,upper].  Cells and vertices are numbered
 * lexicographically, axis 0 fastest (YaspGrid leaf index order [UPSTREAM]). */
int mdb_mesh_structured(mdb_ctx* ctx, int dim, const int64_t* n, const double* lower, const double* upper);
/* Unstructured conforming simplex mesh: what Dune::GmshReader<UGGrid<2>>::read hands to the reference's drivers
 * (test/stokesdarcy2.cc:575-583; fixtures test/stokesdarcy*.msh).  dim = 2, etype = MDB_ETYPE_TRIANGLE; `xyz` holds nv points of `dim`
 * doubles, `conn` ne cells of 3 vertex indices (0-based).  Cell and vertex indices are the caller's order ([UPSTREAM]: the leaf
 * index order of the host grid); local vertex i of a cell carries the P1 shape function i; face f of a cell is opposite to
 * local vertex 2 - f (DUNE simplex numbering: face 0 = {0,1}, 1 = {0,2}, 2 = {1,2}).  The call validates the mesh (indices in
 * range, no degenerate cell, at most two cells per face) and builds the face neighbours and the vertex -> cell adjacency (the
 * grid manager's job in the reference).  Spaces: MDB_FE_P1; operators: MDB_OP_POISSON, MDB_OP_L2, MDB_OP_CVCF_COUPLING,
 * MDB_OP_PROPFLOW_COUPLING (MDB_EINVAL for the others in this build); linear solvers as on structured meshes.  Subdomains: mdb_subdomains with
 * one byte per cell (physical group -> subdomain as in test/stokesdarcy2.cc:596-601). */
enum { MDB_ETYPE_TRIANGLE = 2 };   /* gmsh element type 2 */
int mdb_mesh_unstructured(mdb_ctx* ctx, int dim, int64_t nv, const double* xyz, int64_t ne, int etype, const int64_t* conn);
/* Declare this context's mesh to be a slab of a larger global mesh (multi-GPU; the reference's analogue is
 * the overlap-1 MPI grid of test/testoverlappinginstationarypoisson.cc:232-233).  `global_n` cells per axis in total,
 * this rank's local cells start at `cell_offset`; only the slowest axis (dim-1) may be cut.  Convention: a rank that is
 * not the first one includes ONE ghost cell layer below its owned layers (the top layer owned by the rank below), so
 * that every cell around an owned vertex is local and owned rows are assembled completely without communication.  Local
 * vertex plane 0 of such a rank and the top local plane of every rank but the last are ghost planes (owned by the
 * neighbours).  Faces on a cut are processor boundaries (no boundary terms, globalassembler.hh:275-283).  `lower/upper`
 * given to mdb_mesh_structured are those of the local slab.  Call after mdb_mesh_structured, before mdb_finalize.
 * Child spaces on a partitioned mesh: conforming Qk (Q1 and Q2; a Qk block owns the lattice planes [k, k n) of its slab — the k planes
 * of the ghost cell layer belong to the rank below, the top plane to the rank above — and the halo exchange carries k planes up and
 * one plane down per block).  Cell-blocked QkDG blocks (all children QkDG): a row couples a cell to its face neighbours, so the local
 * slab carries ghost cell layers on BOTH sides; this call then only declares the cut (rank-independent coordinates, processor
 * boundaries) and ownership + halos are declared per cell layer through mdb_partition_lists after mdb_finalize
 * (dune-multidomain_b200/parallel.py: slab2, cell_blocked_ids, connect_lists).  Coupling (mortar) spaces are vertex-attached: their
 * blocks are owned by lattice planes like Q1 blocks, enriched couplings work on a partition.  Mixed conforming / QkDG children and the
 * Neumann-Dirichlet coupling (a one-directional operator driven by sub-block solves) are refused by mdb_finalize. */
int mdb_mesh_partition(mdb_ctx* ctx, const int64_t* global_n, const int64_t* cell_offset,
                       const double* global_lower, const double* global_upper);
/* One byte per cell: bit s set <=> cell belongs to subdomain s (<= 8 subdomains;
 * FewSubDomainsTraits<dim,4> in test/testpoisson.cc:140).  Host pointer, ncells entries. */
int mdb_subdomains(mdb_ctx* ctx, const uint8_t* cell_subdomain_bits);
/* Device-side marking for large meshes: cells whose centre has x[axis] > threshold go to subdomain
 * `sd_above`, the others to `sd_below` (test/testpoisson.cc:152-158). */
/* elementIndexToPhysicalGroup of the Gmsh reader (test/stokesdarcy2.cc:577-578, used by DarcyParameters::A :285-290): one int32 per
 * cell, host pointer.  Default: all zero. */
int mdb_cell_groups(mdb_ctx* ctx, const int32_t* groups);
/* Parity / output export of the marking: one subdomain bit set per cell (what gv.indexSet().subDomains(e) returns). */
int mdb_get_subdomains(mdb_ctx* ctx, uint8_t* cell_subdomain_bits);
int mdb_subdomains_halfspace(mdb_ctx* ctx, int axis, double threshold, int sd_below, int sd_above);

/* ---- function space  (MultiDomainGridFunctionSpace children,
 *      multidomaingridfunctionspace.hh:152,215; lexicographic child blocks) -- */
/* Appends a child space; returns its child index >= 0.  subdomain = -1 puts the child on the
 * MultiDomainGrid view itself, otherwise on SubDomainGrid view `subdomain`. */
int mdb_add_space(mdb_ctx* ctx, int fe_kind, int subdomain);
/* CouplingGridFunctionSpace<MDGV, QkLocalFiniteElementMap<(dim-1)-cube, 1>, SubProblemSubProblemInterface(left, right)>
 * (couplinggridfunctionspace.hh:14-249): a Q1 space on the interface faces whose inside cell satisfies the left condition and
 * whose outside cell satisfies the right one — the mortar space of test/testmortarpoisson.cc:343-375.  Its DOFs are the
 * vertices of those faces, numbered by ascending vertex index (couplinggfsordering.hh:150-200); like every child it
 * becomes one block of the lexicographic MultiDomainGridFunctionSpace ordering.  Returns the child index. */
int mdb_add_coupling_space(mdb_ctx* ctx, int fe_kind, int left_cond_kind, uint32_t left_cond_mask,
                           int right_cond_kind, uint32_t right_cond_mask);

/* ---- participants  (subproblem.hh:44-148, coupling.hh:30-91) -------------- */
/* Returns the subproblem index >= 0.  `children` are indices returned by mdb_add_space. */
int mdb_add_subproblem(mdb_ctx* ctx, int op_id, const void* params, size_t params_bytes,
                       int cond_kind, uint32_t cond_mask, const int* children, int nchildren);
/* Returns the coupling index >= 0.  Fires on a face iff local applies to inside and remote to outside. */
int mdb_add_coupling(mdb_ctx* ctx, int op_id, const void* params, size_t params_bytes,
                     int local_subproblem, int remote_subproblem, int jac_mode);
/* EnrichedCoupling<Local, Remote, Op, couplingLFSIndex> (coupling.hh:123-192): as mdb_add_coupling, with the coupling
 * (mortar) space `coupling_child` returned by mdb_add_coupling_space.  Per face the operator's
 * alpha_enriched_coupling_first(inside, mortar) and _second(outside, mortar) run (residualassemblerfunctors.hh:219-257);
 * Jacobians by NumericalJacobianEnrichedCouplingTo{First,Second}SubProblem (couplingutilities.hh:292-484); pattern =
 * FullEnrichedCoupling{First,Second}SubProblemPattern + FullEnrichedCouplingPattern (:203-286).  Returns a coupling index
 * (same index space as mdb_add_coupling). */
int mdb_add_enriched_coupling(mdb_ctx* ctx, int op_id, const void* params, size_t params_bytes,
                              int local_subproblem, int remote_subproblem, int coupling_child, int jac_mode);

/* PowerCouplingGridFunctionSpace<CouplingGFS, ncomp, VBE> (test/testpowermortarpoisson.cc:384-386): `ncomp` copies of the coupling
 * space of mdb_add_coupling_space, ordered lexicographically — i.e. ncomp consecutive child blocks.  Returns the index of the
 * first block; the others follow.  (The reference counts the power space as ONE child of the MultiDomainGridFunctionSpace; here
 * every component block is a child of its own, the global numbering is the same.) */
int mdb_add_power_coupling_space(mdb_ctx* ctx, int fe_kind, int left_cond_kind, uint32_t left_cond_mask,
                                 int right_cond_kind, uint32_t right_cond_mask, int ncomp);
/* EnrichedCoupling over a power coupling space: the operator of test/testpowermortarpoisson.cc:117-262 loops over the component
 * blocks (`for c_block < PowerLFSU_C::CHILDREN`, :186-189) and applies MortarPoissonCoupling::alpha_generic with each block's
 * multiplier; the finite-difference Jacobians differentiate that sum as a whole (couplingutilities.hh:292-484).
 * `first_coupling_child` comes from mdb_add_power_coupling_space; ncomp = 1 is mdb_add_enriched_coupling. */
int mdb_add_power_enriched_coupling(mdb_ctx* ctx, int op_id, const void* params, size_t params_bytes, int local_subproblem,
                                    int remote_subproblem, int first_coupling_child, int ncomp, int jac_mode);

/* ---- DOF map (ordering update; multidomaingridfunctionspace.hh:308-369) --- */
int mdb_finalize(mdb_ctx* ctx, int64_t* ndof);
/* parity exports.  cell_ptr has ncells+1 entries; cell_dofs has cell_ptr[ncells] entries: the container
 * indices of the MultiDomainLocalFunctionSpace bound to each cell, children concatenated
 * (multidomainlocalfunctionspace.hh:348-358).  Host pointers.  Pass cell_dofs=NULL to get sizes only. */
int mdb_get_dofmap(mdb_ctx* ctx, int64_t* cell_ptr, int64_t* cell_dofs);
int mdb_get_child_offsets(mdb_ctx* ctx, int64_t* offsets /* nchildren+1 */);

/* ---- constraints + interpolation (constraints.hh:539-558, interpolate.hh:76-81) */
/* constraints<R>(gfs, constrainSubProblem(sp_i, bc_i)...).assemble(cg): evaluates the Dirichlet set on the
 * device.  Returns the number of constrained DOFs in *nconstrained. */
int mdb_constraints_assemble(mdb_ctx* ctx, const int* subproblems, const mdb_bctype* bcs, int n,
                             int64_t* nconstrained);
/* Alternatively set the Dirichlet set from a host list of container indices. */
int mdb_set_constraints(mdb_ctx* ctx, int64_t n, const int64_t* dirichlet_dofs);
/* Export: one byte per DOF, 1 = constrained.  Host pointer. */
int mdb_get_constraints(mdb_ctx* ctx, uint8_t* flags);
/* interpolateOnTrialSpace(gfs, x, f_0, sp_0, f_1, sp_1, ...) at time t.  x: host or device, ndof doubles. */
int mdb_interpolate(mdb_ctx* ctx, const int* subproblems, const mdb_function* fns, int n, double time, double* x);
/* copy_nonconstrained_dofs(cu, xold, xnew)  (gridoperator.hh:146) */
int mdb_copy_nonconstrained(mdb_ctx* ctx, const double* xold, double* xnew);

/* ---- grid operators  (gridoperator.hh:155-175) ----------------------------- */
/* GridOperator(gfsu,gfsv,cu,cv,mb,participants...): participant order = argument order. Returns handle >= 0. */
int mdb_gridoperator_create(mdb_ctx* ctx, const int* subproblems, int nsubproblems,
                            const int* couplings, int ncouplings);
/* Jacobian container + fill_pattern (gridoperator.hh:83-87): builds the CSR pattern as the union of the
 * patterns of `gos` (OneStepGridOperator unions dt0 and dt1).  Returns a matrix handle >= 0. */
int mdb_matrix_create(mdb_ctx* ctx, const int* gos, int ngos, int64_t* nnz);
/* parity export in BCRS order (ascending columns per row).  Host pointers: rowptr ndof+1, colidx nnz. */
int mdb_matrix_get_pattern(mdb_ctx* ctx, int mat, int64_t* rowptr, int64_t* colidx);
int mdb_matrix_get_values(mdb_ctx* ctx, int mat, double* values /* host or device, nnz */);
int mdb_matrix_zero(mdb_ctx* ctx, int mat);         /* a = 0.0 */
/* Device pointers of the CSR arrays (int64 rowptr[ndof+1], int32 colidx[nnz], double values[nnz]) for
 * zero-copy consumers. */
int mdb_matrix_device_ptrs(mdb_ctx* ctx, int mat, const int64_t** rowptr, const int32_t** colidx, double** values);

/* setTime / setWeight fan-out (localassembler.hh:214-255) */
int mdb_set_time(mdb_ctx* ctx, double t);

/* GridOperator::residual(x, r) (gridoperator.hh:89-92): r += weight * R(x), then constrained entries of r
 * are set to 0 (residualassemblerengine.hh:553-567).  x, r: host or device. */
int mdb_residual(mdb_ctx* ctx, int go, double weight, const double* x, double* r);
/* GridOperator::jacobian(x, a) (gridoperator.hh:94-97): a += weight * J(x), then every constrained row is
 * replaced by the unit row (jacobianassemblerengine.hh:480-504).  overwrite != 0 treats `a` as zero on entry
 * (fuses the caller's "a = 0" into the assembly kernel). */
int mdb_jacobian(mdb_ctx* ctx, int go, int mat, double weight, const double* x, int overwrite);
/* Note on host vectors: the volume operators of the closed set are linear, so jacobian() reads x only through the
 * finite-difference coupling Jacobian, at the DOFs of the interface cells.  A host x is therefore uploaded on a second
 * stream while the volume kernels run; from pinned (page-locked) memory only the entries that are read are fetched
 * (zero-copy gather), from pageable memory the whole vector is copied.  Lifetime: a host vector is only read DURING the
 * call — every entry point waits for its host-to-device copy (or gather) before it returns, so the caller may overwrite
 * or free the buffer immediately afterwards; the kernels that consume the device copy keep running asynchronously. */

/* ---- linear solve ([UPSTREAM] ISTL solver backends) ----------- */
/* MDB_SOLVER_DIRECT: [UPSTREAM] ISTLBackend_SEQ_SuperLU (test/stokesdarcy2.cc:784-785) — a sparse direct solve.  Here: reverse
 * Cuthill-McKee reordering, banded LU with partial pivoting on the device (csrc/direct.cu), one step of iterative refinement;
 * `precond`, `reduction` and `maxit` are ignored, stats report the defects before / after.  Single-rank contexts. */
enum { MDB_SOLVER_CG = 0, MDB_SOLVER_BICGSTAB = 1, MDB_SOLVER_DIRECT = 2 };
/* Jacobi is the fast path.  SSOR and ILU0 reproduce the reference's sequential ISTL preconditioners (SeqSSOR(A, n, 1.0),
 * SeqILU0) operation for operation on one rank; on a partition they act as block-Jacobi over the ranks (each rank applies
 * the sequential preconditioner of its owned x owned block, so iteration counts depend on the rank count).  [UPSTREAM] ISTLBackend_SEQ_BCGS_SSOR = BiCGSTAB + SSOR(3),
 * ISTLBackend_SEQ_CG_SSOR = CG + SSOR(1)  (test/testpoisson.cc:291-292, test/explicitlycoupledpoisson.cc:297-313). */
/* MDB_PRECOND_MG: one geometric multigrid V-cycle per application (csrc/mg.cu; not in the reference, whose preconditioners are
 * the sequential ISTL sweeps above): coarse operators by re-discretisation on meshes with half the cells per axis (the coarse
 * contexts replay the set-up and the (grid operator, weight) log of the mdb_jacobian calls that built the matrix), per-child trilinear
 * transfers, Chebyshev-Jacobi smoothing.  Structured meshes; Q1 children (Poisson / L2 + flow couplings; also on a slab partition
 * after the collective mdb_mg_setup below) or QkDG children (ConvectionDiffusionDG, nested-space transfers; one rank); constraints
 * from mdb_constraints_assemble.  Anything else is MDB_EINVAL.  The V-cycle is symmetric positive definite: valid inside CG,
 * BiCGSTAB takes it as a right preconditioner. */
enum { MDB_PRECOND_NONE = 0, MDB_PRECOND_JACOBI = 1, MDB_PRECOND_ILU0 = 2, MDB_PRECOND_SSOR = 3, MDB_PRECOND_MG = 4 };
typedef struct mdb_solve_stats {
  int32_t iterations;
  int32_t converged;
  double defect0;          /* initial 2-norm of the defect */
  double defect;           /* final   2-norm of the defect */
  double seconds;          /* device time of the Krylov loop (CUDA events) */
} mdb_solve_stats;
/* Solve A z = rhs to  |defect| <= reduction * |defect0|  starting from z = 0.  rhs, z: host or device. */
/* ISTL_SEQ_Subblock_Backend<Preconditioner, Solver>::apply (test/testpoisson-dirichlet-neumann.cc:264-370), the linear solver of the
 * Dirichlet-Neumann iteration: solves raw(A)[child][child] z_child = rhs_child with the diagonal block of child space `child`
 * and leaves every other entry of z untouched.  The matrix must not couple the block with the others (assemble the block's own
 * grid operator — e.g. {left subproblem, its coupling} — into its own matrix, as the reference's LeftOperator / RightOperator
 * do); MDB_EINVAL otherwise.  Same stats / return codes as mdb_solve. */
int mdb_solve_block(mdb_ctx* ctx, int mat, int child, int method, int precond, double reduction, int maxit, const double* rhs, double* z,
                    mdb_solve_stats* stats);
int mdb_solve(mdb_ctx* ctx, int mat, int method, int precond, double reduction, int maxit,
              const double* rhs, double* z, mdb_solve_stats* stats);
/* Build the multigrid hierarchy of a matrix ahead of the first MDB_PRECOND_MG solve (single rank: optional).  On a slab partition this
 * is REQUIRED and collective: every rank coarsens its slab (owned layers [c0, c1) -> [c0/2, c1/2), one coarse ghost layer below), then
 * the halo windows of every coarse level are connected like the fine level's: mdb_mg_level returns the (borrowed) context of level
 * l >= 1, on which mdb_halo_export / mdb_halo_import are called; the coarse levels share the fine context's communicator. */
int mdb_mg_setup(mdb_ctx* ctx, int mat, int* nlevels);
mdb_ctx* mdb_mg_level(mdb_ctx* ctx, int mat, int level);
/* Levels of the multigrid hierarchy built for this matrix (0 before the first MDB_PRECOND_MG solve). */
int mdb_matrix_mg_levels(mdb_ctx* ctx, int mat);
/* Number of symmetric sweeps n of MDB_PRECOND_SSOR (default 3). */
int mdb_set_ssor_sweeps(mdb_ctx* ctx, int sweeps);
/* [UPSTREAM] PDELab StationaryLinearProblemSolver::apply(x) as the reference's drivers use it (test/testpoisson.cc:295-298,
 * test/testpoisson-dirichlet-neumann.cc:721-822): assemble J(x0) into `mat` and r = R(x0) at the incoming x0, solve J z = r to the
 * relative defect reduction, x <- x0 - z — in one call, on the device.  x (ndof doubles) is a host vector (uploaded once, solution
 * copied back before the call returns) or a device vector (updated in place).  keep_matrix != 0 reuses the values in `mat`.
 * Returns MDB_ENOTCONVERGED like mdb_solve; x is then left unchanged on the host. */
int mdb_stationary_solve(mdb_ctx* ctx, int go, int mat, int method, int precond, double reduction, int maxit, double* x, int keep_matrix,
                         mdb_solve_stats* stats);
/* Run exactly `iters` Krylov iterations without convergence test (benchmarking the loop). */
int mdb_solve_fixed(mdb_ctx* ctx, int mat, int method, int precond, int iters,
                    const double* rhs, double* z, mdb_solve_stats* stats);
/* One time step of the instationary problem (go0 = spatial operator, go1 = temporal/mass operator, mat created from both):
 * [UPSTREAM] OneStepMethod + OneStepGridOperator as in test/testinstationarypoisson.cc:320-403 with the diagonally implicit
 * schemes ImplicitEulerParameter, OneStepThetaParameter(theta), Alexander2Parameter (the reference driver's choice, :320).
 * Per stage: Dirichlet DOFs = interpolation of g[0..n) on the listed subproblems at the stage time, the other DOFs from the
 * previous stage (gridoperator.hh:139-153); J = a_rr J1 + b_rr dt J0; one Newton step with the linear solver given.
 * stats: iterations / seconds summed over the stages.  u_old, u_new: host or device, may not alias. */
enum { MDB_ONESTEP_IMPLICIT_EULER = 0, MDB_ONESTEP_THETA = 1, MDB_ONESTEP_ALEXANDER2 = 2 };
int mdb_onestep_apply(mdb_ctx* ctx, int go0, int go1, int mat, const int* subproblems, const mdb_function* g, int n,
                      double time, double dt, int scheme, double theta, int method, int precond, double reduction, int maxit,
                      const double* u_old, double* u_new, mdb_solve_stats* stats);
/* y = A x (host or device pointers).  On a partitioned context with imported neighbours x must be a device vector: its
 * ghost entries are refreshed by a halo exchange first. */
int mdb_spmv(mdb_ctx* ctx, int mat, const double* x, double* y);

/* ---- multi-GPU plumbing (one process and one context per GPU) -- */
/* NCCL communicator for the Krylov scalars: rank 0 creates the id, every rank calls mdb_comm_init.  With a communicator
 * of more than one rank, mdb_solve / mdb_spmv exchange halos before every SpMV and reduce dot products over owned rows. */
int mdb_nccl_unique_id(void* id128 /* 128 bytes */);
int mdb_comm_init(mdb_ctx* ctx, int nranks, int rank, const void* id128);
/* Halo receive window of this rank (device memory that the neighbours write with peer stores over NVLink).  Export an
 * 80-byte descriptor (cudaIpc handle; a raw pointer when both contexts live in one process), exchange the descriptors
 * out of band (e.g. torch.distributed.all_gather_object) and import the neighbours' ones.  Pass NULL for a missing
 * neighbour (first / last rank). */
int mdb_halo_export(mdb_ctx* ctx, void* blob80, int64_t* window_bytes);
int mdb_halo_import(mdb_ctx* ctx, int lower_rank, const void* lower_blob80, int upper_rank, const void* upper_blob80);
/* Peer-memory all-reduce of the Krylov scalars (optional; without it the scalars go through ncclAllReduce).  Every rank exports the
 * 80-byte descriptor of a small window (after mdb_comm_init), the descriptors of ALL ranks are exchanged out of band and imported in
 * rank order (the own entry is ignored).  From then on the dot products of mdb_solve end in a single kernel that stores this rank's
 * partial sums into every peer's window over NVLink and adds the contributions in rank order: the same bits on every rank, no
 * library call per reduction.  Connect before mdb_mg_setup (the coarse levels share the windows).  At most 16 ranks.  nranks = 0 disconnects
 * (every rank must do the same: the scalars then go through ncclAllReduce again).  The reference's
 * analogue is the MPI_Allreduce inside ISTL's OverlappingSchwarzScalarProduct (test/testoverlappinginstationarypoisson.cc:431-432). */
int mdb_scalars_export(mdb_ctx* ctx, void* blob80);
int mdb_scalars_import(mdb_ctx* ctx, int nranks, const void* blobs80);
/* number of peer-memory all-reduces issued so far on this context's windows (0: not connected, the scalars go through NCCL) */
int64_t mdb_scalars_exchanges(mdb_ctx* ctx);
/* The two halves of one exchange of the device vector x (ndof doubles): push this rank's boundary planes into the
 * neighbours' windows, then wait for the neighbours' planes and copy them into the ghost entries of x.  Every rank must
 * issue the same sequence of pushes and waits. */
int mdb_halo_push(mdb_ctx* ctx, const double* x);
int mdb_halo_wait(mdb_ctx* ctx, double* x);
/* Bound of the device-side wait for a neighbour's planes (default: environment MDB_HALO_TIMEOUT_S, else 30 s).  A timeout is
 * reported once by the call that hit it (MDB_ENCCL) — a running Krylov loop stops at once — and then cleared. */
int mdb_set_halo_timeout(mdb_ctx* ctx, double seconds);
/* Index-list partition: UNSTRUCTURED meshes, and QkDG spaces on a structured slab with two-sided ghost layers (mdb_mesh_partition).
 * Unstructured meshes (recursive coordinate bisection of the cell centroids, a DOF belongs to the lowest rank
 * among the cells of its entity; the reference's analogue is the overlap-1 MPI grid of test/testoverlappinginstationarypoisson.cc:232-233).
 * The context holds the local mesh of one rank: every cell around an owned vertex plus their face neighbours, in global order, so the
 * owned rows are assembled completely without communication.  This call declares the ownership of the local DOFs and, per neighbour
 * rank, the owned entries sent to it and the ghost entries received from it (both sides list them in the same order).  Call after
 * mdb_finalize; then mdb_halo_export / exchange / mdb_halo_import_lists as for slabs.  Up to 16 neighbours. */
int mdb_partition_lists(mdb_ctx* ctx, const uint8_t* owned_flags, int nneigh, const int* neigh_ranks, const int64_t* send_ptr, const int32_t* send_idx,
                        const int64_t* recv_ptr, const int32_t* recv_idx);
/* blobs80: the neighbours' 80-byte window descriptors in neighbour order; peer_cap[i] = doubles per parity of neighbour i's window (its
 * total receive count, at least 1), peer_offset[i] = where this rank's block starts inside it (neighbour i's recv_ptr of this rank),
 * peer_slot[i] = this rank's position in neighbour i's neighbour list. */
int mdb_halo_import_lists(mdb_ctx* ctx, const void* blobs80, const int64_t* peer_cap, const int64_t* peer_offset, const int* peer_slot);
/* Rows this rank owns / all local rows (ghost rows included). */
int mdb_owned_rows(mdb_ctx* ctx, int64_t* owned, int64_t* local_total);
/* Owned rows per child block (nchildren entries). */
int mdb_owned_counts(mdb_ctx* ctx, int64_t* per_child);
/* Export: one byte per local DOF, 1 = owned by this rank; and the global container index of every owned local DOF (-1
 * for ghosts) given the global index of this rank's first owned DOF of every child block (`child_global_base`,
 * nchildren entries: global child offset + owned rows of the lower ranks; NULL = local numbering).  The rank offsets describe the
 * one-rank numbering for Q1 blocks only (lexicographic, cut axis slowest); a Qk (k > 1) block is numbered entity group by entity
 * group, so with a non-NULL base the call is refused for such blocks: map local to global DOFs through the cell maps of
 * mdb_get_dofmap (local cell (i, j, k) = global cell (i, j, k + cell_offset), same local order). */
int mdb_get_ownership(mdb_ctx* ctx, const int64_t* child_global_base, uint8_t* owned_flags, int64_t* global_index);

/* Row classes of the matrix: "simple" rows carry exactly the full 3^dim own-block stencil (their values are streamed by
 * k_jacobian_fill_q1), all others are "irregular" (boundary, subdomain boundary, interface rows). */
int mdb_matrix_stats(mdb_ctx* ctx, int mat, int64_t* simple_rows, int64_t* irregular_rows);
/* Sum of squares of the stored entries of the matrix (device reduction; a cheap checksum of an assembly
 * result that a host caller can read back).  `out` is a host pointer. */
int mdb_matrix_norm2(mdb_ctx* ctx, int mat, double* out);

/* ---- instrumentation ------------------------------------------------------ */
/* Mean device time in ms of the named phase ("jacobian_volume", "jacobian_coupling", "residual_volume",
 * "residual_coupling", "residual_boundary", "pattern", "dofmap", "spmv", "solve", ...) over all its executions
 * since the last mdb_profile_reset, measured with CUDA events recorded on the context's stream (no host
 * synchronisation while recording; this call synchronises).  Negative if never run. */
double mdb_phase_ms(mdb_ctx* ctx, const char* phase);
int mdb_phase_count(mdb_ctx* ctx, const char* phase);
int mdb_profile_reset(mdb_ctx* ctx);
/* Number of kernel launches issued by this context since creation / since the last reset. */
int64_t mdb_launch_count(mdb_ctx* ctx, int reset);
/* Bytes moved between host and device by this context's calls since the last reset (staging copies of host vectors,
 * zero-copy gathers, result read-backs). */
int mdb_bytes_copied(mdb_ctx* ctx, int64_t* h2d, int64_t* d2h, int reset);
/* Raw CUDA stream of the context (cudaStream_t as void*). */
void* mdb_stream(mdb_ctx* ctx);
int mdb_synchronize(mdb_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* MDB200_H */
