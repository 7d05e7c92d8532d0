"""Problem set-ups shared by the oracle and the CUDA backend (both export the same call shapes).

Every function takes an `api` object (tests/orc.Oracle or dune-multidomain_b200.Mdb) and drives it exactly the way
the reference's drivers build their problems.
"""
import importlib
from types import SimpleNamespace

capi = importlib.import_module("dune-multidomain_b200").capi
parallel = importlib.import_module("dune-multidomain_b200").parallel


def testpoisson(api, n, intensity=2.0, split_axis=1, fe=(capi.FE_Q1, capi.FE_Q1), quad_order=4, coupling=capi.OP_CVCF_COUPLING,
                jac_mode=capi.JAC_FD_BITMATCH, f=None, sd_bits=None, lower=None, upper=None, slab=None):
    """test/testpoisson.cc:119-298 (2-D as shipped; dim = len(n) gives the synthetic 3-D extension).

    Two subdomains split at x[split_axis] > 0.5 (:152-158), one Qk space per subdomain (:168-198), Poisson(f,b,j,4)
    on each (:218-232), ContinuousValueContinuousFlowCoupling(4, intensity) left->right (:234-237), Dirichlet
    constraints from DirichletBoundary (:242-246), u = interpolated G (:266).
    """
    dim = len(n)
    if slab is None:
        api.mesh_structured(n, lower, upper)
    else:                       # multi-rank: this rank's slab (+ ghost layer) of the global mesh n
        parallel.apply_slab(api, slab)
    if sd_bits is None:
        api.subdomains_halfspace(split_axis, 0.5, 0, 1)
    else:
        api.subdomains(sd_bits)
    c0 = api.add_space(fe[0], 0)
    c1 = api.add_space(fe[1], 1)
    pp = capi.PoissonParams()
    pp.f = f if f is not None else capi.Function(capi.FN_ZERO)          # testpoisson.cc:38: y = 0
    pp.bc = capi.BCType(capi.BC_TESTPOISSON)
    pp.j = capi.Function(capi.FN_TESTPOISSON_J)
    pp.quad_order = quad_order
    left = api.add_subproblem(capi.OP_POISSON, pp, capi.COND_EQUALITY, 1 << 0, [c0])
    right = api.add_subproblem(capi.OP_POISSON, pp, capi.COND_EQUALITY, 1 << 1, [c1])
    if coupling == capi.OP_CVCF_COUPLING:
        cpar = capi.CVCFParams()
        cpar.quad_order = 4
        cpar.intensity = intensity
    else:
        cpar = capi.PropFlowParams()
        cpar.intensity = intensity
    cp = api.add_coupling(coupling, cpar, left, right, jac_mode)
    ndof = api.finalize()
    ncon = api.constraints_assemble([left, right], [pp.bc, pp.bc])
    go = api.gridoperator_create([left, right], [cp])
    mat = api.matrix_create([go])
    g = capi.gauss()
    return SimpleNamespace(dim=dim, ndof=ndof, ncon=ncon, go=go, mat=mat, left=left, right=right, cp=cp, g=g, sps=[left, right],
                           gfn=[g, g], bc=pp.bc)


def single_domain_poisson(api, n, quad_order=2, fe=capi.FE_Q1, bc=capi.BC_ALL_DIRICHLET, f=None):
    """One subdomain covering everything, one subproblem, no coupling (test/testpoisson-assembly.cc:883-889 marks every
    cell into subdomain 0)."""
    import numpy as np
    api.mesh_structured(n)
    api.subdomains(np.ones(int(np.prod(n)), dtype=np.uint8))
    c0 = api.add_space(fe, 0)
    pp = capi.PoissonParams()
    pp.f = f if f is not None else capi.Function(capi.FN_ZERO)
    pp.bc = capi.BCType(bc)
    pp.j = capi.Function(capi.FN_ZERO)
    pp.quad_order = quad_order
    sp = api.add_subproblem(capi.OP_POISSON, pp, capi.COND_EQUALITY, 1, [c0])
    ndof = api.finalize()
    ncon = api.constraints_assemble([sp], [pp.bc])
    go = api.gridoperator_create([sp], [])
    mat = api.matrix_create([go])
    return SimpleNamespace(ndof=ndof, ncon=ncon, go=go, mat=mat, sp=sp, sps=[sp], bc=pp.bc)


def testpoisson_partitioned(api, n_per_rank, rank, nranks, **kw):
    """Weak-scaling variant of `testpoisson`: a global mesh of n_per_rank[-1] * nranks cell layers along the slowest axis
    (same mesh width everywhere: the domain grows with the rank count), cut into one slab per rank."""
    dim = len(n_per_rank)
    gn = list(n_per_rank)
    gn[-1] = n_per_rank[-1] * nranks
    upper = [1.0] * dim
    upper[-1] = float(nranks)
    s = parallel.slab(gn, rank, nranks, upper=upper)
    P = testpoisson(api, tuple(gn), slab=s, **kw)
    P.slab = s
    return P


def dirichlet_neumann_q1(api, n):
    """BASELINE configs[0] in its conforming-Q1 form: test/testpoisson-dirichlet-neumann.cc with the Q1
    alternative of :966-970 / what test/testpoisson-assembly.cc:930-934 assembles — 2-D unit square (ini: refine 9 =>
    512^2), split x0 > 0.5 (:919-925), Poisson(order 6) with the source F of :73-87 on both sides, monolithic coupling
    ContinuousValueContinuousFlowCoupling(4, 10) (:449, ini:10)."""
    return testpoisson(api, n, intensity=10.0, split_axis=0, quad_order=6, f=capi.Function(capi.FN_DN_SOURCE))


def testmortarpoisson(api, n, gamma_0=1.0, split_axis=0, fe=(capi.FE_Q1, capi.FE_Q1), jac_mode=capi.JAC_FD_BITMATCH, sd_bits=None, ncomp=1,
                      slab=None):
    """test/testmortarpoisson.cc:281-432 (BASELINE configs[3]; 2-D refine 5 as shipped, dim = len(n) for the synthetic 3-D
    form): subdomains x0 < 0.5 -> 0, x0 > 0.5 -> 1 (:325-330), a Q1 space on each (:359-362), a Q1 mortar space
    on the interface faces inside in {0} / outside in {1} as the LAST child (:365-384), Poisson(f,b,j,4) on both sides
    (:400-401; f = 0, :38), EnrichedCoupling<SubProblem0,SubProblem1,MortarPoissonCoupling,2>(gamma_0) (:411-412)."""
    dim = len(n)
    if slab is None:
        api.mesh_structured(n)
    else:                       # multi-rank: this rank's slab (+ ghost layer) of the global mesh n
        parallel.apply_slab(api, slab)
    if sd_bits is None:
        api.subdomains_halfspace(split_axis, 0.5, 0, 1)
    else:
        api.subdomains(sd_bits)
    c0 = api.add_space(fe[0], 0)
    c1 = api.add_space(fe[1], 1)
    # ncomp = 2: test/testpowermortarpoisson.cc — PowerCouplingGridFunctionSpace<CouplingGFS,2,VBE> (:384-386), the operator loops
    # over the component blocks (:186-189)
    if ncomp == 1:
        cc = api.add_coupling_space(capi.FE_Q1, (capi.COND_EQUALITY, 1 << 0), (capi.COND_EQUALITY, 1 << 1))
    else:
        cc = api.add_power_coupling_space(capi.FE_Q1, (capi.COND_EQUALITY, 1 << 0), (capi.COND_EQUALITY, 1 << 1), ncomp)
    pp = capi.PoissonParams()
    pp.f = capi.Function(capi.FN_ZERO)
    pp.bc = capi.BCType(capi.BC_TESTPOISSON)
    pp.j = capi.Function(capi.FN_TESTPOISSON_J)
    pp.quad_order = 4
    left = api.add_subproblem(capi.OP_POISSON, pp, capi.COND_EQUALITY, 1 << 0, [c0])
    right = api.add_subproblem(capi.OP_POISSON, pp, capi.COND_EQUALITY, 1 << 1, [c1])
    mp = capi.MortarParams()
    mp.gamma_0 = gamma_0
    if ncomp == 1:
        cp = api.add_enriched_coupling(capi.OP_MORTAR_POISSON_COUPLING, mp, left, right, cc, jac_mode)
    else:
        cp = api.add_power_enriched_coupling(capi.OP_MORTAR_POISSON_COUPLING, mp, left, right, cc, ncomp, jac_mode)
    ndof = api.finalize()
    ncon = api.constraints_assemble([left, right], [pp.bc, pp.bc])
    go = api.gridoperator_create([left, right], [cp])
    mat = api.matrix_create([go])
    g = capi.gauss()
    return SimpleNamespace(dim=dim, ndof=ndof, ncon=ncon, go=go, mat=mat, left=left, right=right, cp=cp, g=g, sps=[left, right],
                           gfn=[g, g], bc=pp.bc, children=(c0, c1, cc), ncomp=ncomp)


def dirichlet_neumann_dg(api, n, k=2, scaling=10.0, alpha=2.0, method=capi.DG_METHOD_SIPG, weights=capi.DG_WEIGHTS_ON, sd_bits=None, f=None,
                         g=None, bc=None, coupled=True, box=None):
    """BASELINE configs[0] AS SHIPPED: test/testpoisson-dirichlet-neumann.cc main() (:883-977) calls solve() with DGFEM2 =
    QkDGLocalFiniteElementMap<DF,double,2,dim> on both subdomains and dglop() = ConvectionDiffusionDG(params, SIPG, weightsOn, 2.0)
    (:864-881) over ConvectionDiffusionModelProblem (:169-285); the monolithic system (:449-456, :619-624) couples the two
    subproblems with ContinuousValueContinuousFlowCoupling(4, monolithic.coupling.scaling = 10) (ini:10).  Split x0 > 0.5 (:919-925);
    ini mesh.refine = 9 -> 512^2 cells, 2 * 512 * 256 * 9 = 2 359 296 DOFs.  ConformingDirichletConstraints on a DG space
    constrains nothing: the Dirichlet data enter through alpha_boundary."""
    dim = len(n)
    if box is None:
        api.mesh_structured(n)
    else:                       # multi-rank: this rank's sub-box (parallel.slab2: ghost cell layers on both sides) of the global mesh n
        parallel.apply_slab(api, box)
    if sd_bits is None:
        api.subdomains_halfspace(0, 0.5, 0, 1)
    else:
        api.subdomains(sd_bits)
    fe = capi.FE_DGQ1 if k == 1 else capi.FE_DGQ2
    c0, c1 = api.add_space(fe, 0), api.add_space(fe, 1)
    pp = capi.ConvDiffDGParams()
    pp.f = f if f is not None else capi.Function(capi.FN_DN_SOURCE)
    pp.bc = bc if bc is not None else capi.BCType(capi.BC_TESTPOISSON)
    pp.g = g if g is not None else capi.gauss()
    pp.j = capi.Function(capi.FN_TESTPOISSON_J)
    pp.method, pp.weights, pp.intorderadd, pp.alpha = method, weights, 0, alpha
    left = api.add_subproblem(capi.OP_CONVDIFF_DG, pp, capi.COND_EQUALITY, 1 << 0, [c0])
    right = api.add_subproblem(capi.OP_CONVDIFF_DG, pp, capi.COND_EQUALITY, 1 << 1, [c1])
    cps = []
    if coupled:
        cv = capi.CVCFParams()
        cv.quad_order, cv.intensity = 4, scaling
        cps = [api.add_coupling(capi.OP_CVCF_COUPLING, cv, left, right)]
    ndof = api.finalize()
    ncon = api.constraints_assemble([left, right], [pp.bc, pp.bc])
    go = api.gridoperator_create([left, right], cps)
    mat = api.matrix_create([go])
    return SimpleNamespace(dim=dim, ndof=ndof, ncon=ncon, go=go, mat=mat, left=left, right=right, g=pp.g, sps=[left, right],
                           gfn=[pp.g, pp.g], bc=pp.bc, children=(c0, c1), k=k)


def dirichlet_neumann(api, n, alpha=2.0, left_mode=capi.ND_MODE_DIRICHLET, right_mode=capi.ND_MODE_NEUMANN, fe=(capi.FE_Q1, capi.FE_Q1), quad_order=6,
                      scaling=10.0, dg=False):
    """BASELINE configs[0]: test/testpoisson-dirichlet-neumann.cc `solve()` (:371-862) instantiated with conforming Qk spaces and
    Poisson (what test/testpoisson-assembly.cc:919-934 assembles; the shipped main() passes DG-Q2 + ConvectionDiffusionDG, an
    upstream operator — DESIGN.md §9).  2-D unit square (ini: refine 9), split x0 > 0.5 (:919-925), Poisson with the source F of
    :73-87 on both sides, and three sets of participants on the same spaces:
      monolithic   left_sp, right_sp, Coupling(left, right, ContinuousValueContinuousFlowCoupling(4, scaling))      (:449-456, :619-624)
      left op      left_sp, Coupling(left, right, NDCoupling(mode = ini dn.coupling.right, SIPG, weightsOn, alpha))  (:517-528, :566-578)
      right op     right_sp, Coupling(right, left, NDCoupling(mode = ini dn.coupling.left, ...))                     (:505-515, :530-540, :580-592)
      full op      left_sp, dirichlet_coupling, right_sp, neumann_coupling                                           (:542-564)
    NOTE the reference's naming: `nd_coupling_dirichlet` (used by the LEFT operator) is built from ini key dn.coupling.RIGHT
    (= Neumann as shipped) and `nd_coupling_neumann` (RIGHT operator) from dn.coupling.LEFT (= Dirichlet); `left_mode` /
    `right_mode` here are the ini keys dn.coupling.left / dn.coupling.right.
    dg = True: the shipped instantiation — fe = QkDG spaces, both subproblems ConvectionDiffusionDG(params, SIPG, weightsOn, 2.0)
    (:864-881, :968-971)."""
    api.mesh_structured(n)
    api.subdomains_halfspace(0, 0.5, 0, 1)
    c0, c1 = api.add_space(fe[0], 0), api.add_space(fe[1], 1)
    if dg:
        pp = capi.ConvDiffDGParams()
        pp.f, pp.bc, pp.g, pp.j = capi.Function(capi.FN_DN_SOURCE), capi.BCType(capi.BC_TESTPOISSON), capi.gauss(), capi.Function(capi.FN_TESTPOISSON_J)
        pp.method, pp.weights, pp.intorderadd, pp.alpha = capi.DG_METHOD_SIPG, capi.DG_WEIGHTS_ON, 0, 2.0
        op = capi.OP_CONVDIFF_DG
    else:
        pp = capi.PoissonParams()
        pp.f = capi.Function(capi.FN_DN_SOURCE)
        pp.bc = capi.BCType(capi.BC_TESTPOISSON)
        pp.j = capi.Function(capi.FN_TESTPOISSON_J)
        pp.quad_order = quad_order
        op = capi.OP_POISSON
    left = api.add_subproblem(op, pp, capi.COND_EQUALITY, 1 << 0, [c0])
    right = api.add_subproblem(op, pp, capi.COND_EQUALITY, 1 << 1, [c1])
    cv = capi.CVCFParams()
    cv.quad_order = 4
    cv.intensity = scaling
    mono = api.add_coupling(capi.OP_CVCF_COUPLING, cv, left, right)

    def nd(mode):
        p = capi.NDCouplingParams()
        p.mode, p.method, p.weights, p.intorderadd, p.alpha = mode, capi.DG_METHOD_SIPG, capi.DG_WEIGHTS_ON, 0, alpha
        return p
    dirichlet_coupling = api.add_coupling(capi.OP_ND_COUPLING, nd(right_mode), left, right)      # local = left
    neumann_coupling = api.add_coupling(capi.OP_ND_COUPLING, nd(left_mode), right, left)         # local = right
    ndof = api.finalize()
    ncon = api.constraints_assemble([left, right], [pp.bc, pp.bc])
    go = api.gridoperator_create([left, right], [mono])
    full = api.gridoperator_create([left, right], [dirichlet_coupling, neumann_coupling])
    lop = api.gridoperator_create([left], [dirichlet_coupling])
    rop = api.gridoperator_create([right], [neumann_coupling])
    mat = api.matrix_create([go])
    lmat = api.matrix_create([lop])
    rmat = api.matrix_create([rop])
    g = capi.gauss()
    return SimpleNamespace(ndof=ndof, ncon=ncon, go=go, full=full, lop=lop, rop=rop, mat=mat, lmat=lmat, rmat=rmat, left=left, right=right,
                           children=(c0, c1), sps=[left, right], gfn=[g, g], bc=pp.bc)


def dn_iteration(api, P, u, iterations=200, damping=0.75, reduction=1e-8, maxerror=1e-99, left_reduction=(1e-4, 1e-4, 0.2),
                 right_reduction=(1e-4, 1e-4, 0.2), solve_block=None, history=None):
    """The Dirichlet-Neumann loop of test/testpoisson-dirichlet-neumann.cc:721-822 with the ini's parameters
    (poisson-dirichlet-neumann.ini [dn.*]): per sweep a StationaryLinearProblemSolver step with the left operator (BiCGSTAB +
    SSOR(3) on diagonal block 0, ISTL_SEQ_Subblock_Backend :264-370), damping, the same with the right operator on block 1,
    damping, then the residual of the full operator.  reassemble = false: the Jacobians are assembled in the first sweep only
    (apply(u, keep_matrix = i > 0), :806, :813).  *_reduction = (solver.reduction, solver.minreduction, solver.reductiondecay)."""
    import numpy as np
    if solve_block is None:
        solve_block = api.solve_block
    res = np.zeros(P.ndof)
    api.residual(P.full, u, res)
    r = start_r = float(np.linalg.norm(res))
    l_off = 1.0 / (1.0 - left_reduction[0] / left_reduction[1]) if left_reduction[0] != left_reduction[1] else float("inf")
    r_off = 1.0 / (1.0 - right_reduction[0] / right_reduction[1]) if right_reduction[0] != right_reduction[1] else float("inf")
    i = 0
    while r / start_r > reduction and r > maxerror and i < iterations:
        for (op, mat, child, red, off) in ((P.lop, P.lmat, P.children[0], left_reduction, l_off), (P.rop, P.rmat, P.children[1], right_reduction, r_off)):
            u_old = u.copy()
            red_i = red[1] * (1.0 - 1.0 / (off + red[2] * i))
            # StationaryLinearProblemSolver::apply(x, keep_matrix) [UPSTREAM] (vii): r = R(x); solve J z = r; x -= z
            if i == 0:
                api.jacobian(op, mat, u, overwrite=True)
            rr = np.zeros(P.ndof)
            api.residual(op, u, rr)
            z = np.zeros(P.ndof)
            solve_block(mat, child, rr, z, reduction=red_i)
            u = u - z
            u = damping * u + (1.0 - damping) * u_old                       # u *= alpha; u.axpy(1 - alpha, u_old)
        res[:] = 0.0
        api.residual(P.full, u, res)
        r = float(np.linalg.norm(res))
        i += 1
        if history is not None:
            history.append(r)
    return u, i, r


def instationary(api, n, k=2.0, split_axis=0):
    """BASELINE configs[2] in the form the reference's live code supports (SURVEY M3 / §8d reality check 3):
    test/testinstationarypoisson.cc:225-305 — dt0 operator = Poisson (stiffness + source + Neumann) on both subdomains plus
    ProportionalFlowCoupling(k) (test/proportionalflowcoupling.hh:10-91), dt1 operator = L2 mass, one matrix whose pattern is
    the union of both (OneStepGridOperator), time-dependent Dirichlet data g(t) (:94-100 family)."""
    api.mesh_structured(n)
    api.subdomains_halfspace(split_axis, 0.5, 0, 1)
    c0, c1 = api.add_space(capi.FE_Q1, 0), api.add_space(capi.FE_Q1, 1)
    pp = capi.PoissonParams()
    pp.f = capi.Function(capi.FN_DN_SOURCE)
    pp.bc = capi.BCType(capi.BC_TESTPOISSON)
    pp.j = capi.Function(capi.FN_TESTPOISSON_J)
    pp.quad_order = 4
    l = api.add_subproblem(capi.OP_POISSON, pp, capi.COND_EQUALITY, 1, [c0])
    r = api.add_subproblem(capi.OP_POISSON, pp, capi.COND_EQUALITY, 2, [c1])
    lp = capi.L2Params()
    lp.quad_order = 2
    ml = api.add_subproblem(capi.OP_L2, lp, capi.COND_EQUALITY, 1, [c0])
    mr = api.add_subproblem(capi.OP_L2, lp, capi.COND_EQUALITY, 2, [c1])
    cv = capi.PropFlowParams()
    cv.intensity = k
    cp = api.add_coupling(capi.OP_PROPFLOW_COUPLING, cv, l, r)
    ndof = api.finalize()
    ncon = api.constraints_assemble([l, r], [pp.bc, pp.bc])
    go0 = api.gridoperator_create([l, r], [cp])
    go1 = api.gridoperator_create([ml, mr], [])
    mat = api.matrix_create([go0, go1])
    g = capi.Function(capi.FN_INSTAT_G, 1.0, 3.0)
    return SimpleNamespace(ndof=ndof, ncon=ncon, go0=go0, go1=go1, mat=mat, sps=[l, r], gfn=[g, g])


def implicit_euler_step(api, P, u_old, t_new, dt, solve):
    """One step of the one-step-theta scheme with theta = 1 as OneStepMethod / OneStepGridOperator drive the two grid
    operators: new Dirichlet values by interpolation + copy_nonconstrained_dofs (gridoperator.hh:139-153),
    R(u) = r1(u) - r1(u_old) + dt r0(u), J = J1 + dt J0 (setWeight, localassembler.hh:220-223), u_new = u - J^-1 R(u)."""
    import numpy as np
    api.set_time(t_new)
    u = np.zeros(P.ndof)
    api.interpolate(P.sps, P.gfn, u, time=t_new)
    api.copy_nonconstrained(u_old, u)
    api.jacobian(P.go1, P.mat, u, weight=1.0, overwrite=True)
    api.jacobian(P.go0, P.mat, u, weight=dt, overwrite=False)
    R = np.zeros(P.ndof)
    api.residual(P.go1, u, R, weight=1.0)
    api.residual(P.go1, u_old, R, weight=-1.0)
    api.residual(P.go0, u, R, weight=dt)
    z = np.zeros(P.ndof)
    solve(P.mat, R, z)
    return u - z


def one_step_method(api, P, u_old, t, dt, solve, scheme="implicit_euler", theta=1.0):
    """[UPSTREAM] PDELab::OneStepMethod::apply(time, dt, xold, f, xnew) for the diagonally implicit parameter sets
    ImplicitEulerParameter / OneStepThetaParameter / Alexander2Parameter (test/testinstationarypoisson.cc:320-322, :373),
    driven through the two grid operators with weights exactly as OneStepGridOperator does.  Stage r solves
    sum_{i<r} [a_ri r1(x_i) + b_ri dt r0(x_i)] + a_rr r1(x) + b_rr dt r0(x) = 0 with one Newton step."""
    import math
    import numpy as np
    if scheme == "implicit_euler":
        a, b, d = [[-1.0, 1.0]], [[0.0, 1.0]], [0.0, 1.0]
    elif scheme == "theta":
        a, b, d = [[-1.0, 1.0]], [[1.0 - theta, theta]], [0.0, 1.0]
    else:
        al = 1.0 - 0.5 * math.sqrt(2.0)
        a, b, d = [[-1.0, 1.0, 0.0], [-1.0, 0.0, 1.0]], [[0.0, al, 0.0], [0.0, 1.0 - al, al]], [0.0, al, 1.0]
    xs = [np.ascontiguousarray(u_old)]
    for r in range(1, len(a) + 1):
        tr = t + d[r] * dt
        api.set_time(tr)
        x = np.zeros(P.ndof)
        api.interpolate(P.sps, P.gfn, x, time=tr)
        api.copy_nonconstrained(xs[r - 1], x)
        api.jacobian(P.go1, P.mat, x, weight=a[r - 1][r], overwrite=True)
        api.jacobian(P.go0, P.mat, x, weight=b[r - 1][r] * dt, overwrite=False)
        R = np.zeros(P.ndof)
        for i in range(r + 1):
            xi = x if i == r else xs[i]
            api.set_time(t + d[i] * dt)
            if a[r - 1][i] != 0.0:
                api.residual(P.go1, xi, R, weight=a[r - 1][i])
            if b[r - 1][i] != 0.0:
                api.residual(P.go0, xi, R, weight=b[r - 1][i] * dt)
        api.set_time(tr)
        z = np.zeros(P.ndof)
        solve(P.mat, R, z)
        xs.append(x - z)
    return xs[-1]
