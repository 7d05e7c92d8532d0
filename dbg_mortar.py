import sys; sys.path.insert(0,'tests'); sys.path.insert(0,'.')
import importlib, numpy as np, orc, problems, collections
pkg = importlib.import_module("dune-multidomain_b200"); capi = pkg.capi
def report(tag, vd, vo, rp, ci, off, n):
    rows = np.repeat(np.arange(n), np.diff(rp))
    bad = np.isnan(vd) | (np.abs(vd - vo) > 1e-12 * np.abs(vo).max())
    blk = lambda i: np.searchsorted(off, i, side='right') - 1
    print(tag, "bad", bad.sum(), "nan", np.isnan(vd).sum(), dict(collections.Counter(zip(blk(rows[bad]).tolist(), blk(ci[bad]).tolist()))))
for kw in (dict(n=(2, 2, 2)),):
    dev, ora = pkg.Mdb(0), orc.Oracle()
    Pd, Po = problems.testmortarpoisson(dev, **kw), problems.testmortarpoisson(ora, **kw)
    gd = dev.gridoperator_create([Pd.left, Pd.right], []); go = ora.gridoperator_create([Po.left, Po.right], [])
    x = np.zeros(Po.ndof); ora.interpolate(Po.sps, Po.gfn, x)
    rp, ci = ora.matrix_get_pattern(Po.mat); off = ora.get_child_offsets()
    dev.jacobian(gd, Pd.mat, x, overwrite=True); ora.jacobian(go, Po.mat, x, overwrite=True)
    report("volume only", dev.matrix_get_values(Pd.mat), ora.matrix_get_values(Po.mat), rp, ci, off, Po.ndof)
    dev.jacobian(Pd.go, Pd.mat, x, overwrite=True); ora.jacobian(Po.go, Po.mat, x, overwrite=True)
    report("full", dev.matrix_get_values(Pd.mat), ora.matrix_get_values(Po.mat), rp, ci, off, Po.ndof)
    dev.matrix_zero(Pd.mat); ora.matrix_zero(Po.mat)
    dev.jacobian(Pd.go, Pd.mat, x, overwrite=False); ora.jacobian(Po.go, Po.mat, x, overwrite=False)
    report("full additive", dev.matrix_get_values(Pd.mat), ora.matrix_get_values(Po.mat), rp, ci, off, Po.ndof)
