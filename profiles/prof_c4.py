"""Driver for ncu: the deterministic C4 kernels (mean-field at 100 000 cells, gradient) -- one call each."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import __graft_entry__ as g
import bench
pkg = g.load_package()
ctx = pkg.Context(0, bench.SEED)
U = bench.C4_UNITS
L = bench.make_model(U, seed=4)
dm = pkg.DeviceModel.from_host([pkg.NegativeBinomialBernoulliRBM(**L[0])] + [pkg.BernoulliRBM(**l) for l in L[1:]], ctx)
cells, parts = int(os.environ.get("C4_CELLS", 100000)), 8192
gm = np.clip(np.exp(np.random.default_rng(4).normal(-1.5, 1.5, U[0])), 1e-3, 50.0)
x = pkg.Mat(ctx, cells, U[0], pkg.MAT_COUNT).generate_counts(gm, np.full(U[0], 0.5), 0.3, seed=4)
p = pkg.DeviceParticles(dm, parts).init(False)
mu = pkg.DeviceParticles(dm, cells, real_valued=True)
opt = pkg.LoglikelihoodOptimizer().initialized(dm)
lib, ck = pkg.lib(), pkg._lib.check
for _ in range(2):
    ck(lib.scdbm_meanfield(dm.h, x.h, 0.0, 2, mu.h, None, None))
    ck(lib.scdbm_compute_gradient_dbm(opt.h, dm.h, mu.h, p.h, 0.0, 0.0))
    pkg.gibbssample_(p, dm, 1)
ctx.sync()
print("ok")
