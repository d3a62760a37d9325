"""Per-part CUDA-event timings of the C2 step (5000 x 2000 -> 256 -> 64, 100 particles) in ONE process: the whole step,
the Gibbs sweeps, mean-field, gradient and update alone.  C2 is bimodal from process to process (see
r02_ncu_summary.md); running this several times shows which part carries the difference."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
import bench
pkg = g.load_package()
ctx = pkg.Context(0, bench.SEED)
x2 = bench.synth_counts(5000, 2000, seed=2)
L = bench.make_model()
dm = pkg.DeviceModel.from_host([pkg.NegativeBinomialBernoulliRBM(**L[0]), pkg.BernoulliRBM(**L[1])], ctx)
x2m = pkg.Mat.from_host(ctx, x2, pkg.MAT_COUNT)
p = pkg.DeviceParticles(dm, 100).init(False)
bench._spin_up(pkg, ctx, dm)
lib, ck = pkg.lib(), pkg._lib.check
mu = pkg.DeviceParticles(dm, x2m.rows, real_valued=True)
opt = pkg.LoglikelihoodOptimizer().initialized(dm)


def timed(f, reps=40):
    for _ in range(5):
        f()
    best = 1e30
    for _ in range(3):
        ctx.sync(); ctx.timer_start()
        for _ in range(reps):
            f()
        best = min(best, ctx.timer_stop() / reps)
    return best * 1e3


t_step = timed(lambda: ck(lib.scdbm_dbm_pcd_step(dm.h, x2m.h, p.h, mu.h, opt.h, 1e-4, 5, 0.001, None)))
t_gibbs = timed(lambda: pkg.gibbssample_(p, dm, 5))
t_mf = timed(lambda: ck(lib.scdbm_meanfield(dm.h, x2m.h, 0.001, 0, mu.h, None, None)))
t_grad = timed(lambda: ck(lib.scdbm_compute_gradient_dbm(opt.h, dm.h, mu.h, p.h, 0.0, 0.0)))
t_upd = timed(lambda: ck(lib.scdbm_update_parameters(dm.h, opt.h, -1, 0.0)))
print(f"step {t_step:.0f} us = gibbs {t_gibbs:.0f} + mean-field {t_mf:.0f} + gradient {t_grad:.0f} + update {t_upd:.0f} (sum {t_gibbs + t_mf + t_grad + t_upd:.0f})")
