"""Launch counts and CUDA-event timings of the small-problem training paths (BASELINE configs C1 and C2)."""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
import bench
ap = argparse.ArgumentParser()
ap.add_argument("--only", default="", help="c1 | c2 | ops (default: all)")
a = ap.parse_args()
pkg = g.load_package()
ctx = pkg.Context(0, bench.SEED)
x1 = bench.synth_counts(500, 1000, seed=1)
xm = pkg.Mat.from_host(ctx, x1, pkg.MAT_COUNT)
m = pkg.DeviceModel(ctx, [1000, 128], pkg.VIS_NEGBIN)
m.init_layer(0, xm, seed=bench.SEED)
for pcd in ((False, True) if a.only in ("", "c1") else ()):
    tr = pkg.RBMTrainer(m, 100, pcd=pcd)
    bench._spin_up(pkg, ctx, m)                # GPU clocks ramp up from idle
    for _ in range(3):
        pkg.trainrbm_(m, xm, tr, 1e-4, 1)
    ctx.sync(); ctx.launch_count(reset=True)
    best = 1e30
    for _ in range(3):
        ctx.sync(); ctx.timer_start()
        for _ in range(20):
            pkg.trainrbm_(m, xm, tr, 1e-4, 1)
        best = min(best, ctx.timer_stop())
    print(f"C1 pcd={pcd}: {20 / best * 1e3:.0f} epochs/s, {best / 100 * 1e3:.1f} us per minibatch, {ctx.launch_count() / 300:.1f} launches per minibatch")
if a.only not in ("", "c2", "ops"):
    sys.exit(0)
x2 = bench.synth_counts(5000, 2000, seed=2)
L = bench.make_model()
dm = pkg.DeviceModel.from_host([pkg.NegativeBinomialBernoulliRBM(**L[0]), pkg.BernoulliRBM(**L[1])], ctx)
x2m = pkg.Mat.from_host(ctx, x2, pkg.MAT_COUNT)
p = pkg.DeviceParticles(dm, 100).init(False)
if a.only in ("", "c2"):
    bench._spin_up(pkg, ctx, dm)
    pkg.traindbm_(dm, x2m, epochs=3, particles=p, learningrate=1e-4, device=True, ctx=ctx)
    ctx.sync(); ctx.launch_count(reset=True)
    best = 1e30
    for _ in range(3):
        ctx.sync(); ctx.timer_start()
        pkg.traindbm_(dm, x2m, epochs=20, particles=p, learningrate=1e-4, device=True, ctx=ctx)
        best = min(best, ctx.timer_stop())
    print(f"C2: {20 / best * 1e3:.0f} steps/s, {best / 20 * 1e3:.0f} us per step, {ctx.launch_count() / 60:.1f} launches per step")
    # how many mean-field iterations does a step run at this point of training?  (launched in groups of 4)
    import ctypes as C
    mu = pkg.DeviceParticles(dm, x2m.rows, real_valued=True)
    opt2 = pkg.LoglikelihoodOptimizer().initialized(dm)
    it = C.c_int32()
    pkg._lib.check(pkg.lib().scdbm_dbm_pcd_step(dm.h, x2m.h, p.h, mu.h, opt2.h, 1e-4, 5, 0.001, C.byref(it)))
    print(f"    mean-field iterations in the next step: {it.value}")
if a.only not in ("", "ops"):
    sys.exit(0)
# per-operator GPU time at the C1 minibatch shape (back-to-back launches, CUDA events)
import numpy as np
lib, ck = pkg.lib(), pkg._lib.check
v = pkg.Mat.from_host(ctx, x1[:100], pkg.MAT_COUNT)
h = pkg.Mat(ctx, 100, 128, pkg.MAT_REAL); hs = pkg.Mat.from_host(ctx, (np.random.default_rng(0).random((100, 128)) < 0.5).astype(float), pkg.MAT_BINARY)
vm = pkg.Mat(ctx, 100, 1000, pkg.MAT_REAL)
opt = pkg.loglikelihoodoptimizer(m, learningrate=1e-4)
def timed(name, f, reps=200):
    for _ in range(5): f()
    ctx.sync(); ctx.timer_start()
    for _ in range(reps): f()
    print(f"  {name}: {ctx.timer_stop() / reps * 1e3:.1f} us")
for eng, en in ((0, "auto"), (1, "simt"), (2, "tc")):
    ctx.set("engine", eng)
    print("engine", en)
    timed("hidden potential 100x1000->128", lambda: ck(lib.scdbm_hidden(m.h, 0, v.h, h.h, 1.0, pkg.OUT_POTENTIAL, 0, None, 0)))
    timed("visible potential 100x128->1000", lambda: ck(lib.scdbm_visible(m.h, 0, hs.h, vm.h, 1.0, pkg.OUT_POTENTIAL, 0, None, 0)))
    timed("hidden potential (real in)", lambda: ck(lib.scdbm_hidden(m.h, 0, vm.h, h.h, 1.0, pkg.OUT_POTENTIAL, 0, None, 0)))
    timed("gradient", lambda: ck(lib.scdbm_compute_gradient(opt.h, m.h, 0, v.h, vm.h, h.h, h.h, 0)))
ctx.set("engine", 0)
timed("update", lambda: ck(lib.scdbm_update_parameters(m.h, opt.h, -1, 0.0)))
