"""Per-kernel CUDA-event timings of the C3 sweep (1 M chains by default): the NB down-sweep (v | h1), the dual-source
up-sweep (h1 | v, h2), the top layer (h2 | h1), and a whole sweep.  Used while tuning; bench.py reports the same
figures in its roofline block."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
from bench import make_model, SEED, UNITS

ap = argparse.ArgumentParser()
ap.add_argument("--chains", type=int, default=1_000_000)
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--sweeps", type=int, default=10)
a = ap.parse_args()
pkg = g.load_package()
ctx = pkg.Context(0, SEED)
L = make_model()
dm = pkg.DeviceModel.from_host([pkg.NegativeBinomialBernoulliRBM(**L[0]), pkg.BernoulliRBM(**L[1])], ctx)
n = a.chains
p = pkg.DeviceParticles(dm, n).init(False)
pkg.gibbssample_(p, dm, 4)
lib, ck = pkg.lib(), pkg._lib.check
v0, h1, h2 = p.layer(0), p.layer(1), p.layer(2)
vout, h1out, h2out = pkg.Mat(ctx, n, UNITS[0], pkg.MAT_COUNT), pkg.Mat(ctx, n, UNITS[1], pkg.MAT_BINARY), pkg.Mat(ctx, n, UNITS[2], pkg.MAT_BINARY)


def timed(f):
    for i in range(2):
        f(i)
    ctx.timer_start()
    for i in range(a.reps):
        f(100 + i)
    return ctx.timer_stop() / a.reps


t_nb = timed(lambda i: ck(lib.scdbm_visible(dm.h, 0, h1.h, vout.h, 1.0, pkg.OUT_SAMPLE, 1000 + i, None, 0)))
t_up = timed(lambda i: ck(lib.scdbm_dbm_layer(dm.h, 1, v0.h, h2.h, h1out.h, pkg.OUT_SAMPLE, 2000 + i, None, 0)))
t_top = timed(lambda i: ck(lib.scdbm_hidden(dm.h, 1, h1.h, h2out.h, 1.0, pkg.OUT_SAMPLE, 3000 + i, None, 0)))
ctx.timer_start()
pkg.gibbssample_(p, dm, a.sweeps)
t_sweep = ctx.timer_stop() / a.sweeps
fl_nb, fl_up = 2.0 * n * UNITS[0] * UNITS[1], 2.0 * n * UNITS[1] * (UNITS[0] + UNITS[2])
print(f"chains {n}: NB down-sweep {t_nb:.3f} ms ({n * UNITS[0] / t_nb * 1e3:.3e} draws/s, {fl_nb / t_nb / 1e9:.1f} TFLOP/s), "
      f"up-sweep {t_up:.3f} ms ({fl_up / t_up / 1e9:.1f} TFLOP/s), top layer {t_top:.3f} ms, "
      f"whole sweep {t_sweep:.3f} ms -> {n / t_sweep * 1e3:.3e} sweeps*chains/s")
