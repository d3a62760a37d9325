"""Small driver for ncu: a few launches of the fused NB down-sweep (v | h1) and of
the up-sweep (h1 | v, h2) of the C3 model on `--chains` chains."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import __graft_entry__ as g
from bench import make_model, SEED
ap = argparse.ArgumentParser(); ap.add_argument("--chains", type=int, default=148 * 128 * 4); ap.add_argument("--sweeps", type=int, default=3)
a = ap.parse_args()
pkg = g.load_package()
ctx = pkg.Context(0, SEED)
L = make_model()
dm = pkg.DeviceModel.from_host([pkg.NegativeBinomialBernoulliRBM(**L[0]), pkg.BernoulliRBM(**L[1])], ctx)
p = pkg.DeviceParticles(dm, a.chains).init(False)
ctx.timer_start()
pkg.gibbssample_(p, dm, a.sweeps)
ms = ctx.timer_stop()
print(f"{a.sweeps} sweeps x {a.chains} chains: {ms:.3f} ms -> {a.sweeps*a.chains/ms*1e3:.3e} sweeps*chains/s; launches {ctx.launch_count()}")
