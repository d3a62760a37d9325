"""Latency of one tiny sweep (M = 100 rows) as a function of the contraction length: fixed cost vs per-k-block cost."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import __graft_entry__ as g
pkg = g.load_package()
ctx = pkg.Context(0, 1)
ctx.set("engine", 2)
lib, ck = pkg.lib(), pkg._lib.check
rng = np.random.default_rng(0)
for V in (64, 256, 1000, 4000):
    for H in (128,):
        m = pkg.DeviceModel(ctx, [V, H], pkg.VIS_NEGBIN)
        x = rng.poisson(0.5, (100, V)).astype(float); x[0] += 1
        xm = pkg.Mat.from_host(ctx, x, pkg.MAT_COUNT)
        m.init_layer(0, xm, seed=1)
        h = pkg.Mat(ctx, 100, H, pkg.MAT_REAL)
        for what, name in ((pkg.OUT_INPUT, "input"), (pkg.OUT_POTENTIAL, "sigm")):
            f = lambda: ck(lib.scdbm_hidden(m.h, 0, xm.h, h.h, 1.0, what, 0, None, 0))
            for _ in range(5): f()
            ctx.sync(); ctx.timer_start()
            for _ in range(200): f()
            t = ctx.timer_stop() / 200 * 1e3
            print(f"V={V} H={H} {name}: {t:.1f} us")
