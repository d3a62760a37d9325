import sys, os, faulthandler
faulthandler.enable()
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import __graft_entry__ as g
pkg = g.load_package()
ctx = pkg.Context(0, 1)
for rows in (20000, 40000, 400000):
    m = pkg.Mat(ctx, rows, 2000, pkg.MAT_COUNT)
    gm = np.full(2000, 0.5); r = np.full(2000, 0.5)
    m.generate_counts(gm, r, 0.3, seed=1)
    print("gen ok", rows, flush=True)
    out = torch.empty((rows, 2000), dtype=torch.uint16, pin_memory=True).numpy()
    m.to_host_u16(out); print("pinned ok", rows, int(out.sum(dtype=np.int64)), flush=True)
    o2 = m.to_host_u16(); print("pageable ok", rows, int(o2.sum(dtype=np.int64)), flush=True)
