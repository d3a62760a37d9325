"""Times the pieces of bench.py's end-to-end step (host model in, host uint16 cells out)."""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import numpy as np
import __graft_entry__ as g
from bench import make_model, SEED, UNITS
ap = argparse.ArgumentParser(); ap.add_argument("--chains", type=int, default=1000000); ap.add_argument("--burnin", type=int, default=50)
a = ap.parse_args()
pkg = g.load_package()
ctx = pkg.Context(0, SEED)
L = make_model()
hm = [pkg.NegativeBinomialBernoulliRBM(**L[0]), pkg.BernoulliRBM(**L[1])]
out = torch.empty((a.chains, UNITS[0]), dtype=torch.uint16, pin_memory=True).numpy()
def T(label, f):
    ctx.sync(); t = time.perf_counter(); r = f(); ctx.sync(); print(f"{label:28s} {1e3*(time.perf_counter()-t):9.2f} ms", flush=True); return r
for rep in range(3):
    print("rep", rep)
    m = T("DeviceModel.from_host", lambda: pkg.DeviceModel.from_host(hm, ctx))
    q = T("DeviceParticles alloc", lambda: pkg.DeviceParticles(m, a.chains, 0))
    T("init", lambda: q.init(False))
    T("gibbs x burnin", lambda: pkg.gibbssample_(q, m, a.burnin))
    T("to_host_u16", lambda: q.layer(0).to_host_u16(out))
    T("close", lambda: (q.close(), m.close()))

cap = int(0.3 * a.chains * UNITS[0])
ib = torch.empty(cap, dtype=torch.uint16, pin_memory=True).numpy()
vb = torch.empty(cap, dtype=torch.uint16, pin_memory=True).numpy()
for rep in range(2):
    ptr, idx, val = T("to_host_csr (pinned buffers)", lambda: q.layer(0).to_host_csr(buffers=(ib, vb))) if False else (None, None, None)
m = pkg.DeviceModel.from_host(hm, ctx)
q = pkg.DeviceParticles(m, a.chains, 0).init(False)
pkg.gibbssample_(q, m, 5)
for rep in range(2):
    ptr, idx, val = T("to_host_csr (pinned buffers)", lambda: q.layer(0).to_host_csr(buffers=(ib, vb)))
print("nnz", len(idx), "density", len(idx) / (a.chains * UNITS[0]), "bytes", idx.nbytes + val.nbytes + ptr.nbytes)
dense = q.layer(0).to_host_u16(out)
r = 12345
row = np.zeros(UNITS[0], dtype=np.uint16); row[idx[ptr[r]:ptr[r+1]]] = val[ptr[r]:ptr[r+1]]
assert (row == dense[r]).all()

# ---- pipelined loop as in bench.py, host timestamps per phase (no extra syncs)
outs = [out, torch.empty((a.chains, UNITS[0]), dtype=torch.uint16, pin_memory=True).numpy()]
del ib, vb
ctx.sync()
for i in range(4):
    t = [time.perf_counter()]
    m = pkg.DeviceModel.from_host(hm, ctx); t.append(time.perf_counter())
    q = pkg.DeviceParticles(m, a.chains, 0); t.append(time.perf_counter())
    q.init(False); pkg.gibbssample_(q, m, a.burnin); t.append(time.perf_counter())
    q.layer(0).to_host_u16(outs[i % 2], wait=False); t.append(time.perf_counter())
    q.close(); m.close(); t.append(time.perf_counter())
    print("pipelined step", i, " ".join(f"{1e3*(b-a_):8.2f}" for a_, b in zip(t, t[1:])), "(from_host, alloc, launch, dl_async, close) ms", flush=True)
t0 = time.perf_counter(); ctx.wait_downloads(); ctx.sync(); print("final wait", 1e3 * (time.perf_counter() - t0))
