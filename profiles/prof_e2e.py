"""Times the pieces of bench.py's end-to-end step (host model in, host uint16 cells out)."""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
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
