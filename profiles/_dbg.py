import sys; sys.path.insert(0,'/root/repo')
import numpy as np
import __graft_entry__ as g
pkg = g.load_package()
from oracle import model as om
rg = np.random.default_rng(12)
V,H,n = 12,6,4096
means = np.array([.05,.1,.2,.4,.8,1.5,3.,6.,12.,0.3,0.6,25.])
W = -rg.uniform(0.05,0.5,(V,H)); a = np.log(means/(means+0.5)); b = rg.normal(0,.5,H)
res = {}
for eng in (1,2):
    ctx = pkg.Context(0, 77); ctx.set("engine", eng)
    m = pkg.DeviceModel.from_host(pkg.NegativeBinomialBernoulliRBM(W,a,b,np.full(V,0.5)), ctx)
    h = (np.random.default_rng(1).random((n,H))<0.3).astype(float)
    res[eng] = pkg.samplevisible(m, h, tick=5, ctx=ctx)
d = res[1]!=res[2]
print("mismatch frac", d.mean(), "per col", d.mean(0))
print("simt mean", res[1].mean(0)); print("tc mean", res[2].mean(0))
idx = np.argwhere(d)[:20]
for r,c in idx: print(r,c,res[1][r,c],res[2][r,c])
print("rows with mismatch mod 32 hist", np.bincount(np.argwhere(d)[:,0]%32, minlength=32))
