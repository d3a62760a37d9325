"""NumPy Float64 restatement of the reference's model code on the hot path.

Matrices keep the reference's logical shapes: states are (nsamples x nunits),
weights (nvisible x nhidden) (src/bmtypes.jl:17-21,85-92).  Every function
cites the reference lines it follows.  Random draws use `philox.Rng`
addressed by (stream = layer | purpose, tick) -- see oracle/philox.py -- or
explicit uniform arrays (`u=`), which is how the "same uniforms" gate feeds
the CUDA path and this oracle identical numbers.

Tick protocol (shared with the device library): an object that owns a chain
(Particles, the RBM trainer) carries a `clock`; every synchronous sampling
operation uses the current clock value as its tick and then advances it by 1.
"""
import copy
from dataclasses import dataclass, field
import numpy as np

from . import philox as px
from . import sampling as sp
from .sampling import sigm

NB_EPS = 1e-6          # src/gibbssampling.jl:1018
VISBIAS_MAX = -1e-10   # src/optimizers.jl:483


# ----------------------------------------------------------------- types
@dataclass
class BernoulliRBM:
    """src/bmtypes.jl:17-21"""
    weights: np.ndarray
    visbias: np.ndarray
    hidbias: np.ndarray
    kind = "bernoulli"


@dataclass
class NegativeBinomialBernoulliRBM:
    """src/bmtypes.jl:85-92"""
    weights: np.ndarray
    visbias: np.ndarray
    hidbias: np.ndarray
    inversedispersion: np.ndarray
    dropoutmid: np.ndarray = field(default_factory=lambda: np.array([0.5]))
    dropoutshape: np.ndarray = field(default_factory=lambda: np.array([1.0]))
    kind = "nb"


class Particles(list):
    """src/gibbssampling.jl:9 plus the draw clock."""

    def __init__(self, mats, clock=0, row_offset=0):
        super().__init__(mats)
        self.clock = clock
        self.row_offset = row_offset


def nvisiblenodes(rbm):
    return rbm.weights.shape[0]


def nhiddennodes(rbm):
    return rbm.weights.shape[1]


# ----------------------------------------------------- conditionals (L1)
def hiddeninput(rbm, vv):
    """src/gibbssampling.jl:207-212 (Bernoulli), :271-276 (NB)"""
    return vv @ rbm.weights + rbm.hidbias[None, :]


def hiddenpotential(rbm, vv, factor=1.0):
    """src/gibbssampling.jl:327-335"""
    hh = hiddeninput(rbm, vv)
    if factor != 1.0:
        hh = hh * factor
    return sigm(hh)


def visibleinput(rbm, hh):
    """src/gibbssampling.jl:884-890"""
    return hh @ rbm.weights.T + rbm.visbias[None, :]


def nbmean(x, r):
    """src/gibbssampling.jl:1018: r*exp(x)/(1-exp(x)) + 1e-6"""
    ex = np.exp(x)
    return (r[None, :] * ex) / (1.0 - ex) + NB_EPS


def visiblepotential(rbm, hh, factor=1.0):
    """Bernoulli: src/gibbssampling.jl:949-957.  NB: :1013-1020 -- the NB
    method ignores `factor` (quirk B1), replicated."""
    x = visibleinput(rbm, hh)
    if rbm.kind == "nb":
        return nbmean(x, rbm.inversedispersion)
    if factor != 1.0:
        x = x * factor
    return sigm(x)


def samplehidden(rbm, vv, rng, layer, tick, factor=1.0, row_offset=0, u=None):
    """src/gibbssampling.jl:594-610"""
    p = hiddenpotential(rbm, vv, factor)
    if u is None:
        u = rng.uniforms(p.shape[0], p.shape[1], px.stream_id(layer), tick, row_offset)
    return sp.bernoulli(p, u)


def samplevisible(rbm, hh, rng, layer, tick, factor=1.0, row_offset=0, u=None,
                  pshift=0.0, mu_inv_max=sp.MU_INV_MAX):
    """src/gibbssampling.jl:752-756 -> :765-770 (Bernoulli) / :810-820 (NB)"""
    pot = visiblepotential(rbm, hh, factor)
    if rbm.kind == "nb":
        return sp.nb_draw(pot, rbm.inversedispersion, rng, layer, tick, row_offset,
                          pshift=pshift, mu_inv_max=mu_inv_max, u=u)
    if u is None:
        u = rng.uniforms(pot.shape[0], pot.shape[1], px.stream_id(layer), tick, row_offset)
    return sp.bernoulli(pot, u)


# ------------------------------------------------------ particle init
def inithiddennodes(n, rbm, rng, layer, biased, row_offset=0):
    """src/gibbssampling.jl:421-453: `rand!(h, [0.0 1.0])` (fair coin) or
    Bernoulli(sigm(hidbias)) when biased."""
    H = nhiddennodes(rbm)
    u = rng.uniforms(n, H, px.stream_id(layer, px.P_INIT), 0, row_offset)
    if biased:
        return sp.bernoulli(np.broadcast_to(sigm(rbm.hidbias)[None, :], (n, H)), u)
    return sp.bernoulli(np.full((n, H), 0.5), u)


def initvisiblenodes(n, rbm, rng, biased, row_offset=0, reference_faithful=False):
    """Bernoulli: src/gibbssampling.jl:470-480 style fair coin / sigm(bias).
    NB: src/gibbssampling.jl:512-531: v0 ~ NB(1, 0.5), then one gamma-Poisson
    draw with p = r/(v0+r) - 1e-6.  `biased` adds the (log-scale) visible bias
    to the counts in the reference (quirk B6); replicated only when
    `reference_faithful`."""
    V = nvisiblenodes(rbm)
    if rbm.kind != "nb":
        u = rng.uniforms(n, V, px.stream_id(0, px.P_INIT), 0, row_offset)
        p = sigm(rbm.visbias)[None, :] if biased else np.full((1, V), 0.5)
        return sp.bernoulli(np.broadcast_to(p, (n, V)), u)
    u0 = rng.uniforms(n, V, px.stream_id(0, px.P_INIT), 0, row_offset)
    v0 = sp.nb_inversion(np.ones((n, V)), np.ones((n, V)), u0)      # NB(1, 0.5): mean 1
    # second stage: NB with mean v0 (p = r/(v0+r) - 1e-6); v0 == 0 -> mean ~1e-6*r
    u1 = rng.uniforms(n, V, px.stream_id(0, px.P_INIT2), 0, row_offset)
    v = sp.nb_draw(v0, rbm.inversedispersion, rng, 0, 0, row_offset, pshift=1e-6,
                   mu_inv_max=np.inf, u=u1)
    if biased and reference_faithful:
        v = v + rbm.visbias[None, :]
    return v


def initparticles(bm, nparticles, rng, biased=False, row_offset=0, reference_faithful=False):
    """src/gibbssampling.jl:386-419"""
    rbms = bm if isinstance(bm, (list, tuple)) else [bm]
    mats = [initvisiblenodes(nparticles, rbms[0], rng, biased, row_offset, reference_faithful)]
    for i, rbm in enumerate(rbms):
        mats.append(inithiddennodes(nparticles, rbm, rng, i + 1, biased, row_offset))
    return Particles(mats, clock=1, row_offset=row_offset)


# ------------------------------------------------------------ Gibbs (L2)
def gibbssample_rbm(particles, rbm, rng, nsteps=5, upfactor=1.0, downfactor=1.0, **kw):
    """src/gibbssampling.jl:62-70: v | h then h | (new) v."""
    for _ in range(nsteps):
        particles[0] = samplevisible(rbm, particles[1], rng, 0, particles.clock, downfactor,
                                     particles.row_offset, **kw)
        particles.clock += 1
        particles[1] = samplehidden(rbm, particles[0], rng, 1, particles.clock, upfactor,
                                    particles.row_offset)
        particles.clock += 1
    return particles


def gibbssamplenegativebinomial_rbm(particles, rbm, rng, nsteps=5, upfactor=1.0, downfactor=1.0,
                                    mu_inv_max=sp.MU_INV_MAX):
    """src/gibbssampling.jl:82-110: explicit gamma-Poisson with p-1e-6, and the
    hidden layer is sampled from the *old* visible state (`oldvis`)."""
    for _ in range(nsteps):
        oldvis = particles[0].copy()
        particles[0] = samplevisible(rbm, particles[1], rng, 0, particles.clock, downfactor,
                                     particles.row_offset, pshift=1e-6, mu_inv_max=mu_inv_max)
        particles[1] = samplehidden(rbm, oldvis, rng, 1, particles.clock, upfactor,
                                    particles.row_offset)
        particles.clock += 1
    return particles


def dbm_layer_input(dbm, particles, i):
    """Total input of intermediate layer i (0-based, 0 < i < L):
    src/gibbssampling.jl:165-168 / src/dbmtraining.jl:199-203."""
    return visibleinput(dbm[i], particles[i + 1]) + hiddeninput(dbm[i - 1], particles[i - 1])


def gibbssample_dbm(particles, dbm, rng, nsteps=5, temperature=1.0, mu_inv_max=sp.MU_INV_MAX,
                    u=None):
    """src/gibbssampling.jl:153-182: synchronous (Jacobi) update of all layers
    from the previous step's state.  `temperature` scales every weight matrix
    (`copyannealed!`, src/evaluating.jl:320-326,348-355).  `u`, if given, is a
    list (per step) of lists (per layer) of explicit uniform arrays."""
    if temperature != 1.0:
        dbm = [_annealed(rbm, temperature) for rbm in dbm]
    L = len(particles)
    ro = particles.row_offset
    for step in range(nsteps):
        t = particles.clock
        uu = u[step] if u is not None else [None] * L
        new = [None] * L
        new[0] = samplevisible(dbm[0], particles[1], rng, 0, t, 1.0, ro, u=uu[0], mu_inv_max=mu_inv_max)
        for i in range(1, L - 1):
            p = sigm(dbm_layer_input(dbm, particles, i))
            ui = uu[i] if uu[i] is not None else rng.uniforms(p.shape[0], p.shape[1], px.stream_id(i), t, ro)
            new[i] = sp.bernoulli(p, ui)
        new[L - 1] = samplehidden(dbm[-1], particles[L - 2], rng, L - 1, t, 1.0, ro, u=uu[L - 1])
        particles[:] = new
        particles.clock += 1
    return particles


def sampleparticles_native(dbm, nparticles, gen, burnin=10):
    """TIMING twin of `sampleparticles` for a DBM (bench.py's CPU baseline): the same operation sequence as the
    reference (src/gibbssampling.jl:646-650,153-182,512-531,810-820) -- dgemm, NB mean, one `NegativeBinomial(r, p)`
    draw per visible element, `rand() < p` per hidden unit -- with the draws taken from NumPy's own compiled samplers
    (`Generator.negative_binomial`, what Distributions.jl's samplers are to the reference) instead of the
    Philox-addressed inversion the parity tests need.  Never used as a checker."""
    first = dbm[0]
    r = first.inversedispersion[None, :]
    V = first.weights.shape[0]
    v0 = gen.negative_binomial(1.0, 0.5, (nparticles, V)).astype(float)                    # :515
    mu0 = v0 + 1e-6
    parts = [gen.negative_binomial(np.broadcast_to(r, mu0.shape), np.clip(r / (mu0 + r) - 1e-6, 1e-12, 1.0)).astype(float)]
    for rbm in dbm:
        parts.append((gen.random((nparticles, rbm.weights.shape[1])) < 0.5).astype(float))  # :428
    L = len(parts)
    for _ in range(burnin):
        new = [None] * L
        mu = visiblepotential(first, parts[1])                                               # :1018
        new[0] = gen.negative_binomial(np.broadcast_to(r, mu.shape), r / (mu + r)).astype(float)   # :814-818
        for i in range(1, L - 1):
            p = sigm(dbm_layer_input(dbm, parts, i))
            new[i] = (gen.random(p.shape) < p).astype(float)
        p = hiddenpotential(dbm[-1], parts[L - 2])
        new[L - 1] = (gen.random(p.shape) < p).astype(float)
        parts = new
    return parts


def _annealed(rbm, temperature):
    r = copy.copy(rbm)
    r.weights = rbm.weights * temperature
    return r


def sampleparticles(bm, nparticles, rng, burnin=10, row_offset=0, **kw):
    """src/gibbssampling.jl:646-650"""
    particles = initparticles(bm, nparticles, rng, row_offset=row_offset)
    if isinstance(bm, (list, tuple)):
        return gibbssample_dbm(particles, bm, rng, burnin, **kw)
    return gibbssample_rbm(particles, bm, rng, burnin, **kw)


def samples(bm, nsamples, rng, burnin=50, samplelast=True, row_offset=0, **kw):
    """src/gibbssampling.jl:663-689"""
    if samplelast:
        return sampleparticles(bm, nsamples, rng, burnin, row_offset, **kw)[0]
    particles = sampleparticles(bm, nsamples, rng, burnin - 1, row_offset, **kw)
    first = bm[0] if isinstance(bm, (list, tuple)) else bm
    return visiblepotential(first, particles[1])


# ------------------------------------------------------- mean-field (L2)
def meanfield(dbm, x, eps=0.001, maxiter=None, reference_order=True):
    """src/dbmtraining.jl:175-228.  Gauss-Seidel sweep over hidden layers,
    global max-abs-delta stopping rule.  Returns (mu list, iterations, delta).
    `maxiter` runs a fixed number of iterations (parity at fixed count)."""
    nl = len(dbm) + 1
    mu = [x]
    for i in range(1, nl - 1):
        mu.append(hiddenpotential(dbm[i - 1], mu[i - 1], 2.0))
    mu.append(hiddenpotential(dbm[nl - 2], mu[nl - 2]))
    delta, it = 1.0, 0
    while delta > eps and (maxiter is None or it < maxiter):
        delta = 0.0
        for i in range(1, nl - 1):
            new = sigm(hiddeninput(dbm[i - 1], mu[i - 1]) + visibleinput(dbm[i], mu[i + 1]))
            delta = max(delta, np.max(np.abs(mu[i] - new)))
            mu[i] = new
        new = hiddenpotential(dbm[-1], mu[-2])
        delta = max(delta, np.max(np.abs(mu[-1] - new)))
        mu[-1] = new
        it += 1
    return Particles(mu), it, delta


# -------------------------------------------------- gradient / update
@dataclass
class Gradient:
    weights: np.ndarray
    visbias: np.ndarray
    hidbias: np.ndarray


def computegradient(v, vmodel, h, hmodel):
    """src/optimizers.jl:263-310 (generic and NB bodies are identical).
    Division by the sample count is skipped when the count is 1."""
    npos, nneg = v.shape[0], vmodel.shape[0]
    gw = v.T @ h
    if npos != 1:
        gw = gw / npos
    neg = vmodel.T @ hmodel
    if nneg != 1:
        neg = neg / nneg
    gw = gw - neg
    gh = h.mean(axis=0) - hmodel.mean(axis=0)
    gv = v.mean(axis=0) - vmodel.mean(axis=0)
    return Gradient(gw, gv, gh)


def updateparameters(rbm, grad, learningrate):
    """src/optimizers.jl:513-520; NB clamps :478-485."""
    rbm.weights = rbm.weights + learningrate * grad.weights
    rbm.visbias = rbm.visbias + learningrate * grad.visbias
    rbm.hidbias = rbm.hidbias + learningrate * grad.hidbias
    if rbm.kind == "nb":
        rbm.weights = np.minimum(rbm.weights, 0.0)
        rbm.visbias = np.minimum(rbm.visbias, VISBIAS_MAX)
    return rbm


def computegradient_dbm(mu, particles, dbm):
    """src/optimizers.jl:431-442"""
    return [computegradient(mu[i], particles[i], mu[i + 1], particles[i + 1]) for i in range(len(dbm))]


def updateparameters_dbm(dbm, grads, learningrate):
    """src/optimizers.jl:505-510"""
    for rbm, g in zip(dbm, grads):
        updateparameters(rbm, g, learningrate)
    return dbm


# --------------------------------------------------------- RBM training
def initrbm(x, nhidden, rng, rbmtype="bernoulli", inversedispersion=None):
    """src/rbmtraining.jl:332-374.  NB: W = -U(0.0005, 0.002) (:364),
    a = log(mean/(mean+r)) (:394-401), b = 0.  Other: W = randn/sqrt(V),
    a = logit(mean) where mean > 0 (:380-392)."""
    n, V = x.shape
    hidbias = np.zeros(nhidden)
    if rbmtype == "nb":
        u = rng.uniforms(V, nhidden, px.stream_id(0, px.P_INIT), 0xFFFF0000)
        weights = -(0.0005 + (0.002 - 0.0005) * u)
        r = np.full(V, 0.5) if inversedispersion is None else np.asarray(inversedispersion, float)
        means = x.mean(axis=0)
        with np.errstate(divide="ignore"):
            visbias = np.log(means / (means + r))
        return NegativeBinomialBernoulliRBM(weights, visbias, hidbias, r.copy())
    w0 = rng.words(V, nhidden, px.stream_id(0, px.P_INIT), 0xFFFF0000)
    w1 = rng.words(V, nhidden, px.stream_id(0, px.P_INIT2), 0xFFFF0000)
    weights = sp._normal_from_words(w0, w1) / np.sqrt(V)
    visbias = np.zeros(V)
    m = x.mean(axis=0)
    pos = m > 0
    with np.errstate(divide="ignore", invalid="ignore"):
        visbias[pos] = np.log(m[pos] / (1 - m[pos]))
    return BernoulliRBM(weights, visbias, hidbias)


def randombatchmasks(nsamples, batchsize, rng, tick):
    """src/rbmtraining.jl:416-429"""
    if batchsize > nsamples:
        raise ValueError(f"Batchsize ({batchsize}) exceeds number of samples ({nsamples}).")
    perm = rng.permutation(nsamples, tick)
    return [perm[s:s + batchsize] for s in range(0, nsamples, batchsize)]


class RBMTrainer:
    """State that `fitrbm` keeps across epochs (src/rbmtraining.jl:137-151):
    PCD chain state and the draw clock."""

    def __init__(self, rbm, batchsize, rng, pcd=True):
        self.rng = rng
        self.clock = 1
        self.batchsize = batchsize
        self.chainstate = None
        if pcd:
            # `rand(batchsize, nhidden)`: uniforms used as probabilities (quirk B9)
            self.chainstate = rng.uniforms(batchsize, nhiddennodes(rbm), px.stream_id(1, px.P_INIT), 0)


def trainrbm(rbm, x, trainer, learningrate=0.005, cdsteps=1, upfactor=1.0, downfactor=1.0,
             batchmasks=None, mu_inv_max=sp.MU_INV_MAX):
    """One epoch of minibatch CD-k / PCD: src/rbmtraining.jl:498-591.
    Zero-inflation (`estimatedropout!`, :531-534) and `estimatedmean`
    bookkeeping (:575-576) are "next" rows and not part of this restatement."""
    rng = trainer.rng
    n = x.shape[0]
    B = trainer.batchsize
    if batchmasks is None:
        batchmasks = randombatchmasks(n, B, rng, trainer.clock)
        trainer.clock += 1
    pcd = trainer.chainstate is not None
    for mask in batchmasks:
        v = x[mask, :]
        thisb = v.shape[0]
        h = hiddenpotential(rbm, v, upfactor)                       # :556
        hmodel = trainer.chainstate if pcd else h.copy()             # :561-565
        u = rng.uniforms(hmodel.shape[0], hmodel.shape[1], px.stream_id(1), trainer.clock)
        trainer.clock += 1
        hmodel = sp.bernoulli(hmodel, u)                             # :567
        for _ in range(cdsteps - 1):                                 # :570, samplers.jl:47-56
            vmodel = samplevisible(rbm, hmodel, rng, 0, trainer.clock, downfactor, mu_inv_max=mu_inv_max)
            trainer.clock += 1
            hmodel = samplehidden(rbm, vmodel, rng, 1, trainer.clock, upfactor)
            trainer.clock += 1
        vmodel = visiblepotential(rbm, hmodel, downfactor)           # :573
        hmodel = hiddenpotential(rbm, vmodel, upfactor)              # :578
        trainer.last_vmodel = vmodel                                  # rows appended to `estimatedmean` (:575-576)
        if pcd:
            trainer.chainstate = hmodel                              # alias, :562
        grad = computegradient(v, vmodel[:thisb], h, hmodel[:thisb])  # :580-586
        updateparameters(rbm, grad, learningrate)                    # :587
    return rbm


def fitrbm(x, rng, nhidden=None, epochs=10, upfactor=1.0, downfactor=1.0, learningrate=0.005,
           learningrates=None, pcd=True, cdsteps=1, batchsize=1, rbmtype="bernoulli",
           inversedispersion=None, startrbm=None, monitoring=None, mu_inv_max=sp.MU_INV_MAX):
    """src/rbmtraining.jl:89-223 with fisherscoring=0, zeroinflation=false
    (the dispersion/dropout re-estimation are "next" rows)."""
    nhidden = x.shape[1] if nhidden is None else nhidden
    rbm = copy.deepcopy(startrbm) if startrbm is not None else initrbm(x, nhidden, rng, rbmtype, inversedispersion)
    if learningrates is None:
        learningrates = [learningrate] * epochs
    if len(learningrates) < epochs:
        raise ValueError("Not enough `learningrates` for training epochs")
    trainer = RBMTrainer(rbm, batchsize, rng, pcd)
    for epoch in range(epochs):
        trainrbm(rbm, x, trainer, learningrates[epoch], cdsteps, upfactor, downfactor, mu_inv_max=mu_inv_max)
        if monitoring is not None:
            monitoring(rbm, epoch + 1)
    return rbm


# --------------------------------------------------------------- DBM
@dataclass
class TrainLayer:
    """The subset of src/rbmstacking.jl:9-94 on the hot path."""
    nhidden: int = -1
    rbmtype: str = "bernoulli"
    epochs: int = -1
    learningrate: float = -np.inf
    batchsize: int = -1
    pcd: bool = True
    cdsteps: int = 1
    inversedispersion: np.ndarray = None


def stackrbms(x, rng, nhiddens=None, epochs=10, predbm=False, learningrate=0.005, batchsize=1,
              trainlayers=None, mu_inv_max=sp.MU_INV_MAX):
    """src/rbmstacking.jl:228-284 (factors :247-275).  Each layer gets its own
    Philox seed derived from the layer index so that layer fits are
    independent streams."""
    if not trainlayers:
        nhiddens = nhiddens or [x.shape[1]] * 2
        trainlayers = [TrainLayer(nhidden=n) for n in nhiddens]
    dbm = []
    upfactor = 2.0 if predbm else 1.0
    downfactor = 1.0
    hiddenval = x
    for i, tl in enumerate(trainlayers):
        if i > 0:
            hiddenval = hiddenpotential(dbm[i - 1], hiddenval, upfactor)
            if predbm:
                upfactor = downfactor = 2.0
                if i == len(trainlayers) - 1:
                    upfactor = 1.0
            else:
                upfactor = downfactor = 1.0
        lrng = px.Rng((rng.seed + 1000003 * (i + 1)) % (1 << 52))
        dbm.append(fitrbm(hiddenval, lrng, nhidden=tl.nhidden,
                          epochs=tl.epochs if tl.epochs >= 0 else epochs,
                          upfactor=upfactor, downfactor=downfactor,
                          learningrate=tl.learningrate if tl.learningrate >= 0 else learningrate,
                          pcd=tl.pcd, cdsteps=tl.cdsteps,
                          batchsize=tl.batchsize if tl.batchsize > 0 else batchsize,
                          rbmtype=tl.rbmtype, inversedispersion=tl.inversedispersion,
                          mu_inv_max=mu_inv_max))
    return dbm


def defaultfinetuninglearningrates(learningrate, epochs):
    """src/dbmtraining.jl:59-61"""
    return learningrate * 11.0 / (10.0 + np.arange(1, epochs + 1))


def traindbm_step(dbm, x, particles, rng, learningrate, nsteps=5, eps=0.001, mu_inv_max=sp.MU_INV_MAX):
    """src/dbmtraining.jl:287-297"""
    gibbssample_dbm(particles, dbm, rng, nsteps, mu_inv_max=mu_inv_max)
    mu, _, _ = meanfield(dbm, x, eps)
    grads = computegradient_dbm(mu, particles, dbm)
    updateparameters_dbm(dbm, grads, learningrate)
    return dbm


def traindbm(dbm, x, rng, epochs=10, nparticles=100, learningrate=0.005, learningrates=None,
             monitoring=None, mu_inv_max=sp.MU_INV_MAX):
    """src/dbmtraining.jl:250-280"""
    if learningrates is None:
        learningrates = defaultfinetuninglearningrates(learningrate, epochs)
    particles = initparticles(dbm, nparticles, rng)
    for epoch in range(epochs):
        traindbm_step(dbm, x, particles, rng, learningrates[epoch], mu_inv_max=mu_inv_max)
        if monitoring is not None:
            monitoring(dbm, epoch + 1)
    return dbm


def fitdbm(x, rng, nhiddens=None, epochs=10, nparticles=100, learningrate=0.005, learningrates=None,
           learningratepretraining=None, epochspretraining=None, batchsizepretraining=1,
           pretraining=None, monitoring=None, mu_inv_max=sp.MU_INV_MAX):
    """src/dbmtraining.jl:108-149"""
    if not pretraining and not nhiddens:
        nhiddens = [x.shape[1]] * 2
    dbm = stackrbms(x, rng, nhiddens=nhiddens,
                    epochs=epochs if epochspretraining is None else epochspretraining,
                    predbm=True, batchsize=batchsizepretraining,
                    learningrate=learningrate if learningratepretraining is None else learningratepretraining,
                    trainlayers=pretraining, mu_inv_max=mu_inv_max)
    return traindbm(dbm, x, rng, epochs=epochs, nparticles=nparticles, learningrate=learningrate,
                    learningrates=learningrates, monitoring=monitoring, mu_inv_max=mu_inv_max)


# ====================================================================== "next" rows
# ------------------------------------------------------------- conditional Gibbs
def gibbssamplecond(particles, bm, rng, mask, nsteps=5, mu_inv_max=sp.MU_INV_MAX):
    """`gibbssamplecond!` / `gibbssamplenegativebinomial_cond!` / `gibbssamplecondrow!`
    (src/gibbssampling.jl:1042-1237): after every visible update the masked visible entries are
    reset to their initial values.  `mask`: bool vector (columns) or bool matrix (entries)."""
    mask = np.asarray(mask, dtype=bool)
    full = np.broadcast_to(mask[None, :], particles[0].shape) if mask.ndim == 1 else mask
    orig = particles[0].copy()
    ro = particles.row_offset
    if not isinstance(bm, (list, tuple)):
        for _ in range(nsteps):                                  # :1052-1064
            v = samplevisible(bm, particles[1], rng, 0, particles.clock, 1.0, ro, mu_inv_max=mu_inv_max)
            particles.clock += 1
            v[full] = orig[full]
            particles[0] = v
            particles[1] = samplehidden(bm, v, rng, 1, particles.clock, 1.0, ro)
            particles.clock += 1
        return particles
    L = len(particles)
    for _ in range(nsteps):                                      # :1066-1104
        t = particles.clock
        new = [None] * L
        new[0] = samplevisible(bm[0], particles[1], rng, 0, t, 1.0, ro, mu_inv_max=mu_inv_max)
        new[0][full] = orig[full]
        for i in range(1, L - 1):
            p = sigm(dbm_layer_input(bm, particles, i))
            new[i] = sp.bernoulli(p, rng.uniforms(p.shape[0], p.shape[1], px.stream_id(i), t, ro))
        new[L - 1] = samplehidden(bm[-1], particles[L - 2], rng, L - 1, t, 1.0, ro)
        particles[:] = new
        particles.clock += 1
    return particles


# ------------------------------------------------------------- zero-inflation dropout
def estimatedropout(x):
    """`estimatedropout!` (src/rbmtraining.jl:271-287): binomial GLM (logit link) of the per-gene zero
    proportion on log1p(per-gene mean); returns (intercept, slope) = (dropoutmid, dropoutshape).
    GLM.jl's `fit(GeneralizedLinearModel, ..., Binomial())` is IRLS; restated here."""
    logcounts = np.log1p(x.mean(axis=0))
    y = (x == 0).mean(axis=0)
    X = np.stack([np.ones_like(logcounts), logcounts], axis=1)
    beta = np.zeros(2)
    for _ in range(100):
        eta = X @ beta
        mu = sigm(eta)
        w = np.maximum(mu * (1 - mu), 1e-12)
        z = eta + (y - mu) / w
        new = np.linalg.solve(X.T @ (w[:, None] * X), X.T @ (w * z))
        if np.max(np.abs(new - beta)) < 1e-12:
            beta = new
            break
        beta = new
    return float(beta[0]), float(beta[1])


def sampledropout(x, mid, shape, rng, tick, log1p_form=False, row_offset=0):
    """`sampledropout` (src/rbmtraining.jl:289-311): 1 - bernoulli(logistic(k (x' - x0))).  RBM form: x' = x
    and scalar (mid, shape) (:289-299); DBM form: x' = log1p(x) and per-gene parameters (:301-311)."""
    xt = np.log1p(x) if log1p_form else x
    mid = np.broadcast_to(np.asarray(mid, float).reshape(1, -1), (1, x.shape[1]))
    shape = np.broadcast_to(np.asarray(shape, float).reshape(1, -1), (1, x.shape[1]))
    p = 1.0 / (1.0 + np.exp(-shape * (xt - mid)))
    u = rng.uniforms(x.shape[0], x.shape[1], px.stream_id(0, px.P_DROPOUT), tick, row_offset)
    return 1.0 - sp.bernoulli(p, u)


# ------------------------------------------------------------- NB dispersion (Fisher scoring)
def mle_for_theta(y, mu, lam, maxiter=20, convtol=1e-6):
    """`mle_for_θ` (src/rbmtraining.jl:675-719), unit weights: penalised Fisher scoring for the NB
    inverse dispersion of one gene.  `y`, `mu`: (n,) or (n, V) (vectorised over genes)."""
    from scipy.special import digamma, polygamma
    y = np.asarray(y, float)
    mu = np.asarray(mu, float)
    n = y.shape[0]
    theta = n / np.sum((y / mu - 1.0) ** 2, axis=0)
    theta = np.atleast_1d(np.asarray(theta, float)).copy()
    y2 = y.reshape(n, -1)
    mu2 = mu.reshape(n, -1)
    active = np.ones(theta.shape, dtype=bool)
    for _ in range(maxiter):
        theta[active] = np.maximum(theta[active], 0.01)
        th = theta[None, :]
        score = np.sum(-((y2 + th) / (mu2 + th) + np.log(mu2 + th) - 1 - np.log(th) - digamma(th + y2) + digamma(th)), axis=0)
        fisher = np.sum(-(y2 + th) / (mu2 + th) ** 2 + 2 / (mu2 + th) - 1 / th - polygamma(1, th + y2) + polygamma(1, th), axis=0)
        delta = (score + lam * 2 / theta ** 3) / (fisher + lam * 6 / theta ** 4)
        done = np.abs(delta) <= convtol
        step = active & ~done
        theta[step] = theta[step] + delta[step]
        active &= ~done
        if not active.any():
            break
    return theta if y.ndim > 1 else float(theta[0])


def fitrbm_nb(x, rng, nhidden, epochs=10, learningrate=0.005, pcd=True, cdsteps=1, batchsize=1,
              inversedispersion=None, fisherscoring=1, lam=100.0, zeroinflation=True, upfactor=1.0, downfactor=1.0,
              mu_inv_max=sp.MU_INV_MAX):
    """`fitrbm(...; rbmtype = NegativeBinomialBernoulliRBM)` with the reference's NB defaults
    (src/rbmtraining.jl:89-223): every 5th epoch the per-gene inverse dispersion is re-estimated by
    `mle_for_θ(x[:,g], estimatedmean[:,g])` (:172-183, estimatedispersion = "gene"), where `estimatedmean`
    is the epoch's model means in minibatch order (quirk B3: not re-aligned with the rows of x)."""
    rbm = initrbm(x, nhidden, rng, "nb", inversedispersion)
    trainer = RBMTrainer(rbm, batchsize, rng, pcd)
    n = x.shape[0]
    for epoch in range(1, epochs + 1):
        masks = randombatchmasks(n, batchsize, rng, trainer.clock)
        trainer.clock += 1
        est = []
        for mask in masks:
            trainrbm(rbm, x, trainer, learningrate, cdsteps, upfactor, downfactor, batchmasks=[mask], mu_inv_max=mu_inv_max)
            est.append(trainer.last_vmodel)
        if fisherscoring != 0 and epoch % 5 == 0:
            estimatedmean = np.vstack(est)[:n]
            rbm.inversedispersion = mle_for_theta(x, estimatedmean, lam, maxiter=20)
    if zeroinflation:
        mid, shape = estimatedropout(x)
        rbm.dropoutmid, rbm.dropoutshape = np.array([mid]), np.array([shape])
    return rbm
