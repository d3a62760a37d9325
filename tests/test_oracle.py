"""CPU tests of the oracle: third-party known answers, the reference's
relational tests restated on tiny models (test/BMTest.jl), analytic NB
identities, and the committed golden fixtures."""
import copy
import os

import numpy as np
import pytest
from scipy import stats

from oracle import model as om, philox as px, sampling as osp, ais as oais, exact, synth

GOLD = os.path.join(os.path.dirname(__file__), "golden", "dbm_48_20_12.npz")


def test_philox_known_answers():
    """Random123 kat_vectors: philox4x32-10."""
    g = np.load(GOLD)
    for cin, cout in zip(g["philox_kat_in"], g["philox_kat_out"]):
        r = px.philox4x32_10(*[int(x) for x in cin])
        assert [int(x) for x in r] == [int(x) for x in cout]


def test_uniform_mapping_is_fp32_exact():
    w = np.array([0, 1, 511, 512, 0xFFFFFFFF, 0x80000000], dtype=np.uint32)
    u = px.u24(w)
    assert np.all(u > 0) and np.all(u < 1)
    assert np.array_equal(u.astype(np.float32).astype(np.float64), u)


def test_nb_inversion_equals_scipy_quantile():
    g = np.load(GOLD)
    got = osp.nb_inversion(g["nb_mu"], g["nb_r"][None, :], g["nb_u"])
    assert np.array_equal(got, g["nb_ppf"])


@pytest.mark.parametrize("mu0,r0", [(0.2, 0.5), (7.9, 2.0), (20.0, 0.5), (300.0, 5.0), (12.0, 0.05)])
def test_nb_draw_law(mu0, r0):
    """Both branches sample NB(r, r/(mu+r)) (src/gibbssampling.jl:810-820)."""
    n = 60000
    x = osp.nb_draw(np.full((n, 1), mu0), np.array([r0]), px.Rng(9), 0, 3)
    p = r0 / (mu0 + r0)
    ks = [0, 1, 2, 5, 10, 30, 100, 1000]
    emp = np.array([(x <= k).mean() for k in ks])
    assert np.max(np.abs(emp - stats.nbinom.cdf(ks, r0, p))) < 0.01
    assert abs(x.mean() - mu0) < 5 * np.sqrt((mu0 + mu0 ** 2 / r0) / n)


def test_nb_identities():
    """a = log(m/(m+r)) => visiblepotential(h=0) = m + 1e-6; p = 1 - e^x."""
    m = np.array([0.01, 0.3, 2.0, 40.0])
    r = np.array([0.5, 0.5, 2.0, 0.1])
    rbm = om.NegativeBinomialBernoulliRBM(-np.full((4, 3), 0.01), np.log(m / (m + r)), np.zeros(3), r)
    mu = om.visiblepotential(rbm, np.zeros((1, 3)))
    assert np.allclose(mu[0], m + 1e-6, rtol=1e-12)
    x = om.visibleinput(rbm, np.ones((1, 3)))
    mu1 = om.visiblepotential(rbm, np.ones((1, 3)))
    assert np.allclose(r / (mu1[0] - 1e-6 + r), 1 - np.exp(x[0]), rtol=1e-10)
    assert np.allclose(om.visiblepotential(rbm, np.ones((1, 3)), 2.0), mu1)   # quirk B1: factor ignored


def test_potentials_inplace_equals_allocating_shapes():
    """test/BMTest.jl:100-127 restated: conditionals are consistent across row subsets."""
    dbm = synth.random_dbm([12, 6, 4], 1, wscale=(0.02, 0.2))
    g = np.random.default_rng(0)
    v = g.poisson(1.0, (9, 12)).astype(float)
    full = om.hiddenpotential(dbm[0], v, 2.0)
    assert np.allclose(np.vstack([om.hiddenpotential(dbm[0], v[i:i + 1], 2.0) for i in range(9)]), full)


def test_ais_vs_exact_rbm_and_dbm():
    """test/BMTest.jl:319-327,557-562: AIS log Z within 1 % of enumeration."""
    g = np.random.default_rng(0)
    rbm = om.BernoulliRBM(g.normal(0, 0.5, (5, 4)), g.normal(0, .3, 5), g.normal(0, .3, 4))
    lz = exact.exactlogpartitionfunction_rbm(rbm)
    est = oais.logpartitionfunction(rbm, px.Rng(1), ntemperatures=150, nparticles=300)
    assert abs(est - lz) < 0.01 * abs(lz)
    dbm = [om.BernoulliRBM(g.normal(0, 0.5, (4, 3)), g.normal(0, .3, 4), g.normal(0, .3, 3)),
           om.BernoulliRBM(g.normal(0, 0.5, (3, 2)), g.normal(0, .3, 3), g.normal(0, .3, 2))]
    lzd = exact.exactlogpartitionfunction_dbm(dbm)
    estd = oais.logpartitionfunction(dbm, px.Rng(2), ntemperatures=150, nparticles=300)
    assert abs(estd - lzd) < 0.01 * abs(lzd)
    # zero weights: AIS weights are exactly zero, log Z = log Z0 (test/BMTest.jl:178-210 in spirit)
    z = om.BernoulliRBM(np.zeros((5, 4)), rbm.visbias, rbm.hidbias)
    assert np.allclose(oais.aislogimpweights(z, px.Rng(3), ntemperatures=5, nparticles=7), 0.0)
    assert np.isclose(exact.exactlogpartitionfunction_rbm(z), oais.logpartitionfunctionzeroweights(z))


def test_gibbs_frequencies_vs_exact_marginals():
    g = np.random.default_rng(3)
    rbm = om.BernoulliRBM(g.normal(0, 0.7, (5, 3)), g.normal(0, .3, 5), g.normal(0, .3, 3))
    p = om.sampleparticles(rbm, 20000, px.Rng(4), burnin=25)
    assert np.max(np.abs(p[0].mean(0) - exact.bernoulli_rbm_visible_marginal_mean(rbm))) < 0.015
    V, H = 6, 3
    means = np.array([.2, .5, 1, 2, 4, .05])
    nb = om.NegativeBinomialBernoulliRBM(-g.uniform(0.05, 0.6, (V, H)), np.log(means / (means + 0.5)), g.normal(0, .5, H), np.full(V, 0.5))
    q = om.sampleparticles(nb, 30000, px.Rng(5), burnin=30)
    m, v, z = exact.nbrbm_visible_moments(nb, eps=1e-6)
    assert np.all(np.abs(q[0].mean(0) - m) < 5 * np.sqrt(v / 30000))
    assert np.all(np.abs((q[0] == 0).mean(0) - z) < 0.012)


def test_meanfield_order_and_stopping():
    dbm = synth.random_dbm([20, 10, 6, 4], 2, wscale=(0.02, 0.2))
    x = np.random.default_rng(1).poisson(1.0, (15, 20)).astype(float)
    mu, it, delta = om.meanfield(dbm, x, eps=1e-3)
    assert delta <= 1e-3 and it >= 1
    mu2, _, _ = om.meanfield(dbm, x, eps=0.0, maxiter=it)
    for a, b in zip(mu, mu2):
        assert np.array_equal(a, b)
    # fixed point: one more Gauss-Seidel sweep moves by less than the last delta
    mu3, _, d3 = om.meanfield(dbm, x, eps=0.0, maxiter=it + 1)
    assert d3 <= delta + 1e-15
    # global stopping rule: appending a slow row changes the iteration count for every row
    assert mu[0] is x


def test_gradient_and_update_clamps():
    g = np.random.default_rng(2)
    v, vm = g.poisson(1.0, (8, 5)).astype(float), g.gamma(1.0, 1.0, (6, 5))
    h, hm = g.random((8, 3)), g.random((6, 3))
    gr = om.computegradient(v, vm, h, hm)
    assert np.allclose(gr.weights, v.T @ h / 8 - vm.T @ hm / 6)
    assert np.allclose(om.computegradient(v[:1], vm[:1], h[:1], hm[:1]).weights, v[:1].T @ h[:1] - vm[:1].T @ hm[:1])
    rbm = om.NegativeBinomialBernoulliRBM(-np.full((5, 3), 1e-3), np.full(5, -0.5), np.zeros(3), np.full(5, 0.5))
    om.updateparameters(rbm, gr, 10.0)
    assert (rbm.weights <= 0).all() and (rbm.visbias <= -1e-10).all()


def test_pcd_chainstate_semantics():
    """quirk B9 / SURVEY H7: chain state holds probabilities between minibatches; the short
    last batch keeps the full chain but only `thisbatchsize` rows enter the gradient."""
    x, _, _ = synth.counts(25, 12, seed=1)
    x[:, 0] += 1
    rng = px.Rng(7)
    rbm = om.initrbm(x, 5, rng, "nb")
    assert (rbm.weights < 0).all() and (rbm.weights >= -0.002).all() and (rbm.weights <= -0.0005).all()
    tr = om.RBMTrainer(rbm, 10, rng, pcd=True)
    assert tr.chainstate.shape == (10, 5) and ((tr.chainstate > 0) & (tr.chainstate < 1)).all()
    om.trainrbm(rbm, x, tr, 1e-3)
    assert tr.chainstate.shape == (10, 5) and not np.isin(tr.chainstate, (0.0, 1.0)).all()
    with pytest.raises(ValueError, match="exceeds number of samples"):
        om.randombatchmasks(5, 6, rng, 0)
    masks = om.randombatchmasks(25, 10, rng, 3)
    assert [len(m) for m in masks] == [10, 10, 5] and sorted(np.concatenate(masks)) == list(range(25))


def test_fitdbm_equals_stackrbms_plus_traindbm():
    """test/BMTest.jl:699-722 (seed reproducibility)."""
    x, _, _ = synth.counts(40, 16, seed=2)
    x[:, 0] += 1
    tl = [om.TrainLayer(nhidden=8, rbmtype="nb", batchsize=10, learningrate=1e-3), om.TrainLayer(nhidden=4, batchsize=10)]
    a = om.fitdbm(x, px.Rng(5), pretraining=tl, epochs=2, nparticles=8, learningrate=1e-3)
    pre = om.stackrbms(x, px.Rng(5), trainlayers=tl, epochs=2, predbm=True, learningrate=1e-3)
    b = om.traindbm(pre, x, px.Rng(5), epochs=2, nparticles=8, learningrate=1e-3)
    for ra, rb in zip(a, b):
        assert np.array_equal(ra.weights, rb.weights)
    assert np.allclose(om.defaultfinetuninglearningrates(0.005, 3), 0.005 * 11.0 / np.array([11.0, 12.0, 13.0]))


def test_sharding_invariance_of_the_oracle():
    dbm = synth.random_dbm([16, 8, 4], 3, wscale=(0.02, 0.2))
    whole = om.sampleparticles(dbm, 20, px.Rng(1), burnin=3)
    a = om.sampleparticles(dbm, 12, px.Rng(1), burnin=3, row_offset=0)
    b = om.sampleparticles(dbm, 8, px.Rng(1), burnin=3, row_offset=12)
    for w, x, y in zip(whole, a, b):
        assert np.array_equal(w, np.vstack([x, y]))
    wa = oais.aislogimpweights(dbm, px.Rng(2), ntemperatures=5, nparticles=6)
    wb = np.concatenate([oais.aislogimpweights(dbm, px.Rng(2), ntemperatures=5, nparticles=4),
                         oais.aislogimpweights(dbm, px.Rng(2), ntemperatures=5, nparticles=2, row_offset=4)])
    assert np.array_equal(wa, wb)


def test_golden_fixture_regression():
    """The committed fixtures are what the oracle produces today."""
    g = np.load(GOLD)
    dbm = [om.NegativeBinomialBernoulliRBM(g["W0"], g["a0"], g["b0"], g["r0"]), om.BernoulliRBM(g["W1"], g["a1"], g["b1"])]
    rng = px.Rng(int(g["seed"]))
    assert np.array_equal(om.hiddeninput(dbm[0], g["v"]), g["hiddeninput"])
    assert np.array_equal(om.visiblepotential(dbm[0], g["h1"]), g["visiblepotential"])
    assert np.array_equal(om.samplehidden(dbm[0], g["v"], rng, 1, 5), g["samplehidden_t5"])
    assert np.array_equal(om.samplevisible(dbm[0], g["h1"], rng, 0, 6), g["samplevisible_t6"])
    p = om.initparticles(dbm, 32, rng)
    assert np.array_equal(p[0], g["init0"])
    om.gibbssample_dbm(p, dbm, rng, 3)
    assert np.array_equal(p[0], g["gibbs0"]) and np.array_equal(p[2], g["gibbs2"])
    mu, it, _ = om.meanfield(dbm, g["v"], eps=1e-6)
    assert it == int(g["mf_iters"]) and np.array_equal(mu[2], g["mf2"])
    assert np.array_equal(oais.aislogimpweights(dbm, rng, ntemperatures=10, nparticles=16), g["aisw"])


def test_next_rows_oracle():
    """Conditional Gibbs clamps, dropout GLM score equations, Fisher scoring recovers the dispersion."""
    dbm = synth.random_dbm([20, 8, 4], 2, wscale=(0.02, 0.2))
    p = om.initparticles(dbm, 10, px.Rng(1))
    o = p[0].copy()
    mask = np.zeros(20, bool)
    mask[[1, 5, 7]] = True
    om.gibbssamplecond(p, dbm, px.Rng(1), mask, 3)
    assert np.array_equal(p[0][:, mask], o[:, mask]) and p.clock == 4
    x, _, _ = synth.counts(300, 40, seed=3)
    x[0, :] += 1
    b0, b1 = om.estimatedropout(x)
    lc, zf = np.log1p(x.mean(0)), (x == 0).mean(0)
    mu = 1 / (1 + np.exp(-(b0 + b1 * lc)))
    assert abs(np.sum(zf - mu)) < 1e-8 and abs(np.sum((zf - mu) * lc)) < 1e-8      # GLM score equations
    assert b1 < 0                                                                    # more expression, fewer zeros
    m = om.sampledropout(x, 0.5, 1.0, px.Rng(2), 7)
    assert set(np.unique(m)) <= {0.0, 1.0}
    g = np.random.default_rng(5)
    n, theta_true = 4000, np.array([0.5, 2.0, 5.0])
    mu = g.uniform(0.5, 6.0, (n, 3))
    y = g.poisson(g.gamma(theta_true[None, :], mu / theta_true[None, :])).astype(float)
    th = om.mle_for_theta(y, mu, lam=0.0, maxiter=50)
    assert np.all(np.abs(th / theta_true - 1) < 0.15)
    assert np.isclose(om.mle_for_theta(y[:, 0], mu[:, 0], 0.0, maxiter=50), th[0])


def test_philox_counts_generator_statistics_and_sharding():
    """oracle/synth.philox_counts (twin of scdbm_mat_generate_counts): NB(r, r/(s m + r)) moments, shard invariance."""
    from oracle import synth
    g = np.random.default_rng(2)
    V, n = 12, 40000
    gm = np.array([0.05, 0.2, 1.0, 3.0, 7.0, 20.0, 0.0, 0.5, 2.0, 0.1, 12.0, 0.8])
    r = g.uniform(0.3, 2.0, V)
    x, s = synth.philox_counts(n, gm, r, 0.3, px.Rng(9))
    assert abs(s.mean() - np.exp(0.045)) < 0.01 and abs(np.log(s).std() - 0.3) < 0.01
    assert np.all(x[:, 6] == 0)
    sel = gm > 0
    assert np.allclose(x.mean(axis=0)[sel], gm[sel] * np.exp(0.045), rtol=0.08)
    # variance of the NB-lognormal mixture: E[mu] + E[mu^2] (1 + 1/r) - E[mu]^2
    e1, e2 = gm * np.exp(0.045), gm ** 2 * np.exp(0.18)
    want_var = e1 + e2 * (1 + 1 / r) - e1 ** 2
    assert np.allclose(x.var(axis=0)[sel], want_var[sel], rtol=0.25)
    part, _ = synth.philox_counts(50, gm, r, 0.3, px.Rng(9), row_offset=100)
    assert np.array_equal(part, x[100:150])
    ptr, idx, val = synth.to_csr(x[:200])
    assert np.array_equal(synth.from_compressed((200, V), ptr, idx, val), x[:200])
    cptr, cidx, cval = synth.to_csr(x[:200].T)
    assert np.array_equal(synth.from_compressed((200, V), cptr, cidx, cval, major="csc"), x[:200])
