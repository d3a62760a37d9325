"""Parity at the layer shapes of every BASELINE.json configuration (C1, C2, C4, C5; C3 lives in
test_gpu_tc.py::test_full_size_generation_properties), plus the handle-lifetime regressions.

The oracle cannot follow at the full row counts of C4/C5 (100 000 cells, 65 536 particles x 10 000
temperatures), so these tests keep the full *layer* shapes -- K = 20 000 contractions through fp16 hi/lo
planes with a model-wide power-of-two weight scale are where gate G1 could break -- on row slices the
NumPy oracle finishes in seconds.  Every test prints the error / mismatch figures it asserts on.
"""
import copy
import gc

import numpy as np
import pytest

from conftest import to_pkg_model, relerr
from oracle import model as om, philox as px, sampling as osp, ais as oais, synth

pytestmark = pytest.mark.gpu
RTOL = 1e-4   # gate G1
SEED = 20261019
C4_UNITS = [20000, 1024, 256, 64]


@pytest.fixture(scope="module")
def auto(pkg):
    """AUTO engine: what a user gets (tcgen05 for large shapes, CUDA-core kernel for tiny ones)."""
    return pkg.Context(0, SEED)


@pytest.fixture(scope="module")
def c4(pkg, auto):
    dbm = synth.random_dbm(C4_UNITS, 41)
    return dbm, pkg.DeviceModel.from_host(to_pkg_model(pkg, dbm), auto)


# ------------------------------------------------------------ handle lifetime (round-1 use-after-free)
def test_layer_of_a_temporary_particles_object(pkg, auto):
    """`sampleparticles(...).layer(0)` on a temporary: the particles wrapper is collected as soon as the chained
    call returns; the shared matrix handle must stay valid (C ABI reference count + Python keep-alive)."""
    dbm = synth.random_dbm([130, 67, 33], 3, wscale=(0.02, 0.2))
    m = pkg.DeviceModel.from_host(to_pkg_model(pkg, dbm), auto)
    ref = pkg.sampleparticles(m, 300, 3, device=True)
    want = ref.layer(0).to_host_u16()
    mat = pkg.sampleparticles(m, 300, 3, device=True).layer(0)
    gc.collect()
    junk = [pkg.Mat(auto, 300, 130, pkg.MAT_COUNT) for _ in range(4)]   # re-use freed pool memory if there were any
    assert np.array_equal(mat.to_host_u16(), want)
    assert np.array_equal(pkg.sampleparticles(m, 300, 3, device=True).layer(0).to_host(), want.astype(float))
    # at the ABI: destroy the particles first, then use and release the shared handle
    import ctypes as C
    lib = pkg.lib()
    p = pkg.DeviceParticles(m, 64).init(False)
    h = C.c_void_p()
    pkg._lib.check(lib.scdbm_particles_layer(p.h, 1, C.byref(h)))
    p.close()
    r, c, k = C.c_int64(), C.c_int64(), C.c_int32()
    pkg._lib.check(lib.scdbm_mat_shape(h, C.byref(r), C.byref(c), C.byref(k)))
    assert (r.value, c.value, k.value) == (64, 67, pkg.MAT_BINARY)
    out = np.zeros((64, 67), order="F")
    pkg._lib.check(lib.scdbm_mat_download(h, out.ctypes.data_as(pkg._lib.D), 64))
    assert set(np.unique(out)) <= {0.0, 1.0}
    pkg._lib.check(lib.scdbm_mat_destroy(h))
    del junk


def test_trainer_handles_outlive_the_trainer(pkg, auto):
    x, _, _ = synth.counts(60, 40, seed=1)
    x[0, :] += 1.0
    xm = pkg.Mat.from_host(auto, x, pkg.MAT_COUNT)
    m = pkg.DeviceModel(auto, [40, 16], pkg.VIS_NEGBIN)
    m.init_layer(0, xm, seed=SEED)
    tr = pkg.RBMTrainer(m, 20, pcd=True)
    tr.track_mean(True)
    pkg.trainrbm_(m, xm, tr, 1e-3, 1)
    want_chain, want_est = tr.chainstate().to_host(), tr.estimated_mean().to_host()
    chain, est = tr.chainstate(), tr.estimated_mean()
    tr.close()
    del tr
    gc.collect()
    assert np.array_equal(chain.to_host(), want_chain) and np.array_equal(est.to_host(), want_est)
    # mean-field particles: layer 0 is a view of the data matrix and keeps it alive
    dbm = synth.random_dbm([40, 16, 8], 4, wscale=(0.02, 0.1))
    md = pkg.DeviceModel.from_host(to_pkg_model(pkg, dbm), auto)
    mu = pkg.meanfield(md, x, eps=1e-3, device=True, ctx=auto)
    lay0 = mu.layer(0)
    del mu
    gc.collect()
    assert np.array_equal(lay0.to_host(), x)


# ------------------------------------------------------------ C1: NB-RBM CD-1, 500 x 1000, H = 128, B = 100
def test_c1_fitrbm_cd1_epoch(pkg, auto):
    """src/rbmtraining.jl:498-591 at the BASELINE C1 shape: one epoch = 5 minibatches of 100."""
    n, V, H, B = 500, 1000, 128, 100
    x, _, _ = synth.counts(n, V, seed=1)
    x[0, :] += 1.0
    rng = px.Rng(SEED)
    rbm = om.initrbm(x, H, rng, "nb")
    tr = om.RBMTrainer(rbm, B, rng, pcd=False)
    xm = pkg.Mat.from_host(auto, x, pkg.MAT_COUNT)
    m = pkg.DeviceModel(auto, [V, H], pkg.VIS_NEGBIN)
    m.init_layer(0, xm, seed=SEED)
    g0 = m.get_layer(0)
    assert relerr(g0.weights, rbm.weights) < 1e-6 and relerr(g0.visbias, rbm.visbias) < 1e-5
    dt = pkg.RBMTrainer(m, B, pcd=False)
    lr = 1e-4
    for _ in range(2):
        om.trainrbm(rbm, x, tr, lr, 1)
        pkg.trainrbm_(m, xm, dt, lr, 1)
    assert dt.clock == tr.clock
    got = m.get_layer(0)
    d_ref, d_got = rbm.weights - g0.weights, got.weights - g0.weights
    e_step = np.max(np.abs(d_got - d_ref)) / np.abs(d_ref).max()
    e_w = np.max(np.abs(got.weights - rbm.weights)) / np.abs(rbm.weights).max()
    e_a = np.max(np.abs(got.visbias - rbm.visbias)) / np.abs(rbm.visbias).max()
    e_b = np.max(np.abs(got.hidbias - rbm.hidbias)) / max(np.abs(rbm.hidbias).max(), 1e-12)
    print(f"[C1] update error {e_step:.2e} of the largest step; parameters: W {e_w:.2e}, a {e_a:.2e}, b {e_b:.2e}")
    # sampled h flips (accumulation-order ambiguity, gate G2) move single entries of the update a little
    assert e_step < 0.02
    assert e_w < RTOL and e_a < RTOL and e_b < 5e-3


# ------------------------------------------------------------ C2: one traindbm! step, 5000 x 2000 -> 256 -> 64
def test_c2_traindbm_step(pkg, auto):
    """src/dbmtraining.jl:287-297 at the BASELINE C2 shape: Gibbs on 100 particles, mean-field over 5000 cells,
    gradient, update."""
    units, n, P = [2000, 256, 64], 5000, 100
    x, _, _ = synth.counts(n, units[0], seed=2)
    dbm = synth.random_dbm(units, 3)
    m = pkg.DeviceModel.from_host(to_pkg_model(pkg, dbm), auto)
    ref = copy.deepcopy(dbm)
    rng = px.Rng(SEED)
    parts = om.initparticles(ref, P, rng)
    p = pkg.DeviceParticles(m, P).init(False)
    # deterministic part first (G1): mean-field at a fixed iteration count and at eps = 1e-6
    mu_ref, _, _ = om.meanfield(dbm, x, eps=0.0, maxiter=3)
    mu_got = pkg.meanfield(m, x, eps=0.0, maxiter=3, ctx=auto)
    errs = [relerr(a, b) for a, b in zip(mu_got[1:], mu_ref[1:])]
    print(f"[C2] mean-field (3 iterations) relative error per layer {errs}")
    assert max(errs) < RTOL
    lr = 1e-3
    om.traindbm_step(ref, x, parts, rng, lr)
    pkg.traindbm_(m, x, epochs=1, particles=p, learningrates=[lr], ctx=auto)
    for l in range(2):
        got = m.get_layer(l)
        dref = ref[l].weights - dbm[l].weights
        e = np.max(np.abs((got.weights - dbm[l].weights) - dref)) / np.abs(dref).max()
        print(f"[C2] layer {l}: update error {e:.2e} of the largest step")
        assert e < 0.05      # negative phase: 100 sampled particles, a few flipped bits (G2) move the step
        assert np.max(np.abs(got.weights - ref[l].weights)) < RTOL * np.abs(ref[l].weights).max()


# ------------------------------------------------------------ C4: 20000 -> 1024 -> 256 -> 64
def test_c4_conditionals(pkg, auto, c4):
    """Every conditional operator at the C4 layer widths (K = 20 000 through fp16 hi/lo planes), gate G1."""
    dbm, m = c4
    n = 640
    x, _, _ = synth.counts(n, C4_UNITS[0], seed=4)
    x[3, 17] = 5000.0                                          # hi count plane in one row tile
    g = np.random.default_rng(5)
    h = [(g.random((n, u)) < 0.5).astype(float) for u in C4_UNITS[1:]]
    e = {}
    e["hiddeninput"] = relerr(pkg.hiddeninput(m, x, ctx=auto), om.hiddeninput(dbm[0], x), 1e-6)
    for f in (1.0, 2.0):
        e[f"hiddenpotential(f={f})"] = relerr(pkg.hiddenpotential(m, x, f, ctx=auto), om.hiddenpotential(dbm[0], x, f), 1e-7)
    e["visibleinput"] = relerr(pkg.visibleinput(m, h[0], ctx=auto), om.visibleinput(dbm[0], h[0]))
    e["NB mean"] = relerr(pkg.visiblepotential(m, h[0], ctx=auto), om.visiblepotential(dbm[0], h[0]))
    for layer in (1, 2):
        e[f"hiddenpotential L{layer}"] = relerr(pkg.hiddenpotential(m, h[layer - 1], 1.0, layer=layer, ctx=auto),
                                                om.hiddenpotential(dbm[layer], h[layer - 1]))
        e[f"visiblepotential L{layer}"] = relerr(pkg.visiblepotential(m, h[layer], 2.0, layer=layer, ctx=auto),
                                                 om.visiblepotential(dbm[layer], h[layer], 2.0))
    parts = [x] + h
    for layer in (1, 2):
        xin = om.dbm_layer_input(dbm, parts, layer)
        e[f"dbm layer {layer} input"] = relerr(pkg.dbmlayer(m, layer, parts[layer - 1], parts[layer + 1], what=pkg.OUT_INPUT, ctx=auto),
                                               xin, 0.05 * np.abs(xin).max())
        e[f"dbm layer {layer}"] = relerr(pkg.dbmlayer(m, layer, parts[layer - 1], parts[layer + 1], ctx=auto), osp.sigm(xin), 1e-7)
    print("[C4] G1 relative errors:", {k: f"{v:.2e}" for k, v in e.items()})
    assert max(e.values()) < RTOL
    # sampled states from the same Philox stream (G2): NB visible layer (20 000 genes) and h1 | v, h2
    rng = px.Rng(SEED)
    v_got, v_ref = pkg.samplevisible(m, h[0], tick=6, ctx=auto), om.samplevisible(dbm[0], h[0], rng, 0, 6)
    h_got = pkg.dbmlayer(m, 1, x, h[1], what=pkg.OUT_SAMPLE, tick=7, ctx=auto)
    h_ref = osp.bernoulli(osp.sigm(om.dbm_layer_input(dbm, parts, 1)), rng.uniforms(n, C4_UNITS[1], px.stream_id(1), 7))
    mv, mh = float(np.mean(v_got != v_ref)), float(np.mean(h_got != h_ref))
    print(f"[C4] G2 mismatch fractions: NB visible {mv:.2e}, h1 {mh:.2e}")
    assert mv < 1e-3 and mh < 1e-3


def test_c4_meanfield_and_gradient(pkg, auto, c4):
    """Mean-field (src/dbmtraining.jl:175-228) at 1 and 3 iterations and the StackedOptimizer gradient
    (src/optimizers.jl:431-442, K5) at the C4 shapes."""
    dbm, m = c4
    n, P = 1536, 512
    x, _, _ = synth.counts(n, C4_UNITS[0], seed=6)
    for maxiter in (1, 3):
        got = pkg.meanfield(m, x, eps=0.0, maxiter=maxiter, ctx=auto)
        ref, _, _ = om.meanfield(dbm, x, eps=0.0, maxiter=maxiter)
        errs = [relerr(a, b) for a, b in zip(got[1:], ref[1:])]
        print(f"[C4] mean-field, {maxiter} iteration(s): relative error per layer {errs}")
        assert max(errs) < RTOL
    got, it, delta = pkg.meanfield(m, x, eps=1e-3, return_info=True, ctx=auto)
    ref, rit, _ = om.meanfield(dbm, x, eps=1e-3)
    print(f"[C4] mean-field eps=1e-3: {it} iterations (oracle {rit}), delta {delta:.2e}")
    assert it == rit
    # gradient for FIXED states: mu from the oracle, particles random binary / counts
    g = np.random.default_rng(7)
    parts = [g.poisson(0.4, (P, C4_UNITS[0])).astype(float)] + [(g.random((P, u)) < 0.5).astype(float) for u in C4_UNITS[1:]]
    mu = pkg.meanfield(m, x, eps=0.0, maxiter=2, device=True, ctx=auto)
    mu_ref, _, _ = om.meanfield(dbm, x, eps=0.0, maxiter=2)
    p = pkg.DeviceParticles(m, P).upload(parts)
    opt = pkg.loglikelihoodoptimizer(m, learningrate=0.01)
    opt.computegradient_dbm_(mu, p, m)
    grads = om.computegradient_dbm(mu_ref, parts, dbm)
    for l, gr in enumerate(grads):
        W, a, b = opt.gradient(l)
        eW = np.max(np.abs(W - gr.weights)) / np.abs(gr.weights).max()
        ea = np.max(np.abs(a - gr.visbias)) / np.abs(gr.visbias).max()
        eb = np.max(np.abs(b - gr.hidbias)) / np.abs(gr.hidbias).max()
        print(f"[C4] gradient layer {l}: dW {eW:.2e}, da {ea:.2e}, db {eb:.2e} (norm-wise)")
        assert max(eW, ea, eb) < RTOL


# ------------------------------------------------------------ C5: AIS on the C4 model
def test_c5_ais_slice(pkg, auto, c4):
    """aislogimpweights (src/evaluating.jl:164-206) on the 3-layer scDBM: 128 particles x 4 temperature steps.
    Same Philox stream on both sides -> same trajectories up to rare flipped bits; the log-weight increments of a
    particle whose trajectory matches agree to G1."""
    dbm, m = c4
    temps = np.array([0.0, 2.5e-4, 5e-4, 7.5e-4, 1e-3])        # the first steps of collect(0:1e-4:1)-style annealing
    P, burnin = 128, 2
    w = pkg.aislogimpweights(m, temperatures=temps, nparticles=P, burnin=burnin, ctx=auto)
    wr = oais.aislogimpweights(dbm, px.Rng(SEED), temperatures=temps, nparticles=P, burnin=burnin)
    err = np.abs(w - wr) / np.maximum(1.0, np.abs(wr))
    print(f"[C5] AIS log-weights: median relative error {np.median(err):.2e}, max {err.max():.2e}, "
          f"particles beyond 1e-3: {int((err > 1e-3).sum())} of {P}")
    assert np.median(err) < RTOL
    assert np.mean(err > 1e-3) < 0.05
    # sharded == unsharded
    w2 = np.concatenate([pkg.aislogimpweights(m, temperatures=temps, nparticles=P // 2, burnin=burnin, ctx=auto),
                         pkg.aislogimpweights(m, temperatures=temps, nparticles=P // 2, burnin=burnin, row_offset=P // 2, ctx=auto)])
    assert np.allclose(w2, w, rtol=1e-10, atol=1e-10)


# ------------------------------------------------------------ communicator (single rank: the NCCL path itself)
def test_single_rank_communicator_matches_plain_step(pkg):
    """A 1-rank NCCL communicator exercises every collective call of the data-parallel step (allreduce of the gradient
    blocks on the second stream, allreduce(max) of the mean-field delta, scalar reductions); the result must equal
    the plain step bit for bit.  (More ranks need more GPUs: `bench.py --gpus N` and tests/test_multi_gloo.py.)"""
    units = [300, 128, 64]
    dbm = synth.random_dbm(units, 5, wscale=(0.002, 0.02))
    x, _, _ = synth.counts(500, units[0], seed=3)
    res = []
    for use_comm in (False, True):
        c = pkg.Context(0, SEED)
        if use_comm:
            c.comm_init(pkg.Context.comm_unique_id(), 0, 1)
            assert (c.comm_rank, c.comm_size) == (0, 1)
            assert np.array_equal(c.allreduce([1.5, -2.0], "sum"), [1.5, -2.0])
        m = pkg.DeviceModel.from_host(to_pkg_model(pkg, dbm), c)
        xm = pkg.Mat.from_host(c, x, pkg.MAT_COUNT)
        p = pkg.DeviceParticles(m, 64).init(False)
        mu = pkg.DeviceParticles(m, 500, real_valued=True)
        opt = pkg.LoglikelihoodOptimizer().initialized(m)
        import ctypes as C
        its = C.c_int32()
        for _ in range(2):
            if use_comm:
                pkg._lib.check(pkg.lib().scdbm_dbm_pcd_step_dp(m.h, xm.h, p.h, mu.h, opt.h, 0.01, 5, 1e-3, 500, 64, C.byref(its)))
            else:
                pkg._lib.check(pkg.lib().scdbm_dbm_pcd_step(m.h, xm.h, p.h, mu.h, opt.h, 0.01, 5, 1e-3, C.byref(its)))
        res.append(([m.get_layer(l).weights for l in range(2)], its.value))
        if use_comm:
            with pytest.raises(pkg.ScdbmError):          # the plain step refuses to run on a context with a communicator
                pkg._lib.check(pkg.lib().scdbm_dbm_pcd_step(m.h, xm.h, p.h, mu.h, opt.h, 0.01, 5, 1e-3, None))
            w = np.array([0.3, -1.0, 2.0])
            assert abs(pkg.logmeanexp_sharded(w, c) - pkg.logmeanexp(w)) < 1e-12
            c.comm_destroy()
    assert res[0][1] == res[1][1]
    for a, b in zip(res[0][0], res[1][0]):
        assert np.array_equal(a, b)


def test_pair_mode_with_chunked_accumulation(pkg, auto):
    """>= 2*128*148 rows (256-row work tiles) AND a contraction longer than 72 k-blocks (chunk-promoted accumulation):
    the combination the positive phase of C4 runs in (100 000 cells x 20 000 genes).  Ragged rows and width."""
    n, V, H = 2 * 128 * 148 + 77, 5000, 130
    dbm = synth.random_dbm([V, H, 16], 9)
    m = pkg.DeviceModel.from_host(to_pkg_model(pkg, dbm), auto)
    g = np.random.default_rng(3)
    x = g.poisson(0.3, (n, V)).astype(float)
    x[n - 3, 11] = 4000.0
    e_in = relerr(pkg.hiddeninput(m, x, ctx=auto), om.hiddeninput(dbm[0], x), 1e-6)
    e_pot = relerr(pkg.hiddenpotential(m, x, 2.0, ctx=auto), om.hiddenpotential(dbm[0], x, 2.0), 1e-7)
    print(f"[pair + chunks] hiddeninput {e_in:.2e}, hiddenpotential {e_pot:.2e}")
    assert e_in < RTOL and e_pot < RTOL


# ------------------------------------------------------------ two-SM (cta_group::2) sweep kernel
@pytest.fixture()
def two_sm(pkg):
    """A context whose wide Bernoulli / deterministic sweeps run the cta_group::2 kernel (csrc/gemm_tc2.cuh)."""
    c = pkg.Context(0, SEED)
    c.set("two_sm", 1)
    assert c.get("two_sm") == 1.0
    return c


def test_two_sm_deterministic_sweeps(pkg, two_sm):
    """256 x 256 cluster tiles, ragged in rows and width, count data with a >= 2048 entry in the LAST row tile
    (hi-plane pass of one cluster tile only), K long enough for chunk-promoted accumulation; dual-source DBM layer."""
    n, V, H, H2 = 256 * 40 + 77, 5000, 300, 40
    dbm = synth.random_dbm([V, H, H2], 13)
    m = pkg.DeviceModel.from_host(to_pkg_model(pkg, dbm), two_sm)
    g = np.random.default_rng(5)
    x = g.poisson(0.3, (n, V)).astype(float)
    x[n - 3, 11] = 4000.0
    x[300, 4999] = 2049.0
    before = two_sm.get("tc_launches")
    e_in = relerr(pkg.hiddeninput(m, x, ctx=two_sm), om.hiddeninput(dbm[0], x), 1e-6)
    e_pot = relerr(pkg.hiddenpotential(m, x, 2.0, ctx=two_sm), om.hiddenpotential(dbm[0], x, 2.0), 1e-7)
    assert two_sm.get("tc_launches") - before == 2
    h2 = (g.random((n, H2)) < 0.5).astype(float)
    xin = om.dbm_layer_input(dbm, [x, None, h2], 1)
    e_dbm = relerr(pkg.dbmlayer(m, 1, x, h2, what=pkg.OUT_INPUT, ctx=two_sm), xin, 0.05 * np.abs(xin).max())
    e_dbp = relerr(pkg.dbmlayer(m, 1, x, h2, ctx=two_sm), osp.sigm(xin), 1e-7)
    print(f"[two-SM] hiddeninput {e_in:.2e}, hiddenpotential {e_pot:.2e}, dbm layer input {e_dbm:.2e}, potential {e_dbp:.2e}")
    assert max(e_in, e_pot, e_dbm, e_dbp) < RTOL


def test_two_sm_bernoulli_draws_match_single_sm(pkg, two_sm, auto):
    """Same seed, same tick: the two-SM kernel and the single-SM kernel consume the same Philox words against the same
    potentials -- identical samples up to potentials that differ in the last bit next to a uniform (gate G2), and the
    samples match the oracle's draws from the same uniforms."""
    n, V, H, H2 = 256 * 38 + 5, 700, 520, 64
    dbm = synth.random_dbm([V, H, H2], 17)
    g = np.random.default_rng(7)
    x = g.poisson(0.5, (n, V)).astype(float)
    h2 = (g.random((n, H2)) < 0.5).astype(float)
    outs = []
    for c in (two_sm, auto):
        c.set("seed", SEED)
        m = pkg.DeviceModel.from_host(to_pkg_model(pkg, dbm), c)
        outs.append((pkg.samplehidden(m, x, 1.0, tick=77, ctx=c), pkg.dbmlayer(m, 1, x, h2, what=pkg.OUT_SAMPLE, tick=78, ctx=c)))
    for a, b in zip(*outs):
        assert set(np.unique(a)) <= {0.0, 1.0}
        mism = float(np.mean(a != b))
        print(f"[two-SM] Bernoulli draws differing from the single-SM kernel: {mism:.2e}")
        assert mism < 1e-5
    # statistics against the potential (gate G3): per-unit mean within 5 sigma
    pot = om.hiddenpotential(dbm[0], x)
    z = (outs[0][0].mean(0) - pot.mean(0)) / np.sqrt((pot * (1 - pot)).mean(0) / n + 1e-12)
    assert np.abs(z).max() < 5.0


@pytest.mark.parametrize("V,H,n,n2", [(3000, 3300, 16384 + 53, 300), (2000, 2561, 2500, 97)])
def test_two_sm_gradient(pkg, auto, V, H, n, n2):
    """Two-SM gradient kernel (k_grad_cg2): one or more waves of 256 x 256 cluster tiles, ragged in both unit axes (the
    second case has a hidden width that is not a multiple of 4: scalar stores), chunk-promoted accumulation, split-K
    evening out the last wave, count x real positive phase with one 128-sample block that uses the hi count plane,
    count x binary negative phase."""
    g = np.random.default_rng(11)
    v = g.poisson(0.4, (n, V)).astype(float)
    v[700, 5] = 5000.0
    h = g.random((n, H))
    vm = g.poisson(0.4, (n2, V)).astype(float)
    hm = (g.random((n2, H)) < 0.5).astype(float)
    m = pkg.DeviceModel(auto, [V, H], pkg.VIS_NEGBIN)
    opt = pkg.loglikelihoodoptimizer(m, learningrate=0.01)
    mats = [pkg.Mat.from_host(auto, v, pkg.MAT_COUNT), pkg.Mat.from_host(auto, vm, pkg.MAT_COUNT),
            pkg.Mat.from_host(auto, h, pkg.MAT_REAL), pkg.Mat.from_host(auto, hm, pkg.MAT_BINARY)]
    opt.computegradient_(*mats, m)
    W, a, b = opt.gradient(0)
    refW = v.T @ h / n - vm.T @ hm / n2
    refa = v.mean(0) - vm.mean(0)
    refb = h.mean(0) - hm.mean(0)
    eW = np.max(np.abs(W - refW)) / np.abs(refW).max()
    ea = np.max(np.abs(a - refa)) / np.abs(refa).max()
    eb = np.max(np.abs(b - refb)) / np.abs(refb).max()
    print(f"[two-SM gradient] dW {eW:.2e}, da {ea:.2e}, db {eb:.2e} (norm-wise)")
    assert max(eW, ea, eb) < RTOL


def test_two_sm_meanfield(pkg, two_sm):
    """Mean-field at a row count where the layer-1 sweep (hoisted addend, previous value, max |change|) runs the two-SM
    kernel with the transposed accumulator: ragged rows, explicit iteration counts and the stopping rule."""
    units = [1500, 1024, 256, 64]
    n = 256 * 19 + 41
    dbm = synth.random_dbm(units, 23)
    m = pkg.DeviceModel.from_host(to_pkg_model(pkg, dbm), two_sm)
    x, _, _ = synth.counts(n, units[0], seed=8)
    for maxiter in (1, 3):
        got = pkg.meanfield(m, x, eps=0.0, maxiter=maxiter, ctx=two_sm)
        ref, _, _ = om.meanfield(dbm, x, eps=0.0, maxiter=maxiter)
        errs = [relerr(a, b) for a, b in zip(got[1:], ref[1:])]
        print(f"[two-SM] mean-field, {maxiter} iteration(s): relative error per layer {errs}")
        assert max(errs) < RTOL
    got, it, delta = pkg.meanfield(m, x, eps=1e-3, return_info=True, ctx=two_sm)
    ref, rit, rdelta = om.meanfield(dbm, x, eps=1e-3)
    print(f"[two-SM] mean-field eps=1e-3: {it} iterations (oracle {rit}), delta {delta:.3e} (oracle {rdelta:.3e})")
    assert it == rit and abs(delta - rdelta) < 1e-4 * max(rdelta, 1e-6) + 1e-7


def test_c2_gibbs_draws_match_oracle(pkg, auto):
    """100 particles of the C2 model, same Philox counters as the oracle (gate G2).  At this row count the dual-source
    up-sweep h1 | v, h2 (K = 2000 + 64) runs split-K with the Bernoulli draw in the finishing launch."""
    units, P = [2000, 256, 64], 100
    dbm = synth.random_dbm(units, 3)
    m = pkg.DeviceModel.from_host(to_pkg_model(pkg, dbm), auto)
    auto.set("seed", SEED)
    p = pkg.initparticles(m, P, device=True)
    ref = om.initparticles(dbm, P, px.Rng(SEED))
    before = auto.launch_count()
    pkg.gibbssample_(p, m, 3)
    print(f"[C2] launches per Gibbs sweep: {(auto.launch_count() - before) / 3:.1f}")
    om.gibbssample_dbm(ref, dbm, px.Rng(SEED), 3)
    mism = [float(np.mean(a != b)) for a, b in zip(p.to_host(), ref)]
    print(f"[C2] fraction of differing entries per layer after 3 sweeps: {mism}")
    assert max(mism) < 5e-3


def test_programmatic_dependent_launch_changes_nothing(pkg):
    """The kernels of a training chain are launched with programmatic stream serialization and wait
    (griddepcontrol.wait) before touching memory: the same epochs with plain launches must give the same bits."""
    n, V, H, B = 500, 1000, 128, 100
    x, _, _ = synth.counts(n, V, seed=1)
    out = []
    for pdl in (1, 0):
        c = pkg.Context(0, SEED)
        c.set("pdl", pdl)
        assert c.get("pdl") == float(pdl)
        xm = pkg.Mat.from_host(c, x, pkg.MAT_COUNT)
        m = pkg.DeviceModel(c, [V, H], pkg.VIS_NEGBIN)
        m.init_layer(0, xm, seed=SEED)
        tr = pkg.RBMTrainer(m, B, pcd=True)
        for _ in range(6):
            pkg.trainrbm_(m, xm, tr, 1e-3, 1)
        out.append(m.get_layer(0))
        c.set("pdl", 1)
    for f in ("weights", "visbias", "hidbias"):
        assert np.array_equal(getattr(out[0], f), getattr(out[1], f)), f


def test_two_sm_nb_draws_match_single_sm(pkg, two_sm, auto):
    """NB class with a long contraction (K = 1024, the hidden width of C4 / C5): the 20-warp kernel in its two-SM form
    (256-row cluster tiles, half a weight tile per CTA) draws exactly what the single-SM form draws -- same Philox
    counters, same accumulators -- and both agree with the oracle (gate G2)."""
    units, P = [2000, 1024, 64], 256 * 6 + 77
    dbm = synth.random_dbm(units, 29)
    outs = []
    for c, flag in ((two_sm, 1), (auto, 0)):
        c.set("seed", SEED)
        c.set("two_sm", flag)
        m = pkg.DeviceModel.from_host(to_pkg_model(pkg, dbm), c)
        p = pkg.initparticles(m, P, device=True)
        pkg.gibbssample_(p, m, 2)
        outs.append(p.to_host())
    auto.set("two_sm", 1)
    for a, b in zip(*outs):
        assert np.array_equal(a, b)
    ref = om.initparticles(dbm, P, px.Rng(SEED))
    om.gibbssample_dbm(ref, dbm, px.Rng(SEED), 2)
    mism = [float(np.mean(a != b)) for a, b in zip(outs[0], ref)]
    print(f"[two-SM NB] fraction of entries differing from the oracle after 2 sweeps: {mism}")
    assert max(mism) < 5e-3
