"""The tcgen05/TMA/TMEM engine forced on (engine=2), checked against the oracle
and against the CUDA-core engine on the same inputs."""
import numpy as np
import pytest

from conftest import to_pkg_model, relerr
from oracle import model as om, philox as px, sampling as osp, synth, ais as oais

pytestmark = pytest.mark.gpu
RTOL = 1e-4
SEED = 20261019


@pytest.fixture(scope="module")
def tc(pkg):
    c = pkg.Context(0, SEED)
    c.set("engine", pkg.ENGINE_TC)
    return c


@pytest.fixture(scope="module")
def simt(pkg):
    c = pkg.Context(0, SEED)
    c.set("engine", pkg.ENGINE_SIMT)
    return c


@pytest.mark.parametrize("n,units", [(5, (7, 5, 3)), (128, (64, 64, 64)), (129, (130, 67, 33)), (300, (500, 256, 64)),
                                     (1000, (2000, 256, 64))])
def test_tc_conditionals(pkg, tc, n, units):
    dbm = synth.random_dbm(list(units), 3, wscale=(0.002, 0.02))
    m = pkg.DeviceModel.from_host(to_pkg_model(pkg, dbm), tc)
    g = np.random.default_rng(1)
    v = g.poisson(1.5, (n, units[0])).astype(float)
    v[0, 0] = 3000.0
    h1 = (g.random((n, units[1])) < 0.5).astype(float)
    h2 = (g.random((n, units[2])) < 0.5).astype(float)
    tc.launch_count(reset=True)
    assert relerr(pkg.hiddeninput(m, v, ctx=tc), om.hiddeninput(dbm[0], v), 1e-6) < RTOL
    assert relerr(pkg.hiddenpotential(m, v, 2.0, ctx=tc), om.hiddenpotential(dbm[0], v, 2.0)) < RTOL
    assert relerr(pkg.visibleinput(m, h1, ctx=tc), om.visibleinput(dbm[0], h1)) < RTOL
    assert relerr(pkg.visiblepotential(m, h1, ctx=tc), om.visiblepotential(dbm[0], h1)) < RTOL
    assert relerr(pkg.hiddenpotential(m, h1, layer=1, ctx=tc), om.hiddenpotential(dbm[1], h1)) < RTOL
    assert relerr(pkg.visiblepotential(m, h2, 2.0, layer=1, ctx=tc), om.visiblepotential(dbm[1], h2, 2.0)) < RTOL
    mu1 = g.random((n, units[1]))                       # real-valued (mean-field) operand: hi+lo planes
    assert relerr(pkg.hiddenpotential(m, mu1, layer=1, ctx=tc), om.hiddenpotential(dbm[1], mu1)) < RTOL
    ref = osp.sigm(om.dbm_layer_input(dbm, [v, h1, h2], 1))
    assert relerr(pkg.dbmlayer(m, 1, v, h2, ctx=tc), ref, 1e-7) < RTOL
    xin = om.dbm_layer_input(dbm, [v, h1, h2], 1)
    assert relerr(pkg.dbmlayer(m, 1, v, h2, what=pkg.OUT_INPUT, ctx=tc), xin, 0.05 * np.abs(xin).max()) < RTOL
    assert tc.get("tc_launches") >= 8                   # the tensor-core kernel really ran


def test_tc_sampling_matches_simt_and_oracle(pkg, tc, simt):
    units = [500, 256, 64]
    dbm = synth.random_dbm(units, 3)
    pm = to_pkg_model(pkg, dbm)
    mt, ms = pkg.DeviceModel.from_host(pm, tc), pkg.DeviceModel.from_host(pm, simt)
    n = 400
    pt = pkg.DeviceParticles(mt, n).init(False)
    ps = pkg.DeviceParticles(ms, n).init(False)
    ref = om.initparticles(dbm, n, px.Rng(SEED))
    pkg.gibbssample_(pt, mt, 4)
    pkg.gibbssample_(ps, ms, 4)
    om.gibbssample_dbm(ref, dbm, px.Rng(SEED), 4)
    for a, b, c in zip(pt.to_host(), ps.to_host(), ref):
        assert np.mean(a != b) < 2e-3                   # two GPU engines: same draws up to accumulation-order flips
        assert np.mean(a != c) < 5e-3


def test_tc_meanfield_and_ais(pkg, tc):
    units = [300, 128, 64, 32]
    dbm = synth.random_dbm(units, 7, wscale=(0.002, 0.02))
    m = pkg.DeviceModel.from_host(to_pkg_model(pkg, dbm), tc)
    x = np.random.default_rng(8).poisson(1.0, (200, units[0])).astype(float)
    got, it, delta = pkg.meanfield(m, x, eps=1e-6, return_info=True, ctx=tc)
    ref, rit, _ = om.meanfield(dbm, x, eps=1e-6)
    assert abs(it - rit) <= 1
    for a, b in zip(got[1:], ref[1:]):
        assert relerr(a, b) < RTOL
    w = pkg.aislogimpweights(m, ntemperatures=20, nparticles=256, ctx=tc)
    wr = oais.aislogimpweights(dbm, px.Rng(SEED), ntemperatures=20, nparticles=256)
    assert np.mean(np.abs(w - wr) > 1e-3 * np.maximum(1, np.abs(wr))) < 0.05


def test_tc_large_counts_hi_plane_flags(pkg, tc, simt):
    """Counts >= 2048 are rare in scRNA data, so the up-sweep skips the hi count plane unless the row tile
    is flagged by the sampling epilogue.  Here a few genes have means in the thousands: the flags must be
    set and the hi plane used -- the chain must still match the CUDA-core engine and the oracle."""
    units = [320, 128, 64]
    dbm = synth.random_dbm(units, 5, wscale=(0.002, 0.01))
    big = np.array([3, 77, 200, 319])
    dbm[0].visbias[big] = np.log(np.array([3000.0, 5000.0, 8000.0, 4000.0]) / (np.array([3000.0, 5000.0, 8000.0, 4000.0]) + 0.5))
    dbm[0].weights[big] = -1e-6                          # keep these genes' means in the thousands for every h
    pm = to_pkg_model(pkg, dbm)
    mt, ms = pkg.DeviceModel.from_host(pm, tc), pkg.DeviceModel.from_host(pm, simt)
    n = 300
    pt = pkg.DeviceParticles(mt, n).init(False)
    ps = pkg.DeviceParticles(ms, n).init(False)
    ref = om.initparticles(dbm, n, px.Rng(SEED))
    pkg.gibbssample_(pt, mt, 4)
    pkg.gibbssample_(ps, ms, 4)
    om.gibbssample_dbm(ref, dbm, px.Rng(SEED), 4)
    got = pt.to_host()
    assert (got[0][:, big] >= 2048).mean() > 0.1         # the hi plane is really in use
    for a, b, c in zip(got, ps.to_host(), ref):
        assert np.mean(a != b) < 3e-3
        assert np.mean(a != c) < 6e-3
    # a second model without large counts on the same particles' buffers: flags are cleared again
    dbm2 = synth.random_dbm(units, 6, wscale=(0.002, 0.01))
    m2t, m2s = pkg.DeviceModel.from_host(to_pkg_model(pkg, dbm2), tc), pkg.DeviceModel.from_host(to_pkg_model(pkg, dbm2), simt)
    p2t = pkg.DeviceParticles(m2t, n).upload(got)
    p2s = pkg.DeviceParticles(m2s, n).upload(got)
    p2t.clock = p2s.clock = 9
    pkg.gibbssample_(p2t, m2t, 3)
    pkg.gibbssample_(p2s, m2s, 3)
    for a, b in zip(p2t.to_host(), p2s.to_host()):
        assert np.mean(a != b) < 3e-3
    # one more sweep lands in the other particle buffer: its hi plane (in use two sweeps ago) must have been
    # zero-filled lazily -- the Float64 download reads both planes regardless of the flags
    pkg.gibbssample_(p2t, m2t, 1)
    pkg.gibbssample_(p2s, m2s, 1)
    for a, b in zip(p2t.to_host(), p2s.to_host()):
        assert np.mean(a != b) < 3e-3
    assert p2t.to_host()[0].max() < 2048
    assert np.array_equal(p2t.layer(0).to_host_u16().astype(float), p2t.to_host()[0])


@pytest.mark.parametrize("V,H,npos,nneg", [(130, 67, 50, 37), (256, 128, 700, 64), (2000, 256, 1500, 100)])
def test_tc_gradient_gemm(pkg, tc, V, H, npos, nneg):
    """K5: dW = v'h/n+ - vm'hm/n- on the tensor cores (MN-major operands, split-K), vs the oracle."""
    g = np.random.default_rng(9)
    rbm = synth.random_dbm([V, H, 8], 4)[0]
    m = pkg.DeviceModel.from_host(to_pkg_model(pkg, rbm), tc)
    v = g.poisson(1.0, (npos, V)).astype(float)
    v[3, 5] = 3000.0                                     # hi count plane in one row tile only
    h = g.random((npos, H))
    for vm, hm in [(g.gamma(0.5, 2.0, (nneg, V)), g.random((nneg, H))),                                   # CD: real-valued model side
                   (g.poisson(2.0, (nneg, V)).astype(float), (g.random((nneg, H)) < 0.5).astype(float))]:  # PCD: particles
        opt = pkg.loglikelihoodoptimizer(m, learningrate=0.01)
        tc.launch_count(reset=True)
        mats = [pkg.Mat.from_host(tc, v, pkg.MAT_COUNT),
                pkg.Mat.from_host(tc, vm, pkg.MAT_COUNT if np.all(vm == np.floor(vm)) else pkg.MAT_REAL),
                pkg.Mat.from_host(tc, h, pkg.MAT_REAL),
                pkg.Mat.from_host(tc, hm, pkg.MAT_BINARY if np.isin(hm, (0.0, 1.0)).all() else pkg.MAT_REAL)]
        opt.computegradient_(*mats, m)
        assert tc.get("tc_launches") >= 1
        ref = om.computegradient(v, vm, h, hm)
        W, a, b = opt.gradient(0)
        assert np.max(np.abs(W - ref.weights)) < RTOL * np.abs(ref.weights).max()
        assert np.max(np.abs(a - ref.visbias)) < RTOL * np.abs(ref.visbias).max()
        assert np.max(np.abs(b - ref.hidbias)) < RTOL * np.abs(ref.hidbias).max()


def test_tc_traindbm_matches_oracle(pkg, tc):
    import copy
    units = [300, 128, 64]
    x, _, _ = synth.counts(400, units[0], seed=2)
    x[0, :] += 1.0
    dbm = synth.random_dbm(units, 4, wscale=(0.002, 0.02))
    m = pkg.DeviceModel.from_host(to_pkg_model(pkg, dbm), tc)
    ref = copy.deepcopy(dbm)
    om.traindbm(ref, x, px.Rng(SEED), epochs=2, nparticles=256, learningrate=0.01)
    pkg.traindbm_(m, x, epochs=2, nparticles=256, learningrate=0.01, ctx=tc)
    for l in range(2):
        got = m.get_layer(l)
        dref = ref[l].weights - dbm[l].weights
        assert np.max(np.abs((got.weights - dbm[l].weights) - dref)) < 0.05 * np.abs(dref).max() + 1e-9


def test_tc_bernoulli_visible_dbm(pkg, tc):
    """BernoulliRBM first layer (src/bmtypes.jl:17-21) through the tensor-core engine: Gibbs, mean-field, PCD."""
    import copy
    g = np.random.default_rng(2)
    units = [130, 67, 33]
    dbm = [om.BernoulliRBM(g.normal(0, 0.3, (units[i], units[i + 1])), g.normal(0, .3, units[i]), g.normal(0, .3, units[i + 1]))
           for i in range(2)]
    m = pkg.DeviceModel.from_host(to_pkg_model(pkg, dbm), tc)
    n = 200
    p = pkg.initparticles(m, n, device=True)
    ref = om.initparticles(dbm, n, px.Rng(SEED))
    for a, b in zip(p.to_host(), ref):
        assert np.array_equal(a, b)
    pkg.gibbssample_(p, m, 4)
    om.gibbssample_dbm(ref, dbm, px.Rng(SEED), 4)
    for a, b in zip(p.to_host(), ref):
        assert np.mean(a != b) < 5e-3
    x = (g.random((150, units[0])) < 0.3).astype(float)
    got = pkg.meanfield(m, x, eps=1e-6, ctx=tc)
    mf, _, _ = om.meanfield(dbm, x, eps=1e-6)
    for a, b in zip(got[1:], mf[1:]):
        assert relerr(a, b) < RTOL
    r2 = copy.deepcopy(dbm)
    om.traindbm(r2, x, px.Rng(SEED), epochs=2, nparticles=64, learningrate=0.05)
    pkg.traindbm_(m, x, epochs=2, nparticles=64, learningrate=0.05, ctx=tc)
    for l in range(2):
        dref = r2[l].weights - dbm[l].weights
        assert np.max(np.abs((m.get_layer(l).weights - dbm[l].weights) - dref)) < 0.05 * np.abs(dref).max() + 1e-9


def test_tc_pair_mode_large_m(pkg, tc, simt):
    """With >= 2*128*148 rows the deterministic / Bernoulli kernels process 256-row work tiles (two MMA row
    tiles per weight-tile load).  Ragged row count, count operand with a flagged hi plane in one tile only."""
    n = 2 * 128 * 148 + 77
    units = (130, 67, 33)
    dbm = synth.random_dbm(list(units), 3, wscale=(0.002, 0.02))
    g = np.random.default_rng(1)
    v = g.poisson(1.5, (n, units[0])).astype(float)
    v[n - 5, 7] = 3000.0
    h1 = (g.random((n, units[1])) < 0.5).astype(float)
    h2 = (g.random((n, units[2])) < 0.5).astype(float)
    m = pkg.DeviceModel.from_host(to_pkg_model(pkg, dbm), tc)
    assert relerr(pkg.hiddenpotential(m, v, 2.0, ctx=tc), om.hiddenpotential(dbm[0], v, 2.0)) < RTOL
    assert relerr(pkg.visiblepotential(m, h1, ctx=tc), om.visiblepotential(dbm[0], h1)) < RTOL
    ref = osp.sigm(om.dbm_layer_input(dbm, [v, h1, h2], 1))
    assert relerr(pkg.dbmlayer(m, 1, v, h2, ctx=tc), ref, 1e-7) < RTOL
    got = pkg.samplehidden(m, v, tick=5, ctx=tc)                      # Philox-fed Bernoulli kernel, pair mode
    assert np.mean(got != om.samplehidden(dbm[0], v, px.Rng(SEED), 1, 5)) < 5e-4
    got2 = pkg.dbmlayer(m, 1, v, h2, what=pkg.OUT_SAMPLE, tick=6, ctx=tc)
    ms = pkg.DeviceModel.from_host(to_pkg_model(pkg, dbm), simt)
    assert np.mean(got2 != pkg.dbmlayer(ms, 1, v, h2, what=pkg.OUT_SAMPLE, tick=6, ctx=simt)) < 5e-4


def test_full_size_generation_properties(pkg, tc):
    """BASELINE configuration C3 at (nearly) full size -- the oracle cannot follow here, so size-independent
    properties: (1) any shard of chains generated on its own (global row offset) equals the same rows of the
    big run bit for bit and equals the oracle on a slice it can afford; (2) the run is reproducible;
    (3) dense uint16, Float64 and row-compressed downloads describe the same matrix (checksums);
    (4) two disjoint halves of the chains agree in per-gene mean and zero fraction within sampling error."""
    import bench
    n, burnin, V = 400_000, 6, bench.UNITS[0]
    L = bench.make_model()
    host = [pkg.NegativeBinomialBernoulliRBM(**L[0]), pkg.BernoulliRBM(**L[1])]
    m = pkg.DeviceModel.from_host(host, tc)
    p = pkg.sampleparticles(m, n, burnin, row_offset=0, device=True)
    big = p.layer(0).to_host_u16()
    # (1) shards: last rows of the run (ragged start inside a 128-row tile) on the device and on the oracle
    lo, cnt = n - 1000 - 37, 1000
    shard = pkg.sampleparticles(m, cnt, burnin, row_offset=lo, device=True).layer(0).to_host_u16()
    assert np.array_equal(shard, big[lo:lo + cnt])
    dbm = [om.NegativeBinomialBernoulliRBM(L[0]["weights"], L[0]["visbias"], L[0]["hidbias"], L[0]["inversedispersion"]),
           om.BernoulliRBM(L[1]["weights"], L[1]["visbias"], L[1]["hidbias"])]
    rng = px.Rng(SEED)
    ref = om.initparticles(dbm, 64, rng, row_offset=lo)
    om.gibbssample_dbm(ref, dbm, rng, burnin)
    assert np.mean(ref[0] != big[lo:lo + 64]) < 6e-3
    # (2) reproducible
    p2 = pkg.sampleparticles(m, n, burnin, row_offset=0, device=True)
    s1, s2 = p.layer(0).column_stats(), p2.layer(0).column_stats()
    assert all(np.array_equal(a, b) for a, b in zip(s1, s2))
    # (3) one matrix, three wire formats
    ptr, idx, val = p.layer(0).to_host_csr()
    assert ptr[-1] == np.count_nonzero(big) and int(val.sum(dtype=np.int64)) == int(big.sum(dtype=np.int64))
    assert np.array_equal(np.diff(ptr), np.count_nonzero(big, axis=1))
    r = 123_456
    row = np.zeros(V, dtype=np.uint16)
    row[idx[ptr[r]:ptr[r + 1]]] = val[ptr[r]:ptr[r + 1]]
    assert np.array_equal(row, big[r])
    mean, var, zf = s1
    assert np.allclose(mean, big.mean(axis=0, dtype=np.float64), rtol=1e-9, atol=1e-12)
    assert np.allclose(zf, (big == 0).mean(axis=0), rtol=1e-9, atol=1e-12)
    # (4) halves agree: |mean_a - mean_b| <= 6 SE, zero fractions likewise (statistics reduced on the device)
    h = n // 2
    ma, va, za = pkg.sampleparticles(m, h, burnin, row_offset=0, device=True).layer(0).column_stats()
    mb, vb, zb = pkg.sampleparticles(m, h, burnin, row_offset=h, device=True).layer(0).column_stats()
    assert np.allclose(0.5 * (ma + mb), mean, rtol=1e-9, atol=1e-12)          # the halves are the big run's rows
    se = np.sqrt(va / h + vb / h) + 1e-9
    assert np.all(np.abs(ma - mb) <= 6 * se)
    sz = np.sqrt(za * (1 - za) / h + zb * (1 - zb) / h) + 1e-9
    assert np.all(np.abs(za - zb) <= 6 * sz)
    assert 0.7 < (big == 0).mean() < 0.95                  # 10x-like sparsity
