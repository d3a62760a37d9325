"""Generates tests/golden/*.npz from the oracle with fixed seeds.

The reference (Julia) cannot run in this image and its own tests hold no
golden vectors for the NB path, so these fixtures pin the
*oracle's* behaviour: tests/test_oracle.py re-derives them (regression guard
for the checker) and tests/test_gpu_golden.py checks the CUDA path against
them.  Third-party anchors: the Random123 Philox known-answer vectors and
scipy.stats.nbinom quantiles stored alongside.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np
from scipy import stats

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import model as om, philox as px, sampling as osp, ais as oais, exact, synth  # noqa: E402

SEED = 424242


def main():
    rng = px.Rng(SEED)
    units = [48, 20, 12]
    dbm = synth.random_dbm(units, 17, wscale=(0.02, 0.2))
    g = np.random.default_rng(5)
    n = 40
    v = g.poisson(1.2, (n, units[0])).astype(float)
    h1 = (g.random((n, units[1])) < 0.5).astype(float)
    h2 = (g.random((n, units[2])) < 0.5).astype(float)
    out = dict(
        seed=SEED, units=np.array(units),
        W0=dbm[0].weights, a0=dbm[0].visbias, b0=dbm[0].hidbias, r0=dbm[0].inversedispersion,
        W1=dbm[1].weights, a1=dbm[1].visbias, b1=dbm[1].hidbias,
        v=v, h1=h1, h2=h2,
        hiddeninput=om.hiddeninput(dbm[0], v),
        hiddenpotential2=om.hiddenpotential(dbm[0], v, 2.0),
        visiblepotential=om.visiblepotential(dbm[0], h1),
        dbmlayer1=osp.sigm(om.dbm_layer_input(dbm, [v, h1, h2], 1)),
        samplehidden_t5=om.samplehidden(dbm[0], v, rng, 1, 5),
        samplevisible_t6=om.samplevisible(dbm[0], h1, rng, 0, 6),
    )
    p = om.initparticles(dbm, 32, rng)
    out.update(init0=p[0], init1=p[1], init2=p[2])
    om.gibbssample_dbm(p, dbm, rng, 3)
    out.update(gibbs0=p[0], gibbs1=p[1], gibbs2=p[2])
    mu, it, delta = om.meanfield(dbm, v, eps=1e-6)
    out.update(mf1=mu[1], mf2=mu[2], mf_iters=it)
    grad = om.computegradient(v, np.asarray(p[0][:20]), mu[1], np.asarray(p[1][:20]))
    out.update(gradW=grad.weights, grada=grad.visbias, gradb=grad.hidbias)
    out.update(aisw=oais.aislogimpweights(dbm, rng, ntemperatures=10, nparticles=16))
    # third-party anchors
    out.update(philox_kat_in=np.array([[0, 0, 0, 0, 0, 0], [0xffffffff] * 6,
                                       [0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344, 0xa4093822, 0x299f31d0]], dtype=np.uint64),
               philox_kat_out=np.array([[0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8],
                                        [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd],
                                        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]], dtype=np.uint64))
    uq = (np.floor(g.random((64, 8)) * 2 ** 23) + 0.5) / 2 ** 23
    muq = np.exp(g.normal(-0.5, 1.2, (64, 8))).clip(1e-3, 8.0)
    rq = g.uniform(0.2, 3.0, 8)
    out.update(nb_u=uq, nb_mu=muq, nb_r=rq, nb_ppf=stats.nbinom.ppf(uq, rq[None, :], rq[None, :] / (muq + rq[None, :])))
    np.savez_compressed(os.path.join(HERE, "dbm_48_20_12.npz"), **out)
    print("wrote", os.path.join(HERE, "dbm_48_20_12.npz"))


if __name__ == "__main__":
    main()
