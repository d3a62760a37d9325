"""Synthetic "10x-like" count data and random valid model parameters
(SURVEY.md section 8d).  Uses numpy's default_rng so that bench and tests
generate the same matrices on any machine."""
import numpy as np

from . import model as md


def counts(n, V, seed, r=0.5):
    """x_cg ~ NB(r_g, p = r_g/(s_c m_g + r_g)); m_g ~ LogNormal(-1.5, 1.5)
    clipped to [1e-3, 50]; s_c ~ LogNormal(0, 0.3)."""
    g = np.random.default_rng(seed)
    m = np.clip(np.exp(g.normal(-1.5, 1.5, V)), 1e-3, 50.0)
    s = np.exp(g.normal(0.0, 0.3, n))
    rr = np.full(V, r) if np.isscalar(r) else np.asarray(r, float)
    mu = s[:, None] * m[None, :]
    lam = g.gamma(rr[None, :], mu / rr[None, :])
    return g.poisson(lam).astype(np.float64), m, rr


def random_dbm(units, seed, r=0.5, wscale=(5e-4, 2e-3)):
    """Random valid parameters of an NB-visible DBM with layer sizes `units`
    (W1 = -U(5e-4, 2e-3), a = log(m/(m+r)), upper W ~ N(0, 1/fan_in), b = 0)."""
    g = np.random.default_rng(seed)
    V = units[0]
    m = np.clip(np.exp(g.normal(-1.5, 1.5, V)), 1e-3, 50.0)
    rr = np.full(V, r)
    dbm = [md.NegativeBinomialBernoulliRBM(-g.uniform(wscale[0], wscale[1], (V, units[1])),
                                           np.log(m / (m + rr)), np.zeros(units[1]), rr)]
    for i in range(1, len(units) - 1):
        dbm.append(md.BernoulliRBM(g.normal(0, 1.0 / np.sqrt(units[i]), (units[i], units[i + 1])),
                                   g.normal(0, 0.1, units[i]), g.normal(0, 0.1, units[i + 1])))
    return dbm
