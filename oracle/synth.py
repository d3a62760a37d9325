"""Synthetic "10x-like" count data and random valid model parameters
(SURVEY.md section 8d).  Uses numpy's default_rng so that bench and tests
generate the same matrices on any machine."""
import numpy as np

from . import model as md
from . import philox as px
from . import sampling as sp


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


def philox_counts(n, gene_mean, r, sigma, rng, row_offset=0, mu_inv_max=sp.MU_INV_MAX):
    """Twin of the device generator `scdbm_mat_generate_counts` (SURVEY 8d data, Philox-addressed so any shard
    of cells can be generated where it is consumed): size factor s_c = exp(sigma z_c) with the Box-Muller normal
    z_c = sqrt(-2 ln u1) cos(2 pi u2) from the words of element (col 0, row c) on stream (0, P_SIZE), tick 0;
    x_cg = nb_draw(s_c m_g, r_g) with the grouped uniforms of stream (0, P_DATA), tick 0 (gamma-Poisson branch
    above `mu_inv_max` on the element streams of layer 0, tick 0).  Returns (x, s)."""
    gene_mean, r = np.asarray(gene_mean, float), np.asarray(r, float)
    V = gene_mean.shape[0]
    rows = np.arange(n, dtype=np.uint64) + np.uint64(row_offset)
    w = rng.element_words(rows, np.zeros(n, dtype=np.uint64), px.stream_id(0, px.P_SIZE), 0)
    z = np.sqrt(-2.0 * np.log(px.u24(w[0]))) * np.cos(2.0 * np.pi * px.u24(w[1]))
    s = np.exp(sigma * z)
    u = rng.uniforms(n, V, px.stream_id(0, px.P_DATA), 0, row_offset)
    x = sp.nb_draw(s[:, None] * gene_mean[None, :], r, rng, 0, 0, row_offset, mu_inv_max=mu_inv_max, u=u)
    return x, s


def to_csr(x):
    """Dense (samples x units) -> (indptr, indices, values), columns ascending within a row."""
    x = np.asarray(x)
    nz = x != 0
    indptr = np.concatenate([[0], np.cumsum(nz.sum(axis=1))]).astype(np.int64)
    rows, cols = np.nonzero(nz)
    return indptr, cols.astype(np.int32), x[rows, cols]


def from_compressed(shape, ptr, idx, val, major="csr"):
    out = np.zeros(shape)
    for s in range(len(ptr) - 1):
        seg = slice(ptr[s], ptr[s + 1])
        if major == "csr":
            out[s, idx[seg]] = val[seg]
        else:
            out[idx[seg], s] = val[seg]
    return out
