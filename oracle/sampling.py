"""Elementwise draws of the hot path: Bernoulli and Negative-Binomial.

Reference sites
  * Bernoulli: `float(rand() < p)` -- strict `<` (src/gibbssampling.jl:40-45,
    fused form :852-857).
  * NB matrix form: p = r/(mu+r); v ~ Distributions.NegativeBinomial(r, p)
    (src/gibbssampling.jl:810-820).
  * NB vector / RBM-specific form: p = r/(mu+r) - 1e-6;
    v ~ Poisson(Gamma(shape r, scale (1-p)/p)) (src/gibbssampling.jl:91-99,
    796-808).  Both are the same law up to the 1e-6 shift (quirk B5);
    `pshift` selects it.

Distributions.jl is an un-vendored, unpinned dependency (Project.toml:7); its
`rand(NegativeBinomial(r,p))` is, in every release, an exact sampler of
NB(r,p) realised as Poisson(Gamma(r, (1-p)/p)).  The oracle therefore
restates the *law* with a published exact algorithm whose uniform consumption
is explicit, so that the CUDA kernel can be fed the very same uniforms:

  mu <= mu_inv_max : single-uniform sequential inversion ("chop-down";
                     Devroye 1986, III.2) from P0 = p**r with
                     P_{k+1} = P_k * (1-p) * (k+r)/(k+1);
  otherwise        : Gamma(r) by Marsaglia & Tsang (2000) (with the U**(1/r)
                     boost for r < 1) scaled by (1-p)/p, then Poisson(lam) by
                     inversion (lam < 10) or Hoermann's PTRS (1993).
"""
import numpy as np
from scipy.special import gammaln

from . import philox as px

KMAX = 65535          # counts saturate here (device count planes hold 16 bits)
MAX_ATTEMPTS = 64     # rejection-loop cap (probability of hitting it < 1e-30)
MU_INV_MAX = 8.0      # default switch between inversion and gamma-Poisson
PMIN = 2.0 ** -26     # inversion stops once the pmf term falls below this (the
                      # uniforms carry 23 bits; guards fp32 round-off run-away)


def sigm(x):
    """src/gibbssampling.jl:832-842"""
    return 1.0 / (1.0 + np.exp(-x))


def bernoulli(p, u):
    """src/gibbssampling.jl:40-45 with the uniform made explicit."""
    return (u < p).astype(np.float64)


def nb_params(mu, r, pshift=0.0):
    """p = r/(mu+r) [- 1e-6]  (src/gibbssampling.jl:814 / :91,800)."""
    p = r / (mu + r) - pshift
    return p, 1.0 - p


def _invert_pmf(u, p0, ratio_fn, kmax=KMAX):
    """Chop-down inversion: smallest k with CDF(k) > u, vectorised.
    ratio_fn(k) returns P_{k+1}/P_k for the still-active elements."""
    u = u.copy()
    P = p0.copy()
    k = np.zeros(u.shape, dtype=np.int64)
    active = (u >= P) & (P > PMIN)
    while active.any():
        idx = np.nonzero(active)
        u[idx] -= P[idx]
        P[idx] *= ratio_fn(k[idx], idx)
        k[idx] += 1
        active[idx] = (u[idx] >= P[idx]) & (k[idx] < kmax) & (P[idx] > PMIN)
    return k


def nb_inversion(mu, r, u, pshift=0.0):
    """Single-uniform NB(r,p) draw by sequential inversion."""
    mu, r, u = np.broadcast_arrays(np.asarray(mu, float), np.asarray(r, float), np.asarray(u, float))
    p, q = nb_params(mu, r, pshift)
    p0 = np.exp(r * np.log(p))
    return _invert_pmf(u, p0, lambda k, idx: q[idx] * (k + r[idx]) / (k + 1.0)).astype(np.float64)


def _normal_from_words(w0, w1):
    """Box-Muller (cosine branch) from two 24-bit uniforms."""
    u1, u2 = px.u24(w0), px.u24(w1)
    return np.sqrt(-2.0 * np.log(u1)) * np.cos(2.0 * np.pi * u2)


def gamma_mt(shape, rng, rows, cols, layer, tick):
    """Gamma(shape, 1) by Marsaglia-Tsang; uniforms come from Philox element
    addressing with purpose P_GAMMA and the attempt number in the stream."""
    a = np.asarray(shape, float).copy()
    n = a.shape[0]
    out = np.zeros(n)
    small = a < 1.0
    d = np.where(small, a + 1.0, a) - 1.0 / 3.0
    c = 1.0 / np.sqrt(9.0 * d)
    boost = np.ones(n)
    todo = np.ones(n, dtype=bool)
    for attempt in range(MAX_ATTEMPTS):
        idx = np.nonzero(todo)[0]
        if idx.size == 0:
            break
        w0, w1, w2, w3 = rng.element_words(rows[idx], cols[idx], px.stream_id(layer, px.P_GAMMA, attempt), tick)
        if attempt == 0:
            ub = px.u24(w3)
            boost[idx] = np.where(small[idx], ub ** (1.0 / a[idx]), 1.0)
        z = _normal_from_words(w0, w1)
        t = 1.0 + c[idx] * z
        v = t * t * t
        u2 = px.u24(w2)
        with np.errstate(invalid="ignore", divide="ignore"):
            ok = (v > 0.0) & (np.log(u2) < 0.5 * z * z + d[idx] - d[idx] * v + d[idx] * np.log(np.where(v > 0, v, 1.0)))
        acc = idx[ok]
        out[acc] = d[acc] * v[ok] * boost[acc]
        todo[acc] = False
    out[todo] = a[todo]   # unreachable in practice
    return out


def poisson_draw(lam, rng, rows, cols, layer, tick):
    """Poisson(lam): inversion for lam < 10, PTRS (Hoermann 1993) otherwise."""
    lam = np.asarray(lam, float)
    n = lam.shape[0]
    out = np.zeros(n)
    small = lam < 10.0
    # --- inversion branch: one uniform (word 0 of attempt 0)
    idx = np.nonzero(small)[0]
    if idx.size:
        w0, _, _, _ = rng.element_words(rows[idx], cols[idx], px.stream_id(layer, px.P_POIS, 0), tick)
        l = lam[idx]
        out[idx] = _invert_pmf(px.u24(w0), np.exp(-l), lambda k, sub: l[sub] / (k + 1.0))
    # --- PTRS branch
    todo = ~small
    l = lam
    with np.errstate(invalid="ignore", divide="ignore"):
        slam = np.sqrt(l)
        loglam = np.log(np.where(l > 0, l, 1.0))
        b = 0.931 + 2.53 * slam
        a = -0.059 + 0.02483 * b
        invalpha = 1.1239 + 1.1328 / (b - 3.4)
        vr = 0.9277 - 3.6224 / (b - 2.0)
    for attempt in range(MAX_ATTEMPTS):
        idx = np.nonzero(todo)[0]
        if idx.size == 0:
            break
        w0, w1, _, _ = rng.element_words(rows[idx], cols[idx], px.stream_id(layer, px.P_POIS, attempt), tick)
        U = px.u24(w0) - 0.5
        V = px.u24(w1)
        us = 0.5 - np.abs(U)
        k = np.floor((2.0 * a[idx] / us + b[idx]) * U + l[idx] + 0.43)
        fast = (us >= 0.07) & (V <= vr[idx])
        rej = (k < 0) | ((us < 0.013) & (V > us))
        with np.errstate(invalid="ignore", divide="ignore"):
            lhs = np.log(V) + np.log(invalpha[idx]) - np.log(a[idx] / (us * us) + b[idx])
            rhs = -l[idx] + k * loglam[idx] - gammaln(np.where(k >= 0, k, 0) + 1.0)
        ok = fast | (~rej & (lhs <= rhs))
        acc = idx[ok]
        out[acc] = np.minimum(k[ok], KMAX)
        todo[acc] = False
    out[todo] = np.floor(lam[todo])
    return out


def nb_draw(mu, r, rng, layer, tick, row_offset=0, pshift=0.0, mu_inv_max=MU_INV_MAX, u=None):
    """Draw an [n x V] matrix of NB(r_j, p_ij) counts.

    `mu` [n x V] conditional means, `r` [V] inverse dispersions.  Uniforms for
    the inversion branch come from `rng` grouped addressing (stream
    (layer, P_GIBBS)) or from the explicit array `u` (gate G2); the
    gamma-Poisson branch always uses `rng` element addressing.
    """
    mu = np.asarray(mu, float)
    n, V = mu.shape
    rr = np.broadcast_to(np.asarray(r, float)[None, :], mu.shape)
    if u is None:
        u = rng.uniforms(n, V, px.stream_id(layer, px.P_GIBBS), tick, row_offset)
    out = np.zeros_like(mu)
    inv = mu <= mu_inv_max
    if inv.any():
        out[inv] = nb_inversion(mu[inv], rr[inv], u[inv], pshift)
    if (~inv).any():
        ii, jj = np.nonzero(~inv)
        rows = ii.astype(np.uint64) + np.uint64(row_offset)
        cols = jj.astype(np.uint64)
        p, q = nb_params(mu[ii, jj], rr[ii, jj], pshift)
        g = gamma_mt(rr[ii, jj], rng, rows, cols, layer, tick)
        lam = g * (q / p)
        out[ii, jj] = poisson_draw(lam, rng, rows, cols, layer, tick)
    return np.minimum(out, KMAX)
