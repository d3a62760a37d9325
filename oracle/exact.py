"""Exact ground truth by enumeration for tiny models (tests only).

Restates the *purpose* of src/evaluating.jl:546-565,633-671,1183-1211
(exact log partition function of small Bernoulli RBMs / DBMs by enumerating
states) and adds the closed-form NB-RBM mixture used by the free-running
statistics gate: given h, v_i ~ NB(r_i, 1-exp(x_i)),
and sum_v C(v+r-1, v) e^{xv} = (1-e^x)^{-r}.
"""
import itertools
import numpy as np


def _states(n):
    return np.array(list(itertools.product([0.0, 1.0], repeat=n)))


def exactlogpartitionfunction_rbm(rbm):
    """log Z of a Bernoulli RBM: sum over hidden states, visible summed out."""
    hs = _states(rbm.weights.shape[1])
    x = hs @ rbm.weights.T + rbm.visbias[None, :]
    logp = hs @ rbm.hidbias + np.sum(np.logaddexp(0.0, x), axis=1)
    m = logp.max()
    return m + np.log(np.exp(logp - m).sum())


def exactlogpartitionfunction_dbm(dbm):
    """log Z of a Bernoulli DBM by brute force over all layers' states."""
    sizes = [dbm[0].weights.shape[0]] + [r.weights.shape[1] for r in dbm]
    layers = [_states(n) for n in sizes]
    total = -np.inf
    # energy terms: sum_l s_l' W_l s_{l+1} + biases
    from .ais import combinedbiases
    biases = combinedbiases(dbm)
    for combo in itertools.product(*[range(len(l)) for l in layers]):
        s = [layers[i][c] for i, c in enumerate(combo)]
        e = sum(s[i] @ dbm[i].weights @ s[i + 1] for i in range(len(dbm)))
        e += sum(b @ si for b, si in zip(biases, s))
        total = np.logaddexp(total, e)
    return total


def nbrbm_hidden_distribution(rbm):
    """p(h) of an NB-Bernoulli RBM (visible summed out analytically)."""
    hs = _states(rbm.weights.shape[1])
    x = hs @ rbm.weights.T + rbm.visbias[None, :]
    logp = hs @ rbm.hidbias - np.sum(rbm.inversedispersion[None, :] * np.log1p(-np.exp(x)), axis=1)
    logz = np.logaddexp.reduce(logp)
    return hs, np.exp(logp - logz), logz


def nbrbm_visible_moments(rbm, eps=0.0):
    """Exact per-gene mean, variance and zero fraction of the NB-RBM's
    visible marginal (mixture over h of NB(r, p=1-e^x)); `eps` adds the
    reference's +1e-6 to the conditional mean if requested."""
    hs, ph, _ = nbrbm_hidden_distribution(rbm)
    r = rbm.inversedispersion[None, :]
    x = hs @ rbm.weights.T + rbm.visbias[None, :]
    ex = np.exp(x)
    mu = r * ex / (1.0 - ex) + eps
    p = r / (mu + r)
    var_c = mu + mu * mu / r
    mean = ph @ mu
    second = ph @ (var_c + mu * mu)
    zero = ph @ np.exp(r * np.log(p))
    return mean, second - mean ** 2, zero


def bernoulli_rbm_visible_marginal_mean(rbm):
    hs = _states(rbm.weights.shape[1])
    x = hs @ rbm.weights.T + rbm.visbias[None, :]
    logp = hs @ rbm.hidbias + np.sum(np.logaddexp(0.0, x), axis=1)
    ph = np.exp(logp - np.logaddexp.reduce(logp))
    return ph @ (1.0 / (1.0 + np.exp(-x)))
