"""Host-side Philox4x32-10 for the one piece of integer randomness the host
owns: the minibatch permutation (`randperm`, src/rbmtraining.jl:426).  Same
generator and addressing as csrc/philox.cuh."""
import numpy as np

_M0, _M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_W0, _W1 = 0x9E3779B9, 0xBB67AE85
_MASK = np.uint64(0xFFFFFFFF)
P_PERM = 6


def _philox(c0, c1, c2, c3, k0, k1):
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint64) for c in np.broadcast_arrays(c0, c1, c2, c3))
    for _ in range(10):
        p0, p1 = _M0 * c0, _M1 * c2
        c0, c1, c2, c3 = ((p1 >> np.uint64(32)) ^ c1 ^ np.uint64(k0), p1 & _MASK,
                          (p0 >> np.uint64(32)) ^ c3 ^ np.uint64(k1), p0 & _MASK)
        k0, k1 = (k0 + _W0) & 0xFFFFFFFF, (k1 + _W1) & 0xFFFFFFFF
    return c0, c1, c2, c3


def permutation(n, seed, tick):
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    groups = np.arange((n + 3) // 4, dtype=np.uint64)
    stream = (0 & 0xFF) | (P_PERM << 8)
    w = np.stack(_philox(groups, np.uint64(0), np.uint64(int(tick) & 0xFFFFFFFF), np.uint64(stream),
                         seed & 0xFFFFFFFF, seed >> 32), axis=-1).reshape(-1)[:n]
    key = w * np.uint64(n) + np.arange(n, dtype=np.uint64)
    return np.argsort(key, kind="stable").astype(np.int32)
