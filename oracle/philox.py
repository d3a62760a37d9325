"""Philox4x32-10 counter-based RNG (Salmon et al., SC'11; Random123 v1.14
`philox.h`) restated in NumPy, plus the element->counter addressing shared
with the CUDA kernels (scdbm.jl_b200/csrc/philox.cuh).

The reference uses Julia's global RNG (`rand()` at src/gibbssampling.jl:35,
42,49,854); that stream depends on the Julia version and cannot be
reproduced, so oracle and GPU agree on this stateless generator instead.

Addressing of the uniform used for element (row i, column j) of a state
matrix:  counter = (j >> 2, i, tick, stream), key = (seed_lo, seed_hi),
word = output[j & 3],  u = ((word >> 9) + 0.5) * 2**-23   (23 random bits;
24 significant bits, hence exact in fp32; 2**-24 <= u <= 1 - 2**-24).
"""
import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = np.uint32(0x9E3779B9)
W1 = np.uint32(0xBB67AE85)
MASK32 = np.uint64(0xFFFFFFFF)

# stream purposes (bits 8..15 of the stream word); bits 0..7 = layer index,
# bits 16..31 = rejection-sampler attempt number
P_GIBBS = 0      # conditional draw of a layer inside a Gibbs sweep
P_INIT = 1       # particle initialisation, first draw
P_INIT2 = 2      # particle initialisation, second draw (NB visible init)
P_GAMMA = 3      # Marsaglia-Tsang gamma attempts (NB rejection branch)
P_POIS = 4       # Poisson stage of the NB rejection branch
P_DROPOUT = 5    # zero-inflation mask
P_PERM = 6       # minibatch permutation
P_DATA = 7       # synthetic count data: NB draw
P_SIZE = 8       # synthetic count data: per-cell size factor


def stream_id(layer, purpose=P_GIBBS, attempt=0):
    return (int(layer) & 0xFF) | ((int(purpose) & 0xFF) << 8) | ((int(attempt) & 0xFFFF) << 16)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10.  All arguments broadcastable uint32 arrays.
    Returns four uint32 arrays."""
    c0 = np.asarray(c0, dtype=np.uint64)
    c1 = np.asarray(c1, dtype=np.uint64)
    c2 = np.asarray(c2, dtype=np.uint64)
    c3 = np.asarray(c3, dtype=np.uint64)
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    k0 = np.uint64(int(k0) & 0xFFFFFFFF)
    k1 = np.uint64(int(k1) & 0xFFFFFFFF)
    for _ in range(10):
        p0 = M0 * c0            # < 2**64, no overflow
        p1 = M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK32
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK32
        c0, c1, c2, c3 = (hi1 ^ c1 ^ k0), lo1, (hi0 ^ c3 ^ k1), lo0
        k0 = (k0 + np.uint64(W0)) & MASK32
        k1 = (k1 + np.uint64(W1)) & MASK32
    return (c0.astype(np.uint32), c1.astype(np.uint32),
            c2.astype(np.uint32), c3.astype(np.uint32))


def u24(word):
    """uint32 word -> uniform in (0,1): 23 random bits plus a half, i.e. 24
    significant bits, exactly representable in fp32, so GPU and oracle
    compare the *same* number."""
    return ((np.asarray(word, dtype=np.uint32) >> np.uint32(9)).astype(np.float64) + 0.5) * (2.0 ** -23)


class Rng:
    """Stateless uniform source: (seed) -> uniforms addressed by
    (stream, tick, row, col)."""

    def __init__(self, seed=0):
        seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        self.seed = seed
        self.k0 = seed & 0xFFFFFFFF
        self.k1 = (seed >> 32) & 0xFFFFFFFF

    def words(self, nrows, ncols, stream, tick, row_offset=0):
        """uint32 words for an [nrows x ncols] state matrix (grouped
        addressing: one Philox call per 4 consecutive columns)."""
        ngroups = (ncols + 3) // 4
        rows = (np.arange(nrows, dtype=np.uint64) + np.uint64(row_offset))[:, None]
        grp = np.arange(ngroups, dtype=np.uint64)[None, :]
        r = philox4x32_10(grp, rows, np.uint64(int(tick) & 0xFFFFFFFF),
                          np.uint64(int(stream) & 0xFFFFFFFF), self.k0, self.k1)
        out = np.stack(r, axis=-1).reshape(nrows, ngroups * 4)
        return out[:, :ncols]

    def uniforms(self, nrows, ncols, stream, tick, row_offset=0):
        return u24(self.words(nrows, ncols, stream, tick, row_offset))

    def element_words(self, rows, cols, stream, tick):
        """Per-element addressing (counter = (col, row, tick, stream)) used by
        the rejection samplers: four words for each listed element."""
        rows = np.asarray(rows, dtype=np.uint64)
        cols = np.asarray(cols, dtype=np.uint64)
        return philox4x32_10(cols, rows, np.uint64(int(tick) & 0xFFFFFFFF),
                             np.uint64(int(stream) & 0xFFFFFFFF), self.k0, self.k1)

    def permutation(self, n, tick):
        """Random permutation of 0..n-1 (stands in for `randperm`,
        src/rbmtraining.jl:426): argsort of per-index Philox words."""
        w = self.words(1, n, stream_id(0, P_PERM), tick)[0]
        key = w.astype(np.uint64) * np.uint64(n) + np.arange(n, dtype=np.uint64)
        return np.argsort(key, kind="stable")
