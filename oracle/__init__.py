"""CPU oracle for the scDBM.jl Gibbs / CD / PCD / mean-field / AIS hot path.

TEST INFRASTRUCTURE ONLY.  This package is a NumPy Float64 restatement of the
reference's Julia code (file:line cited per function, paths relative to the
reference checkout).  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s CPU-baseline legs may import it; the product path
(``scdbm.jl_b200``) never does and fails loudly without its CUDA library.

PARITY UNPINNED: the reference's own test-suite never exercises the
Negative-Binomial path (it imports upstream BoltzmannMachines, not scDBM) and
holds no golden vectors; Julia is not installed here, so the reference cannot
be executed either.  The oracle's authority is (1) line-by-line fidelity to
the cited reference lines, (2) the reference's *relational* tests restated on
tiny models (exact enumeration vs AIS, exact marginals vs Gibbs frequencies,
fitdbm == stackrbms + traindbm under one seed), (3) analytic NB identities and
third-party known answers (Random123 Philox KAT, scipy.stats.nbinom).

Randomness: the reference draws from Julia's global RNG and Distributions.jl
(unpinned versions).  Neither stream can be reproduced, so every stochastic
function here consumes *counter-based Philox4x32-10 uniforms* addressed by
(seed, stream, clock tick, row, column) -- the same addressing the CUDA
kernels use -- or an explicit uniform array supplied by the caller.
"""
from . import philox, sampling, model, ais, exact, synth  # noqa: F401
