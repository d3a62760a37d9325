"""world_size-2 `gloo` tests (CPU) of the N>1 host logic: chain sharding by
global row offset (no collective) and the data-parallel gradient reduction
(sum of un-normalised partial sums == single-rank gradient)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import model as om, philox as px, synth


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import __graft_entry__ as g
    pkg = g.load_package()
    dbm = synth.random_dbm([16, 8, 4], 3, wscale=(0.02, 0.2))
    # --- chain sharding: each rank generates its shard addressed by global index
    shards = pkg.mostevenbatches(21, world)
    off = sum(shards[:rank])
    mine = om.sampleparticles(dbm, shards[rank], px.Rng(1), burnin=3, row_offset=off)[0]
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    # --- data-parallel PCD gradient: rows of x and particles split by rank
    x, _, _ = synth.counts(20, 16, seed=2)
    mu, _, _ = om.meanfield(dbm, x)
    part = om.sampleparticles(dbm, 10, px.Rng(2), burnin=2)
    rx, rp = np.array_split(np.arange(20), world)[rank], np.array_split(np.arange(10), world)[rank]
    partial = mu[0][rx].T @ mu[1][rx] / 20.0 - part[0][rp].T @ part[1][rp] / 10.0
    t = torch.from_numpy(partial.copy())
    dist.all_reduce(t)
    # --- communicator bootstrap: rank 0 makes the 128-byte id, every rank ends up with the same bytes
    def bcast(obj):
        box = [obj]
        dist.broadcast_object_list(box, src=0)
        return box[0]
    uid = pkg.exchange_unique_id(lambda: bytes(range(128)), rank, bcast)
    # --- logmeanexp over particle shards == logmeanexp of the whole vector (src/evaluating.jl:935-943)
    w = np.random.default_rng(7).normal(0, 3, 21)

    class FakeCtx:   # the reductions scdbm_comm_allreduce_f64 performs, over gloo
        def allreduce(self, values, op="sum"):
            t2 = torch.tensor(np.atleast_1d(values).astype(np.float64))
            dist.all_reduce(t2, op=dist.ReduceOp.MAX if op == "max" else dist.ReduceOp.SUM)
            return t2.numpy()
    lme = pkg.logmeanexp_sharded(w[off:off + shards[rank]], FakeCtx())
    # --- global mean-field stopping rule: max over ranks of the per-shard delta == delta over all rows
    mu1, _, _ = om.meanfield(dbm, x, eps=0.0, maxiter=1)
    mu2, _, _ = om.meanfield(dbm, x, eps=0.0, maxiter=2)
    local_delta = max(float(np.abs(a[rx] - b[rx]).max()) for a, b in zip(mu1[1:], mu2[1:]))
    gd = FakeCtx().allreduce([local_delta], "max")[0]
    full_delta = max(float(np.abs(a - b).max()) for a, b in zip(mu1[1:], mu2[1:]))
    if rank == 0:
        q.put((np.vstack(gathered), t.numpy(), om.computegradient(mu[0], part[0], mu[1], part[1]).weights,
               om.sampleparticles(dbm, 21, px.Rng(1), burnin=3)[0], uid, lme, pkg.logmeanexp(w), gd, full_delta))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_and_allreduce():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    sharded, reduced, full_grad, whole, uid, lme, lme_ref, gd, full_delta = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert np.array_equal(sharded, whole)            # generation is invariant to the sharding
    assert np.allclose(reduced, full_grad, rtol=1e-12, atol=1e-15)
    assert uid == bytes(range(128))
    assert abs(lme - lme_ref) < 1e-12
    assert gd == full_delta
