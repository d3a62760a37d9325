#!/usr/bin/env python
"""Data-parallel DBM PCD (`traindbm!`, src/dbmtraining.jl:287-297) over N GPUs:
data rows and persistent particles are sharded by rank; per step one NCCL
allreduce(sum) of the contiguous gradient block {dW_l, da_l, db_l}, issued by the
library itself (scdbm_comm_init / scdbm_dbm_pcd_step_dp); torch.distributed only
carries the 128-byte communicator id and the final timing reduction.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29511 examples/dp_pcd.py --cells 20000 --units 2000 256 64 --steps 5 [--check]

--check: rank 0 also runs the same steps on one GPU over all rows / particles and
compares the parameters (fp32 reduction-order tolerance).
"""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402
from bench import make_model  # noqa: E402


def counts(n, V, seed):
    g = np.random.default_rng(seed)
    m = np.clip(np.exp(g.normal(-1.5, 1.5, V)), 1e-3, 50.0)
    s = np.exp(g.normal(0.0, 0.3, n))
    lam = g.gamma(0.5, s[:, None] * m[None, :] / 0.5)
    x = g.poisson(lam).astype(np.float64)
    x[0, :] += 1.0
    return np.minimum(x, 60000.0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cells", type=int, default=20000)
    ap.add_argument("--units", type=int, nargs="+", default=[2000, 256, 64])
    ap.add_argument("--particles", type=int, default=4096)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--check", action="store_true")
    a = ap.parse_args()
    world, rank, local = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    os.environ["NCCL_DEBUG"] = "WARN"
    dist.init_process_group("nccl", device_id=dev)
    pkg = entry.load_package()
    layers = make_model(a.units)
    host = [pkg.NegativeBinomialBernoulliRBM(**layers[0])] + [pkg.BernoulliRBM(**l) for l in layers[1:]]
    x = counts(a.cells, a.units[0], seed=4)
    rows = np.array_split(np.arange(a.cells), world)[rank]
    pshard = pkg.mostevenbatches(a.particles, world)
    poff = sum(pshard[:rank])

    ctx = pkg.Context(local, 20261019)
    dm = pkg.DeviceModel.from_host(host, ctx)
    xm = pkg.Mat.from_host(ctx, x[rows], pkg.MAT_COUNT)
    p = pkg.DeviceParticles(dm, pshard[rank], row_offset=poff).init(False)
    pkg.init_comm_from_torch(ctx)
    # one untimed step: NCCL sets up its channels on the first collective, descriptors and workspaces are created
    pkg.traindbm_(dm, xm, epochs=1, learningrate=1e-3, particles=p, global_nsamples=a.cells,
                  global_nparticles=a.particles, device=True, ctx=ctx)
    dist.barrier()
    torch.cuda.synchronize()
    ctx.timer_start()
    pkg.traindbm_(dm, xm, epochs=a.steps, learningrate=1e-3, particles=p, global_nsamples=a.cells,
                  global_nparticles=a.particles, device=True, ctx=ctx)
    ms = ctx.timer_stop()
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    res = {"n_gpus": world, "cells": a.cells, "units": a.units, "particles": a.particles, "steps": a.steps,
           "ms_per_step": float(t.item()) / a.steps, "pcd_steps_per_sec": a.steps / (float(t.item()) * 1e-3),
           "allreduce_bytes_per_step": 4 * sum(u * v + u + v for u, v in zip(a.units[:-1], a.units[1:]))}
    if a.check and rank == 0:
        ctx1 = pkg.Context(local, 20261019)
        dm1 = pkg.DeviceModel.from_host(host, ctx1)
        x1 = pkg.Mat.from_host(ctx1, x, pkg.MAT_COUNT)
        p1 = pkg.DeviceParticles(dm1, a.particles).init(False)
        pkg.traindbm_(dm1, x1, epochs=a.steps + 1, learningrate=1e-3, particles=p1, device=True, ctx=ctx1)
        errs = []
        for l in range(len(host)):
            g, r = dm.get_layer(l), dm1.get_layer(l)
            d0 = np.abs(r.weights - host[l].weights).max()
            errs.append(float(np.abs(g.weights - r.weights).max() / max(d0, 1e-30)))
        res["max_update_mismatch_vs_1gpu"] = errs
    if rank == 0:
        print(json.dumps(res), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
