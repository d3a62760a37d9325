#!/usr/bin/env python
"""AIS (`aislogimpweights`, src/evaluating.jl:164-206) on the 3-layer scDBM of config C5
(20000 -> 1024 -> 256 -> 64): particles sharded over GPUs (the reference's `batchparallelized`,
src/misc.jl:38-47), temperatures sequential; the final logmeanexp needs max + sum-exp over ranks.

    python examples/ais_c5.py --particles 65536 --temperatures 100          (1 GPU)
    python -m torch.distributed.run --nproc-per-node N ... examples/ais_c5.py ...
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402
from bench import make_model  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--units", type=int, nargs="+", default=[20000, 1024, 256, 64])
    ap.add_argument("--particles", type=int, default=65536)
    ap.add_argument("--temperatures", type=int, default=100, help="number of transitions timed (slice of the 10,000 of C5)")
    ap.add_argument("--burnin", type=int, default=5)
    a = ap.parse_args()
    world, rank, local = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
    import torch
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    pkg = entry.load_package()
    layers = make_model(a.units)
    host = [pkg.NegativeBinomialBernoulliRBM(**layers[0])] + [pkg.BernoulliRBM(**l) for l in layers[1:]]
    ctx = pkg.Context(local, 20261019)
    dm = pkg.DeviceModel.from_host(host, ctx)
    shard = pkg.mostevenbatches(a.particles, world)
    off = sum(shard[:rank])
    # a slice of the default schedule collect(0:1e-4:1): the first `temperatures` transitions
    temps = np.arange(a.temperatures + 1) * 1e-4
    pkg.aislogimpweights(dm, temperatures=temps[:3], nparticles=shard[rank], burnin=a.burnin, row_offset=off, ctx=ctx)  # warm-up
    ctx.sync()
    t0 = time.perf_counter()
    w = pkg.aislogimpweights(dm, temperatures=temps, nparticles=shard[rank], burnin=a.burnin, row_offset=off, ctx=ctx)
    dt = time.perf_counter() - t0
    print(f"[rank {rank}] {shard[rank]} particles, {a.temperatures} transitions: {dt:.3f} s", file=sys.stderr, flush=True)
    # logmeanexp over all ranks (src/evaluating.jl:935-943): two scalar allreduces through the library's communicator
    if world > 1:
        pkg.init_comm_from_torch(ctx)
    lme = pkg.logmeanexp_sharded(w, ctx)
    if world > 1:
        td = torch.tensor([dt], dtype=torch.float64, device="cuda")
        dist.all_reduce(td, op=dist.ReduceOp.MAX)
        dt = float(td.item())
    if rank == 0:
        print(json.dumps({"n_gpus": world, "units": a.units, "particles": a.particles, "transitions": a.temperatures, "burnin": a.burnin,
                          "seconds": dt, "particle_temperatures_per_sec": a.particles * a.temperatures / dt,
                          "logmeanexp_slice": float(lme)}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
