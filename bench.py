#!/usr/bin/env python
"""bench.py -- Gibbs-generation throughput of the 2-layer scDBM (BASELINE.json
config "Generation of 1M synthetic cells by Gibbs sampling from the 2-layer
scDBM"): metric = Gibbs sweeps x chains / sec.

  python bench.py --gpus N --steps K --warmup W            # CUDA path (this repo)
  python bench.py --impl reference --gpus N --steps K ...  # CPU restatement of the reference

One "step" = one `samples(dbm, chains; burnin)` generation pass: particle
initialisation + `burnin` synchronous Gibbs sweeps of all layers over
`chains` chains (src/gibbssampling.jl:646-650,153-182).  Chains are sharded
over GPUs with no collective (weak scaling: `--chains` per GPU).

The reference is Julia and cannot run here; `--impl reference` times the
NumPy/OpenBLAS restatement (oracle/) on the host cores, one worker process per
core over disjoint chain shards (the reference's own `batchparallelized`
pattern, src/misc.jl:38-47).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

UNITS = [2000, 256, 64]
C4_UNITS = [20000, 1024, 256, 64]
MODEL_SEED = 3
SEED = 20261019


def make_model(units=UNITS, seed=MODEL_SEED, r=0.5):
    """Random valid parameters of the C3 model: W1 = -U(5e-4, 2e-3),
    a = log(m/(m+r)) with 10x-like gene means, upper W ~ N(0, 1/fan_in)."""
    g = np.random.default_rng(seed)
    V = units[0]
    m = np.clip(np.exp(g.normal(-1.5, 1.5, V)), 1e-3, 50.0)
    rr = np.full(V, r)
    layers = [dict(weights=-g.uniform(5e-4, 2e-3, (V, units[1])), visbias=np.log(m / (m + rr)),
                   hidbias=np.zeros(units[1]), inversedispersion=rr)]
    for i in range(1, len(units) - 1):
        layers.append(dict(weights=g.normal(0, 1.0 / np.sqrt(units[i]), (units[i], units[i + 1])),
                           visbias=g.normal(0, 0.1, units[i]), hidbias=g.normal(0, 0.1, units[i + 1])))
    return layers


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"tflops_burst": p["bf16_tflops"], "tflops_sustained": p["bf16_tflops_sustained"],
                "hbm_gbs": p["hbm_gbs"], "source": "measured (MEASURED_PEAKS.json)"}
    return {"tflops_burst": 1590.0, "tflops_sustained": 1400.0, "hbm_gbs": 6650.0, "source": "fallback (B200_PROFILING.md)"}


def kernel_source_hash():
    """sha256 over the CUDA sources the library is built from: ties a committed ncu capture to the code it profiled."""
    import hashlib
    h = hashlib.sha256()
    d = os.path.join(ROOT, "scdbm.jl_b200", "csrc")
    for f in sorted(os.listdir(d)):
        h.update(f.encode())
        h.update(open(os.path.join(d, f), "rb").read())
    return h.hexdigest()[:16]


def static_profile():
    """ncu counters are NOT measured by this run: they come from a committed capture (profiles/r02_ncu_static.json,
    written by profiles/ncu_to_static.py) and are reported only while the kernel sources are byte-identical to the
    ones that were profiled."""
    path = os.path.join(ROOT, "profiles", "r02_ncu_static.json")
    if not os.path.exists(path):
        return None
    sp = json.load(open(path))
    if sp.get("kernel_source_sha") != kernel_source_hash():
        return {"stale": True, "capture": sp.get("capture"), "note": "kernel sources changed since the capture; counters withheld"}
    return sp


def sampler_max_mhz():
    out = subprocess.run(["nvidia-smi", "--query-gpu=clocks.max.sm", "--format=csv,noheader,nounits", "-i", "0"],
                         capture_output=True, text=True, timeout=10).stdout.strip().splitlines()
    return float(out[0]) if out else 0.0


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        super().__init__(daemon=True)
        self.gpu, self.rows, self.stop_flag = gpu, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.gpu)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(float(r[1]) for r in self.rows)
        reasons = set()
        for r in self.rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.rows[0][2]), "reasons": sorted(reasons),
                "samples": len(self.rows)}


# ------------------------------------------------------------------ CUDA arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    import __graft_entry__ as entry
    pkg = entry.load_package()

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run")
    torch.cuda.set_device(local)
    if world > 1:
        # NCCL writes its version banner to stdout at communicator creation; this script must print ONE JSON
        # line, so fd 1 points at stderr until the first collective has run
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)

    n, burnin = args.chains, args.burnin
    layers = make_model()
    ctx = pkg.Context(local, SEED)
    host_model = [pkg.NegativeBinomialBernoulliRBM(**layers[0]), pkg.BernoulliRBM(**layers[1])]
    dm = pkg.DeviceModel.from_host(host_model, ctx)
    row_offset = rank * n

    def barrier():
        ctx.sync()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def step(p):
        p.init(False)
        pkg.gibbssample_(p, dm, burnin)

    p = pkg.DeviceParticles(dm, n, row_offset)
    for _ in range(args.warmup):
        step(p)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    ctx.launch_count(reset=True)
    ctx.timer_start()
    for _ in range(args.steps):
        step(p)
    ms = ctx.timer_stop()
    launches = ctx.launch_count()
    tc_launches = int(ctx.get("tc_launches"))
    barrier()
    sampler.stop_flag = True
    sampler.join(timeout=10)
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = world * n * burnin * args.steps / (ms * 1e-3)

    # ---- roofline of the dominant kernel: the fused NB down-sweep (v | h1), timed alone
    h1, vout = p.layer(1), pkg.Mat(ctx, n, UNITS[0], pkg.MAT_COUNT)
    lib = pkg.lib()
    reps = 5
    for _ in range(2):
        pkg._lib.check(lib.scdbm_visible(dm.h, 0, h1.h, vout.h, 1.0, pkg.OUT_SAMPLE, 12345, None, 0))
    ctx.timer_start()
    for i in range(reps):
        pkg._lib.check(lib.scdbm_visible(dm.h, 0, h1.h, vout.h, 1.0, pkg.OUT_SAMPLE, 1000 + i, None, 0))
    k_ms = ctx.timer_stop() / reps
    pk = peaks()
    flops = 2.0 * n * UNITS[1] * UNITS[0]
    achieved = flops / (k_ms * 1e-3) / 1e12
    # algorithmic HBM bytes of one launch: read h1 (fp16), write the lo count plane (fp16; the hi plane is written
    # lazily, flag-driven, and stays untouched at scRNA-seq counts) -- computed for THIS run, not measured
    nb_bytes = float(n) * (UNITS[1] * 2 + UNITS[0] * 2)
    roofline = {"kernel": "fused NB down-sweep GEMM (v|h1: tcgen05 GEMM + NB-mean + Philox NB draw), k_gemm_tc<224,1,28>",
                "bound": "tensor", "achieved": achieved, "peak": pk["tflops_burst"], "unit": "TFLOP/s",
                "frac": achieved / pk["tflops_burst"],
                "traffic": None, "algorithmic_bytes_per_launch": nb_bytes,
                "algorithmic_gbs": nb_bytes / (k_ms * 1e-3) / 1e9, "hbm_frac_algorithmic": nb_bytes / (k_ms * 1e-3) / 1e9 / pk["hbm_gbs"],
                "traffic_note": "`traffic` (dram bytes from ncu) is not measured by this run: see static_profile",
                "peak_source": pk["source"] + ", burst (kernel timed alone)",
                "launch_ms": k_ms, "flops_per_launch": flops, "nb_draws_per_sec": n * UNITS[0] / (k_ms * 1e-3),
                "note": "sampled sweeps are bound by the draw epilogue (issue slots / dependency latency), not by the tensor pipe; see DESIGN.md"}
    sp = static_profile()
    if sp is not None:
        roofline["static_profile"] = sp
        if not sp.get("stale") and n == sp.get("chains"):
            nbp = sp.get("nb_downsweep", {})
            roofline["traffic"] = nbp.get("dram_bytes_per_launch")
            # what actually bounds the draw epilogue: warp-instruction issue.  Instruction count from the ncu capture of
            # these sources, time from THIS run; peak = 4 schedulers x SMs x the SM clock nvidia-smi reports as maximum.
            wi, num_sms, mhz = nbp.get("warp_instructions"), int(ctx.get("num_sms")), 1965.0
            try:
                mhz = float(sampler_max_mhz()) or mhz
            except Exception:
                pass
            if wi:
                peak_ips = 4.0 * num_sms * mhz * 1e6
                roofline["issue_slots"] = {"warp_instructions_per_launch": wi, "achieved_gwips": wi / (k_ms * 1e-3) / 1e9,
                                           "peak_gwips": peak_ips / 1e9, "frac": wi / (k_ms * 1e-3) / peak_ips,
                                           "thread_instructions_per_draw": nbp.get("warp_instructions_per_draw_x32"),
                                           "note": "instruction count (polls of waiting warps included) from the static ncu capture, "
                                                   "launch time live; 7 warps per scheduler, 64 registers per thread"}
    # the deterministic part of the sweep for comparison: the dual-source up-sweep (h1 | v, h2), tensor-bound
    h1out = pkg.Mat(ctx, n, UNITS[1], pkg.MAT_BINARY)
    v0, h2 = p.layer(0), p.layer(2)
    for _ in range(2):
        pkg._lib.check(lib.scdbm_dbm_layer(dm.h, 1, v0.h, h2.h, h1out.h, pkg.OUT_SAMPLE, 777, None, 0))
    ctx.timer_start()
    for i in range(reps):
        pkg._lib.check(lib.scdbm_dbm_layer(dm.h, 1, v0.h, h2.h, h1out.h, pkg.OUT_SAMPLE, 2000 + i, None, 0))
    u_ms = ctx.timer_stop() / reps
    uflops = 2.0 * n * UNITS[1] * (UNITS[0] + UNITS[2])
    roofline["upsweep"] = {"kernel": "dual-source up-sweep GEMM (h1|v,h2: count x W hi/lo = 2 plane products + Bernoulli draw), k_gemm_cg2<2> (cta_group::2, 256x256 cluster tiles)",
                           "bound": "tensor", "launch_ms": u_ms, "achieved": uflops / (u_ms * 1e-3) / 1e12, "peak": pk["tflops_burst"],
                           "unit": "TFLOP/s", "frac": uflops / (u_ms * 1e-3) / 1e12 / pk["tflops_burst"],
                           "issued_tflops": 2.0 * uflops / (u_ms * 1e-3) / 1e12, "issued_frac": 2.0 * uflops / (u_ms * 1e-3) / 1e12 / pk["tflops_burst"],
                           "note": "achieved/frac count algorithmic FLOPs (2MNK); issued_* count the 2 fp16 plane products (W hi + W lo) the tensor pipe executes"}
    del vout, h1out

    # ---- end to end through the public API: host model in, host uint16 cells out
    e2e = None
    if rank == 0 or world > 1:
        # two pinned result buffers: batch i travels to the host (copy engine) while batch i+1 is sampled
        outs = [torch.empty((n, UNITS[0]), dtype=torch.uint16, pin_memory=True).numpy() for _ in range(2)]
        h2d = sum(a.nbytes for l in layers for a in l.values())
        d2h = outs[0].nbytes

        def e2e_step(i):
            m = pkg.DeviceModel.from_host(host_model, ctx)          # H2D: parameters
            q = pkg.sampleparticles(m, n, burnin, row_offset, device=True)
            q.layer(0).to_host_u16(outs[i % 2], wait=False)           # D2H: generated cells (all of them, every step)
            q.close()
            m.close()

        e2e_step(0)
        ctx.wait_downloads()
        barrier()
        t0 = time.perf_counter()
        ksteps = max(1, args.steps)
        for i in range(ksteps):
            e2e_step(i)
        ctx.wait_downloads()                                          # the last batch's copy is inside the timed region
        barrier()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e2e = {"value": world * n * burnin * ksteps / float(dt.item()), "unit": "sweeps*chains/s",
               "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "steps": ksteps,
               "overlap": "the device->host copy of batch i (copy engine, own stream) overlaps the sampling of batch i+1; "
                          "two pinned result buffers; the last copy completes inside the timed region",
               "cells_per_sec": world * n * ksteps / float(dt.item())}

        del outs
    if e2e is not None and world == 1:
        # the same step with the row-compressed wire format (uint16 column indices + uint16 counts, ~20 % dense)
        cap = int(0.3 * n * UNITS[0])
        ib = torch.empty(cap, dtype=torch.uint16, pin_memory=True).numpy()
        vb = torch.empty(cap, dtype=torch.uint16, pin_memory=True).numpy()
        csr_bytes = [0]

        def csr_step():
            m = pkg.DeviceModel.from_host(host_model, ctx)
            q = pkg.sampleparticles(m, n, burnin, row_offset, device=True)
            ptr, idx, val = q.layer(0).to_host_csr(buffers=(ib, vb))
            csr_bytes[0] = ptr.nbytes + idx.nbytes + val.nbytes
            q.close()
            m.close()

        csr_step()
        barrier()
        t0 = time.perf_counter()
        csteps = max(1, min(args.steps, 2))
        for _ in range(csteps):
            csr_step()
        barrier()
        dtc = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(dtc, op=dist.ReduceOp.MAX)
        e2e["csr"] = {"value": world * n * burnin * csteps / float(dtc.item()), "unit": "sweeps*chains/s", "steps": csteps,
                      "d2h_bytes_per_step": int(csr_bytes[0]),
                      "note": "scdbm_mat_csr_prepare + scdbm_mat_download_csr into pinned buffers, synchronous (not overlapped)"}
        del ib, vb

    del p, dm, h1, v0, h2
    extra = {}
    if not args.no_secondary:
        comm_ready = False
        if world > 1:
            pkg.init_comm_from_torch(ctx)             # the library's own NCCL communicator (id carried by torch.distributed)
            comm_ready = True
        extra["strong_c3"] = strong_c3(pkg, ctx, torch, dist, world, rank, args, host_model, barrier)
        extra["dp_pcd_c4"] = dp_pcd_c4(pkg, ctx, torch, dist, world, rank, args, barrier, pk)
        extra["ais_c5"] = ais_c5(pkg, ctx, torch, dist, world, rank, args, barrier, pk)
        if comm_ready:
            ctx.comm_destroy()

    line = {
        "metric": "Gibbs sweeps x chains / sec (synthetic-cell generation, 2-layer scDBM)",
        "value": value, "unit": "sweeps*chains/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "fp16x2-split operands (value = hi + lo plane), fp32 accumulate/epilogue", "data": "synthetic (random valid parameters, seed 3)",
        "config": {"workload": f"C3: samples(dbm {UNITS[0]}->{UNITS[1]}->{UNITS[2]}, {n} chains per GPU, burnin={burnin}); "
                               "chains sharded over GPUs, no collective",
                   "chains_per_gpu": n, "burnin": burnin, "units": UNITS,
                   "l2": "inputs larger than L2 (particle state %.1f GB per GPU)" % (n * (2048 * 2 * 2 + 256 * 2 + 64 * 2) * 2 / 1e9)},
        "cells_per_sec": value / burnin,
        "gpu_launches": launches, "tc_launches": tc_launches,
        "e2e": e2e, "roofline": roofline, "clocks": sampler.summary(),
    }
    line.update(extra)
    if rank == 0 and world == 1 and not args.no_training:
        line["training"] = training_metrics(pkg, ctx, with_cpu=not args.no_cpu_baseline)
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(layers, burnin=5, chains=args.cpu_chains, workers=os.cpu_count() or 1)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()



# ------------------------------------------------- other BASELINE configs, fixed TOTAL work (strong scaling over --gpus)
def _max_over_ranks(torch, dist, world, x):
    t = torch.tensor([x], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def strong_c3(pkg, ctx, torch, dist, world, rank, args, host_model, barrier):
    """C3 exactly as BASELINE.json words it: 1 M cells TOTAL, chains sharded over the GPUs (no collective)."""
    total, burnin = args.chains, args.burnin
    shard = pkg.mostevenbatches(total, world)
    dm = pkg.DeviceModel.from_host(host_model, ctx)
    p = pkg.DeviceParticles(dm, shard[rank], sum(shard[:rank]))
    p.init(False)
    pkg.gibbssample_(p, dm, burnin)
    barrier()
    ctx.timer_start()
    reps = 2
    for _ in range(reps):
        p.init(False)
        pkg.gibbssample_(p, dm, burnin)
    ms = _max_over_ranks(torch, dist, world, ctx.timer_stop()) / reps
    barrier()
    return {"metric": "Gibbs sweeps x chains / sec", "scaling": "strong", "chains_total": total, "burnin": burnin,
            "ms_per_pass": ms, "value": total * burnin / (ms * 1e-3), "unit": "sweeps*chains/s", "collective": "none"}


def dp_pcd_c4(pkg, ctx, torch, dist, world, rank, args, barrier, pk):
    """C4: 3-layer scDBM 20000->1024->256->64, PCD (`traindbm!`) on 100 000 cells TOTAL with 8192 persistent
    particles TOTAL; rows and particles sharded over the GPUs, one NCCL allreduce of the gradient block per step
    (inside the library), global mean-field stopping rule.  Data are generated on the device per shard."""
    cells, parts, steps = args.c4_cells, args.c4_particles, args.c4_steps
    layers = make_model(C4_UNITS, seed=4)
    host = [pkg.NegativeBinomialBernoulliRBM(**layers[0])] + [pkg.BernoulliRBM(**l) for l in layers[1:]]
    dm = pkg.DeviceModel.from_host(host, ctx)
    rows = pkg.mostevenbatches(cells, world)
    pshard = pkg.mostevenbatches(parts, world)
    g = np.random.default_rng(4)
    gene_mean = np.clip(np.exp(g.normal(-1.5, 1.5, C4_UNITS[0])), 1e-3, 50.0)
    x = pkg.Mat(ctx, rows[rank], C4_UNITS[0], pkg.MAT_COUNT).generate_counts(gene_mean, np.full(C4_UNITS[0], 0.5), 0.3, seed=4,
                                                                             row_offset=sum(rows[:rank]))
    p = pkg.DeviceParticles(dm, pshard[rank], row_offset=sum(pshard[:rank])).init(False)
    lib, ck = pkg.lib(), pkg._lib.check
    mu = pkg.DeviceParticles(dm, rows[rank], real_valued=True)
    opt = pkg.LoglikelihoodOptimizer().initialized(dm)
    import ctypes as C
    its = C.c_int32()

    def step():
        if world > 1:
            ck(lib.scdbm_dbm_pcd_step_dp(dm.h, x.h, p.h, mu.h, opt.h, 1e-4, 5, 1e-3, cells, parts, C.byref(its)))
        else:
            ck(lib.scdbm_dbm_pcd_step(dm.h, x.h, p.h, mu.h, opt.h, 1e-4, 5, 1e-3, C.byref(its)))

    for _ in range(2):
        step()
    barrier()
    ctx.launch_count(reset=True)
    ctx.timer_start()
    for _ in range(steps):
        step()
    ms = _max_over_ranks(torch, dist, world, ctx.timer_stop()) / steps
    launches = ctx.launch_count() / steps
    barrier()
    # the exchange step alone: allreduce(sum) of the whole gradient block on the context's stream
    ar_ms = None
    nfl = opt.buffer()[1]
    if world > 1:
        for _ in range(2):
            ck(lib.scdbm_allreduce_gradient(opt.h))
        barrier()
        ctx.timer_start()
        for _ in range(5):
            ck(lib.scdbm_allreduce_gradient(opt.h))
        ar_ms = _max_over_ranks(torch, dist, world, ctx.timer_stop()) / 5
        barrier()
    S = sum(a * b for a, b in zip(C4_UNITS[:-1], C4_UNITS[1:]))
    S_hid = sum(a * b for a, b in zip(C4_UNITS[1:-1], C4_UNITS[2:]))
    mf = int(its.value)
    flops = 5 * parts * 4 * S + cells * (2 * S + mf * 4 * S_hid) + 2 * (cells + parts) * S       # SURVEY 8d, x W1 hoisted
    out = {"metric": "DBM PCD steps / sec (traindbm!)", "scaling": "strong", "units": C4_UNITS, "cells_total": cells,
           "particles_total": parts, "steps": steps, "ms_per_step": ms, "value": 1e3 / ms, "unit": "steps/s",
           "meanfield_iterations": mf, "launches_per_step": launches,
           "collective": "ncclAllReduce(sum) of {dW, da, db} on a second stream, in row blocks under the gradient GEMM; "
                         "ncclAllReduce(max) of one word per mean-field iteration" if world > 1 else "none (1 GPU, same total work)",
           "allreduce_bytes_per_step": int(nfl * 4) if world > 1 else 0, "allreduce_alone_ms": ar_ms,
           "allreduce_busbw_gbs": (2.0 * (world - 1) / world * nfl * 4 / (ar_ms * 1e-3) / 1e9) if ar_ms else None,
           "algorithmic_flops_per_step": flops, "achieved_tflops_aggregate": flops / (ms * 1e-3) / 1e12,
           "frac_of_tensor_peak_sustained": flops / (ms * 1e-3) / 1e12 / (world * pk["tflops_sustained"])}
    if rank == 0 and world == 1:
        out["kernels"] = c4_kernel_rooflines(pkg, ctx, dm, x, mu, p, opt, cells, parts, pk)
    return out


def c4_kernel_rooflines(pkg, ctx, dm, x, mu, p, opt, cells, parts, pk):
    """Live CUDA-event timings of the deterministic GEMMs at the C4 shapes (K4 mean-field, K5 gradient)."""
    lib, ck = pkg.lib(), pkg._lib.check
    V, H1 = C4_UNITS[0], C4_UNITS[1]
    res = {}

    def timed(f, reps=3):
        f()
        ctx.timer_start()
        for _ in range(reps):
            f()
        return ctx.timer_stop() / reps

    # K4: the positive phase.  One fixed iteration = the hoisted x W1 GEMM (n x 20000 x 1024, count operand) + the
    # initialisation pass + one Gauss-Seidel sweep over the hidden layers; the hoisted GEMM dominates
    t1 = timed(lambda: ck(lib.scdbm_meanfield(dm.h, x.h, 0.0, 1, mu.h, None, None)))
    t3 = timed(lambda: ck(lib.scdbm_meanfield(dm.h, x.h, 0.0, 5, mu.h, None, None)))
    per_iter = (t3 - t1) / 4
    S_hid = sum(a * b for a, b in zip(C4_UNITS[1:-1], C4_UNITS[2:]))
    f_init = 2.0 * cells * (V * H1 + S_hid)
    res["K4_meanfield"] = {"kernel": "k_gemm_cg2<3> (two-SM cta_group::2 tiles; hoisted x*W1, chunk-promoted fp32 accumulation) + hidden-layer sweeps",
                           "bound": "tensor", "init_ms": t1 - per_iter, "per_iteration_ms": per_iter,
                           "achieved": f_init / ((t1 - per_iter) * 1e-3) / 1e12, "peak": pk["tflops_burst"], "unit": "TFLOP/s",
                           "frac": f_init / ((t1 - per_iter) * 1e-3) / 1e12 / pk["tflops_burst"],
                           "per_iteration_tflops": 4.0 * cells * S_hid / (per_iter * 1e-3) / 1e12,
                           "note": "count x (W hi + W lo) = 2 plane products issued per algorithmic FLOP"}
    # K5: StackedOptimizer gradient, all layers (layer 0: (cells + particles) x 20000 x 1024)
    tg = timed(lambda: ck(lib.scdbm_compute_gradient_dbm(opt.h, dm.h, mu.h, p.h, 0.0, 0.0)))
    S = sum(a * b for a, b in zip(C4_UNITS[:-1], C4_UNITS[1:]))
    fg = 2.0 * (cells + parts) * S
    res["K5_gradient"] = {"kernel": "k_grad_tc (MN-major operands, two TMEM accumulators, chunk-promoted) + column sums",
                          "bound": "tensor", "launch_ms": tg, "achieved": fg / (tg * 1e-3) / 1e12, "peak": pk["tflops_burst"],
                          "unit": "TFLOP/s", "frac": fg / (tg * 1e-3) / 1e12 / pk["tflops_burst"],
                          "note": "count x real (positive phase) = 2 plane products, count x binary (negative) = 1"}
    return res


def ais_c5(pkg, ctx, torch, dist, world, rank, args, barrier, pk):
    """C5 slice: `aislogimpweights` on the 3-layer scDBM, 65 536 particles TOTAL sharded over the GPUs, the first
    `--ais-temps` transitions of the 10 000-step schedule collect(0:1e-4:1); the final logmeanexp is reduced over ranks
    with two scalar allreduces."""
    parts, T, burnin = args.ais_particles, args.ais_temps, 5
    layers = make_model(C4_UNITS, seed=4)
    host = [pkg.NegativeBinomialBernoulliRBM(**layers[0])] + [pkg.BernoulliRBM(**l) for l in layers[1:]]
    dm = pkg.DeviceModel.from_host(host, ctx)
    shard = pkg.mostevenbatches(parts, world)
    off = sum(shard[:rank])
    temps = np.arange(T + 1) * 1e-4
    pkg.aislogimpweights(dm, temperatures=temps[:3], nparticles=shard[rank], burnin=burnin, row_offset=off, ctx=ctx)   # warm-up
    barrier()
    ctx.timer_start()
    w = pkg.aislogimpweights(dm, temperatures=temps, nparticles=shard[rank], burnin=burnin, row_offset=off, ctx=ctx)
    ms = _max_over_ranks(torch, dist, world, ctx.timer_stop())
    lme = pkg.logmeanexp_sharded(w, ctx)
    barrier()
    S = sum(a * b for a, b in zip(C4_UNITS[:-1], C4_UNITS[1:]))
    flops = T * (burnin * parts * 4.0 * S + 2.0 * parts * S)
    out = {"metric": "AIS particles x temperatures / sec", "scaling": "strong", "units": C4_UNITS, "particles_total": parts,
           "transitions": T, "burnin": burnin, "seconds": ms * 1e-3, "value": parts * T / (ms * 1e-3), "unit": "particle*temperatures/s",
           "collective": "two scalar ncclAllReduce (max, sum) for logmeanexp" if world > 1 else "none",
           "logmeanexp_slice": float(lme), "algorithmic_flops": flops, "achieved_tflops_aggregate": flops / (ms * 1e-3) / 1e12,
           "nb_draws_per_sec": T * burnin * parts * C4_UNITS[0] / (ms * 1e-3)}
    if rank == 0 and world == 1:
        # K8: the AIS ratio sweeps alone (h1 x W1' -> 20000 softplus differences per particle, row-summed in the epilogue,
        # plus the even hidden layer): scdbm_ais_run with burnin = 0 launches nothing else per temperature, so the
        # difference of a 12- and a 2-temperature call is 10 ratio steps (allocation, init and download cancel)
        P = min(parts, 16384)
        lib, ck = pkg.lib(), pkg._lib.check

        def ratio_call(ntemps):
            t = np.ascontiguousarray(0.5 + 1e-4 * np.arange(ntemps))
            outw = np.zeros(P)
            ck(lib.scdbm_ais_run(dm.h, t.ctypes.data_as(pkg._lib.D), ntemps, P, 0, 0, outw.ctypes.data_as(pkg._lib.D)))

        def timed(ntemps):
            ratio_call(ntemps)
            best = 1e30
            for _ in range(3):
                ctx.sync()
                ctx.timer_start()
                ratio_call(ntemps)
                best = min(best, ctx.timer_stop())
            return best

        t_ratio = (timed(12) - timed(2)) / 10
        fr = 2.0 * P * (C4_UNITS[0] * C4_UNITS[1] + C4_UNITS[1] * C4_UNITS[2] + C4_UNITS[2] * C4_UNITS[3])
        out["K8_ais_ratio"] = {"kernel": "k_gemm_tc<128,0,16> EPI_AIS_RATIO: visible layer summed out from h1 (N = 20000, K = 1024) + "
                                         "even hidden layer; one temperature step without Gibbs transitions",
                               "bound": "tensor", "particles": P, "per_temperature_ms": t_ratio, "achieved": fr / (t_ratio * 1e-3) / 1e12,
                               "peak": pk["tflops_burst"], "unit": "TFLOP/s", "frac": fr / (t_ratio * 1e-3) / 1e12 / pk["tflops_burst"],
                               "note": "binary x (W hi + W lo) = 2 plane products; 2 softplus evaluations (4 MUFU) per output element"}
    return out

# ---------------------------------------- secondary metric: CD / PCD epochs per second
def synth_counts(n, V, seed, r=0.5):
    g = np.random.default_rng(seed)
    m = np.clip(np.exp(g.normal(-1.5, 1.5, V)), 1e-3, 50.0)
    s = np.exp(g.normal(0.0, 0.3, n))
    x = g.poisson(g.gamma(r, s[:, None] * m[None, :] / r)).astype(np.float64)
    x[0, :] += 1.0            # no all-zero gene (log(0) in initrbm)
    return x


def _spin_up(pkg, ctx, model, seconds=0.4):
    """GPU clocks up: Gibbs sweeps over scratch particles of `model` (the model and the training state are untouched)."""
    scratch = pkg.DeviceParticles(model, 20000).init(False)
    t_end = time.perf_counter() + seconds
    while time.perf_counter() < t_end:
        pkg.gibbssample_(scratch, model, 5)
        ctx.sync()
    scratch.close()


def training_metrics(pkg, ctx, with_cpu=True):
    """BASELINE.json's second metric, "PCD epochs/sec", on configs C1 and C2 (reported beside the
    headline; timed with CUDA events, data resident)."""
    out = {}
    # C1: NB-Bernoulli RBM, fitrbm CD-1, 500 cells x 1000 genes, 128 hidden, batchsize 100
    x1 = synth_counts(500, 1000, seed=1)
    xm = pkg.Mat.from_host(ctx, x1, pkg.MAT_COUNT)
    m = pkg.DeviceModel(ctx, [1000, 128], pkg.VIS_NEGBIN)
    m.init_layer(0, xm, seed=SEED)
    tr = pkg.RBMTrainer(m, 100, pcd=False)
    # the host work above leaves the GPU idle long enough for its clocks to drop, and these launch-bound loops are over
    # before they have ramped up again (a 3-epoch warm-up alone gave 840 ... 1490 C2 steps/s from run to run): keep the GPU
    # busy for 0.4 s with work that does not touch the training state, then 3 untimed epochs
    _spin_up(pkg, ctx, m)
    for _ in range(3):
        pkg.trainrbm_(m, xm, tr, 1e-4, 1)
    best = 1e30
    for _ in range(3):                      # best of 3 repetitions (host launch jitter)
        ctx.sync()
        ctx.timer_start()
        for _ in range(20):
            pkg.trainrbm_(m, xm, tr, 1e-4, 1)
        best = min(best, ctx.timer_stop())
    out["C1_rbm_cd1_epochs_per_sec"] = 20 / (best * 1e-3)
    # C2: DBM 2000 -> 256 -> 64, 5000 cells, full-batch PCD step (one traindbm! "epoch"), 100 particles
    x2 = synth_counts(5000, 2000, seed=2)
    layers = make_model()
    dm = pkg.DeviceModel.from_host([pkg.NegativeBinomialBernoulliRBM(**layers[0]), pkg.BernoulliRBM(**layers[1])], ctx)
    x2m = pkg.Mat.from_host(ctx, x2, pkg.MAT_COUNT)
    p = pkg.DeviceParticles(dm, 100).init(False)
    _spin_up(pkg, ctx, dm)
    pkg.traindbm_(dm, x2m, epochs=3, particles=p, learningrate=1e-4, device=True, ctx=ctx)
    best = 1e30
    for _ in range(3):
        ctx.sync()
        ctx.timer_start()
        pkg.traindbm_(dm, x2m, epochs=20, particles=p, learningrate=1e-4, device=True, ctx=ctx)
        best = min(best, ctx.timer_stop())
    out["C2_dbm_pcd_epochs_per_sec"] = 20 / (best * 1e-3)
    out["note"] = "GPU kept busy for 0.4 s (clocks), 3 untimed epochs, then best of 3 x 20 epochs, CUDA events, data resident"
    if with_cpu:
        from oracle import model as om, philox as px
        rng = px.Rng(SEED)
        rbm = om.initrbm(x1, 128, rng, "nb")
        otr = om.RBMTrainer(rbm, 100, rng, pcd=False)
        om.trainrbm(rbm, x1, otr, 1e-4)
        t0 = time.perf_counter()
        for _ in range(5):
            om.trainrbm(rbm, x1, otr, 1e-4)
        out["C1_cpu_port_epochs_per_sec"] = 5 / (time.perf_counter() - t0)
        dbm = [om.NegativeBinomialBernoulliRBM(**layers[0]), om.BernoulliRBM(**layers[1])]
        parts = om.initparticles(dbm, 100, rng)
        t0 = time.perf_counter()
        for _ in range(2):
            om.traindbm_step(dbm, x2, parts, rng, 1e-4)
        out["C2_cpu_port_epochs_per_sec"] = 2 / (time.perf_counter() - t0)
        out["cpu_note"] = "oracle (NumPy/OpenBLAS, all BLAS threads), not Julia"
    return out


# --------------------------------------------------------- CPU baseline arm
def _cpu_worker(job):
    layers, chains, burnin, row_offset = job
    os.environ["OPENBLAS_NUM_THREADS"] = "1"
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(1)
    except Exception:
        pass
    from oracle import model as om
    dbm = [om.NegativeBinomialBernoulliRBM(**layers[0]), om.BernoulliRBM(**layers[1])]
    t0 = time.perf_counter()
    # timing twin of the oracle: the reference's operation sequence with NumPy's compiled NB / uniform samplers
    # (the Philox-addressed inversion the parity tests use costs ~3x more per draw and would flatter the GPU)
    om.sampleparticles_native(dbm, chains, np.random.default_rng(SEED + row_offset), burnin)
    return time.perf_counter() - t0


def cpu_baseline(layers, burnin, chains, workers):
    """The oracle (NumPy/OpenBLAS restatement of the reference, not Julia) on a
    bounded sample: `chains` chains x `burnin` sweeps per worker process."""
    import multiprocessing as mp
    jobs = [(layers, chains, burnin, w * chains) for w in range(workers)]
    t0 = time.perf_counter()
    with mp.get_context("fork").Pool(workers) as pool:
        pool.map(_cpu_worker, jobs)
    dt = time.perf_counter() - t0
    return {"value": workers * chains * burnin / dt, "unit": "sweeps*chains/s", "cores": workers, "kind": "port",
            "sample": f"{workers} worker processes x {chains} chains x {burnin} sweeps of the same 2000->256->64 model "
                      f"(CPU restatement of the reference in NumPy/OpenBLAS with NumPy's compiled NB sampler, not Julia); wall {dt:.1f} s"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    layers = make_model()
    workers = os.cpu_count() or 1
    # warm-up and steps each run a bounded sample; keep the whole run within a few minutes
    chains, burnin = args.cpu_chains, 5
    for _ in range(min(args.warmup, 1)):
        cpu_baseline(layers, burnin, max(chains // 4, 8), workers)
    t0 = time.perf_counter()
    vals = [cpu_baseline(layers, burnin, chains, workers) for _ in range(args.steps)]
    dt = time.perf_counter() - t0
    v = float(np.mean([x["value"] for x in vals]))
    cb = dict(vals[-1])
    cb["value"] = v
    line = {"impl": "reference", "metric": "Gibbs sweeps x chains / sec (synthetic-cell generation, 2-layer scDBM)",
            "value": v, "unit": "sweeps*chains/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic (random valid parameters, seed 3)",
            "config": {"workload": f"C3 bounded sample: {workers} x {chains} chains x {burnin} sweeps per step, same model",
                       "units": UNITS},
            "cpu_baseline": cb,
            "e2e": {"value": v, "unit": "sweeps*chains/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--chains", type=int, default=1_000_000, help="chains per GPU (C3: 1M cells)")
    ap.add_argument("--burnin", type=int, default=50)
    ap.add_argument("--cpu-chains", type=int, default=4000, help="chains per worker in the CPU baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-training", action="store_true", help="skip the secondary CD/PCD epochs/sec metrics")
    ap.add_argument("--no-secondary", action="store_true", help="skip the strong-scaling C3 / C4 data-parallel PCD / C5 AIS legs")
    ap.add_argument("--c4-cells", type=int, default=100_000, help="C4: cells in total (sharded over the GPUs)")
    ap.add_argument("--c4-particles", type=int, default=8192)
    ap.add_argument("--c4-steps", type=int, default=5)
    ap.add_argument("--ais-particles", type=int, default=65536, help="C5: particles in total (sharded over the GPUs)")
    ap.add_argument("--ais-temps", type=int, default=20, help="C5: transitions timed (slice of the 10 000)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
