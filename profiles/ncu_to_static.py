"""Turns an `ncu --set full` capture of `profiles/prof_visible.py --chains N --sweeps 3` into
profiles/r02_ncu_static.json: the counters bench.py may quote (roofline.static_profile) -- tagged with the sha256 of
the CUDA sources they were measured on, so that bench.py withholds them once the kernels change.

    python profiles/ncu_to_static.py gpurun_out/prof_r02c.ncu-rep --chains 1000000 [--commit <hash>]
"""
import argparse, csv, io, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench

ap = argparse.ArgumentParser()
ap.add_argument("report")
ap.add_argument("--chains", type=int, required=True)
ap.add_argument("--commit", default=None)
a = ap.parse_args()
raw = subprocess.run(["ncu", "-i", a.report, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}


def val(r, name, scale=None):
    v, u = float(r[ix[name]]), units[ix[name]]
    if scale == "bytes":
        v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
    if scale == "ms":
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1, "s": 1e3}[u]
    return v


def kernel(r):
    return {"kernel": r[ix["Kernel Name"]],
            "duration_ms_under_ncu": val(r, "gpu__time_duration.sum", "ms"),
            "dram_bytes_per_launch": val(r, "dram__bytes_read.sum", "bytes") + val(r, "dram__bytes_write.sum", "bytes"),
            "dram_read_bytes": val(r, "dram__bytes_read.sum", "bytes"), "dram_write_bytes": val(r, "dram__bytes_write.sum", "bytes"),
            "issue_active_pct": val(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
            "tensor_pipe_pct": val(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
            "alu_pipe_pct": val(r, "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
            "fma_pipe_pct": val(r, "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
            "fmaheavy_pipe_pct": val(r, "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed"),
            "xu_pipe_pct": val(r, "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
            "threads_per_instruction": val(r, "smsp__thread_inst_executed_per_inst_executed.ratio"),
            "warp_instructions": val(r, "smsp__inst_executed.sum"),
            "registers_per_thread": val(r, "launch__registers_per_thread")}


out = {"capture": os.path.basename(a.report), "commit": a.commit, "kernel_source_sha": bench.kernel_source_hash(), "chains": a.chains,
       "command": f"ncu --set full --clock-control none --import-source on -k regex:k_gemm_ -s 6 -c 3 python profiles/prof_visible.py --chains {a.chains} --sweeps 3",
       "note": "counters of ONE capture, cold-cache and serialised under ncu; not measured by the bench run that quotes them"}
for r in rows[2:]:
    name = r[ix["Kernel Name"]]
    key = "nb_downsweep" if "224, 1, 28" in name else "upsweep" if ("k_gemm_cg2<2>" in name.replace("(int)", "") or "128, 2, 16" in name) else "top_layer" if "64, 2, 16" in name else None
    if key and key not in out:
        out[key] = kernel(r)
        if key == "nb_downsweep":
            out[key]["warp_instructions_per_draw_x32"] = out[key]["warp_instructions"] * 32 / (a.chains * bench.UNITS[0])
json.dump(out, open(os.path.join(ROOT, "profiles", "r02_ncu_static.json"), "w"), indent=1)
print(json.dumps(out, indent=1)[:1500])
