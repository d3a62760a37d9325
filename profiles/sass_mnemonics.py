"""SASS mnemonic counts of the built library (cuobjdump -sass): which tcgen05 / TMA / cluster / griddepcontrol forms the
kernels really contain.  Writes profiles/r02_sass_mnemonics.txt."""
import os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
txt = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, "scdbm.jl_b200", "libscdbm_b200.so")], capture_output=True, text=True, check=True).stdout
pats = [("UTCHMMA (tcgen05.mma)", r"UTCHMMA"), ("UTCHMMA.2CTA (cta_group::2)", r"UTCHMMA\.2CTA"), ("UTCQMMA", r"UTCQMMA"),
        ("UTMALDG (TMA load)", r"UTMALDG"), ("UTMALDG.2D.2CTA", r"UTMALDG\.2D\.2CTA"), ("UTMASTG", r"UTMASTG"),
        ("LDTM (tcgen05.ld)", r"\bLDTM"), ("STTM (tcgen05.st)", r"\bSTTM"), ("UTCBAR (tcgen05.commit)", r"UTCBAR"),
        ("UTCBAR.2CTA.MULTICAST", r"UTCBAR\.2CTA\.MULTICAST"), ("UCGABAR_ARV/WAIT (barrier.cluster)", r"UCGABAR_"),
        ("UTCATOMSWS (tcgen05.alloc)", r"UTCATOMSWS"), ("SYNCS (mbarrier)", r"\bSYNCS"),
        ("ACQBULK (griddepcontrol.wait)", r"ACQBULK"), ("PREEXIT (griddepcontrol.launch_dependents)", r"PREEXIT"),
        ("HMMA. (mma.sync)", r"\bHMMA\."), ("HGMMA (wgmma)", r"HGMMA"), ("LDGSTS (cp.async)", r"LDGSTS"),
        ("MUFU", r"\bMUFU"), ("VOTE", r"\bVOTE"), ("REDUX", r"\bREDUX")]
out = ["# SASS mnemonic counts of scdbm.jl_b200/libscdbm_b200.so (cuobjdump -sass), sm_100a, nvcc 12.9",
       f"# built from the sources with sha {bench.kernel_source_hash()}"]
out += [f"{name:50s} {len(re.findall(p, txt))}" for name, p in pats]
out += ["", "# per kernel: tcgen05.mma (of which .2CTA) / TMA loads (of which .2CTA) / TMEM loads / TMEM stores"]
for f in re.split(r"\n\s*Function : ", txt)[1:]:
    name = f.split("\n", 1)[0].strip()
    if re.search(r"k_gemm_tc|k_gemm_cg2|k_grad_tc|k_grad_cg2", name):
        c = lambda p: len(re.findall(p, f))
        out.append(f"{name:100s} {c('UTCHMMA'):4d} ({c(r'UTCHMMA[.]2CTA')}) {c('UTMALDG'):4d} ({c(r'UTMALDG[.]2D[.]2CTA')}) {c('LDTM'):4d} {c('STTM'):4d}")
open(os.path.join(ROOT, "profiles", "r02_sass_mnemonics.txt"), "w").write("\n".join(out) + "\n")
print("\n".join(out))
