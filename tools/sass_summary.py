"""Per-function SASS mnemonic counts of libradar_depth_b200.so -> profiles/rNN_sass_summary.txt (the tracked proof that the
hot kernels are tcgen05 / TMEM / TMA code: UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG = cp.async.bulk.tensor, ...).
usage: python tools/sass_summary.py [out.txt]      (needs cuobjdump; runs in the build container, no GPU)"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "radar_depth_b200", "libradar_depth_b200.so")
out_path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r02_sass_summary.txt")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", txt)
out = ["# SASS evidence of the Blackwell-native paths in radar_depth_b200/libradar_depth_b200.so (sm_100a)",
       "# command: python tools/sass_summary.py  (cuobjdump -sass, per-function mnemonic counts)",
       "# UTCHMMA = tcgen05.mma (kind::f16), UTCBAR = tcgen05.commit, LDTM = tcgen05.ld, UTMALDG = cp.async.bulk.tensor (TMA tile load),",
       "# UTMALDG.4D = the 128-byte-swizzled [slot][64 channels] boxes of the weight-gradient kernel (SWIZZLE_128B MN-major operands), .5D = 16-byte chunk planes,",
       "# UBLKCP = cp.async.bulk (weight tiles), SYNCS = mbarrier ops, UTCATOMSWS = TMEM allocation, REDG = fp32 vector reductions", ""]
keys = ["UTCHMMA", "UTCBAR", "LDTM", "UTMALDG", "UTMALDG.4D", "UTMALDG.5D", "UBLKCP", "SYNCS", "UTCATOMSWS", "LDGSTS", "REDG", "HMMA", "ATOMG"]
tot = collections.Counter()
for f in funcs[1:]:
    name = f.split("\n", 1)[0].strip()
    c = collections.Counter({k: len(re.findall(r"\b" + re.escape(k), f)) for k in keys})
    tot.update(c)
    if c["UTCHMMA"] or c["UTMALDG"] or c["LDTM"]:
        out += [name, "    " + "  ".join(f"{k}={v}" for k, v in c.items() if v)]
out += ["", "library totals: " + "  ".join(f"{k}={v}" for k, v in tot.items() if v), "",
        "# first tcgen05 / TMA instructions of the bf16 forward kernel:"]
m = [f for f in funcs[1:] if "conv_fprop_kernelI13__nv_bfloat16" in f.split("\n", 1)[0]]
if m:
    lines = [ln.strip() for ln in m[0].split("\n") if re.search(r"UTCHMMA|UTMALDG|LDTM|UTCBAR|UBLKCP", ln)]
    out += ["    " + re.sub(r"\s+", " ", ln).strip() for ln in lines[:14]]
open(out_path, "w").write("\n".join(out) + "\n")
print(out_path)
