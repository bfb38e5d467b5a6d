"""Summarise `nvcc -Xptxas -v` output (csrc/build.log): registers, spills, shared memory per kernel.
usage: python tools/ptxas_summary.py [pattern]"""
import re
import subprocess
import sys

log = open(__file__.rsplit("/", 2)[0] + "/cfd_demo_b200/csrc/build.log").read().split("\n")
pat = sys.argv[1] if len(sys.argv) > 1 else ""
names, rows = [], []
for i, l in enumerate(log):
    m = re.search(r"Compiling entry function '(\S+)' for", l)
    if not m:
        continue
    blob = " ".join(log[i + 1:i + 4])
    regs = re.search(r"Used (\d+) registers", blob)
    spill = re.search(r"(\d+) bytes spill stores, (\d+) bytes spill loads", blob)
    smem = re.search(r"(\d+) bytes smem", blob)
    names.append(m.group(1))
    rows.append((int(regs.group(1)) if regs else -1, spill.group(1) if spill else "?", spill.group(2) if spill else "?",
                 smem.group(1) if smem else "0"))
dem = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
for n, r in zip(dem, rows):
    short = re.sub(r"\(.*", "", n).replace("void cfdk::", "")
    if pat in short:
        print(f"{short:55s} regs {r[0]:4d}  spill st/ld {r[1]:>4s}/{r[2]:>4s}  smem {r[3]}")
