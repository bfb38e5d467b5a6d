"""Static SASS instruction counts per kernel of libcfd_b200.so (cuobjdump -sass): total, fp64, loads/stores, calls.
usage: python tools/sass_count.py <pattern> [...]"""
import collections
import re
import subprocess
import sys

so = __file__.rsplit("/", 2)[0] + "/cfd_demo_b200/libcfd_b200.so"
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", txt)[1:]
names = [f.split("\n", 1)[0].strip() for f in funcs]
dem = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
for name, body in zip(dem, funcs):
    short = re.sub(r"\(.*", "", name).replace("void cfdk::", "")
    if not any(p in short for p in sys.argv[1:]):
        continue
    ops = re.findall(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", body)
    c = collections.Counter(o.split(".")[0] for o in ops)
    fp64 = sum(v for k, v in c.items() if k in ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX"))
    mem = sum(v for k, v in c.items() if k in ("LDG", "STG", "LD", "ST", "LDS", "STS", "LDC", "LDL", "STL"))
    top = ", ".join(f"{k} {v}" for k, v in c.most_common(12))
    print(f"{short:42s} total {len(ops):5d}  fp64 {fp64:4d}  mem {mem:4d}  CALL {c.get('CALL', 0):3d}  BRA {c.get('BRA', 0):3d} | {top}")
