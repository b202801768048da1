#!/usr/bin/env python
"""Static evidence from the built library (no GPU needed): which SASS instructions the kernels contain.
    python profiles/sass_evidence.py > profiles/r2_sass_evidence.txt
UBLKCP = cp.async.bulk (TMA 1-D bulk copy), SYNCS = mbarrier operations, REDUX = warp reductions, ATOMS = shared-memory
atomics (the ring protocol), ATOMG/RED = global atomics (queues), LDL/STL = local-memory (spill) traffic."""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "freesasa_b200", "csrc", "libfsb200.so")
sass = subprocess.run(["cuobjdump", "-sass", LIB], check=True, capture_output=True, text=True).stdout
print(f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)}  (arch: {re.search(r'arch = (sm_\w+)', sass).group(1)})")
print(f"# {'kernel':58s} {'instr':>7s} {'UBLKCP':>7s} {'SYNCS':>6s} {'REDUX':>6s} {'VOTE':>5s} {'SHFL':>5s} {'ATOMS':>6s} {'ATOMG':>6s} {'MUFU':>5s} {'DFMA..':>7s} {'LDL':>4s} {'STL':>4s}")
for chunk in sass.split("Function : ")[1:]:
    name = chunk.split("\n", 1)[0].strip()
    short = re.sub(r"_ZN6fsb200\d+_GLOBAL__N__[0-9a-f]+_\d+_\w+_cu_[0-9a-f]+", "", name)
    ops = re.findall(r"/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", chunk)
    c = lambda pat: sum(1 for o in ops if re.match(pat, o))
    print(f"  {short[:58]:58s} {len(ops):7d} {c(r'UBLKCP'):7d} {c(r'SYNCS'):6d} {c(r'REDUX'):6d} {c(r'VOTE'):5d} {c(r'SHFL'):5d} "
          f"{c(r'ATOMS'):6d} {c(r'ATOMG|RED'):6d} {c(r'MUFU'):5d} {c(r'D(FMA|ADD|MUL|SETP)'):7d} {c(r'LDL'):4d} {c(r'STL'):4d}")
print()
print("# the TMA staging of a fill (fill_slot in integrate.cu): excerpt of k_integrate<0,float>")
m = re.search(r"Function : \S*k_integrateILi0EfE.*?\n(.*?)(?=Function : |\Z)", sass, re.S)
lines = m.group(1).split("\n")
for i, l in enumerate(lines):
    if "UBLKCP" in l:
        for k in lines[max(0, i - 6):i + 3]:
            if "/*" in k and not re.match(r"\s*/\* 0x", k):
                print("   " + k.strip()[:120])
        print("   ...")
        break
