#!/usr/bin/env python3
"""Registers / stack / spill bytes of every kernel of the library from the log of a verbose build.

    python -m juqbox_b200.build --force -v > build_v.log 2>&1
    python tools/ptxas_table.py build_v.log > profiles/r02_ptxas_v.txt
"""
import re
import subprocess
import sys

txt = open(sys.argv[1]).read()
rows = {}
for m in re.finditer(r"Function properties for (\S+)\n\s*(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n"
                     r"ptxas info\s*: Used (\d+) registers", txt):
    name, stack, st, ld, regs = m.group(1), int(m.group(2)), int(m.group(3)), int(m.group(4)), int(m.group(5))
    rows[name] = (regs, stack, st, ld)
names = list(rows)
dem = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
clean = lambda s: re.sub(r"\(anonymous namespace\)::|<unnamed>::", "", s).replace("void ", "")
print("# ptxas -v of every kernel in libjuqbox_b200.so (nvcc 12.9, -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo), round 2 final build.")
print("# columns: registers | stack frame B | spill stores B | spill loads B | kernel (demangled)")
print("# 40-byte stack frames without spills belong to the FP64 sincos slow path (__internal_trig_reduction_slowpathd), not to register spills.")
print("# jq_traj_kernel<Lane, UPL, MINB, JT, OBJ, GLT, NW, PIPE, SEG>: SEG = true are the segment sweeps of the time-parallel evaluation.")
for n, d in sorted(zip(names, dem), key=lambda t: clean(t[1])):
    r = rows[n]
    print(f"{r[0]:4d} | {r[1]:5d} | {r[2]:5d} | {r[3]:5d} | {clean(d)}")
