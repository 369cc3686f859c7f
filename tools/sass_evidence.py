#!/usr/bin/env python
"""Static SASS evidence: mnemonic counts per kernel of the built library (no GPU needed).

    python tools/sass_evidence.py > profiles/<tag>_sass_evidence.txt

Counts are of the static code, not of executed instructions.  What to look for: UTMALDG + SYNCS in the sweep (TMA boxes
arriving on mbarriers), ATOMG.E.CAS.128 in the raster (the 128-bit z-buffer exchange), the .SYS loads / stores of the flag
barrier and of the kernels that write into another rank's memory.
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "contour3d_b200", "libcontour3d_b200.so")

LEGEND = [
    ("UTMALDG", "TMA tensor-map load (cp.async.bulk.tensor)"),
    ("SYNCS", "mbarrier operations"),
    ("ATOMG.E.CAS.128", "128-bit global compare-and-swap"),
    ("ATOMG", "global atomics (all kinds)"),
    ("ATOMS", "shared-memory atomics"),
    ("REDUX", "warp-wide integer reduce"),
    ("SYS", "loads / stores / fences at system scope (.SYS): the flag barrier between ranks, stores into a peer's memory"),
    ("NANOSLEEP", "back-off in a spin loop"),
    ("MUFU.RCP", "fp32 reciprocal"),
    ("DFMA", "fp64 FMA (the kernel TUs are built with -fmad=false: div / sqrt expansions only; k_work_plan is plain arithmetic)"),
    ("DADD", "fp64 add"),
    ("DMUL", "fp64 mul"),
    ("LDG", "global loads"),
    ("LDS", "shared loads"),
    ("STG", "global stores"),
]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels = collections.OrderedDict()
    name = None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = m.group(1)
            kernels[name] = []
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m and name is not None:
            kernels[name].append(m.group(1))
    print(f"SASS of contour3d_b200/libcontour3d_b200.so (sm_100a): cuobjdump -sass, mnemonic counts per kernel (static code, not executed counts)")
    print("legend:")
    for key, what in LEGEND:
        print(f"  {key:<18} {what}")
    print()
    for kernel, ops in kernels.items():
        counts = []
        for key, _ in LEGEND:
            if key == "SYS":
                n = sum(1 for o in ops if ".SYS" in o)
            elif key in ("ATOMG", "LDG", "STG", "LDS", "ATOMS", "REDUX", "SYNCS", "UTMALDG", "NANOSLEEP"):
                n = sum(1 for o in ops if o.split(".")[0] == key)
            else:
                n = sum(1 for o in ops if o.startswith(key))
            if n:
                counts.append(f"{key} {n}")
        print(kernel)
        print(f"   {len(ops)} instructions: " + ", ".join(counts))
    return 0


if __name__ == "__main__":
    sys.exit(main())
