#!/usr/bin/env python3
"""Extract the SASS of one kernel from a built library: `sass_fn.py lib.so substring [--ops]`.
Prints the instruction text without addresses / encodings (for diffing two builds) or, with --ops, an opcode histogram.
Used to check that refactors of hardware-validated kernels leave their instruction stream unchanged."""
import collections
import re
import subprocess
import sys


def function_sass(lib, needle):
    txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    out, on = [], False
    for line in txt.splitlines():
        if "Function :" in line:
            on = needle in line
            continue
        if on:
            m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?)\s*;", line)
            if m:
                out.append(m.group(1))
    return out


if __name__ == "__main__":
    ins = function_sass(sys.argv[1], sys.argv[2])
    if "--ops" in sys.argv:
        h = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", i).split()[0] for i in ins)
        for k, v in h.most_common():
            print(f"{v:6d} {k}")
        print(f"{len(ins):6d} total")
    else:
        print("\n".join(ins))
