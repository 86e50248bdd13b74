#!/usr/bin/env python3
"""Compare two builds of the library kernel by kernel at the SASS level: `sass_diff.py old.so new.so`.
Lists kernels whose instruction stream changed, disappeared or is new -- the check used before committing a refactor of
hardware-validated kernels while no GPU is available (an identical instruction stream cannot change behaviour)."""
import collections
import re
import subprocess
import sys


def kernels(lib):
    txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    out, cur = {}, None
    for line in txt.splitlines():
        if "Function :" in line:
            cur = line.split("Function :")[1].strip()
            out[cur] = []
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?)\s*;", line)
        if m and cur:
            out[cur].append(m.group(1))
    return out


if __name__ == "__main__":
    a, b = kernels(sys.argv[1]), kernels(sys.argv[2])
    def opcodes(ins):   # mnemonic + modifiers, operands and predicates dropped: invariant under register renaming / scheduling
        return collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", i).split()[0] for i in ins)

    differ = [k for k in a if k in b and a[k] != b[k]]
    changed = [k for k in differ if opcodes(a[k]) != opcodes(b[k])]
    realloc = [k for k in differ if opcodes(a[k]) == opcodes(b[k])]
    print(f"{len(a)} -> {len(b)} kernels; identical: {sum(1 for k in a if k in b and a[k] == b[k])}; "
          f"same opcode mix, different register allocation / order: {len(realloc)}")
    for title, ks in (("changed", changed), ("removed", [k for k in a if k not in b]), ("new", [k for k in b if k not in a])):
        for k in ks:
            print(f"  {title}: {k}  ({len(a.get(k, []))} -> {len(b.get(k, []))} instructions)")
    sys.exit(1 if changed else 0)
