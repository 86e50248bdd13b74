"""Warp-stall samples of one kernel by SOURCE LINE: joins the SASS page of an ncu report (ncu -i rep --page source --csv) with the
line table of the shipped cubin (nvdisasm -g).  Usage: python scripts/ncu_lines.py report.ncu-rep kernel_name_substring [top_n]"""
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

rep, kname = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.environ.get("DSHEG_LIB") or os.path.join(root, "diffsheg_b200", "libdiffsheg_b200.so")
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
body = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
col = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
# line table: instruction index -> (file, line)
with tempfile.TemporaryDirectory() as td:
    subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=td, capture_output=True)
    cub = [os.path.join(td, f) for f in os.listdir(td) if f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "-g", "-c", cub], capture_output=True, text=True).stdout
lines, cur, on = [], ("?", 0), False
for ln in dis.splitlines():
    if ln.startswith(".text."):
        on = kname in ln
        continue
    if not on:
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln):
        lines.append(cur)
if len(lines) != len(body):
    print(f"warning: {len(lines)} instructions in the cubin vs {len(body)} in the report", file=sys.stderr)
agg = {}
tot = 0
for i, r in enumerate(body):
    n = int(r[col["# Samples"]] or 0)
    tot += n
    key = lines[i] if i < len(lines) else ("?", 0)
    a = agg.setdefault(key, {"n": 0, "inst": 0, **{s: 0 for s in stalls}})
    a["n"] += n
    a["inst"] += int(r[col["Instructions Executed"]] or 0)
    for s in stalls:
        a[s] += int(r[col[s]] or 0)
print(f"total samples {tot}")
src_cache = {}
for key, a in sorted(agg.items(), key=lambda kv: -kv[1]["n"])[:top]:
    f, l = key
    path = os.path.join(root, "diffsheg_b200", "csrc", f)
    if path not in src_cache:
        src_cache[path] = open(path).read().splitlines() if os.path.exists(path) else []
    text = src_cache[path][l - 1].strip()[:100] if 0 < l <= len(src_cache[path]) else ""
    top3 = sorted(((a[s], s[6:]) for s in stalls), reverse=True)[:3]
    print(f"{100.0 * a['n'] / max(tot, 1):5.1f}%  {f}:{l:<4d} inst {a['inst']:>9d}  {', '.join(f'{s} {v}' for v, s in top3 if v)}  | {text}")
