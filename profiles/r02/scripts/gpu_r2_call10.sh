#!/bin/bash
# Round-2 GPU call 10: first hardware run of attn_ws (warp-specialised TMA attention, Y parked in tensor memory), DSHEG_ATTN=ws
mkdir -p gpurun_out; O=gpurun_out
export DSHEG_PROF_TABLE=1
DSHEG_ATTN=ws timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider > $O/c10_pytest_ws.log 2>&1; echo "pytest ws rc=$?" > $O/c10_rc.txt
DSHEG_ATTN=ws timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-ref-cuda > $O/c10_bench_ws.json 2> $O/c10_bench_ws.err; echo "bench ws rc=$?" >> $O/c10_rc.txt
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-ref-cuda > $O/c10_bench_tma.json 2> $O/c10_bench_tma.err; echo "bench tma rc=$?" >> $O/c10_rc.txt
DSHEG_ATTN=ws timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_ws_kernel -s 17 -c 1 -o $O/c10_attn_ws python scripts/prof_denoise.py --batch 950 --calls 2 > $O/c10_ncu.log 2>&1
DSHEG_ATTN=ws timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/prof_denoise.py --batch 3 --calls 1 > $O/c10_memcheck_B3.log 2>&1; echo "memcheck rc=$?" >> $O/c10_rc.txt
DSHEG_ATTN=ws timeout 400 compute-sanitizer --tool racecheck --error-exitcode 9 python scripts/prof_denoise.py --batch 2 --calls 1 > $O/c10_racecheck_B2.log 2>&1; echo "racecheck rc=$?" >> $O/c10_rc.txt
for f in $O/c10_memcheck_B3.log $O/c10_racecheck_B2.log; do tail -n 12 $f > ${f%.log}.tail.log; rm -f $f; done
cat $O/c10_rc.txt; tail -5 $O/c10_pytest_ws.log; grep "attention\|qkv\|sa_out" $O/c10_bench_ws.err $O/c10_bench_tma.err | head -8
python - <<'PY'
import json
for f in ("gpurun_out/c10_bench_ws.json", "gpurun_out/c10_bench_tma.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["value"]), "frames/s", round(d["ms_per_step"], 1), "ms  attn", round(d["roofline_attention"]["achieved"]), "GB/s frac", round(d["roofline_attention"]["frac"], 3), d["clocks"], d.get("parity", {}).get("relmax"))
    except Exception as e:
        print(f, "failed", e)
PY
