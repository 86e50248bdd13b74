#!/bin/bash
# Round-2 GPU call 11: attn_ws v2 (A warps a whole sample ahead: thread-local column sums, A^T parked in tensor memory, two phases),
# tf32 GEMM with 8 epilogue warps and two CTAs per SM
mkdir -p gpurun_out; O=gpurun_out
export DSHEG_PROF_TABLE=1
DSHEG_ATTN=ws timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider > $O/c11_pytest_ws.log 2>&1; echo "pytest ws rc=$?" > $O/c11_rc.txt
DSHEG_ATTN=ws timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-ref-cuda > $O/c11_bench_ws.json 2> $O/c11_bench_ws.err; echo "bench ws rc=$?" >> $O/c11_rc.txt
DSHEG_ATTN=ws timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_ws_kernel -s 17 -c 1 -o $O/c11_attn_ws python scripts/prof_denoise.py --batch 950 --calls 2 > $O/c11_ncu.log 2>&1
timeout 600 python bench.py --precision tf32 --steps 2 --no-ref-cuda --no-cpu-baseline > $O/c11_bench_tf32.json 2> $O/c11_bench_tf32.err; echo "bench tf32 rc=$?" >> $O/c11_rc.txt
DSHEG_ATTN=ws timeout 400 compute-sanitizer --tool racecheck --racecheck-report all --error-exitcode 9 python scripts/prof_denoise.py --batch 2 --calls 1 > $O/c11_racecheck_B2.log 2>&1; echo "racecheck rc=$?" >> $O/c11_rc.txt
DSHEG_ATTN=ws timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/prof_denoise.py --batch 3 --calls 1 > $O/c11_memcheck_B3.log 2>&1; echo "memcheck rc=$?" >> $O/c11_rc.txt
head -c 20000 $O/c11_racecheck_B2.log > $O/c11_racecheck_B2.head.log; tail -n 12 $O/c11_racecheck_B2.log > $O/c11_racecheck_B2.tail.log; rm -f $O/c11_racecheck_B2.log
tail -n 8 $O/c11_memcheck_B3.log > $O/c11_memcheck_B3.tail.log; rm -f $O/c11_memcheck_B3.log
cat $O/c11_rc.txt; tail -5 $O/c11_pytest_ws.log; grep "attention\|qkv\|sa_out" $O/c11_bench_ws.err | head -4; grep "dsheg profile" $O/c11_bench_tf32.err | head -16
python - <<'PY'
import json
for f in ("gpurun_out/c11_bench_ws.json", "gpurun_out/c11_bench_tf32.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["value"]), "frames/s", round(d["ms_per_step"], 1), "ms  attn", round(d["roofline_attention"]["achieved"]), "GB/s frac", round(d["roofline_attention"]["frac"], 3), d["clocks"], d.get("parity", {}).get("relmax"))
    except Exception as e:
        print(f, "failed", e)
PY
