#!/bin/bash
# Round-2 GPU call 12: attn_ws v3 (8 A warps with 32 x 16 tiles, 24 warps at 80 registers, all Y tiles parked in tensor memory, LayerNorm by 16-column runs)
mkdir -p gpurun_out; O=gpurun_out
export DSHEG_PROF_TABLE=1
DSHEG_ATTN=ws timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider > $O/c12_pytest_ws.log 2>&1; echo "pytest ws rc=$?" > $O/c12_rc.txt
DSHEG_ATTN=ws timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-ref-cuda > $O/c12_bench_ws.json 2> $O/c12_bench_ws.err; echo "bench ws rc=$?" >> $O/c12_rc.txt
DSHEG_ATTN=ws timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_ws_kernel -s 17 -c 1 -o $O/c12_attn_ws python scripts/prof_denoise.py --batch 950 --calls 2 > $O/c12_ncu.log 2>&1
DSHEG_ATTN=ws timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/prof_denoise.py --batch 3 --calls 1 > $O/c12_memcheck_B3.log 2>&1; echo "memcheck rc=$?" >> $O/c12_rc.txt
tail -n 8 $O/c12_memcheck_B3.log > $O/c12_memcheck_B3.tail.log; rm -f $O/c12_memcheck_B3.log
cat $O/c12_rc.txt; tail -5 $O/c12_pytest_ws.log; grep "attention\|qkv\|sa_out" $O/c12_bench_ws.err | head -4
python - <<'PY'
import json
for f in ("gpurun_out/c12_bench_ws.json",):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["value"]), "frames/s", round(d["ms_per_step"], 1), "ms  attn", round(d["roofline_attention"]["achieved"]), "GB/s frac", round(d["roofline_attention"]["frac"], 3), d["clocks"], d.get("parity", {}).get("relmax"))
    except Exception as e:
        print(f, "failed", e)
PY
