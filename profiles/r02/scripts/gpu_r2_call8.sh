#!/bin/bash
# Round-2 GPU call 8: tf32 mode v2 (CTA-pair tf32 GEMM, TF32 mma.sync attention): parity, sanitizers, speed
mkdir -p gpurun_out; O=gpurun_out
export DSHEG_PROF_TABLE=1
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "tf32" > $O/c8_tf32_tests.log 2>&1; echo "tf32 tests rc=$?" > $O/c8_rc.txt
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/prof_denoise.py --batch 3 --calls 1 --precision tf32 > $O/c8_memcheck_tf32_B3.log 2>&1; echo "memcheck tf32 B3 rc=$?" >> $O/c8_rc.txt
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/prof_denoise.py --batch 50 --calls 1 --precision tf32 > $O/c8_memcheck_tf32_B50.log 2>&1; echo "memcheck tf32 B50 rc=$?" >> $O/c8_rc.txt
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python scripts/prof_denoise.py --batch 2 --calls 1 --precision tf32 > $O/c8_racecheck_tf32_B2.log 2>&1; echo "racecheck tf32 B2 rc=$?" >> $O/c8_rc.txt
timeout 900 python bench.py --precision tf32 --steps 3 --no-ref-cuda --no-cpu-baseline > $O/c8_bench_tf32.json 2> $O/c8_bench_tf32.err; echo "bench tf32 rc=$?" >> $O/c8_rc.txt
timeout 900 python scripts/parity_report.py --modes tf32 > $O/c8_parity_tf32.json 2> $O/c8_parity_tf32.err; echo "parity rc=$?" >> $O/c8_rc.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32_kernel -s 30 -c 6 -o $O/c8_gemm_tf32 python scripts/prof_denoise.py --batch 950 --calls 1 --precision tf32 > $O/c8_prof_tf32.log 2>&1
cat $O/c8_rc.txt; tail -3 $O/c8_tf32_tests.log; grep "dsheg profile" $O/c8_bench_tf32.err | head -16
python - <<'PY'
import json
d = json.loads(open("gpurun_out/c8_bench_tf32.json").read().strip().splitlines()[-1])
print(round(d["value"]), "frames/s", round(d["ms_per_step"], 1), "ms", d.get("parity"))
PY
