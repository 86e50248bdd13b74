#!/bin/bash
# Round-2 GPU call 9 (first call of the resumed session): the whole GPU suite on the restored tree, the pending hardware run of the
# tf32 mode v2 (CTA-pair tf32 GEMM + TF32 mma.sync attention), the attention-redesign micro-benchmarks, one full ncu capture of attn_tma.
mkdir -p gpurun_out; O=gpurun_out
export DSHEG_PROF_TABLE=1
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > $O/c9_pytest_gpu.log 2>&1; echo "pytest rc=$?" > $O/c9_rc.txt
timeout 120 ./build_variants/ubench_attn_parts > $O/c9_ubench.txt 2>&1; echo "ubench rc=$?" >> $O/c9_rc.txt
timeout 600 python bench.py --precision tf32 --steps 3 --no-ref-cuda --no-cpu-baseline > $O/c9_bench_tf32.json 2> $O/c9_bench_tf32.err; echo "bench tf32 rc=$?" >> $O/c9_rc.txt
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/prof_denoise.py --batch 3 --calls 1 --precision tf32 > $O/c9_memcheck_tf32_B3.log 2>&1; echo "memcheck tf32 B3 rc=$?" >> $O/c9_rc.txt
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/prof_denoise.py --batch 50 --calls 1 --precision tf32 > $O/c9_memcheck_tf32_B50.log 2>&1; echo "memcheck tf32 B50 rc=$?" >> $O/c9_rc.txt
timeout 600 python scripts/parity_report.py --modes tf32 > $O/c9_parity_tf32.json 2> $O/c9_parity_tf32.err; echo "parity rc=$?" >> $O/c9_rc.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_tma_kernel -s 17 -c 1 -o $O/c9_attn_tma python scripts/prof_denoise.py --batch 950 --calls 2 > $O/c9_ncu_attn.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32_kernel -s 30 -c 4 -o $O/c9_gemm_tf32 python scripts/prof_denoise.py --batch 950 --calls 1 --precision tf32 > $O/c9_ncu_tf32.log 2>&1
for f in $O/c9_memcheck_*.log; do tail -n 5 $f > ${f%.log}.tail.log; rm -f $f; done
cat $O/c9_rc.txt; tail -3 $O/c9_pytest_gpu.log; cat $O/c9_ubench.txt; grep "dsheg profile" $O/c9_bench_tf32.err | head -20
python - <<'PY'
import json
d = json.loads(open("gpurun_out/c9_bench_tf32.json").read().strip().splitlines()[-1])
print(round(d["value"]), "frames/s", round(d["ms_per_step"], 1), "ms", d.get("parity"))
PY
