#!/bin/bash
# FIRST gpurun call of round 2: everything written after round 1's GPU budget ran out, in one box session.
#   1. full racecheck log of the CTA-pair GEMMs (B=24), see profiles/r01/NOTES_next_round.md "Open: racecheck"
#   2. the gated tests (postprocess kernels, attention v4) with DSHEG_RUN_UNVALIDATED=1
#   3. attention v3 vs v4 inside the real loop (bench.py, same box, back to back) + memcheck/racecheck of v4
#   4. post-processing bandwidth
mkdir -p gpurun_out
timeout 300 compute-sanitizer --tool racecheck python scripts/prof_denoise.py --batch 24 --calls 1 > gpurun_out/r2_racecheck_pairs_B24.log 2>&1
echo "racecheck pairs rc=$?" > gpurun_out/r2_rc.txt
DSHEG_RUN_UNVALIDATED=1 timeout 600 python -m pytest tests/test_postprocess.py tests/test_gpu_parity.py -m gpu -q -s -k "postprocess or gpu_inv or gpu_axis or op_attention_bf16" > gpurun_out/r2_unvalidated_tests.log 2>&1
echo "unvalidated tests rc=$?" >> gpurun_out/r2_rc.txt
DSHEG_ATTN=v4 timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/prof_denoise.py --batch 3 --calls 1 > gpurun_out/r2_v4_memcheck.log 2>&1
echo "v4 memcheck rc=$?" >> gpurun_out/r2_rc.txt
DSHEG_ATTN=v4 timeout 300 compute-sanitizer --tool racecheck --error-exitcode 9 python scripts/prof_denoise.py --batch 2 --calls 1 > gpurun_out/r2_v4_racecheck.log 2>&1
echo "v4 racecheck rc=$?" >> gpurun_out/r2_rc.txt
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/r2_bench_attn_v3.json 2> gpurun_out/r2_bench_attn_v3.err
DSHEG_ATTN=v4 timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/r2_bench_attn_v4.json 2> gpurun_out/r2_bench_attn_v4.err
timeout 300 python scripts/bench_postprocess.py > gpurun_out/r2_postprocess_bw.txt 2>&1
cat gpurun_out/r2_rc.txt
grep -c "Race reported\|hazard" gpurun_out/r2_racecheck_pairs_B24.log; grep -E "Write access|Read access" gpurun_out/r2_racecheck_pairs_B24.log | sed 's/(CUtensorMap.*//' | sort | uniq -c | sort -rn | head -8
tail -3 gpurun_out/r2_unvalidated_tests.log
python - <<'PY'
import json
for v in ("v3", "v4"):
    try:
        d = json.loads(open(f"gpurun_out/r2_bench_attn_{v}.json").read().strip().splitlines()[-1])
        print(v, round(d["value"]), "frames/s; attention", d.get("roofline_attention", {}).get("achieved"), "GB/s")
    except Exception as e:  # noqa: BLE001
        print(v, "failed:", e)
PY
cat gpurun_out/r2_postprocess_bw.txt
