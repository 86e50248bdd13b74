#!/bin/bash
# Round-2 GPU call 7: attn_tma v3 (A^T as register-resident B fragments, no pinned ring slots, balanced Y product)
mkdir -p gpurun_out; O=gpurun_out
export DSHEG_PROF_TABLE=1
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "static_shift or cross_attention or bisecting or rows_are_independent or denoise_matches or headline" > $O/c7_attn_tests.log 2>&1; echo "attn tests rc=$?" > $O/c7_rc.txt
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-ref-cuda > $O/c7_bench.json 2> $O/c7_bench.err; echo "bench rc=$?" >> $O/c7_rc.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_tma_kernel -s 17 -c 1 -o $O/c7_attn_tma python scripts/prof_denoise.py --batch 950 --calls 2 > $O/c7_ncu.log 2>&1
timeout 300 compute-sanitizer --tool racecheck --error-exitcode 9 python scripts/prof_denoise.py --batch 2 --calls 1 > $O/c7_racecheck_B2.log 2>&1; echo "racecheck rc=$?" >> $O/c7_rc.txt
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/prof_denoise.py --batch 3 --calls 1 > $O/c7_memcheck_B3.log 2>&1; echo "memcheck rc=$?" >> $O/c7_rc.txt
cat $O/c7_rc.txt; tail -3 $O/c7_attn_tests.log; python scripts/gpu_round2_summary.py 2>/dev/null | grep c7_; grep "attention\|qkv" $O/c7_bench.err | head -4
