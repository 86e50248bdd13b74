#!/bin/bash
# Round-2 GPU call 4: attn_tma v2 (constant table by the producer warp, A^T in the K' slot, 10-deep ring, L2 prefetch of the next Q')
mkdir -p gpurun_out
O=gpurun_out
export DSHEG_PROF_TABLE=1
DSHEG_RUN_UNVALIDATED=1 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "static_shift_numerators and tma" > $O/r4_tma_op.log 2>&1; echo "tma op rc=$?" > $O/r4_rc.txt
DSHEG_FIRST_RUN_BATCH=950 timeout 300 python scripts/first_hw_run.py --variant tma+expo > $O/r4_tma_inloop.log 2>&1; echo "tma in-loop rc=$?" >> $O/r4_rc.txt
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
DSHEG_ATTN=tma DSHEG_EXPO=1 timeout 300 $B > $O/r4_bench_attn_tma.json 2> $O/r4_bench_attn_tma.err
DSHEG_ATTN=tma DSHEG_EXPO=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_tma_kernel -s 17 -c 1 -o $O/r4_attn_tma python scripts/prof_denoise.py --batch 950 --calls 2 > $O/r4_ncu_attn_tma.log 2>&1
DSHEG_RUN_UNVALIDATED=1 timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider -x > $O/r4_pytest_gpu.log 2>&1; echo "pytest gpu rc=$?" >> $O/r4_rc.txt
cat $O/r4_rc.txt; tail -4 $O/r4_tma_op.log; grep -E "^(PASS|FAIL)" $O/r4_tma_inloop.log | cut -c1-200; python scripts/gpu_round2_summary.py 2>/dev/null | grep r4_; tail -5 $O/r4_pytest_gpu.log
