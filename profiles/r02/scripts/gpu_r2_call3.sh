#!/bin/bash
# Round-2 GPU call 3: first hardware run of the persistent TMA-staged attention kernel (attn_tma.cuh)
mkdir -p gpurun_out
O=gpurun_out
export DSHEG_PROF_TABLE=1
DSHEG_RUN_UNVALIDATED=1 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "static_shift_numerators and tma" > $O/r3_tma_op.log 2>&1; echo "tma op rc=$?" > $O/r3_rc.txt
DSHEG_FIRST_RUN_BATCH=950 timeout 300 python scripts/first_hw_run.py --variant tma+expo > $O/r3_tma_inloop.log 2>&1; echo "tma in-loop rc=$?" >> $O/r3_rc.txt
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
DSHEG_ATTN=tma DSHEG_EXPO=1 timeout 300 $B > $O/r3_bench_attn_tma.json 2> $O/r3_bench_attn_tma.err
DSHEG_ATTN=tma DSHEG_EXPO=1 DSHEG_FUSE_LNMS=1 DSHEG_ATTN_AUD=1 timeout 300 $B > $O/r3_bench_all_tma.json 2> $O/r3_bench_all_tma.err
DSHEG_ATTN=tma DSHEG_EXPO=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_tma_kernel -s 17 -c 1 -o $O/r3_attn_tma python scripts/prof_denoise.py --batch 950 --calls 2 > $O/r3_ncu_attn_tma.log 2>&1
cat $O/r3_rc.txt; tail -4 $O/r3_tma_op.log; grep -E "^(PASS|FAIL)" $O/r3_tma_inloop.log | cut -c1-200; python scripts/gpu_round2_summary.py 2>/dev/null 
grep -h "attn\|attention" $O/r3_bench_attn_tma.err | head
