#!/bin/bash
# Round-2 GPU call 2: ncu --set full of the attention candidates at the headline batch (why do they all plateau near 2.2 TB/s?),
# the post-processing kernels with verbose output, the parity report, smoke(), then the round-2 "second" script.
mkdir -p gpurun_out
O=gpurun_out
P="python scripts/prof_denoise.py --batch 950 --calls 2"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_v3_kernel -s 17 -c 1 -o $O/r2_attn_v3 $P > $O/r2_ncu_attn_v3.log 2>&1
DSHEG_ATTN=v5c4 DSHEG_EXPO=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_v5_kernel -s 17 -c 1 -o $O/r2_attn_v5c4_expo $P > $O/r2_ncu_attn_v5c4.log 2>&1
DSHEG_ATTN=v6 DSHEG_EXPO=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_v6_kernel -s 17 -c 1 -o $O/r2_attn_v6_expo $P > $O/r2_ncu_attn_v6.log 2>&1
DSHEG_ATTN=v5c1 DSHEG_EXPO=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_v5_kernel -s 17 -c 1 -o $O/r2_attn_v5c1_expo $P > $O/r2_ncu_attn_v5c1.log 2>&1
DSHEG_RUN_UNVALIDATED=1 timeout 300 python -m pytest tests/test_postprocess.py -m gpu -q -s -p no:cacheprovider > $O/r2_postprocess_tests.log 2>&1; echo "postprocess rc=$?" > $O/r2c_rc.txt
timeout 900 python scripts/parity_report.py > $O/r2_parity_report.json 2> $O/r2_parity_report.err; echo "parity_report rc=$?" >> $O/r2c_rc.txt
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2_smoke.log 2>&1; echo "smoke rc=$?" >> $O/r2c_rc.txt
bash scripts/gpu_round2_second.sh > $O/r2_second.log 2>&1
cat $O/r2c_rc.txt; tail -5 $O/r2_postprocess_tests.log; tail -3 $O/r2_smoke.log; tail -60 $O/r2_second.log
