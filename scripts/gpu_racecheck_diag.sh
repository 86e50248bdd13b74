#!/bin/bash
# Diagnostic (not shipped) build in which EVERY lane arrives on the A^T hand-over mbarriers of attn_ws: does racecheck then see the edges?
mkdir -p gpurun_out; O=gpurun_out
DSHEG_LIB=$PWD/build_variants/libdsheg_all_lanes_arrive.so timeout 200 compute-sanitizer --tool racecheck --error-exitcode 9 python scripts/prof_denoise.py --batch 2 --calls 1 > $O/diag_racecheck_all_lanes.log 2>&1; echo "racecheck (every lane arrives) rc=$?" > $O/diag_rc.txt
grep "Race reported\|hazard detected" $O/diag_racecheck_all_lanes.log | sed 's/+0x[0-9a-f]*//g; s/(CUtensorMap_st[^)]*)//g; s/0x[0-9a-f]* in block ([0-9,]*)//' | sort | uniq -c | sort -rn | head -12 > $O/diag_racecheck_all_lanes.kinds.txt
head -c 3000 $O/diag_racecheck_all_lanes.log > $O/diag_racecheck_all_lanes.head.log; tail -n 6 $O/diag_racecheck_all_lanes.log > $O/diag_racecheck_all_lanes.tail.log; rm -f $O/diag_racecheck_all_lanes.log
cat $O/diag_rc.txt; cat $O/diag_racecheck_all_lanes.kinds.txt; cat $O/diag_racecheck_all_lanes.tail.log
