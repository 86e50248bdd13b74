#!/bin/bash
# Last hardware call of the round: the every-lane-arrives form of attn_ws as the default (racecheck-clean) against the single-arrive build
mkdir -p gpurun_out; O=gpurun_out
export DSHEG_PROF_TABLE=1
timeout 100 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "static_shift or cross_attention or rows_are_independent or headline" > $O/last_pytest.log 2>&1; echo "pytest rc=$?" > $O/last_rc.txt
timeout 100 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-ref-cuda > $O/last_bench_all_lanes.json 2> $O/last_bench_all_lanes.err; echo "bench all-lanes rc=$?" >> $O/last_rc.txt
DSHEG_LIB=$PWD/build_variants/libdsheg_lane0_arrive.so timeout 100 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-ref-cuda > $O/last_bench_lane0.json 2> $O/last_bench_lane0.err; echo "bench lane0 rc=$?" >> $O/last_rc.txt
cat $O/last_rc.txt; tail -2 $O/last_pytest.log; grep "attention" $O/last_bench_all_lanes.err $O/last_bench_lane0.err
python - <<'PY'
import json
for v in ("all_lanes", "lane0"):
    try:
        d = json.loads(open(f"gpurun_out/last_bench_{v}.json").read().strip().splitlines()[-1])
        print(v, round(d["value"]), "frames/s", round(d["ms_per_step"], 1), "ms attn", round(d["roofline_attention"]["achieved"]), "GB/s", d["clocks"]["sm_mhz"])
    except Exception as e:
        print(v, "failed", e)
PY
