#!/bin/bash
# What the driver runs at round end, on the committed tree: GPU tests, smoke, the default bench line (no environment switches)
mkdir -p gpurun_out; O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > $O/final_pytest_gpu.log 2>&1; echo "pytest rc=$?" > $O/final_rc.txt
timeout 300 python __graft_entry__.py smoke > $O/final_smoke.log 2>&1; echo "smoke rc=$?" >> $O/final_rc.txt
timeout 600 python bench.py > $O/final_bench.json 2> $O/final_bench.err; echo "bench rc=$?" >> $O/final_rc.txt
cat $O/final_rc.txt; tail -2 $O/final_pytest_gpu.log; tail -1 $O/final_smoke.log
python - <<'PY'
import json
d = json.loads(open("gpurun_out/final_bench.json").read().strip().splitlines()[-1])
print(round(d["value"]), "frames/s", round(d["ms_per_step"], 1), "ms e2e", round(d["e2e"]["value"]), "gemm", round(d["roofline"]["achieved"]), d["roofline"]["frac"], d["roofline"]["traffic"], "attn", round(d["roofline_attention"]["achieved"]), d["roofline_attention"]["frac"], d["roofline_attention"]["traffic"], d["clocks"], d["gpu_launches"])
PY
