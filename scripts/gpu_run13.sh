mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -s > gpurun_out/t13_all.log 2>&1; echo "all rc=$?" > gpurun_out/rc13.txt
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke13.log 2>&1; echo "smoke rc=$?" >> gpurun_out/rc13.txt
timeout 900 python scripts/bench_configs.py 1 4 > gpurun_out/configs13.log 2>&1; echo "configs rc=$?" >> gpurun_out/rc13.txt
DSHEG_GRAPHS=0 timeout 900 python scripts/bench_configs.py 1 > gpurun_out/configs13_nograph.log 2>&1
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench13.log 2>&1; echo "bench rc=$?" >> gpurun_out/rc13.txt
cat gpurun_out/rc13.txt; grep -E "passed|failed|rror" gpurun_out/t13_all.log | tail -5; tail -3 gpurun_out/smoke13.log; cat gpurun_out/configs13.log | cut -c1-220; echo nograph; cat gpurun_out/configs13_nograph.log | cut -c1-200
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench13.log").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], "gemm", d["roofline"]["achieved"], d["roofline"]["ms_per_step"], "attn", d["roofline_attention"]["ms_per_step"], "row", d["rowwise"]["ms_per_step"])
PY
