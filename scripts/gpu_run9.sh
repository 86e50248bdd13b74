mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -s -k "bf16 and not op_linear" > gpurun_out/t9_bf16.log 2>&1; echo "bf16 rc=$?" > gpurun_out/rc9.txt
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench9.log 2>&1; echo "bench rc=$?" >> gpurun_out/rc9.txt
timeout 1200 python scripts/bench_configs.py > gpurun_out/configs9.log 2>&1; echo "configs rc=$?" >> gpurun_out/rc9.txt
cat gpurun_out/rc9.txt; grep -E "passed|failed|rror" gpurun_out/t9_bf16.log | tail -3; cat gpurun_out/configs9.log | tail -8
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench9.log").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], "gemm", d["roofline"]["achieved"], d["roofline"]["ms_per_step"], "attn", d["roofline_attention"]["ms_per_step"], d["roofline_attention"]["achieved"], "row", d["rowwise"])
PY
