mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -s -k "bf16 and not op_linear" > gpurun_out/t6_bf16.log 2>&1; echo "bf16 rc=$?" > gpurun_out/rc6.txt
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench6.log 2>&1; echo "bench rc=$?" >> gpurun_out/rc6.txt
cat gpurun_out/rc6.txt; grep -E "passed|failed|rror" gpurun_out/t6_bf16.log | tail -5; grep "attention bf16" gpurun_out/t6_bf16.log | head -3
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench6.log").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], "gemm", d["roofline"]["achieved"], d["roofline"]["ms_per_step"], "attn", d["roofline_attention"]["ms_per_step"], d["roofline_attention"]["achieved"], "row", d["rowwise"])
PY
