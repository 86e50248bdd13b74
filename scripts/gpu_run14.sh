mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -s -k "attention or long_form or (denoise and bf16) or harmonize" > gpurun_out/t14.log 2>&1; echo "t rc=$?" > gpurun_out/rc14.txt
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench14.log 2>&1; echo "bench rc=$?" >> gpurun_out/rc14.txt
cat gpurun_out/rc14.txt; grep -E "passed|failed|rror" gpurun_out/t14.log | tail -4; grep -E "long-form|attention bf16 tensor-core Bn3" gpurun_out/t14.log
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench14.log").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], "gemm", d["roofline"]["achieved"], d["roofline"]["ms_per_step"], "attn", d["roofline_attention"]["ms_per_step"], d["roofline_attention"]["achieved"], "row", d["rowwise"]["ms_per_step"])
PY
