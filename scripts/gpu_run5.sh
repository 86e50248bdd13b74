mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -s -k "op_linear and bf16" > gpurun_out/t5_tc.log 2>&1; echo "tc rc=$?" > gpurun_out/rc5.txt
timeout 600 python scripts/bench_gemm.py > gpurun_out/gemm_sweep5.log 2>&1; echo "sweep rc=$?" >> gpurun_out/rc5.txt
timeout 900 python -m pytest tests -m gpu -q -s -k "bf16 and not op_linear" > gpurun_out/t5_bf16.log 2>&1; echo "bf16 rc=$?" >> gpurun_out/rc5.txt
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench5.log 2>&1; echo "bench rc=$?" >> gpurun_out/rc5.txt
cat gpurun_out/rc5.txt; grep -E "passed|failed|Error|error" gpurun_out/t5_tc.log | tail -5; cat gpurun_out/gemm_sweep5.log; grep -E "passed|failed" gpurun_out/t5_bf16.log | tail -3
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench5.log").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], "gemm", d["roofline"]["achieved"], d["roofline"]["ms_per_step"], "attn", d["roofline_attention"]["ms_per_step"], d["roofline_attention"]["achieved"], "row", d["rowwise"])
PY
