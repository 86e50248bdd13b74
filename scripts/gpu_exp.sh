mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -s -k "(op_linear and bf16) or full_size or (denoise and bf16) or (ddim25 and bf16)" > gpurun_out/t17.log 2>&1; echo "t rc=$?" > gpurun_out/rc17.txt
timeout 600 python scripts/bench_gemm.py > gpurun_out/gemm_sweep17.log 2>&1
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench17.log 2>&1; echo "bench rc=$?" >> gpurun_out/rc17.txt
cat gpurun_out/rc17.txt; grep -E "passed|failed|rror" gpurun_out/t17.log | tail -3
cut -c1-150 gpurun_out/gemm_sweep17.log | head -7
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench17.log").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], "gemm", d["roofline"]["achieved"], d["roofline"]["ms_per_step"], "attn", d["roofline_attention"]["ms_per_step"], "row", d["rowwise"]["ms_per_step"])
PY
