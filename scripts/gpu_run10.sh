mkdir -p gpurun_out
timeout 600 python scripts/bench_gemm.py > gpurun_out/gemm_sweep10_pf1.log 2>&1
DSHEG_TC_PREFETCH=0 timeout 600 python scripts/bench_gemm.py > gpurun_out/gemm_sweep10_pf0.log 2>&1
timeout 600 python -m pytest tests -m gpu -q -s -k "(denoise and bf16) or (op_linear and bf16)" > gpurun_out/t10.log 2>&1; echo "t rc=$?" > gpurun_out/rc10.txt
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench10.log 2>&1; echo "bench rc=$?" >> gpurun_out/rc10.txt
timeout 1500 python scripts/bench_configs.py > gpurun_out/configs10.log 2>&1; echo "configs rc=$?" >> gpurun_out/rc10.txt
cat gpurun_out/rc10.txt; grep -E "passed|failed" gpurun_out/t10.log | tail -2
echo "--- prefetch on"; cut -c1-140 gpurun_out/gemm_sweep10_pf1.log; echo "--- prefetch off"; cut -c1-140 gpurun_out/gemm_sweep10_pf0.log
tail -8 gpurun_out/configs10.log
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench10.log").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], "gemm", d["roofline"]["achieved"], d["roofline"]["ms_per_step"], "attn", d["roofline_attention"]["ms_per_step"], d["roofline_attention"]["achieved"], "row", d["rowwise"])
PY
