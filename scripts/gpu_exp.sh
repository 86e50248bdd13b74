mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -s -k "edge_shapes or repaint_flags or long_form" > gpurun_out/t15.log 2>&1; echo "t rc=$?" > gpurun_out/rc15.txt
timeout 600 python scripts/bench_gemm.py > gpurun_out/gemm_sweep15.log 2>&1
DSHEG_TC_PREFETCH=2 timeout 600 python scripts/bench_gemm.py > gpurun_out/gemm_sweep15_skipw.log 2>&1
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench15.log 2>&1; echo "bench rc=$?" >> gpurun_out/rc15.txt
cat gpurun_out/rc15.txt; grep -E "passed|failed|rror" gpurun_out/t15.log | tail -3; grep "parity\]" gpurun_out/t15.log | cut -c1-150
echo "--- normal"; cut -c1-150 gpurun_out/gemm_sweep15.log | head -7; echo "--- skip W fill (timing model of resident W)"; cut -c1-150 gpurun_out/gemm_sweep15_skipw.log | head -7
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench15.log").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], "gemm", d["roofline"]["achieved"], d["roofline"]["ms_per_step"], "attn", d["roofline_attention"]["ms_per_step"], "row", d["rowwise"])
PY
