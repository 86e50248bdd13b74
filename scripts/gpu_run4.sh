mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -s -k "bf16 and not op_linear" > gpurun_out/t4_bf16.log 2>&1; echo "bf16 rc=$?" > gpurun_out/rc4.txt
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench4.log 2>&1; echo "bench rc=$?" >> gpurun_out/rc4.txt
# each bench_gemm call = 2 warm-up + 1 timed launch: capture the timed one of each shape
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -o gpurun_out/gemm_v2 python scripts/bench_gemm_few.py > gpurun_out/prof4.log 2>&1
cat gpurun_out/rc4.txt; tail -3 gpurun_out/t4_bf16.log; tail -3 gpurun_out/prof4.log
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench4.log").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], "gemm", d["roofline"]["achieved"], d["roofline"]["ms_per_step"], "attn", d["roofline_attention"]["ms_per_step"], d["roofline_attention"]["achieved"], "row", d["rowwise"])
PY
