mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -s -k "op_linear and bf16" > gpurun_out/t11_tc.log 2>&1; echo "tc rc=$?" > gpurun_out/rc11.txt
timeout 600 python scripts/bench_gemm.py > gpurun_out/gemm_sweep11.log 2>&1; echo "sweep rc=$?" >> gpurun_out/rc11.txt
timeout 900 python -m pytest tests -m gpu -q -s -k "(bf16 and not op_linear) or full_size" > gpurun_out/t11_bf16.log 2>&1; echo "bf16 rc=$?" >> gpurun_out/rc11.txt
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench11.log 2>&1; echo "bench rc=$?" >> gpurun_out/rc11.txt
DSHEG_TC_CG=1 timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench11_cg1.log 2>&1
cat gpurun_out/rc11.txt; grep -E "passed|failed|rror" gpurun_out/t11_tc.log | tail -4; cut -c1-150 gpurun_out/gemm_sweep11.log; grep -E "passed|failed|rror|differ" gpurun_out/t11_bf16.log | tail -4
python - <<'PY'
import json
for f in ("gpurun_out/bench11.log","gpurun_out/bench11_cg1.log"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["value"], d["ms_per_step"], "gemm", d["roofline"]["achieved"], d["roofline"]["ms_per_step"], "attn", d["roofline_attention"]["ms_per_step"], "row", d["rowwise"]["ms_per_step"])
    except Exception as e: print(f, "ERR", e, open(f).read()[-600:])
PY
