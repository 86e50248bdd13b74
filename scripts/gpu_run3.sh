mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -s -k "op_linear and bf16" > gpurun_out/t3_tc.log 2>&1; echo "tc rc=$?" > gpurun_out/rc3.txt
timeout 600 python scripts/bench_gemm.py > gpurun_out/gemm_sweep.log 2>&1; echo "sweep rc=$?" >> gpurun_out/rc3.txt
timeout 900 python -m pytest tests -m gpu -q -s -k "bf16 and not op_linear" > gpurun_out/t3_bf16.log 2>&1; echo "bf16 rc=$?" >> gpurun_out/rc3.txt
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench3.log 2>&1; echo "bench rc=$?" >> gpurun_out/rc3.txt
DSHEG_TC_BN=128 timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench3_bn128.log 2>&1
cat gpurun_out/rc3.txt; grep -c parity gpurun_out/t3_tc.log; tail -3 gpurun_out/t3_tc.log; cat gpurun_out/gemm_sweep.log; tail -3 gpurun_out/t3_bf16.log
python - <<'PY'
import json
for f in ("gpurun_out/bench3.log","gpurun_out/bench3_bn128.log"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["value"], d["ms_per_step"], d["roofline"]["achieved"], d["roofline"]["ms_per_step"], d["roofline_attention"]["ms_per_step"], d["rowwise"]["ms_per_step"])
    except Exception as e: print(f, "ERR", e)
PY
