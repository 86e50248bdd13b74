mkdir -p gpurun_out
DSHEG_TC_PREFETCH=1 timeout 600 python scripts/bench_gemm.py > gpurun_out/gemm_sweep16_pf.log 2>&1
DSHEG_TC_PREFETCH=1 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench16_pf.log 2>&1
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench16.log 2>&1
bash scripts/gpu_profile.sh > gpurun_out/profile16.log 2>&1
echo "--- pair prefetch"; cut -c1-150 gpurun_out/gemm_sweep16_pf.log | head -7
python - <<'PY'
import json
for f in ("gpurun_out/bench16_pf.log","gpurun_out/bench16.log"):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, d["value"], d["ms_per_step"], "gemm", d["roofline"]["achieved"], d["roofline"]["ms_per_step"], "attn", d["roofline_attention"]["ms_per_step"], "row", d["rowwise"]["ms_per_step"])
PY
tail -4 gpurun_out/profile16.log
