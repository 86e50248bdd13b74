mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -s > gpurun_out/t7_multi.log 2>&1; echo "multi rc=$?" > gpurun_out/rc7.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/bench7_n2.log 2>&1; echo "bench2 rc=$?" >> gpurun_out/rc7.txt
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench7_n1.log 2>&1; echo "bench1 rc=$?" >> gpurun_out/rc7.txt
timeout 900 python bench.py --ref-cuda 950 > gpurun_out/refcuda.log 2>&1; echo "refcuda rc=$?" >> gpurun_out/rc7.txt
cat gpurun_out/rc7.txt; tail -3 gpurun_out/t7_multi.log; tail -1 gpurun_out/refcuda.log
python - <<'PY'
import json
for f in ("gpurun_out/bench7_n1.log","gpurun_out/bench7_n2.log"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["n_gpus"], d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "gemm", d["roofline"]["achieved"], d["roofline"]["ms_per_step"], "attn", d["roofline_attention"]["ms_per_step"], d["roofline_attention"]["achieved"], "row", d["rowwise"]["ms_per_step"])
    except Exception as e: print(f, "ERR", e, open(f).read()[-800:])
PY
