#!/bin/bash
# N-GPU validation (run under `gpurun --gpus N`): sharded == single-GPU parity test, then the weak-scaling bench via torchrun
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -s > gpurun_out/multi_pytest.log 2>&1; echo "multi rc=$?" > gpurun_out/multi_rc.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/multi_bench_n$N.json 2> gpurun_out/multi_bench_n$N.err; echo "bench rc=$?" >> gpurun_out/multi_rc.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 1 --warmup 0 > gpurun_out/multi_bench_ref_n$N.json 2> gpurun_out/multi_bench_ref_n$N.err; echo "ref rc=$?" >> gpurun_out/multi_rc.txt
cat gpurun_out/multi_rc.txt; tail -2 gpurun_out/multi_pytest.log; tail -c 400 gpurun_out/multi_bench_n$N.json; echo; cut -c1-200 gpurun_out/multi_bench_ref_n$N.json
