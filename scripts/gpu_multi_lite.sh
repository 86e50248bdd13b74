#!/bin/bash
# N-GPU check of the final build on a tight budget (run under `gpurun --gpus N`): sharded == single-GPU parity test + the config-2 scaling bench
N=${1:-2}
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -s > $O/multi_pytest_n$N.log 2>&1; echo "multi pytest rc=$?" > $O/multi_rc_n$N.txt
timeout 600 $TR --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline --no-ref-cuda > $O/scale_config2_n$N.json 2> $O/scale_config2_n$N.err; echo "config 2 rc=$?" >> $O/multi_rc_n$N.txt
cat $O/multi_rc_n$N.txt; tail -2 $O/multi_pytest_n$N.log
python - <<PY
import json
d = json.loads(open("gpurun_out/scale_config2_n$N.json").read().strip().splitlines()[-1])
print("config 2 N=$N", round(d["value"]), "frames/s", round(d["ms_per_step"], 1), "ms  e2e", round(d["e2e"]["value"]), d["scaling"], d["config"]["per_gpu_batch"], d["clocks"])
PY
