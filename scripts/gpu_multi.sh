#!/bin/bash
# N-GPU validation (run under `gpurun --gpus N`): sharded == single-GPU parity test, then the scaling benches via torchrun:
# config 2 (weak: 950 per GPU) and config 5 (strong: global 4096 split over the GPUs); at N = 8 also config 4 (8 long clips, one per GPU).
N=${1:-2}
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -s > $O/multi_pytest_n$N.log 2>&1; echo "multi pytest rc=$?" > $O/multi_rc_n$N.txt
timeout 900 $TR --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 > $O/scale_config2_n$N.json 2> $O/scale_config2_n$N.err; echo "config 2 rc=$?" >> $O/multi_rc_n$N.txt
timeout 900 $TR --master-port 29512 bench.py --gpus $N --config 5 --steps 3 --warmup 3 > $O/scale_config5_n$N.json 2> $O/scale_config5_n$N.err; echo "config 5 rc=$?" >> $O/multi_rc_n$N.txt
if [ "$N" = "8" ]; then
  timeout 900 $TR --master-port 29513 bench.py --gpus $N --config 4 --steps 3 --warmup 3 > $O/scale_config4_n$N.json 2> $O/scale_config4_n$N.err; echo "config 4 rc=$?" >> $O/multi_rc_n$N.txt
fi
timeout 600 $TR --master-port 29514 bench.py --impl reference --gpus $N --steps 1 --warmup 0 > $O/scale_ref_n$N.json 2> $O/scale_ref_n$N.err; echo "ref rc=$?" >> $O/multi_rc_n$N.txt
cat $O/multi_rc_n$N.txt; tail -2 $O/multi_pytest_n$N.log
python - <<PY
import json
for c in ("2", "5", "4"):
    try:
        d = json.loads(open("gpurun_out/scale_config%s_n$N.json" % c).read().strip().splitlines()[-1])
        print("config", c, "N=$N", round(d["value"]), "frames/s", round(d["ms_per_step"], 1), "ms  e2e", round(d["e2e"]["value"]), d["scaling"], d["config"]["per_gpu_batch"], d["clocks"])
    except Exception as e:
        print("config", c, "failed:", e)
PY
