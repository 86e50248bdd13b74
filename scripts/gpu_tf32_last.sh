#!/bin/bash
# One short run for the persistent tf32 GEMM: bench line (incl. its live parity against the fp32 oracle), then the tf32 tests
mkdir -p gpurun_out; O=gpurun_out
export DSHEG_PROF_TABLE=1
timeout 70 python bench.py --precision tf32 --steps 1 --warmup 1 --no-ref-cuda --no-cpu-baseline > $O/t32p_bench.json 2> $O/t32p_bench.err; echo "bench tf32 rc=$?" > $O/t32p_rc.txt
timeout 40 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "tf32" > $O/t32p_tests.log 2>&1; echo "tf32 tests rc=$?" >> $O/t32p_rc.txt
cat $O/t32p_rc.txt; tail -2 $O/t32p_tests.log; grep "dsheg profile" $O/t32p_bench.err | head -14
python - <<'PY'
import json
d = json.loads(open("gpurun_out/t32p_bench.json").read().strip().splitlines()[-1])
print(round(d["value"]), "frames/s", round(d["ms_per_step"], 1), "ms", d.get("parity"))
PY
