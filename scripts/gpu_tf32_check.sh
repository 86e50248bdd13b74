#!/bin/bash
# tf32 mode check (run under gpurun): its parity tests, the bench line and a full ncu capture of one layer's tf32 GEMMs
mkdir -p gpurun_out; O=gpurun_out
export DSHEG_PROF_TABLE=1
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "tf32" > $O/t32_tests.log 2>&1; echo "tf32 tests rc=$?" > $O/t32_rc.txt
timeout 600 python bench.py --precision tf32 --steps 2 --no-ref-cuda --no-cpu-baseline > $O/t32_bench.json 2> $O/t32_bench.err; echo "bench tf32 rc=$?" >> $O/t32_rc.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32_kernel -s 30 -c 7 -o $O/t32_gemm python scripts/prof_denoise.py --batch 950 --calls 1 --precision tf32 > $O/t32_ncu.log 2>&1
cat $O/t32_rc.txt; tail -3 $O/t32_tests.log; grep "dsheg profile" $O/t32_bench.err | head -16
python - <<'PY'
import json
d = json.loads(open("gpurun_out/t32_bench.json").read().strip().splitlines()[-1])
print(round(d["value"]), "frames/s", round(d["ms_per_step"], 1), "ms", d.get("parity", {}).get("relmax"))
PY
