#!/bin/bash
# Round-2 GPU call 16: DSHEG_ZIGZAG=1 experiment (every large kernel starts with the rows its operand's producer wrote last: L2 reuse)
mkdir -p gpurun_out; O=gpurun_out
export DSHEG_PROF_TABLE=1
DSHEG_ZIGZAG=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "rows_are_independent or headline or denoise_matches or bisecting" > $O/c16_pytest_zz.log 2>&1; echo "pytest zigzag rc=$?" > $O/c16_rc.txt
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-ref-cuda > $O/c16_bench_default.json 2> $O/c16_bench_default.err; echo "bench default rc=$?" >> $O/c16_rc.txt
DSHEG_ZIGZAG=1 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-ref-cuda > $O/c16_bench_zz.json 2> $O/c16_bench_zz.err; echo "bench zigzag rc=$?" >> $O/c16_rc.txt
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-ref-cuda > $O/c16_bench_default2.json 2> $O/c16_bench_default2.err; echo "bench default (again) rc=$?" >> $O/c16_rc.txt
DSHEG_ZIGZAG=1 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-ref-cuda > $O/c16_bench_zz2.json 2> $O/c16_bench_zz2.err; echo "bench zigzag (again) rc=$?" >> $O/c16_rc.txt
cat $O/c16_rc.txt; tail -3 $O/c16_pytest_zz.log
for v in default zz; do echo "== $v"; grep "attention\|qkv \|sa_out\|ffn1\|ffn2\|ffn_out\|feat" $O/c16_bench_$v.err | head -9; done
python - <<'PY'
import json
for v in ("default", "zz", "default2", "zz2"):
    f = f"gpurun_out/c16_bench_{v}.json"
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["value"]), "frames/s", round(d["ms_per_step"], 1), "ms  gemm", round(d["roofline"]["achieved"]), "TF/s  attn", round(d["roofline_attention"]["achieved"]), "GB/s", d["clocks"])
    except Exception as e:
        print(f, "failed", e)
PY
