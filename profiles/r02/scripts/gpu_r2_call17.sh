#!/bin/bash
# Round-2 GPU call 17: the default build with the alternating row walk on, against an experiment build with L2 eviction hints on the
# TMA loads (activations / residuals evict-first, weights evict-last): A / B / A / B
mkdir -p gpurun_out; O=gpurun_out
export DSHEG_PROF_TABLE=1
V=$PWD/build_variants/libdsheg_l2hints.so
DSHEG_LIB=$V timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "rows_are_independent or headline or denoise_matches" > $O/c17_pytest_hints.log 2>&1; echo "pytest hints rc=$?" > $O/c17_rc.txt
for r in 1 2; do
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-ref-cuda > $O/c17_bench_default$r.json 2> $O/c17_bench_default$r.err; echo "bench default $r rc=$?" >> $O/c17_rc.txt
DSHEG_LIB=$V timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-ref-cuda > $O/c17_bench_hints$r.json 2> $O/c17_bench_hints$r.err; echo "bench hints $r rc=$?" >> $O/c17_rc.txt
done
cat $O/c17_rc.txt; tail -3 $O/c17_pytest_hints.log
for v in default1 hints1; do echo "== $v"; grep "attention\|qkv \|sa_out\|ffn1\|ffn2\|ffn_out\|feat" $O/c17_bench_$v.err | head -9; done
python - <<'PY'
import json
for v in ("default1", "hints1", "default2", "hints2"):
    f = f"gpurun_out/c17_bench_{v}.json"
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["value"]), "frames/s", round(d["ms_per_step"], 1), "ms  gemm", round(d["roofline"]["achieved"]), "TF/s  attn", round(d["roofline_attention"]["achieved"]), "GB/s", d["clocks"])
    except Exception as e:
        print(f, "failed", e)
PY
