#!/bin/bash
# Round-2 GPU call 13: attn_ws v4 (mbarrier waits sleep with a suspend-time hint instead of spinning; ping-pong fragments in the A loop);
# experiment library with sleeping waits in the GEMM / attn_tma too (build_variants/libdsheg_sleep.so)
mkdir -p gpurun_out; O=gpurun_out
export DSHEG_PROF_TABLE=1
DSHEG_ATTN=ws timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "static_shift or cross_attention or bisecting or rows_are_independent or denoise_matches or headline or attention" > $O/c13_pytest_ws.log 2>&1; echo "pytest ws rc=$?" > $O/c13_rc.txt
DSHEG_ATTN=ws timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-ref-cuda > $O/c13_bench_ws.json 2> $O/c13_bench_ws.err; echo "bench ws rc=$?" >> $O/c13_rc.txt
DSHEG_LIB=$PWD/build_variants/libdsheg_sleep.so timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-ref-cuda > $O/c13_bench_tma_sleep.json 2> $O/c13_bench_tma_sleep.err; echo "bench tma+sleep rc=$?" >> $O/c13_rc.txt
DSHEG_LIB=$PWD/build_variants/libdsheg_sleep.so DSHEG_ATTN=ws timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-ref-cuda > $O/c13_bench_ws_sleep.json 2> $O/c13_bench_ws_sleep.err; echo "bench ws+sleep rc=$?" >> $O/c13_rc.txt
DSHEG_ATTN=ws timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_ws_kernel -s 17 -c 1 -o $O/c13_attn_ws python scripts/prof_denoise.py --batch 950 --calls 2 > $O/c13_ncu.log 2>&1
cat $O/c13_rc.txt; tail -3 $O/c13_pytest_ws.log
for v in ws tma_sleep ws_sleep; do echo "== $v"; grep "attention\|qkv \|sa_out\|ffn1\|ffn2\|ffn_out\|feat" $O/c13_bench_$v.err | head -9; done
python - <<'PY'
import json
for v in ("ws", "tma_sleep", "ws_sleep"):
    f = f"gpurun_out/c13_bench_{v}.json"
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["value"]), "frames/s", round(d["ms_per_step"], 1), "ms  gemm", round(d["roofline"]["achieved"]), "TF/s  attn", round(d["roofline_attention"]["achieved"]), "GB/s frac", round(d["roofline_attention"]["frac"], 3), d["clocks"], d.get("parity", {}).get("relmax"))
    except Exception as e:
        print(f, "failed", e)
PY
