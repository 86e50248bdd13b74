#!/bin/bash
# Round-2 GPU call 15: attn_ws v6 (K, V half tiles in the ring, the last A warp to finish a half refills it; warp-uniform direct-write decision)
mkdir -p gpurun_out; O=gpurun_out
export DSHEG_PROF_TABLE=1
DSHEG_ATTN=ws timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "static_shift or cross_attention or bisecting or rows_are_independent or denoise_matches or headline or attention" > $O/c15_pytest_ws.log 2>&1; echo "pytest ws rc=$?" > $O/c15_rc.txt
DSHEG_ATTN=ws timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-ref-cuda > $O/c15_bench_ws.json 2> $O/c15_bench_ws.err; echo "bench ws rc=$?" >> $O/c15_rc.txt
DSHEG_ATTN=ws timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_ws_kernel -s 17 -c 1 -o $O/c15_attn_ws python scripts/prof_denoise.py --batch 950 --calls 2 > $O/c15_ncu.log 2>&1
DSHEG_ATTN=ws timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/prof_denoise.py --batch 3 --calls 1 > $O/c15_memcheck_B3.log 2>&1; echo "memcheck rc=$?" >> $O/c15_rc.txt
tail -n 8 $O/c15_memcheck_B3.log > $O/c15_memcheck_B3.tail.log; rm -f $O/c15_memcheck_B3.log
cat $O/c15_rc.txt; tail -3 $O/c15_pytest_ws.log
for v in ws; do echo "== $v"; grep "attention\|qkv \|sa_out\|ffn1\|ffn2\|ffn_out\|feat" $O/c15_bench_$v.err | head -9; done
python - <<'PY'
import json
for v in ("ws",):
    f = f"gpurun_out/c15_bench_{v}.json"
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["value"]), "frames/s", round(d["ms_per_step"], 1), "ms  gemm", round(d["roofline"]["achieved"]), "TF/s  attn", round(d["roofline_attention"]["achieved"]), "GB/s frac", round(d["roofline_attention"]["frac"], 3), d["clocks"], d.get("parity", {}).get("relmax"))
    except Exception as e:
        print(f, "failed", e)
PY
