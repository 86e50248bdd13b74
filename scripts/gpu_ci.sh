#!/bin/bash
# What the driver runs at round end, plus the reference arm and the evidence files of profiles/: GPU tests, smoke, both bench arms,
# the other BASELINE configurations, the parity report, sanitizers, ncu launch list + full captures.  Run under gpurun (1 GPU).
mkdir -p gpurun_out
O=gpurun_out
export DSHEG_PROF_TABLE=1
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > $O/ci_pytest_gpu.log 2>&1; echo "pytest rc=$?" > $O/ci_rc.txt
timeout 300 python __graft_entry__.py smoke > $O/ci_smoke.log 2>&1; echo "smoke rc=$?" >> $O/ci_rc.txt
timeout 900 python bench.py > $O/ci_bench.json 2> $O/ci_bench.err; echo "bench rc=$?" >> $O/ci_rc.txt
timeout 600 python bench.py --impl reference > $O/ci_bench_reference.json 2> $O/ci_bench_reference.err; echo "ref rc=$?" >> $O/ci_rc.txt
timeout 900 python bench.py --precision tf32 --no-ref-cuda --no-cpu-baseline > $O/ci_bench_tf32.json 2> $O/ci_bench_tf32.err; echo "bench tf32 rc=$?" >> $O/ci_rc.txt
timeout 900 python scripts/parity_report.py --modes fp32,bf16,tf32 > $O/ci_parity_report.json 2> $O/ci_parity_report.err; echo "parity_report rc=$?" >> $O/ci_rc.txt
: > $O/ci_configs.jsonl
for c in 1a 1b 1c 4 5; do timeout 600 python bench.py --config $c --steps 3 --no-cpu-baseline >> $O/ci_configs.jsonl 2>> $O/ci_configs.err; done
timeout 900 python bench.py --config 3 --steps 1 --no-cpu-baseline >> $O/ci_configs.jsonl 2>> $O/ci_configs.err
echo "configs rc=$? lines=$(wc -l < $O/ci_configs.jsonl)" >> $O/ci_rc.txt
# sanitizers: the default bf16 path (single-CTA GEMM variants at B = 3 / 2; CTA pairs + ACT_LNMS at B = 50) and the tf32 engine
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/prof_denoise.py --batch 3 --calls 1 > $O/ci_memcheck_B3.log 2>&1; echo "memcheck B3 rc=$?" >> $O/ci_rc.txt
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python scripts/prof_denoise.py --batch 2 --calls 1 > $O/ci_racecheck_B2.log 2>&1; echo "racecheck B2 rc=$? (attn_ws hands A^T over through mbarriers, which racecheck does not model: DESIGN 5.2)" >> $O/ci_rc.txt
grep "Race reported\|hazard detected" $O/ci_racecheck_B2.log | sed 's/+0x[0-9a-f]*//g; s/(CUtensorMap_st[^)]*)//g' | sort | uniq -c | sort -rn | head -20 > $O/ci_racecheck_B2.kinds.txt
DSHEG_ATTN=v3 DSHEG_EXPO=0 timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python scripts/prof_denoise.py --batch 2 --calls 1 > $O/ci_racecheck_B2_nows.log 2>&1; echo "racecheck B2 without attn_ws (every other kernel of the call) rc=$?" >> $O/ci_rc.txt
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/prof_denoise.py --batch 50 --calls 1 > $O/ci_memcheck_B50.log 2>&1; echo "memcheck B50 rc=$?" >> $O/ci_rc.txt
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/prof_denoise.py --batch 3 --calls 1 --precision tf32 > $O/ci_memcheck_tf32.log 2>&1; echo "memcheck tf32 rc=$?" >> $O/ci_rc.txt
# ncu: launch list of one denoiser call at the headline batch; full captures of one layer's GEMMs and of the attention kernel
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 200 --csv --log-file $O/ci_launches.csv python scripts/prof_denoise.py --batch 950 --calls 2 > $O/ci_prof_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attn_ws_kernel -s 17 -c 1 -o $O/ci_attn_ws python scripts/prof_denoise.py --batch 950 --calls 2 > $O/ci_prof_attn.log 2>&1
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 133 -c 7 -o $O/ci_gemm python scripts/prof_denoise.py --batch 950 --calls 2 > $O/ci_prof_gemm.log 2>&1
head -c 4000 $O/ci_racecheck_B2.log > $O/ci_racecheck_B2.head.log
for n in memcheck_B3 memcheck_B50 memcheck_tf32 racecheck_B2 racecheck_B2_nows; do tail -n 6 $O/ci_$n.log > $O/ci_$n.tail.log; rm -f $O/ci_$n.log; done
cat $O/ci_rc.txt; grep -E "passed|failed|rror" $O/ci_pytest_gpu.log | tail -4; tail -2 $O/ci_smoke.log
python - <<'PY'
import json
for f in ("gpurun_out/ci_bench.json", "gpurun_out/ci_bench_tf32.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["value"]), "frames/s", round(d["ms_per_step"], 1), "ms  e2e", round(d["e2e"]["value"]), " gemm", round(d["roofline"]["achieved"]), "TF/s frac", round(d["roofline"]["frac"], 3),
              " attn", round(d["roofline_attention"]["achieved"]), "GB/s frac", round(d["roofline_attention"]["frac"], 3), " parity", d.get("parity"), " ref_cuda", d.get("ref_cuda", {}).get("rows"), d["clocks"])
    except Exception as e:
        print(f, "failed", e)
for ln in open("gpurun_out/ci_configs.jsonl"):
    try:
        d = json.loads(ln); print(d["config"]["id"], round(d["value"]), "frames/s", round(d["ms_per_step"], 1), "ms", d.get("parity", {}).get("relmax"))
    except Exception as e:
        print("config line failed", e)
PY
