#!/bin/bash
# SECOND gpurun call of the next round (about 35 GPU-minutes): sanitizers over the new kernels, the programmatic-dependent-launch build,
# the single-clip configurations, post-processing bandwidth and the variants the first call left out.
# Usage: gpurun --timeout 2700 -- 'bash scripts/gpu_round2_second.sh'   (after scripts/gpu_round2_first.sh; same build_variants/)
mkdir -p gpurun_out
export DSHEG_PROF_TABLE=1
O=gpurun_out
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
: > $O/r2b_rc.txt

# ---- sanitizers: new attention kernels + ACT_EXPO epilogue (memcheck B = 3, racecheck B = 2: single-CTA GEMM variants), ACT_LNMS and the
#      CTA-pair GEMMs at B = 24 (racecheck: full log for the open tcgen05.alloc item in profiles/r01/NOTES_next_round.md)
for a in v5c1 v5c4 v6; do
  DSHEG_ATTN=$a DSHEG_EXPO=1 timeout 200 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/prof_denoise.py --batch 3 --calls 1 > $O/r2_${a}_memcheck.log 2>&1; echo "$a memcheck rc=$?" >> $O/r2b_rc.txt
  DSHEG_ATTN=$a DSHEG_EXPO=1 timeout 200 compute-sanitizer --tool racecheck --error-exitcode 9 python scripts/prof_denoise.py --batch 2 --calls 1 > $O/r2_${a}_racecheck.log 2>&1; echo "$a racecheck rc=$?" >> $O/r2b_rc.txt
done
DSHEG_ATTN=v5c1 timeout 200 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/prof_denoise.py --batch 3 --calls 1 > $O/r2_v5c1_plain_memcheck.log 2>&1; echo "v5c1 (no expo) memcheck rc=$?" >> $O/r2b_rc.txt
DSHEG_FUSE_LNMS=1 DSHEG_ATTN=v6 DSHEG_EXPO=1 timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/prof_denoise.py --batch 24 --calls 1 > $O/r2_lnms_memcheck_B24.log 2>&1; echo "lnms+v6+expo memcheck B=24 rc=$?" >> $O/r2b_rc.txt
timeout 300 compute-sanitizer --tool racecheck python scripts/prof_denoise.py --batch 24 --calls 1 > $O/r2_racecheck_pairs_B24.log 2>&1; echo "racecheck pairs rc=$?" >> $O/r2b_rc.txt

# ---- programmatic dependent launch build (griddepcontrol in every bf16 hot-path kernel): parity first, then the latency-bound single-clip
#      configs (B = 1: about 165 dependent kernels per call) and the headline
PDL=$PWD/build_variants/libdiffsheg_b200_pdl.so
DSHEG_LIB=$PDL timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "denoise or loop or rows_are_independent" > $O/r2_pdl_parity.log 2>&1; echo "pdl parity rc=$?" >> $O/r2b_rc.txt
timeout 200 python scripts/bench_configs.py 1 > $O/r2_configs1_default.jsonl 2>&1
DSHEG_LIB=$PDL timeout 200 python scripts/bench_configs.py 1 > $O/r2_configs1_pdl.jsonl 2>&1
# single clip: 128-wide tiles double the CTAs that stream W and deepen the ring (5 stages) -- candidate heuristic for tiles < SMs / 4
DSHEG_TC_BN=128 timeout 200 python scripts/bench_configs.py 1 > $O/r2_configs1_bn128.jsonl 2>&1
DSHEG_TC_BN=128 DSHEG_LIB=$PDL timeout 200 python scripts/bench_configs.py 1 > $O/r2_configs1_bn128_pdl.jsonl 2>&1
DSHEG_FUSE_LNMS=1 timeout 200 python scripts/bench_configs.py 1 > $O/r2_configs1_lnms.jsonl 2>&1   # single-CTA form of the fused LayerNorm GEMM: 17 launches less per call
DSHEG_LIB=$PDL timeout 300 $B > $O/r2_bench_gemm_pdl.json 2> $O/r2_bench_gemm_pdl.err

# ---- the variants the first call left out
DSHEG_ATTN=v4 timeout 300 $B > $O/r2_bench_attn_v4.json 2> $O/r2_bench_attn_v4.err
DSHEG_ATTN=v5c2 DSHEG_EXPO=1 timeout 300 $B > $O/r2_bench_attn_v5c2_expo.json 2> $O/r2_bench_attn_v5c2_expo.err
DSHEG_ATTN=v5c4 DSHEG_QSOFT=1 timeout 300 $B > $O/r2_bench_attn_v5c4_qsoft.json 2> $O/r2_bench_attn_v5c4_qsoft.err
DSHEG_LIB=$PWD/build_variants/libdiffsheg_b200_split64.so timeout 300 $B > $O/r2_bench_gemm_split64.json 2> $O/r2_bench_gemm_split64.err
DSHEG_TC_PREFETCH=3 DSHEG_LIB=$PWD/build_variants/libdiffsheg_b200_split73.so timeout 300 $B > $O/r2_bench_gemm_split73_prefetch3.json 2> $O/r2_bench_gemm_split73_prefetch3.err
timeout 200 python scripts/bench_postprocess.py > $O/r2_postprocess_bw.txt 2>&1

# ---- summary
cat $O/r2b_rc.txt
python scripts/gpu_round2_summary.py
grep -c "Race reported\|hazard" $O/r2_racecheck_pairs_B24.log; grep -E "Write access|Read access" $O/r2_racecheck_pairs_B24.log | sed 's/(CUtensorMap.*//' | sort | uniq -c | sort -rn | head -8
cat $O/r2_postprocess_bw.txt
echo "== single clip (config 1): default / PDL / BN=128 / both"; for f in default pdl bn128 bn128_pdl lnms; do echo "-- $f"; cut -c1-200 $O/r2_configs1_$f.jsonl; done; tail -2 $O/r2_pdl_parity.log
