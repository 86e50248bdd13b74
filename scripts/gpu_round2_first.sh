#!/bin/bash
# FIRST gpurun call of the next round: everything written while no GPU was available, in ONE box session (~50 min; split it at the section comments if the budget is tight).
# All of it is CPU-validated on the thread-level emulator (tests/test_emu_*.py); this call gives the hardware verdict and the
# numbers that decide which candidates become defaults.
#   1. first_hw_run.py   : post-processing kernels, attention v4 / v5c1 / v5c2 / v5c4 (+ qsoft / expo / lnms combinations) -- op parity and
#                          agreement with the default inside dsheg_denoise at B = 3 (one subprocess per variant)
#   2. GEMM candidates   : isolated shape sweep + full bench.py per experiment build (DSHEG_LIB) and prefetch mode
#   3. bench.py          : default vs the best attention variant, back to back on this box
#   4. racecheck         : full log of the CTA-pair GEMMs at B = 24 (open item in profiles/r01/NOTES_next_round.md)
# NB the variant libraries are loaded through the same ctypes table as the default one: after ANY change under diffsheg_b200/csrc or
# include/, rebuild them HERE (bash scripts/build_variants.sh, about 5 min) before the gpurun call -- a stale variant fails to load.
# Usage: bash scripts/build_variants.sh   (here, no GPU needed; the .so files travel)   then
#        gpurun --timeout 3600 -- 'bash scripts/gpu_round2_first.sh'
mkdir -p gpurun_out
export DSHEG_PROF_TABLE=1   # per-kernel-name table of every profiled region on stderr (the .err file of each bench run)
O=gpurun_out
[ -f build_variants/libdiffsheg_b200_split73.so ] && [ -f build_variants/libdiffsheg_b200_pdl.so ] || bash scripts/build_variants.sh > $O/r2_build_variants.log 2>&1
DSHEG_FIRST_RUN_BATCH=0 timeout 1300 python scripts/first_hw_run.py > $O/r2_first_hw_run.log 2>&1; echo "first_hw_run rc=$?" > $O/r2_rc.txt

# ---- GEMM candidates: isolated sweep (dsheg_bench_gemm, 10 iterations per shape) and the real loop
for v in default k512deep split73 split64 epipacked; do
  lib=""; [ $v != default ] && lib="$PWD/build_variants/libdiffsheg_b200_$v.so"
  DSHEG_LIB=$lib timeout 200 python scripts/bench_gemm.py > $O/r2_gemm_sweep_$v.txt 2>&1
  DSHEG_LIB=$lib timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/r2_bench_gemm_$v.json 2> $O/r2_bench_gemm_$v.err
done
DSHEG_TC_PREFETCH=3 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/r2_bench_gemm_prefetch3.json 2> $O/r2_bench_gemm_prefetch3.err
DSHEG_TC_PREFETCH=3 DSHEG_LIB=$PWD/build_variants/libdiffsheg_b200_split73.so timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/r2_bench_gemm_split73_prefetch3.json 2> $O/r2_bench_gemm_split73_prefetch3.err

# ---- attention variants in the real loop
for a in v4 v5c1 v5c2 v5c4; do
  DSHEG_ATTN=$a timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/r2_bench_attn_$a.json 2> $O/r2_bench_attn_$a.err
done
# Q row-softmax moved into the QKV GEMM epilogue (ACT_QSOFT) + attn_v5<CL, QPRE>: watch BOTH the attention and the gemm column
for a in v5c1 v5c4; do
  DSHEG_ATTN=$a DSHEG_QSOFT=1 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/r2_bench_attn_${a}_qsoft.json 2> $O/r2_bench_attn_${a}_qsoft.err
done

# Q AND K softmax numerators with static, pack-time-proven shifts from the QKV epilogue (ACT_EXPO) + attn_v5<CL, 2>: the attention
# kernel loses every exp / max outside its LayerNorm pass (static SASS 2872 -> 2096 for CL = 1); again watch BOTH columns
for a in v5c1 v5c2 v5c4 v6; do
  DSHEG_ATTN=$a DSHEG_EXPO=1 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/r2_bench_attn_${a}_expo.json 2> $O/r2_bench_attn_${a}_expo.err
done

# ffn.linear2 + LayerNorm / modulate / SiLU in ONE GEMM (ACT_LNMS: a CTA pair keeps both 256-column halves of its rows in TMEM): the
# ln_mod_silu pass ("rowwise" column, about 41 ms per step) disappears, the ffn2 GEMM loses one ring stage and gains an exposed epilogue
DSHEG_FUSE_LNMS=1 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/r2_bench_gemm_lnms.json 2> $O/r2_bench_gemm_lnms.err
for a in v5c4 v6; do
  DSHEG_ATTN=$a DSHEG_EXPO=1 DSHEG_FUSE_LNMS=1 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/r2_bench_all_${a}_expo_lnms.json 2> $O/r2_bench_all_${a}_expo_lnms.err
done

# ---- programmatic dependent launch build (griddepcontrol in every bf16 hot-path kernel): parity first, then the latency-bound single-clip
#      configs (B = 1: about 165 dependent kernels per call) and the headline
PDL=$PWD/build_variants/libdiffsheg_b200_pdl.so
DSHEG_LIB=$PDL timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "denoise or loop or rows_are_independent" > $O/r2_pdl_parity.log 2>&1; echo "pdl parity rc=$?" >> $O/r2_rc.txt
timeout 200 python scripts/bench_configs.py 1 > $O/r2_configs1_default.jsonl 2>&1
DSHEG_LIB=$PDL timeout 200 python scripts/bench_configs.py 1 > $O/r2_configs1_pdl.jsonl 2>&1
# single clip: 128-wide tiles double the CTAs that stream W and deepen the ring (5 stages) -- candidate heuristic for tiles < SMs / 4
DSHEG_TC_BN=128 timeout 200 python scripts/bench_configs.py 1 > $O/r2_configs1_bn128.jsonl 2>&1
DSHEG_TC_BN=128 DSHEG_LIB=$PDL timeout 200 python scripts/bench_configs.py 1 > $O/r2_configs1_bn128_pdl.jsonl 2>&1
DSHEG_LIB=$PDL timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/r2_bench_gemm_pdl.json 2> $O/r2_bench_gemm_pdl.err

# ---- sanitizers: v5 variants (memcheck + racecheck at small batch), CTA-pair GEMM racecheck (full log)
for a in v5c1 v5c4 v6; do
  DSHEG_ATTN=$a DSHEG_EXPO=1 timeout 200 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/prof_denoise.py --batch 3 --calls 1 > $O/r2_${a}_memcheck.log 2>&1; echo "$a memcheck rc=$?" >> $O/r2_rc.txt
  DSHEG_ATTN=$a DSHEG_EXPO=1 timeout 200 compute-sanitizer --tool racecheck --error-exitcode 9 python scripts/prof_denoise.py --batch 2 --calls 1 > $O/r2_${a}_racecheck.log 2>&1; echo "$a racecheck rc=$?" >> $O/r2_rc.txt
done
timeout 300 compute-sanitizer --tool racecheck python scripts/prof_denoise.py --batch 24 --calls 1 > $O/r2_racecheck_pairs_B24.log 2>&1; echo "racecheck pairs rc=$?" >> $O/r2_rc.txt
timeout 200 python scripts/bench_postprocess.py > $O/r2_postprocess_bw.txt 2>&1

# ---- summary
cat $O/r2_rc.txt
grep -E "^(PASS|FAIL)" $O/r2_first_hw_run.log | cut -c1-220
python - <<'PY'
import glob, json, os
for f in sorted(glob.glob("gpurun_out/r2_bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f"{os.path.basename(f):44s} {d['value']:10.0f} frames/s  gemm {d['roofline']['achieved']:7.1f} TF/s ({d['roofline']['ms_per_step']:6.1f} ms)  rowwise {d.get('rowwise', {}).get('ms_per_step', 0):5.1f} ms"
              f"  attention {d['roofline_attention']['achieved']:6.0f} GB/s ({d['roofline_attention']['ms_per_step']:6.1f} ms)  sm {d['clocks']['sm_mhz']}")
    except Exception as e:  # noqa: BLE001
        print(os.path.basename(f), "failed:", e)
PY
for v in default k512deep split73 split64 epipacked; do echo "== gemm sweep $v"; grep -E "qkv|sa_out|ffn1|ffn2 " $O/r2_gemm_sweep_$v.txt | cut -c1-140; done
grep -c "Race reported\|hazard" $O/r2_racecheck_pairs_B24.log; grep -E "Write access|Read access" $O/r2_racecheck_pairs_B24.log | sed 's/(CUtensorMap.*//' | sort | uniq -c | sort -rn | head -8
cat $O/r2_postprocess_bw.txt
echo "== single clip (config 1), default vs PDL build"; for f in default pdl bn128 bn128_pdl; do echo "-- $f"; cut -c1-200 $O/r2_configs1_$f.jsonl; done; tail -2 $O/r2_pdl_parity.log
