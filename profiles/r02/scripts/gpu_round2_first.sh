#!/bin/bash
# FIRST gpurun call of the next round (about 40 GPU-minutes): the hardware verdict and the headline numbers for everything written
# while no GPU was available.  All of it is CPU-validated on the thread-level emulator (tests/test_emu_*.py).
#   1. first_hw_run.py : op-level parity of every new kernel (post-processing, attention v4 / v5 / v6, ACT_EXPO, ACT_LNMS) and agreement
#                        with the default path inside dsheg_denoise at B = 3 / B = 24 -- one subprocess per candidate
#   2. bench.py        : the real loop (B = 950) per GEMM experiment build, attention variant and fusion, back to back on this box
# scripts/gpu_round2_second.sh (sanitizers, programmatic dependent launch, single-clip configs, the remaining variants) is the second call.
# NB the variant libraries are loaded through the same ctypes table as the default one: after ANY change under diffsheg_b200/csrc or
# include/, rebuild them HERE (bash scripts/build_variants.sh, about 5 min) before the gpurun call -- a stale variant fails to load.
# Usage: bash scripts/build_variants.sh   (here, no GPU needed; the .so files travel)   then
#        gpurun --timeout 3000 -- 'bash scripts/gpu_round2_first.sh'
mkdir -p gpurun_out
export DSHEG_PROF_TABLE=1   # per-kernel-name table of every profiled region on stderr (the .err file of each bench run)
O=gpurun_out
[ -f build_variants/libdiffsheg_b200_split73.so ] && [ -f build_variants/libdiffsheg_b200_pdl.so ] || bash scripts/build_variants.sh > $O/r2_build_variants.log 2>&1
DSHEG_FIRST_RUN_BATCH=0 timeout 1500 python scripts/first_hw_run.py > $O/r2_first_hw_run.log 2>&1; echo "first_hw_run rc=$?" > $O/r2_rc.txt
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"

# ---- GEMM candidates: isolated sweep (dsheg_bench_gemm, 10 iterations per shape) and the real loop
for v in default k512deep split73 epipacked; do
  lib=""; [ $v != default ] && lib="$PWD/build_variants/libdiffsheg_b200_$v.so"
  DSHEG_LIB=$lib timeout 200 python scripts/bench_gemm.py > $O/r2_gemm_sweep_$v.txt 2>&1
  DSHEG_LIB=$lib timeout 300 $B > $O/r2_bench_gemm_$v.json 2> $O/r2_bench_gemm_$v.err
done
DSHEG_TC_PREFETCH=3 timeout 300 $B > $O/r2_bench_gemm_prefetch3.json 2> $O/r2_bench_gemm_prefetch3.err

# ---- attention: the instruction-diet kernels as 1 / 2 / 4 CTAs per sample ...
for a in v5c1 v5c2 v5c4; do
  DSHEG_ATTN=$a timeout 300 $B > $O/r2_bench_attn_$a.json 2> $O/r2_bench_attn_$a.err
done
# ... and with Q AND K softmax numerators (static, pack-time-proven shifts) from the QKV epilogue (ACT_EXPO): attn_v5<CL, 2> has no
# exp / max left outside its LayerNorm pass, attn_v6 additionally runs 32 warps per SM.  Watch BOTH columns (attention down, gemm up?)
for a in v5c1 v5c4 v6 v6c2 v6c1; do
  DSHEG_ATTN=$a DSHEG_EXPO=1 timeout 300 $B > $O/r2_bench_attn_${a}_expo.json 2> $O/r2_bench_attn_${a}_expo.err
done

# ---- ffn.linear2 + LayerNorm / modulate / SiLU in ONE GEMM (ACT_LNMS): the "rowwise" column (about 41 ms per step) should disappear,
#      the ffn2 GEMM loses one ring stage and gains an exposed epilogue
DSHEG_FUSE_LNMS=1 timeout 300 $B > $O/r2_bench_gemm_lnms.json 2> $O/r2_bench_gemm_lnms.err
# the audio encoder's attention (D = 128, 8 heads of 16) with all heads at once instead of the generic head-by-head kernel (336 us per call)
DSHEG_ATTN_AUD=1 timeout 300 $B > $O/r2_bench_attn_aud.json 2> $O/r2_bench_attn_aud.err
for a in v5c4 v6; do
  DSHEG_ATTN=$a DSHEG_EXPO=1 DSHEG_FUSE_LNMS=1 DSHEG_ATTN_AUD=1 timeout 300 $B > $O/r2_bench_all_${a}_expo_lnms_aud.json 2> $O/r2_bench_all_${a}_expo_lnms_aud.err
done

# ---- summary
cat $O/r2_rc.txt
grep -E "^(PASS|FAIL)" $O/r2_first_hw_run.log | cut -c1-220
python scripts/gpu_round2_summary.py
for v in default k512deep split73 epipacked; do echo "== gemm sweep $v"; grep -E "qkv|sa_out|ffn1|ffn2" $O/r2_gemm_sweep_$v.txt | cut -c1-140; done
