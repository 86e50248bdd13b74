#!/bin/bash
# ncu evidence for profiles/: launch list of one denoiser call at the headline batch, full captures of one layer's
# GEMMs and of the attention kernel (B200_PROFILING.md recipe).  Run under gpurun (1 GPU).
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 340 -c 175 --csv --log-file gpurun_out/launches.csv python scripts/prof_denoise.py --batch 950 --calls 3 > gpurun_out/prof_launches.log 2>&1
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 266 -c 7 -o gpurun_out/gemm python scripts/prof_denoise.py --batch 950 --calls 3 > gpurun_out/prof_gemm.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attn_ws_kernel -s 34 -c 1 -o gpurun_out/attn python scripts/prof_denoise.py --batch 950 --calls 3 > gpurun_out/prof_attn.log 2>&1
tail -n 2 gpurun_out/prof_launches.log gpurun_out/prof_gemm.log gpurun_out/prof_attn.log
