mkdir -p gpurun_out
# launch list of the third denoiser call at the headline batch
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 330 -c 170 --csv --log-file gpurun_out/launches_v6.csv python scripts/prof_denoise.py --batch 950 --calls 3 > gpurun_out/prof12a.log 2>&1
# full captures: exp layer 1 GEMMs (pair kernels) and one attention kernel of the third call
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 266 -c 7 -o gpurun_out/gemm_v6 python scripts/prof_denoise.py --batch 950 --calls 3 > gpurun_out/prof12b.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attn_v3_kernel -s 34 -c 1 -o gpurun_out/attn_v3 python scripts/prof_denoise.py --batch 950 --calls 3 > gpurun_out/prof12c.log 2>&1
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench12.log 2>&1; echo "bench rc=$?" > gpurun_out/rc12.txt
tail -2 gpurun_out/prof12a.log gpurun_out/prof12b.log gpurun_out/prof12c.log; cat gpurun_out/rc12.txt; tail -c 600 gpurun_out/bench12.log
