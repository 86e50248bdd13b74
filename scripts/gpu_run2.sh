mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -s -k "attention or (denoise and bf16) or (ddim25 and bf16)" > gpurun_out/t2_attn.log 2>&1; echo "attn rc=$?" > gpurun_out/rc2.txt
# launch list of one denoiser call at the headline batch (prepare = 9 launches, 2 warm-up calls skipped)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 420 -c 215 --csv --log-file gpurun_out/launches_v1.csv python scripts/prof_denoise.py --batch 950 --calls 3 > gpurun_out/prof_a.log 2>&1
# full capture: the 7 GEMMs of exp layer 0 (third call) and one attention kernel
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 259 -c 7 -o gpurun_out/gemm_v0 python scripts/prof_denoise.py --batch 950 --calls 3 > gpurun_out/prof_b.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attn_v2_kernel -s 34 -c 1 -o gpurun_out/attn_v2 python scripts/prof_denoise.py --batch 950 --calls 3 > gpurun_out/prof_c.log 2>&1
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench2.log 2>&1; echo "bench rc=$?" >> gpurun_out/rc2.txt
cat gpurun_out/rc2.txt; tail -4 gpurun_out/t2_attn.log; tail -2 gpurun_out/prof_a.log gpurun_out/prof_b.log gpurun_out/prof_c.log; tail -c 1500 gpurun_out/bench2.log
