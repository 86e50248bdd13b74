#!/bin/bash
# compute-sanitizer over a small denoiser call + sampler steps (memcheck, then racecheck on shared memory)
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/prof_denoise.py --batch 3 --calls 1 > gpurun_out/sanitize_memcheck.log 2>&1; echo "memcheck rc=$?" > gpurun_out/sanitize_rc.txt
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python scripts/prof_denoise.py --batch 2 --calls 1 > gpurun_out/sanitize_racecheck.log 2>&1; echo "racecheck rc=$?" >> gpurun_out/sanitize_rc.txt
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "ddim_step or undo_ddpm" > gpurun_out/sanitize_steps.log 2>&1; echo "steps rc=$?" >> gpurun_out/sanitize_rc.txt
cat gpurun_out/sanitize_rc.txt; tail -5 gpurun_out/sanitize_memcheck.log; tail -8 gpurun_out/sanitize_racecheck.log; tail -3 gpurun_out/sanitize_steps.log
