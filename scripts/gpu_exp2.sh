bash scripts/gpu_exp.sh
bash scripts/gpu_profile.sh > gpurun_out/profile_final.log 2>&1; tail -3 gpurun_out/profile_final.log
