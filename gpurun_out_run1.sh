mkdir -p gpurun_out; nvidia-smi > gpurun_out/smi.txt 2>&1
python -c "import torch; print(torch.cuda.get_device_name(0))" > gpurun_out/dev.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q -s -k "not bf16 and not full_size" > gpurun_out/t_fp32.log 2>&1; echo "fp32 rc=$?" > gpurun_out/rc.txt
timeout 600 python -m pytest tests -m gpu -q -s -k "op_linear and bf16" > gpurun_out/t_tc.log 2>&1; echo "tc_op rc=$?" >> gpurun_out/rc.txt
DSHEG_GEMM_ENGINE=simt timeout 900 python -m pytest tests -m gpu -q -s -k "bf16 and not op_linear" > gpurun_out/t_bf16_simt.log 2>&1; echo "bf16_simt rc=$?" >> gpurun_out/rc.txt
timeout 900 python -m pytest tests -m gpu -q -s -k "(bf16 and not op_linear) or full_size" > gpurun_out/t_bf16_tc.log 2>&1; echo "bf16_tc rc=$?" >> gpurun_out/rc.txt
timeout 900 python bench.py --steps 2 --warmup 3 > gpurun_out/bench1.log 2>&1; echo "bench rc=$?" >> gpurun_out/rc.txt
cat gpurun_out/rc.txt; tail -5 gpurun_out/t_fp32.log; tail -5 gpurun_out/t_tc.log; tail -3 gpurun_out/t_bf16_tc.log; tail -2 gpurun_out/bench1.log
