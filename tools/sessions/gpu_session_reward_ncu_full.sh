#!/bin/bash
# ncu --set full of the two heaviest kernels of the reward path's vision pass (d = 64 attention, CTA-pair GEMM of the MLP)
mkdir -p gpurun_out
bash tools/sessions/gpu_session_reward_ncu.sh > /dev/null 2>&1      # writes gpurun_out/_reward_once.py
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:attn_tc_kernel -s 14 -c 1 -o gpurun_out/u_attn64 python gpurun_out/_reward_once.py > gpurun_out/u_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_tc_kernel -s 40 -c 3 -o gpurun_out/u_gemm python gpurun_out/_reward_once.py >> gpurun_out/u_ncu.log 2>&1
ls -la gpurun_out/u_*.ncu-rep
