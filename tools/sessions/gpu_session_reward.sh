#!/bin/bash
# reward path (row f4): CLIP vision tower + preprocessing + reward head parity and timing
mkdir -p gpurun_out
python tests/reward_checks.py > gpurun_out/reward_checks.txt 2>&1; grep -v Warning gpurun_out/reward_checks.txt | tail -40
python tools/time_reward.py 8 > gpurun_out/reward_timing.txt 2>&1; python tools/time_reward.py 1 >> gpurun_out/reward_timing.txt 2>&1; grep -v "Warning" gpurun_out/reward_timing.txt | tail -12
