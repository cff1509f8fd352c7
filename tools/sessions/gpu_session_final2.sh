#!/bin/bash
# Final evidence of round 2 on the shipped build: gpu suite (incl. compute-sanitizer), ncu launch lists / DRAM traffic / --set full.
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/y_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/y_pytest.log
grep -E "passed|failed|FAILED|rc=" gpurun_out/y_pytest.log | tail -8
K='regex:gemm_tc_kernel|attn_tc_kernel|gn_|layernorm_kernel|rela_|conv_in|conv_out|plms_|upsample2x|im2col|timestep|posnet|softmax_rows|vae_'
LTT_NO_AUTOTUNE=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -s 23500 -c 900 --csv \
   --log-file gpurun_out/y_launches_bench.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-decode > gpurun_out/y_ncu_bench.log 2>&1
echo "ncu bench rc=$?"; python tools/ncu_summarize.py gpurun_out/y_launches_bench.csv | head -12
LTT_NO_GRAPH=1 LTT_CUPROF=1 timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv \
   --log-file gpurun_out/y_dram_a1.csv python tools/profile_forward.py 1 64 3 1.0 > gpurun_out/y_ncu_dram.log 2>&1
python tools/ncu_dram_traffic.py gpurun_out/y_dram_a1.csv gpurun_out/y_dram_traffic.json "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --profile-from-start off over one warm [cond;uncond] UNet evaluation (gate 1, B=1, 64x64 latent), tools/profile_forward.py, round 2 final build" | tail -8
LTT_NO_GRAPH=1 LTT_CUPROF=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --profile-from-start off --csv \
   --log-file gpurun_out/y_launches_eval_a1.csv python tools/profile_forward.py 1 64 3 1.0 > gpurun_out/y_ncu_eval.log 2>&1
python tools/ncu_summarize.py gpurun_out/y_launches_eval_a1.csv | head -8
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/y_ops python tools/ncu_ops.py attn40 attn80 conv0 geglu0 lin0 > gpurun_out/y_ncu_ops.log 2>&1
echo "ncu ops rc=$?"; ls -la gpurun_out/y_ops.ncu-rep
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/y_bench_B1.json 2> gpurun_out/y_bench_B1.err; cut -c1-200 gpurun_out/y_bench_B1.json
