#!/bin/bash
# full gpu suite on the build with the CLIP text tower + bench lines
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/gputests.txt 2>&1; tail -15 gpurun_out/gputests.txt
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_B1.json 2> gpurun_out/bench_B1.err; cut -c1-300 gpurun_out/bench_B1.json
