#!/bin/bash
# strong scaling of BASELINE config 5: 512x512, 50 PLMS steps, 6 boxes, global batch 64 over N GPUs (one process per GPU)
N=$1
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
   bench.py --gpus $N --global-batch 64 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/scale_strong_N$N.json 2> gpurun_out/scale_strong_N$N.err
echo "rc=$?"; tail -3 gpurun_out/scale_strong_N$N.err | cut -c1-300
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/scale_strong_N$N.json") if l.startswith("{")][-1])
print("N=$N", d["value"], "img/s", d["ms_per_step"], "ms/step", "sampler-only", d["sampler_only"]["value"], "e2e", d["e2e"]["value"], d["config"]["per_gpu_batch"], d["scaling"])
PY
