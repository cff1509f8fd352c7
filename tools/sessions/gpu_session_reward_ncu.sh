#!/bin/bash
# ncu launch list of one reward evaluation (8 samples: text pass + vision pass over 16 images + head)
mkdir -p gpurun_out
cat > gpurun_out/_reward_once.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import torch
from layoutllm_t2i_b200.clip import ClipTextEncoder, ClipVisionEncoder
from layoutllm_t2i_b200.reward import reward_head
from oracle import clip_text_oracle as co, clip_vision_oracle as cv
B = 8
tcfg, vcfg = co.default_clip_text_config(), cv.default_clip_vision_config()
sd = {k: v.cuda() for k, v in co.random_state_dict(tcfg, seed=0).items()}
sd.update({k: v.cuda() for k, v in cv.random_state_dict(vcfg, seed=1).items()})
aes = {k: v.cuda() for k, v in cv.aesthetic_state_dict(768, seed=2).items()}
text, vision = ClipTextEncoder(tcfg, 0), ClipVisionEncoder(vcfg, 0)
text.load_state_dict(sd); vision.load_state_dict(sd); text.finalize(); vision.finalize()
ids = co.synthetic_ids(tcfg, [12] * B, L=14, seed=3).cuda()
imgs = torch.randint(0, 256, (2 * B, 512, 512, 3), dtype=torch.uint8, device="cuda")
def once():
    t = text.encode_ids(ids, want_hidden=False, want_embeds=True)[2]
    e = vision.encode(vision.preprocess(imgs))[2]
    return reward_head(t, e[:B], e[B:], aes)[0]
for _ in range(3): once()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
once()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
PY
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --profile-from-start off --csv --log-file gpurun_out/reward_launches.csv python gpurun_out/_reward_once.py > gpurun_out/reward_ncu.log 2>&1
python - <<'PY'
import csv, collections, re
rows = [r for r in csv.reader(open('gpurun_out/reward_launches.csv')) if len(r) > 10]
hdr = rows[0]; ki, vi = hdr.index('Kernel Name'), hdr.index('Metric Value')
agg = collections.OrderedDict()
for r in rows[1:]:
    name = re.sub(r'\(.*', '', r[ki]); v = float(r[vi].replace(',', '')) / 1000.0
    a = agg.setdefault(name, [0.0, 0]); a[0] += v; a[1] += 1
tot = sum(a[0] for a in agg.values())
print(f"{sum(a[1] for a in agg.values())} launches, {tot/1000:.3f} ms of kernel time (serialised by ncu)")
for n, a in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"{a[0]:10.1f} us {100*a[0]/tot:5.1f}% {a[1]:5d} {a[0]/a[1]:8.2f}  {n}")
PY
