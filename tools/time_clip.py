"""Conditioning prep timing: ONE batched pass of the sm_100a CLIP text tower over all strings of an
image (prompt, "", box phrases, relation phrases) against the reference's call pattern on the same GPU --
eager fp32 transformers, one call per string: CLIPTextModel for prompt / "" / relations (FrozenCLIPEmbedder,
modules.py:157-169) and the FULL CLIPModel with a dummy 224 x 224 vision pass per box phrase (get_clip_feature,
txt2img.py:147-156).  Random-init weights of the clip-vit-large-patch14 architecture (no checkpoints offline).
usage: python tools/time_clip.py [n_boxes] [n_relations]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "stubs"))
import torch  # noqa: E402

from layoutllm_t2i_b200.clip import ClipTextEncoder, prepare_conditioning  # noqa: E402
from ltt_test_stubs import HashTokenizer  # noqa: E402
from oracle import clip_text_oracle as co  # noqa: E402

n_boxes = int(sys.argv[1]) if len(sys.argv) > 1 else 6
n_rel = int(sys.argv[2]) if len(sys.argv) > 2 else 3
DEV = "cuda"
cfg = co.default_clip_text_config()
sd = {k: v.to(DEV) for k, v in co.random_state_dict(cfg, seed=0).items()}
enc = ClipTextEncoder(cfg, 0)
enc.load_state_dict(sd)
enc.finalize()
tok = HashTokenizer(cfg["vocab_size"])
tok77 = lambda t: tok(t, truncation=True, max_length=77, padding="max_length")["input_ids"]  # noqa: E731
prompt = "a cat sitting on a wooden bench next to a red bicycle in the park"
phrases = ["object number %d" % i for i in range(n_boxes)]
locations = [[0.1, 0.1, 0.5, 0.5]] * n_boxes
relations = ["PAD", "cat sitting on bench", "bench next to bicycle", "cat sitting on bench", "bench next to bicycle"][:n_rel]


def timed(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


ours = timed(lambda: prepare_conditioning(enc, tok77, prompt, phrases, locations, relations, batch=1))
ids = tok77([prompt, ""] + phrases + relations).to(DEV)
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for _ in range(3):
    enc.encode_ids(ids)
ev0.record()
for _ in range(20):
    enc.encode_ids(ids)
ev1.record()
torch.cuda.synchronize()
dev_ms = ev0.elapsed_time(ev1) / 20
n0 = enc.launch_count
enc.encode_ids(ids)
per_pass = enc.launch_count - n0
print(f"strings per image: {ids.shape[0]} (prompt, negative prompt, {n_boxes} box phrases, {len(relations)} relation phrases)")
print(f"sm_100a tower, ONE batched pass [B={ids.shape[0]}, 77]: {dev_ms:.3f} ms on the device; prepare_conditioning end to end "
      f"(tokenizer stub + pass + tensor assembly): {ours:.3f} ms; launches per pass: {per_pass}")

try:
    from transformers import CLIPConfig, CLIPModel, CLIPTextConfig, CLIPTextModel
    keys = ("vocab_size", "max_position_embeddings", "hidden_size", "num_attention_heads", "num_hidden_layers", "intermediate_size",
            "layer_norm_eps", "hidden_act", "projection_dim", "eos_token_id")
    tc = CLIPTextConfig(**{k: cfg[k] for k in keys}, bos_token_id=cfg["vocab_size"] - 2, pad_token_id=1)
    text_model = CLIPTextModel(tc).to(DEV).eval()
    text_model.load_state_dict({k: v for k, v in sd.items() if k.startswith("text_model.")}, strict=False)
    full = CLIPModel(CLIPConfig(text_config=tc.to_dict(), vision_config=dict(hidden_size=1024, intermediate_size=4096, num_hidden_layers=24,
                                                                              num_attention_heads=16, patch_size=14, image_size=224),
                                projection_dim=768)).to(DEV).eval()
    pix = torch.ones(1, 3, 224, 224, device=DEV)

    @torch.no_grad()
    def reference_pattern():
        text_model(input_ids=tok77([prompt]).to(DEV)).last_hidden_state                       # context
        if relations:
            text_model(input_ids=tok77(relations).to(DEV)).pooler_output                      # relation phrases (one call)
        text_model(input_ids=tok77([""]).to(DEV)).last_hidden_state                           # uc
        for p in phrases:                                                                     # get_clip_feature per phrase
            t = tok(p, padding=True)
            full(input_ids=t["input_ids"].to(DEV), attention_mask=t["attention_mask"].to(DEV), pixel_values=pix).text_model_output.pooler_output

    ref = timed(reference_pattern, reps=5, warm=2)
    print(f"reference call pattern (eager fp32 transformers {__import__('transformers').__version__}, {3 if relations else 2} text calls + "
          f"{n_boxes} CLIPModel calls incl. dummy vision pass): {ref:.1f} ms  -> {ref / ours:.1f}x")
    from oracle.ref_loader import true_fp32
    with torch.no_grad(), true_fp32():
        o = text_model(input_ids=ids)
    z, pooled = enc.encode_ids(ids)
    print(f"relative L2, engine vs transformers fp32 (TF32 off) on the GPU: last_hidden_state "
          f"{float((z - o.last_hidden_state).norm() / o.last_hidden_state.norm()):.2e}, pooler_output "
          f"{float((pooled - o.pooler_output).norm() / o.pooler_output.norm()):.2e}")
except Exception as ex:  # noqa: BLE001
    print("reference arm unavailable:", repr(ex)[:200])
