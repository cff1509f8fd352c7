"""oracle/clip_text_oracle.py against the third-party implementation the reference calls (fixtures made by
tests/gen_golden_clip.py from the installed `transformers`), and the properties the batched conditioning pass rests on."""
import os

import pytest
import torch

from oracle import clip_text_oracle as co

GOLD = os.path.join(os.path.dirname(__file__), "golden", "clip_text.pt")


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


@pytest.mark.parametrize("name", ["tiny", "tiny_eos"])
def test_oracle_matches_transformers_fixture(name):
    g = torch.load(GOLD)[name]
    sd = co.random_state_dict(g["cfg"], seed=g["seed"], outliers=True)
    for tag, c in g["cases"].items():
        with torch.no_grad():
            z, pooled = co.clip_text_forward(sd, g["cfg"], c["ids"], c["mask"])
            emb = co.text_features(sd, g["cfg"], c["ids"], c["mask"])
        assert rel(z, c["last_hidden_state"]) < 2e-6, tag
        assert rel(pooled, c["pooler_output"]) < 2e-6, tag
        assert rel(emb, c["text_embeds"]) < 2e-6, tag


def test_oracle_matches_transformers_live():
    pytest.importorskip("transformers")
    import sys
    sys.path.insert(0, os.path.dirname(__file__))
    from gen_golden_clip import hf_text_model
    cfg = co.tiny_clip_text_config()
    sd = co.random_state_dict(cfg, seed=11)
    m, _ = hf_text_model(cfg, sd)
    ids = co.synthetic_ids(cfg, [4, 9, 0], seed=5)
    with torch.no_grad():
        o = m(input_ids=ids)
        z, pooled = co.clip_text_forward(sd, cfg, ids)
    assert rel(z, o.last_hidden_state) < 2e-6 and rel(pooled, o.pooler_output) < 2e-6


def test_batched_padded_pass_equals_per_phrase_calls():
    """Causal attention: rows up to the end-of-text token do not see the padding behind it, so ONE pass over phrases
    padded to 77 tokens (no attention mask) returns the pooled vectors of the reference's per-phrase unpadded calls
    (txt2img.py:147-156, modules.py:174-182) and of a masked padded batch (txt2img.py:454-457)."""
    cfg = co.tiny_clip_text_config()
    sd = co.random_state_dict(cfg, seed=2)
    words = [1, 3, 8, 20, 75]
    ids = co.synthetic_ids(cfg, words, seed=9)
    with torch.no_grad():
        _, pooled = co.clip_text_forward(sd, cfg, ids)
        mask = torch.zeros_like(ids)
        for r, n in enumerate(words):
            mask[r, :n + 2] = 1
        _, pooled_masked = co.clip_text_forward(sd, cfg, ids, mask)
        for r, n in enumerate(words):
            _, one = co.clip_text_forward(sd, cfg, ids[r:r + 1, :n + 2])
            assert rel(pooled[r:r + 1], one) < 1e-5
    assert rel(pooled, pooled_masked) < 1e-5


def test_eos_position_rules():
    ids = torch.tensor([[998, 5, 999, 999], [998, 999, 0, 0]])
    assert co.eos_positions(ids, 2).tolist() == [2, 1]
    assert co.eos_positions(ids, 999).tolist() == [2, 1]
    assert co.eos_positions(torch.tensor([[998, 7, 3, 999]]), 999).tolist() == [3]


# ------------------------------------------------------------------------------------------------- vision tower / reward head
def test_vision_oracle_matches_transformers_fixture():
    from oracle import clip_vision_oracle as cv
    g = torch.load(os.path.join(os.path.dirname(__file__), "golden", "clip_vision.pt"))
    sd = cv.random_state_dict(g["cfg"], seed=g["seed"])
    with torch.no_grad():
        z, pooled = cv.clip_vision_forward(sd, g["cfg"], g["pixel_values"])
        emb = cv.image_features(sd, g["cfg"], g["pixel_values"])
    assert rel(z, g["last_hidden_state"]) < 2e-6 and rel(pooled, g["pooler_output"]) < 2e-6 and rel(emb, g["image_embeds"]) < 2e-6


def test_reward_head_oracle_matches_reference_pieces():
    """AestheticMLP / normalized are the reference's own code (tools/aesthetic.py, staged under oracle/_ref); the reward
    expression is models/policy.py:115-139 evaluated with them."""
    from oracle import clip_vision_oracle as cv
    from oracle import ref_loader as rl
    if not rl.available():
        pytest.skip("oracle/_ref not staged")
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "stubs"))
    with rl.reference_tree():
        sys.path.insert(0, rl.REF_ROOT)
        try:
            from tools.aesthetic import AestheticMLP, normalized
        finally:
            sys.path.remove(rl.REF_ROOT)
    sd = cv.aesthetic_state_dict(64, seed=3)
    m = AestheticMLP(64).eval()
    m.load_state_dict(sd)
    gen = torch.Generator().manual_seed(0)
    t, p, g = (torch.randn(5, 64, generator=gen) for _ in range(3))
    miou, laysim = torch.rand(5, generator=gen), torch.rand(5, generator=gen)
    with torch.no_grad():
        import torch.nn.functional as F
        tn, pn, gn = F.normalize(t, dim=-1), F.normalize(p, dim=-1), F.normalize(g, dim=-1)
        clip_reward = (tn * pn).sum(dim=-1) + (gn * pn).sum(dim=-1)
        aes = m(torch.from_numpy(normalized(pn.numpy())).float()).flatten()
        want = clip_reward + aes * 0.1 + miou * 10 + laysim * 10
        got, got_clip, got_aes = cv.reward_forward(t, p, g, sd, miou, laysim)
    assert rel(got_clip, clip_reward) < 1e-6 and rel(got_aes, aes) < 1e-5 and rel(got, want) < 1e-6


# ------------------------------------------------------------------------------------------------- image preprocessing
def test_preprocess_oracle_matches_fixture():
    """tests/golden/clip_preprocess.npz = outputs of transformers' PIL-backed CLIP image processor: bit-exact."""
    import numpy as np
    from oracle import clip_preprocess_oracle as cp
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "clip_preprocess.npz"))
    imgs = [g[f"img{i}"] for i in range(5)]
    assert np.array_equal(cp.preprocess(imgs, size=int(g["size"])), g["pixel_values"])


def test_preprocess_resize_matches_pillow_live():
    import numpy as np
    PIL = pytest.importorskip("PIL.Image")
    from oracle import clip_preprocess_oracle as cp
    rng = np.random.default_rng(1)
    for h, w in ((512, 512), (64, 80), (300, 224), (97, 211), (640, 480)):
        a = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        rh, rw, _, _ = cp.output_geometry(h, w)
        assert np.array_equal(np.asarray(PIL.fromarray(a).resize((rw, rh), resample=PIL.BICUBIC)), cp.pil_resize_bicubic(a, rw, rh)), (h, w)
