"""The oracle (oracle/*.py) against golden vectors produced by the REFERENCE
itself (tests/gen_golden.py, run in the build container).  CPU only."""
import os

import numpy as np
import pytest
import torch

from oracle import plms_oracle as po
from oracle import unet_oracle as uo

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm()).item()


@pytest.fixture(scope="module")
def mods():
    return torch.load(os.path.join(GOLD, "modules.pt"), weights_only=False)


def test_relation_fusion_matches_reference(mods):
    o = mods["rela"]
    sd = {"p." + k: v for k, v in uo.seeded_module_sd(o["shapes"], seed=21, gates=(0.7, -0.4)).items()}
    x = uo.seeded_randn((3, o["h"] * o["w"], o["C"]), 31)
    relations = uo.seeded_randn((3, 10, 768), 32)
    y = uo.relation_fusion(sd, "p", x, relations, o["boxes"], o["masks"], o["h"], o["w"], 8)
    assert rel(y, o["y"]) < 1e-5
    # sample 1 starts with a degenerate box -> the scan breaks at slot 0 -> output == LN3(x)
    ln = uo.layer_norm(sd, "p.norm3", x)
    assert rel(y[1], ln[1]) < 1e-6
    assert rel(y[0], ln[0]) > 1e-3


def test_box_rect_rules():
    boxes = torch.zeros(1, 30, 4)
    masks = torch.zeros(1, 30)
    boxes[0, 0] = torch.tensor([0.1, 0.1, 0.5, 0.5])
    boxes[0, 1] = torch.tensor([0.26, 0.2, 0.30, 0.9])     # l == r at w=16 (not at 64) -> break
    boxes[0, 2] = torch.tensor([0.0, 0.0, 1.0, 1.0])       # valid but after the break -> ignored
    masks[0, :3] = 1
    assert uo.box_pixel_rects(boxes, masks, 16, 16) == [[(1, 8, 1, 8)]]
    assert uo.box_pixel_rects(boxes, masks, 64, 64)[0][2] == (0, 64, 0, 64)
    masks[0, 0] = 0   # sum(mask)=2 -> slots 0,1 considered regardless of which entries are set
    assert len(uo.box_pixel_rects(boxes, masks, 64, 64)[0]) == 2


def test_gated_self_attention_matches_reference(mods):
    o = mods["fuser"]
    sd = {"p." + k: v for k, v in uo.seeded_module_sd(o["shapes"], seed=22, gates=(0.5, -0.4)).items()}
    x = uo.seeded_randn((3, 120, 64), 31)
    objs = uo.seeded_randn((3, 30, 768), 33)
    assert rel(uo.gated_self_attention(sd, "p", x, objs, 8, 1.0), o["y_scale1"]) < 1e-5
    assert rel(uo.gated_self_attention(sd, "p", x, objs, 8, 0.35), o["y_scale035"]) < 1e-5
    assert torch.equal(uo.gated_self_attention(sd, "p", x, objs, 8, 0.0), x)


def test_position_net_matches_reference(mods):
    o = mods["posnet"]
    r = mods["rela"]
    sd = {"position_net." + k: v for k, v in uo.seeded_module_sd(o["shapes"], seed=23).items()}
    emb = uo.seeded_randn((3, 30, 768), 34)
    assert rel(uo.position_net(sd, r["boxes"], r["masks"], emb), o["y"]) < 1e-6
    z = uo.position_net(sd, torch.zeros(3, 30, 4), torch.zeros(3, 30), torch.zeros(3, 30, 768))
    assert rel(z, o["y_null"]) < 1e-6
    assert (z - z[0, 0]).abs().max() < 1e-6      # null grounding: one constant row


def test_timestep_embedding_matches_reference(mods):
    o = mods["temb"]
    assert rel(uo.timestep_embedding(o["t"], 320), o["y"]) < 1e-6


def test_tiny_unet_matches_reference():
    g = torch.load(os.path.join(GOLD, "tiny_unet.pt"), weights_only=False)
    cfg = g["cfg"]
    sd = uo.synthetic_state_dict(cfg, seed=g["seed"])
    syn = uo.synthetic_inputs(**g["syn_args"])
    b, i, box = g["box_override"]
    syn["grounding"]["boxes"][b, i] = torch.tensor(box)
    for t in (981, 1):
        for s in (1, 0):
            for name in ("cond", "unc"):
                inp = dict(x=syn["x"], timesteps=torch.full((2,), t, dtype=torch.long), relations=syn["relations"],
                           context=syn["context"] if name == "cond" else syn["uc"])
                if name == "cond":
                    inp["grounding_input"] = syn["grounding"]
                y = uo.unet_forward(sd, cfg, inp, scale=float(s))
                assert rel(y, g[f"eps_{name}_t{t}_s{s}"]) < 1e-5, (t, s, name)
    gg = torch.Generator().manual_seed(g["first_conv_seed"])
    fc = dict(weight=0.2 * torch.randn(64, 4, 3, 3, generator=gg), bias=0.02 * torch.randn(64, generator=gg))
    inp = dict(x=syn["x"], timesteps=torch.full((2,), 481, dtype=torch.long), relations=syn["relations"],
               context=syn["context"], grounding_input=syn["grounding"])
    assert rel(uo.unet_forward(sd, cfg, inp, scale=0.0, first_conv=fc), g["eps_cond_t481_s0_fc"]) < 1e-5


def test_state_dict_grammar_full_config():
    spec = uo.state_dict_spec(uo.default_unet_config())
    assert len(spec) == 1238                                    # SURVEY.md section 3.2 [probe]
    n = sum(int(np.prod(s)) if len(s) else 1 for _, s, _ in spec)
    assert n == 1_261_457_796


def test_plms_schedule_and_loop_match_reference():
    g = torch.load(os.path.join(GOLD, "plms.pt"), weights_only=False)
    acp = po.alphas_cumprod()
    assert torch.equal(acp, g["alphas_cumprod"])
    ts, a_t, a_prev, s1m = po.plms_tables(50, acp)
    assert np.array_equal(ts, g["ddim_timesteps"].numpy())
    assert np.allclose(a_t, g["ddim_alphas"].numpy(), rtol=0, atol=0)
    assert np.allclose(a_prev, g["ddim_alphas_prev"].numpy(), rtol=0, atol=0)
    assert np.allclose(s1m, g["ddim_sqrt_one_minus_alphas"].numpy(), rtol=1e-7)
    assert abs(a_t[49] - 0.0057755) < 1e-6 and abs(a_prev[0] - 0.99915) < 1e-5
    assert po.alpha_schedule(50) == [1] * 15 + [0] * 35

    import sys
    sys.path.insert(0, os.path.dirname(__file__))
    from gen_golden import stub_eps
    log = []

    def model(x, t, cond, scale, restored):
        log.append((int(t[0]), cond, float(scale), restored))
        return stub_eps(x, t.float().view(-1, 1, 1, 1) / 1000.0, cond, scale, restored)

    y = po.plms_sample(model, g["x0"].clone())
    assert rel(y, g["y"]) < 1e-6
    assert log == [tuple(e) for e in g["log"]]
    assert len(log) == 102 and sum(1 for e in log if e[2] == 1.0) == 32
