"""Host logic of the conditioning prep and of the reward mirror against the UNMODIFIED reference functions, on CPU.

`prepare_conditioning` / `relation_phrases` (layoutllm_t2i_b200/clip.py) against `prepare_batch`,
`prepare_relation_phrases` and the `text_encoder.encode` calls of `generate_one_image` (txt2img.py:173-244,268-277, imported
from its byte-identical copy under oracle/_ref); `Reward.nn_close_set` / `label_to_id` (layoutllm_t2i_b200/reward.py) against
models/policy.py:77-102.  Both sides get the SAME deterministic stand-in for the text tower (a function of the token ids up
to the end-of-text token only -- the property the batched pass relies on), so the outputs must be identical, not close.
"""
import os
import sys
from types import SimpleNamespace

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")
sys.path.insert(0, os.path.join(ROOT, "tests", "stubs"))

needs_ref = pytest.mark.skipif(not os.path.isfile(os.path.join(REF, "txt2img.py")),
                               reason="oracle/_ref not staged (python oracle/make_ref.py in the build container)")
W = 768


def _row_vec(ids_row, salt):
    """Deterministic [W] vector from the tokens up to (and including) the first end-of-text token."""
    ids_row = [int(v) for v in ids_row]
    eos = max(ids_row)
    cut = ids_row[:ids_row.index(eos) + 1]
    seed = (hash((salt,) + tuple(cut)) & 0x7FFFFFFF)
    return torch.randn(W, generator=torch.Generator().manual_seed(seed))


class FakeTower:
    """ClipTextEncoder's surface (`encode_ids`) without a GPU."""
    device = torch.device("cpu")
    cfg = dict(projection_dim=0)

    def encode_ids(self, ids, want_hidden=True, want_embeds=False):
        hid = torch.stack([torch.stack([_row_vec(r, ("h", t)) for t in range(len(r))]) for r in ids.tolist()]) if want_hidden else None
        pooled = torch.stack([_row_vec(r, "p") for r in ids.tolist()])
        return hid, pooled


class FakeEmbedder:
    """FrozenCLIPEmbedder's surface on the fake tower (what the reference functions call)."""

    def __init__(self, tok):
        self.tok, self.tower = tok, FakeTower()

    def encode(self, text, return_pooler_output=False):
        ids = self.tok(text, truncation=True, max_length=77, padding="max_length")["input_ids"]
        z, p = self.tower.encode_ids(ids)
        return (z, p) if return_pooler_output else z


class FakeClipModel:
    def __init__(self):
        self.tower = FakeTower()

    def __call__(self, input_ids=None, attention_mask=None, pixel_values=None, **kw):
        return SimpleNamespace(text_model_output=SimpleNamespace(pooler_output=self.tower.encode_ids(input_ids, want_hidden=False)[1]))

    def get_text_features(self, input_ids=None, attention_mask=None, **kw):
        return self.tower.encode_ids(input_ids, want_hidden=False)[1]


def _txt2img():
    sys.path.insert(0, os.path.dirname(__file__))
    import test_callers as tc
    return tc._import_callers()[0]


@needs_ref
@pytest.mark.parametrize("n_boxes,prompt,with_none", [
    (3, "a cat on a sofa near a lamp", False),
    (0, "an empty street at night", False),
    (30, "a dog under a table beside a chair near a window on a rug", True),
    (5, "a plain prompt without any relation word", False),
])
def test_prepare_conditioning_equals_reference_functions(n_boxes, prompt, with_none):
    import sng_parser
    from ltt_test_stubs import HashProcessor, HashTokenizer
    from layoutllm_t2i_b200.clip import prepare_conditioning, relation_phrases
    txt2img = _txt2img()
    tok = HashTokenizer()
    g = torch.Generator().manual_seed(n_boxes)
    phrases = ["thing %d" % i for i in range(n_boxes)]
    if with_none:
        phrases[7] = None
    boxes = [torch.rand(4, generator=g).tolist() for _ in range(n_boxes)]
    if n_boxes == 5:
        boxes = boxes[:3]                  # fewer locations than phrases: prepare_batch's zip stops at the shorter list
    batch, max_relas = 2, 10
    # ---- the reference's own sequence (txt2img.py:268-277)
    emb = FakeEmbedder(tok)
    meta = dict(prompt=prompt, phrases=phrases, locations=boxes)
    ref_batch = txt2img.prepare_batch(meta, FakeClipModel(), HashProcessor(), batch, device="cpu")
    ref_context = emb.encode([prompt] * batch)
    ref_rel = txt2img.prepare_relation_phrases(prompt, batch, max_relas, emb, device="cpu")
    ref_uc = emb.encode(batch * [""])
    # ---- one batched pass
    rels = relation_phrases(sng_parser.parse(prompt), max_relas)
    out = prepare_conditioning(FakeTower(), lambda t: tok(t, truncation=True, max_length=77, padding="max_length")["input_ids"],
                               prompt, phrases, boxes, rels, batch=batch, max_relas=max_relas)
    assert torch.equal(out["context"], ref_context) and torch.equal(out["uc"], ref_uc)
    assert torch.equal(out["relations"], ref_rel)
    for k in ("boxes", "masks", "text_masks", "image_masks", "text_embeddings", "image_embeddings"):
        assert out[k].shape == ref_batch[k].shape and torch.equal(out[k], ref_batch[k]), k


@needs_ref
def test_relation_phrase_list_equals_reference_loop():
    """The strings the reference's two loops hand to text_encoder.encode (txt2img.py:218-238), captured from the call."""
    import sng_parser
    from layoutllm_t2i_b200.clip import relation_phrases
    txt2img = _txt2img()
    seen = []

    class Spy:
        def encode(self, texts, return_pooler_output=False):
            seen.append(list(texts))
            return torch.zeros(len(texts), 77, W), torch.zeros(len(texts), W)
    for prompt, max_relas in (("a cat on a sofa near a lamp", 10), ("a b on c near d under e beside f", 5), ("nothing here", 5)):
        seen.clear()
        txt2img.prepare_relation_phrases(prompt, 1, max_relas, Spy(), device="cpu")
        want = seen[0] if seen else []
        assert relation_phrases(sng_parser.parse(prompt), max_relas) == want, prompt


@needs_ref
def test_reward_label_logic_equals_reference():
    """nn_close_set / label_to_id (models/policy.py:77-102) on the reference class itself (constructed without
    from_pretrained) and on the mirror, sharing one stand-in text tower."""
    from ltt_test_stubs import HashTokenizer
    from oracle import ref_loader as rl
    import layoutllm_t2i_b200.reward as rw
    with rl.reference_tree():
        sys.path.insert(0, rl.REF_ROOT)
        try:
            from models.policy import Reward as RefReward
        finally:
            sys.path.remove(rl.REF_ROOT)
    tok = HashTokenizer()

    class Tok:       # AutoTokenizer returns a BatchEncoding with .to(); the mirror and the reference both index / ** it
        def __call__(self, texts, padding=True, return_tensors="pt"):
            d = tok(texts, padding=padding)

            class Enc(dict):
                def to(self, device):
                    return self
            return Enc(d)
    ref = object.__new__(RefReward)
    torch.nn.Module.__init__(ref)
    ref.tokenizer, ref.model, ref.device = Tok(), FakeClipModel(), "cpu"
    ref.args = SimpleNamespace(img_dir="x/train2014")
    ref.emb_labels()
    mine = object.__new__(rw.Reward)
    mine.text, mine.tokenizer, mine.labels = FakeTower(), Tok(), list(rw.COCO_LABELS)
    mine.text.encode_ids = lambda ids, want_hidden=False, want_embeds=True: (None, None, FakeTower().encode_ids(ids, want_hidden=False)[1])
    mine.label2index = {l: i for i, l in enumerate(mine.labels)}
    mine.emb_labels()
    assert mine.labels == ref.labels and torch.equal(mine.labels_emb, ref.labels_emb)
    layouts = [([[0.1, 0.1, 0.4, 0.5], [0.2, 0.3, 0.9, 0.8]], ["person", "a racing bike"]),
               ([[0.0, 0.0, 1.0, 1.0]], ["sofa in a living room"]), ([], [])]
    got, want = mine.nn_close_set(layouts), ref.nn_close_set(layouts)
    assert [l for _, l in got] == [l for _, l in want]
    gi, wi = mine.label_to_id(got), ref.label_to_id(want)
    assert all((a[0] == b[0]).all() and (a[1] == b[1]).all() for a, b in zip(gi, wi))
