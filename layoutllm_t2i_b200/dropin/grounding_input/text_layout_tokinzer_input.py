"""GroundingNetInput for the text+layout tokenizer (reference GLIGEN/grounding_input/text_layout_tokinzer_input.py:6-62).
Host-side pass-through; the tensors it returns are what UNetModel hands to the engine's conditioning call."""
import torch


class GroundingNetInput:
    def __init__(self):
        self.set = False
        self.in_dim = 768

    def prepare(self, batch, text_encoder=None):
        boxes, masks = batch["boxes"], batch["masks"]
        self.set = True
        self.batch, self.max_box = boxes.shape[0], boxes.shape[1]
        self.device = boxes.device
        if "text_embeddings" in batch:
            emb = batch["text_embeddings"]
        else:   # per-phrase CLIP tokens (reference :26-38); needs the caller's text encoder
            emb = torch.zeros(self.batch, self.max_box, self.in_dim, device=self.device)
            counts = masks.sum(dim=-1).tolist()
            for b, line in enumerate(batch["labels"]):
                phrases = line.split("|")
                for i in range(int(counts[b])):
                    if i < len(phrases):
                        emb[b, i] = text_encoder.encode_one_token(phrases[i])
        self.dtype = emb.dtype
        return {"boxes": boxes, "masks": masks, "positive_embeddings": emb}

    def get_null_input(self, batch=None, device=None, dtype=None):
        assert self.set, "not set yet, cannot call this funcion"
        batch = self.batch if batch is None else batch
        device = self.device if device is None else device
        dtype = self.dtype if dtype is None else dtype
        z = lambda *shape: torch.zeros(*shape, dtype=dtype, device=device)
        return {"boxes": z(batch, self.max_box, 4), "masks": z(batch, self.max_box),
                "positive_embeddings": z(batch, self.max_box, self.in_dim)}
