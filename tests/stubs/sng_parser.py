"""Test stand-in for `sng_parser` (scene-graph parser, absent offline): "<a> on|near|under <b>" -> one relation."""


def parse(prompt):
    words = prompt.replace(",", " ").split()
    ents, rels = [], []
    for i, w in enumerate(words):
        if w in ("on", "near", "under", "beside") and 0 < i < len(words) - 1:
            ents += [dict(lemma_head=words[i - 1]), dict(lemma_head=words[i + 1])]
            rels.append(dict(subject=len(ents) - 2, object=len(ents) - 1, relation=w))
    return dict(entities=ents, relations=rels)
