"""Import hook of the drop-in tree: the module names this tree owns always resolve here.

`txt2img.py` appends `./GLIGEN` to `sys.path` (txt2img.py:15), so PYTHONPATH order is enough for it; but
`GLIGEN/interface.py` (the entry point of `train_rl.py`, train_rl.py:19) does `sys.path.insert(0, dirname(__file__))`
(interface.py:2), which puts the reference's own `ldm/` AHEAD of anything on PYTHONPATH.  A meta-path finder is
consulted before `sys.path`, so the callers stay unmodified: every `ldm.*` / `grounding_input.*` leaf module that exists
in this directory is served from here, everything else (VAE internals, CLIP encoder, other GLIGEN modalities ...) keeps
resolving through the normal path search to the reference's files.

Installed automatically by `sitecustomize.py` when this directory is on PYTHONPATH at interpreter start-up, or
explicitly with `import _ltt_dropin_hook; _ltt_dropin_hook.install()`.
"""
import importlib.abc
import importlib.util
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOTS = ("ldm", "grounding_input")


def _owned():
    out = {}
    for root in _ROOTS:
        for d, _, files in os.walk(os.path.join(_HERE, root)):
            for f in files:
                if f.endswith(".py") and f != "__init__.py":
                    rel = os.path.relpath(os.path.join(d, f), _HERE)[:-3]
                    out[rel.replace(os.sep, ".")] = os.path.join(d, f)
    return out


class DropinFinder(importlib.abc.MetaPathFinder):
    def __init__(self):
        self.owned = _owned()

    def find_spec(self, fullname, path=None, target=None):
        file = self.owned.get(fullname)
        if file is None:
            return None
        return importlib.util.spec_from_file_location(fullname, file)


def install() -> DropinFinder:
    for f in sys.meta_path:
        if isinstance(f, DropinFinder):
            return f
    if _HERE not in sys.path:          # `_ltt_fallthrough` and the namespace portions of ldm/ live here
        sys.path.insert(0, _HERE)
    finder = DropinFinder()
    sys.meta_path.insert(0, finder)
    return finder
