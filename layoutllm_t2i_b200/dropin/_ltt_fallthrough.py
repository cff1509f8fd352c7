"""Name fall-through for the drop-in tree: a module that shadows a reference file resolves the names it does not
define itself from the reference's file of the same relative path, if one is importable further down sys.path.
(The reference's VAE, for instance, imports `LinearAttention` from `ldm.modules.attention`.)"""
import importlib.util
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))


def install(module_name: str, module_globals: dict) -> None:
    rel = os.path.join(*module_name.split(".")) + ".py"
    cache = {}

    def _shadowed():
        if cache.get("mod") is None:          # a miss is not cached: sys.path may gain the reference later
            for base in sys.path:
                base = os.path.abspath(base or ".")
                cand = os.path.join(base, rel)
                if base != _HERE and os.path.isfile(cand):
                    spec = importlib.util.spec_from_file_location("_ltt_shadowed_." + module_name, cand)
                    mod = importlib.util.module_from_spec(spec)
                    spec.loader.exec_module(mod)
                    cache["mod"] = mod
                    break
        return cache.get("mod")

    def __getattr__(name):
        if name.startswith("__"):
            raise AttributeError(name)
        mod = _shadowed()
        if mod is not None and hasattr(mod, name):
            return getattr(mod, name)
        raise AttributeError(f"module {module_name!r} has no attribute {name!r} "
                             f"(not part of the B200 hot path and no reference copy on sys.path)")

    module_globals["__getattr__"] = __getattr__
