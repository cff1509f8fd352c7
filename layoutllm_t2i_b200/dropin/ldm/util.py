"""`ldm.util` of the drop-in tree: the config -> object lookup the checkpoint's saved config relies on
(reference GLIGEN/ldm/util.py:71-86).  Other helper names fall through to the reference's file."""
import importlib

import _ltt_fallthrough


def get_obj_from_str(string, reload=False):
    module_name, attr = string.rsplit(".", 1)
    module = importlib.import_module(module_name)
    if reload:
        module = importlib.reload(module)
    return getattr(module, attr)


def instantiate_from_config(config):
    if "target" not in config:
        if config in ("__is_first_stage__", "__is_unconditional__"):
            return None
        raise KeyError("Expected key `target` to instantiate.")
    return get_obj_from_str(config["target"])(**config.get("params", dict()))


def default(val, d):
    if val is not None:
        return val
    return d() if callable(d) else d


_ltt_fallthrough.install(__name__, globals())
