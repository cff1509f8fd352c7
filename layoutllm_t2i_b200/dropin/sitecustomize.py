"""Picked up by Python's `site` start-up when this directory is on PYTHONPATH: installs the drop-in import hook so that
the unmodified reference callers (txt2img.py, train_rl.py -> GLIGEN/interface.py) import the B200 implementations of
the hot-path modules whatever they do to sys.path afterwards (see _ltt_dropin_hook.py)."""
try:
    import _ltt_dropin_hook
    _ltt_dropin_hook.install()
except Exception as _ex:  # noqa: BLE001 -- never break interpreter start-up
    import sys
    print(f"[layoutllm_t2i_b200] drop-in import hook not installed: {_ex!r}", file=sys.stderr)
