"""Test stand-in for `kornia` (absent offline): imported at module level by the reference's ldm/modules/encoders/modules.py,
used only by FrozenClipImageEmbedder (not on the LayoutLLM-T2I path)."""
