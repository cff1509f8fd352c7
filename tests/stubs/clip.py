"""Test stand-in for OpenAI `clip` (absent offline); never called on the hot path."""
