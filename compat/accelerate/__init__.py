"""Shim for `accelerate` (only `accelerate.utils.set_seed` is used by inference_ID-Booth.py:8)."""
