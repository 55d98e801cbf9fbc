"""Operator namespace (the reference's det3d/ops/__init__.py:1-10 is entirely commented out as well)."""
