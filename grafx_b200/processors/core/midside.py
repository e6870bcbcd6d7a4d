"""lr_to_ms / ms_to_lr (grafx/processors/core/midside.py:4-17) on the CUDA kernel."""
from ...functional import lr_to_ms, ms_to_lr  # noqa: F401
