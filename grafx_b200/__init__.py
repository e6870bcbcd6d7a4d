"""grafx_b200 -- B200-native (sm_100a CUDA) drop-in for the hot path of sh-lee97/grafx:
`grafx.processors.*` forward passes and `grafx.render.render_grafx`.

Layout:
  csrc/        hand-written CUDA kernels + the C ABI (include/grafx_b200.h)
  _cabi.py     ctypes binding of that ABI
  functional.py  tensor-level wrappers around the ABI entry points
  processors/  nn.Modules mirroring grafx.processors (same names / kwargs / forward signatures)
  render/      render_grafx over a RenderData plan, batch sharding across GPUs
"""
__version__ = "0.1.0"
