"""CUDA-graph capture of a whole render plan.

`render_grafx` issues a few dozen launches per call (design kernels, tables, the processors, node sums, the
parameter expansions); for short audio or small batches the host-side cost of issuing them exceeds the device time.
A render plan is static -- same graph, same shapes, same order every step of a training / inference loop -- so the
launch sequence is captured once into a CUDA graph and replayed with one `cudaGraphLaunch`.

Inputs live in static device tensors (`.input_signals`, `.parameters`); `__call__` copies the caller's tensors
into them on the current stream, replays the graph and returns the static outputs (valid until the next call).
Upstream has no equivalent (its render loop is eager PyTorch, render/graph.py:77-177).
"""
from __future__ import annotations

from typing import Mapping

import torch

from .graph import _map_tensors, render_grafx


def _copy_tree(dst, src):
    if isinstance(dst, torch.Tensor):
        dst.copy_(src, non_blocking=True)
        return
    assert dst.keys() == src.keys(), (sorted(dst.keys()), sorted(src.keys()))
    for k in dst:
        _copy_tree(dst[k], src[k])


class CapturedRender:
    """Captures `render_grafx(processors, input_signals, per_type_parameters, render_data, common_parameters)`.

    `input_signals` / the parameter trees given here fix shapes, dtypes and the device; their values are the
    first inputs.  Every processor the plan uses must be a grafx_b200 processor on that device (anything that
    synchronises with the host inside `forward` cannot be captured and raises at construction)."""

    def __init__(self, processors: Mapping, input_signals: torch.Tensor, per_type_parameters: Mapping, render_data,
                 common_parameters=None, warmup: int = 2):
        if not input_signals.is_cuda:
            raise RuntimeError("CapturedRender needs CUDA tensors (there is no CPU path)")
        self.processors = processors
        self.render_data = render_data
        clone = lambda t: t.detach().clone()  # noqa: E731
        self.input_signals = clone(input_signals)
        self.parameters = _map_tensors(per_type_parameters, clone)
        self.common_parameters = None if common_parameters is None else _map_tensors(common_parameters, clone)
        self.device = input_signals.device

        def run():
            return render_grafx(self.processors, self.input_signals, self.parameters, self.render_data,
                                common_parameters=self.common_parameters)

        with torch.cuda.device(self.device):
            # plans, tables and per-device constants are built on first use: do that outside the capture
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(max(1, warmup)):
                    run()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            from .. import _cabi

            n0 = _cabi.lib().gfx_kernel_launch_count()
            run()
            torch.cuda.synchronize()
            #: kernels of libgrafx_b200 inside one replay (the library's own launch counter over one eager pass)
            self.launches_per_replay = int(_cabi.lib().gfx_kernel_launch_count() - n0)
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self.output_signals, self.intermediates_list, self.signal_buffer = run()

    def __call__(self, input_signals: torch.Tensor | None = None, per_type_parameters: Mapping | None = None,
                 common_parameters=None):
        if input_signals is not None:
            assert input_signals.shape == self.input_signals.shape, "the captured plan has fixed shapes"
            self.input_signals.copy_(input_signals, non_blocking=True)
        if per_type_parameters is not None:
            _copy_tree(self.parameters, per_type_parameters)
        if common_parameters is not None:
            _copy_tree(self.common_parameters, common_parameters)
        self.graph.replay()
        return self.output_signals, self.intermediates_list, self.signal_buffer
