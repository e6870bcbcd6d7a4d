"""render_grafx -- drop-in for grafx.render.graph.render_grafx (render/graph.py:16-177) with the
tensor primitives of render/core.py (:6-140) folded in.

Same signature and return value: `(output_signals, intermediates_list, signal_buffer)`; accepts
3-D `[|V0|, C, L]` or 4-D `[B, |V0|, C, L]` sources and any RenderData-like plan (the reference's
own objects work: only attributes are read).  `parameters_grad` / `input_signal_grad` are accepted for compatibility.
Under `torch.no_grad()` (or with detached tensors) the loop writes into one signal buffer in place; when autograd
expects a gradient (a tensor that requires grad, grad mode on) the same plan runs functionally
(`_render_grafx_functional`): differentiable for the processors that have a backward pass, a loud error for the rest.

What runs where: processors are the CUDA kernels of this package; node-axis aggregation
(`sum` / `scatter`) is csrc/elementwise.cu:node_sum_kernel reading and writing slices of the
signal buffer in place; slice reads are views; the remaining copies are torch device copies.
"""
from __future__ import annotations

from typing import Mapping

import torch
import torch.nn as nn

from .. import functional as F_

UTILITY_TYPES = ("in", "out", "mix")  # grafx/data/configs.py


def _read(x, access, dim):
    """read_tensor_or_tensor_dict (render/core.py:36-73)."""
    if isinstance(x, torch.Tensor):
        if access.method == "slice":
            return x.narrow(dim, access.idx[0], access.idx[1] - access.idx[0])
        if access.method == "index":
            return x.index_select(dim, access.idx.to(x.device))
        raise Exception(f"The provided read method is not available: {access.method}.")
    if isinstance(x, (dict, nn.ParameterDict, nn.ModuleDict)):
        return {k: _read(v, access, dim) for k, v in x.items()}
    if isinstance(x, list):
        return x[access.idx[0]]
    raise TypeError(type(x))


def _map_tensors(x, fn):
    if isinstance(x, torch.Tensor):
        return fn(x)
    return {k: _map_tensors(v, fn) for k, v in x.items()}


def _any_requires_grad(x) -> bool:
    if x is None:
        return False
    if isinstance(x, torch.Tensor):
        return x.requires_grad
    if isinstance(x, (list, tuple)):
        return any(_any_requires_grad(v) for v in x)
    return any(_any_requires_grad(v) for v in x.values())


def _flatten2(x):
    return x.reshape(-1, *x.shape[2:])


def render_grafx(processors: Mapping, input_signals: torch.Tensor, per_type_parameters: Mapping, render_data,
                 common_parameters=None, parameters_grad=True, input_signal_grad=False):
    """Layout: the signal buffer is allocated NODE-major, `[|V|, (B,) C, L]`, so that the nodes of a render
    order are one contiguous block of rows: processors read their inputs as views of the buffer and (ours)
    write their outputs straight into the destination slice -- no gather / scatter copies.  For 4-D sources
    the returned output and buffer are the `[B, ...]`-first views of those tensors (same shape and values as
    upstream, render/graph.py:63-75,177; they are not contiguous)."""
    method = render_data.method
    ndim = input_signals.ndim
    if ndim == 3:
        batch_size = None
        num_sources, channels, audio_len = input_signals.shape
    elif ndim == 4:
        batch_size, num_sources, channels, audio_len = input_signals.shape
        # node-major flattening of the batch: row = node * B + b (upstream: b * n + node; the order of the
        # rows inside a render order is an implementation detail, signals and parameters use the same one)
        expand = lambda t: t.detach().unsqueeze(1).expand(t.shape[0], batch_size, *t.shape[1:])  # noqa: E731
    else:
        raise Exception(f"input_signal has shape of {input_signals.shape} ({ndim} ndims), which is not 3 or 4 dims.")
    node_dim = 0
    post = _flatten2 if ndim == 4 else (lambda t: t)
    batched = (lambda t: post(expand(t))) if ndim == 4 else (lambda t: t)
    if torch.is_grad_enabled() and (input_signals.requires_grad or _any_requires_grad(per_type_parameters)
                                    or _any_requires_grad(common_parameters)):
        # the loop below writes processor outputs into slices of one signal buffer in place, which records nothing for
        # autograd: training mode takes the functional path instead (no in-place write; processors without a backward
        # pass raise there -- a loud error beats a result whose gradients are silently zero)
        return _render_grafx_functional(processors, input_signals, per_type_parameters, render_data, common_parameters)
    input_signals = input_signals.detach()

    # create_signal_buffer (render/core.py:6-33)
    one_by_one = method == "one-by-one"
    pending_sources = None
    if one_by_one:
        assert ndim == 3, "the one-by-one list buffer has no batch axis upstream either"
        signal_buffer = [x[None] for x in input_signals] + [None] * (render_data.num_nodes - num_sources)
    else:
        shape = (render_data.num_nodes, channels, audio_len) if ndim == 3 else (render_data.num_nodes, batch_size, channels, audio_len)
        signal_buffer = torch.empty(shape, device=input_signals.device, dtype=torch.float32)
        sources = signal_buffer.narrow(0, 0, num_sources)
        x_in = input_signals if input_signals.dtype == torch.float32 else input_signals.float()
        if ndim == 4 and (x_in.stride(3) != 1 or x_in.stride(2) != audio_len):
            x_in = x_in.contiguous()

        def fill_sources():
            if ndim == 3:
                sources.copy_(x_in)
            else:
                F_.node_copy(x_in.transpose(0, 1), sources)

        if x_in.is_contiguous() and x_in.numel() > 0 and _first_order_folds(processors, render_data, num_sources):
            # the first processor reads the caller's sources and fills this slice itself (F_.source_fold)
            pending_sources = x_in if ndim == 4 else x_in.unsqueeze(0)
        else:
            fill_sources()

    intermediates_list = []
    output_signals = None
    for i in range(1, int(render_data.max_order) + 1):
        it = render_data.iter_list[i]
        node_type = it.node_type
        is_proc = node_type in processors
        if not is_proc and node_type not in UTILITY_TYPES:
            raise Exception(f"Wrong node type given: {node_type}")
        dest = it.dest_write
        dest_view = None
        if (not one_by_one) and dest.method == "slice":
            dest_view = signal_buffer.narrow(0, dest.idx[0], dest.idx[1] - dest.idx[0])
        direct_dest = dest_view if ((not is_proc) and len(it.source_reads) == 1) else None

        inputs = []
        wrote_direct = False
        for read, agg in zip(it.source_reads, it.aggregations):
            src = _read(signal_buffer, read, node_dim)
            if agg.method == "sum":
                out_view = direct_dest if (direct_dest is not None and direct_dest.shape[0] == 1) else None
                src = _node_sum(src, None, 1, out_view, ndim)
                wrote_direct = out_view is not None
            elif agg.method == "scatter":
                n_dst = int(agg.idx.max()) + 1
                out_view = direct_dest if (direct_dest is not None and direct_dest.shape[0] == n_dst) else None
                src = _node_sum(src, agg.idx, n_dst, out_view, ndim)
                wrote_direct = out_view is not None
            elif agg.method != "none":
                raise Exception(f"The provided aggregation method is not available: {agg.method}.")
            inputs.append(post(src))

        if is_proc:
            # 4-D sources: every node's parameter rows repeat over the batch of renders; a processor that says so
            # (`accepts_parameter_repeat`) gets them un-expanded inside `shared_parameters(B)` and shares each row over a run
            # of B batch items in its kernels, the others get the expansion upstream makes (render/graph.py:132-147)
            share = ndim == 4 and common_parameters is None
            plain = share and getattr(processors[node_type], "accepts_parameter_repeat", lambda: False)()
            parameters = _map_tensors(_read(per_type_parameters[node_type], it.parameter_read, 0),
                                      (lambda t: t.detach()) if plain else batched)
            common_i = {}
            if common_parameters is not None:
                common_i = _map_tensors(_read(common_parameters, dest, 0), batched)
                if isinstance(common_i, torch.Tensor):
                    common_i = {"parameter": common_i}
            if isinstance(parameters, torch.Tensor):
                parameters = {"parameter": parameters}
            def run():
                with F_.output_into(post(dest_view) if dest_view is not None else None), \
                        F_.shared_parameters(batch_size if share else 1):
                    return processors[node_type](*inputs, **parameters, **common_i)

            if pending_sources is not None:
                with F_.source_fold(pending_sources, post(sources)) as fold:
                    output = run()
                pending_sources = None
                if not fold.used:
                    # the processor did not take the offer (its first op is not the cascade kernel for these shapes):
                    # fill the slice the usual way, run it again, and do not offer again to this module
                    fill_sources()
                    processors[node_type]._gfx_no_source_fold = True
                    output = run()
            else:
                output = run()
            if isinstance(output, tuple):
                output_signals, intermediates = output
                intermediates_list.append(intermediates)
            else:
                output_signals = output
        else:
            output_signals = inputs

        if isinstance(output_signals, list):
            if len(output_signals) == 1:
                output_signals = output_signals[0]
            elif ndim == 3:
                output_signals = torch.stack(output_signals, -3).view(-1, channels, audio_len)
            else:  # same node order as upstream ((node, input) pairs), batch axis kept second
                output_signals = torch.stack([o.view(-1, batch_size, channels, audio_len) for o in output_signals], 1)
        if ndim == 4:
            output_signals = output_signals.reshape(-1, batch_size, channels, audio_len)

        # inplace_write_tensor (render/core.py:80-98)
        if one_by_one:
            signal_buffer[dest.idx[0]] = output_signals
        elif wrote_direct or (dest_view is not None and output_signals.data_ptr() == dest_view.data_ptr()
                              and output_signals.shape == dest_view.shape):
            pass  # the kernel already wrote the slice of the buffer; output_signals is that view
        elif dest.method == "slice":
            dest_view.copy_(output_signals)
        elif dest.method == "index":
            signal_buffer.index_copy_(0, dest.idx.to(signal_buffer.device), output_signals)
        else:
            raise Exception(f"The provided inplace write method is not available: {dest.method}.")
    if ndim == 4:
        return output_signals.transpose(0, 1), intermediates_list, signal_buffer.transpose(0, 1)
    return output_signals, intermediates_list, signal_buffer


def _first_order_folds(processors, render_data, num_sources) -> bool:
    """True when the first render order is ONE processor call on exactly the source slice and that processor starts
    with the biquad cascade on its input (`folds_source_read`): the cascade kernel then reads the caller's sources and
    writes the buffer's source slice on the way (gfx_biquad_cascade_ex_f32) instead of a separate copy pass."""
    if int(render_data.max_order) < 1:
        return False
    it = render_data.iter_list[1]
    proc = processors[it.node_type] if it.node_type in processors else None
    if proc is None or getattr(proc, "_gfx_no_source_fold", False):
        return False
    wants = getattr(proc, "folds_source_read", None)
    if wants is None or not wants():
        return False
    if len(it.source_reads) != 1 or it.aggregations[0].method != "none":
        return False
    read = it.source_reads[0]
    return read.method == "slice" and int(read.idx[0]) == 0 and int(read.idx[1]) == num_sources


def _render_grafx_functional(processors, input_signals, per_type_parameters, render_data, common_parameters):
    """render_grafx for training mode: the same plan evaluated without in-place writes, so that autograd sees every step
    (upstream clones its reads for the same reason, render/graph.py:108).  Node signals are kept as the list of blocks
    the render orders produced (node-major, `[nodes, (B,) C, L]`); a read is a block, a narrow of one, or -- irregular
    plans -- an index into their concatenation; aggregation is torch's `sum` / `index_add`; processors get fresh
    outputs (those with a backward pass, grafx_b200/autograd.py, carry the graph; the others raise).  Returns the same
    triple as the forward path; the signal buffer is the concatenation of the blocks."""
    method = render_data.method
    ndim = input_signals.ndim
    if ndim == 4:
        batch_size, num_sources, channels, audio_len = input_signals.shape
        expand = lambda t: t.unsqueeze(1).expand(t.shape[0], batch_size, *t.shape[1:])  # noqa: E731
        flat = lambda t: t.reshape(-1, *t.shape[2:])  # noqa: E731
        batched = lambda t: flat(expand(t))  # noqa: E731
        src0 = input_signals.to(torch.float32).transpose(0, 1)      # node-major view [V0, B, C, L]
    elif ndim == 3:
        batch_size = None
        num_sources, channels, audio_len = input_signals.shape
        flat = batched = lambda t: t  # noqa: E731
        src0 = input_signals.to(torch.float32)
    else:
        raise Exception(f"input_signal has shape of {input_signals.shape} ({ndim} ndims), which is not 3 or 4 dims.")
    num_nodes = int(render_data.num_nodes)
    blocks = {0: src0}                      # first node index -> block of consecutive nodes
    owner = [0] * num_sources + [None] * (num_nodes - num_sources)

    def whole_buffer():
        parts, i = [], 0
        while i < num_nodes:
            if owner[i] is None:  # nodes not rendered yet read as zeros (upstream's buffer starts as zeros)
                j = i
                while j < num_nodes and owner[j] is None:
                    j += 1
                parts.append(src0.new_zeros((j - i,) + tuple(src0.shape[1:])))
                i = j
            else:
                b = blocks[owner[i]]
                parts.append(b.narrow(0, i - owner[i], min(b.shape[0] - (i - owner[i]), num_nodes - i)))
                i += parts[-1].shape[0]
        return torch.cat(parts, 0)

    def read(access):
        if access.method == "slice":
            a, b = int(access.idx[0]), int(access.idx[1])
            o = owner[a]
            if o is not None and all(owner[i] == o for i in range(a, b)):
                return blocks[o].narrow(0, a - o, b - a)
            return whole_buffer().narrow(0, a, b - a)
        if access.method == "index":
            return whole_buffer().index_select(0, access.idx.to(src0.device))
        raise Exception(f"The provided read method is not available: {access.method}.")

    def write(access, value):
        if access.method == "slice":
            a, b = int(access.idx[0]), int(access.idx[1])
            assert value.shape[0] == b - a
            blocks[a] = value
            for i in range(a, b):
                owner[i] = a
        elif access.method == "index":
            idx = access.idx.tolist()
            for k, node in enumerate(idx):
                blocks[node] = value.narrow(0, k, 1)
                owner[node] = node
        else:
            raise Exception(f"The provided inplace write method is not available: {access.method}.")

    intermediates_list, output_signals = [], None
    for i in range(1, int(render_data.max_order) + 1):
        it = render_data.iter_list[i]
        node_type = it.node_type
        is_proc = node_type in processors
        if not is_proc and node_type not in UTILITY_TYPES:
            raise Exception(f"Wrong node type given: {node_type}")
        inputs = []
        for rd, agg in zip(it.source_reads, it.aggregations):
            src = read(rd)
            if agg.method == "sum":
                src = src.sum(0, keepdim=True)
            elif agg.method == "scatter":
                idx = agg.idx.to(src.device)
                n_dst = int(agg.idx.max()) + 1
                src = src.new_zeros((n_dst,) + tuple(src.shape[1:])).index_add(0, idx, src)
            elif agg.method != "none":
                raise Exception(f"The provided aggregation method is not available: {agg.method}.")
            inputs.append(flat(src).contiguous())
        if is_proc:
            parameters = _map_tensors(_read(per_type_parameters[node_type], it.parameter_read, 0), batched)
            common_i = {}
            if common_parameters is not None:
                common_i = _map_tensors(_read(common_parameters, it.dest_write, 0), batched)
                if isinstance(common_i, torch.Tensor):
                    common_i = {"parameter": common_i}
            if isinstance(parameters, torch.Tensor):
                parameters = {"parameter": parameters}
            output = processors[node_type](*inputs, **parameters, **common_i)
            if isinstance(output, tuple):
                output_signals, intermediates = output
                intermediates_list.append(intermediates)
            else:
                output_signals = output
        else:
            output_signals = inputs
        if isinstance(output_signals, list):
            if len(output_signals) == 1:
                output_signals = output_signals[0]
            elif ndim == 3:
                output_signals = torch.stack(output_signals, -3).view(-1, channels, audio_len)
            else:
                output_signals = torch.stack([o.view(-1, batch_size, channels, audio_len) for o in output_signals], 1)
        if ndim == 4:
            output_signals = output_signals.reshape(-1, batch_size, channels, audio_len)
        write(it.dest_write, output_signals)
    if method == "one-by-one":
        assert ndim == 3
        buf = whole_buffer()
        return output_signals, intermediates_list, [buf[n:n + 1] for n in range(num_nodes)]
    signal_buffer = whole_buffer()
    if ndim == 4:
        return output_signals.transpose(0, 1), intermediates_list, signal_buffer.transpose(0, 1)
    return output_signals, intermediates_list, signal_buffer


def _node_sum(src, index, n_dst, out_view, ndim):
    """node-axis (dim 0) sum / scatter-sum of a node-major view [Q, (B,) C, L] -> [n_dst, (B,) C, L]."""
    if ndim == 3:
        return F_.node_sum(src, 0, index, n_dst, out=out_view)
    # [Q, B, C, L]: the kernel wants [batch, node, C*L] strides -> hand it the transposed views
    out_t = out_view.transpose(0, 1) if out_view is not None else None
    res = F_.node_sum(src.transpose(0, 1), 1, index, n_dst, out=out_t)
    return res.transpose(0, 1)
