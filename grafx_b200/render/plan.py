"""Render plan containers -- same field names as grafx.render.prepare (prepare.py:10-90), so a
`RenderData` produced by the reference's `prepare_render` and one built here are interchangeable
(render_grafx only reads attributes).  Building a plan from a graph (scheduling, node
relabelling: grafx.render.order / grafx.data) is host-side integer work outside the hot path;
`mixing_console_plan` covers the serial-chains-into-a-bus graphs of BASELINE config 5 and
`plan_from_dict` rebuilds any plan recorded from the reference (tests/golden/render_*.npz)."""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple, Union

import torch

IDX = Union[Tuple[int, ...], torch.Tensor, list]


@dataclass
class _TensorAccessData:
    method: str  # "slice" | "index" | "none"
    idx: IDX


@dataclass
class _AggregationData:
    method: str  # "none" | "sum" | "scatter"
    idx: Optional[IDX] = None


@dataclass
class _SingleRenderData:
    node_type: str
    source_reads: List[_TensorAccessData]
    aggregations: List[_AggregationData]
    parameter_read: _TensorAccessData
    dest_write: _TensorAccessData


@dataclass
class RenderData:
    method: str
    num_nodes: int
    max_order: int
    siso_only: bool
    iter_list: List[_SingleRenderData]


def _access(method, idx):
    if method == "index" and not isinstance(idx, torch.Tensor):
        idx = torch.tensor(list(idx), dtype=torch.long)
    elif method == "slice":
        idx = (int(idx[0]), int(idx[1]))
    return _TensorAccessData(method, idx)


def plan_from_dict(plan: dict, method: str = "beam") -> RenderData:
    iters = []
    for it in plan["iters"]:
        aggs = []
        for m, idx in it["aggs"]:
            aggs.append(_AggregationData(m, None if idx is None else torch.tensor(idx, dtype=torch.long)))
        iters.append(_SingleRenderData(it["type"], [_access(*a) for a in it["reads"]], aggs, _access(*it["param"]),
                                       _access(*it["write"])))
    return RenderData(method, int(plan["num_nodes"]), int(plan["max_order"]), True, iters)


def mixing_console_plan(num_tracks: int, chain: Sequence[str], method: str = "beam") -> RenderData:
    """Plan of `num_tracks` parallel chains  in -> chain[0] -> ... -> chain[-1]  summed into one
    `out` node: what compute_render_order + prepare_render produce for that graph (one render
    order per processor type, contiguous slices, a single `sum` at the bus).  Buffer layout:
    [inputs | chain[0] outputs | ... | chain[-1] outputs | out]."""
    T = int(num_tracks)
    iters = [_SingleRenderData("in", [_TensorAccessData("none", ())], [_AggregationData("none")],
                               _TensorAccessData("slice", (0, T)), _TensorAccessData("slice", (0, T)))]
    for i, t in enumerate(chain):
        iters.append(_SingleRenderData(t, [_TensorAccessData("slice", (i * T, (i + 1) * T))], [_AggregationData("none")],
                                       _TensorAccessData("slice", (0, T)),
                                       _TensorAccessData("slice", ((i + 1) * T, (i + 2) * T))))
    n = len(chain)
    iters.append(_SingleRenderData("out", [_TensorAccessData("slice", (n * T, (n + 1) * T))], [_AggregationData("sum")],
                                   _TensorAccessData("slice", (0, 1)),
                                   _TensorAccessData("slice", ((n + 1) * T, (n + 1) * T + 1))))
    return RenderData(method, (n + 1) * T + 1, n + 1, True, iters)
