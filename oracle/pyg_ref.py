"""Plain-torch CPU restatement of the torch_geometric 2.7.0 pieces segger composes.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  PARITY UNPINNED: the
torch_geometric source is not available offline; semantics follow SURVEY.md
Appendix A.1-A.3/A.6 and the reference call sites
(/root/reference/src/segger/models/ist_encoder.py:109-134, 261, 282-286).

Everything here is written with the same ATen primitives PyG dispatches to on
CPU (index_select, scatter_reduce_, index_add_, mm) so that timing it is a fair
"reference CPU path" baseline.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import torch
from torch import Tensor
from torch.nn import functional as F

EdgeType = Tuple[str, str, str]


def mangle(key) -> str:
    """PyG ``ModuleDict.to_internal_key``: tuple keys -> '<a___b___c>' (Appendix A.6)."""
    if isinstance(key, tuple):
        return "<" + "___".join(key) + ">"
    return key


def segment_softmax(a: Tensor, index: Tensor, num_nodes: int) -> Tensor:
    """``torch_geometric.utils.softmax(src, index, num_nodes=N)`` (Appendix A.1).

    m_i = amax over incoming edges (detached, 0 for empty segments),
    p = exp(a - m[i]), s_i = sum p + 1e-16, alpha = p / s[i].
    """
    shape = (num_nodes,) + tuple(a.shape[1:])
    idx = index.view(-1, *([1] * (a.dim() - 1))).expand_as(a)
    m = a.new_zeros(shape).scatter_reduce_(0, idx, a.detach(), reduce="amax", include_self=False)
    p = (a - m.index_select(0, index)).exp()
    s = a.new_zeros(shape).scatter_add_(0, idx, p) + 1e-16
    return p / s.index_select(0, index)


def gatv2_aggregate(
    x_l: Tensor,          # [N_s, H, C] source-side projection (also the message payload)
    x_r: Tensor,          # [N_d, H, C] target-side projection
    edge_index: Tensor,   # [2, E]  row 0 = source j, row 1 = target i
    att: Tensor,          # [1, H, C]
    bias: Optional[Tensor],  # [H*C]
    negative_slope: float = 0.2,
    dropout_p: float = 0.0,
    training: bool = False,
    keep_mask: Optional[Tensor] = None,   # [E, H] bool: injected dropout mask (test hook)
    return_alpha: bool = False,
):
    """Attention + aggregation part of GATv2Conv.forward (Appendix A.1)."""
    N_d, H, C = x_r.shape
    j = edge_index[0].long()
    i = edge_index[1].long()
    e = F.leaky_relu(x_l.index_select(0, j) + x_r.index_select(0, i), negative_slope)  # [E,H,C]
    a = (e * att).sum(-1)                                                                 # [E,H]
    alpha = segment_softmax(a, i, N_d)
    alpha_pre = alpha
    if keep_mask is not None:
        alpha = alpha * keep_mask.to(alpha.dtype) / (1.0 - dropout_p)
    else:
        alpha = F.dropout(alpha, p=dropout_p, training=training)
    msg = x_l.index_select(0, j) * alpha.unsqueeze(-1)                                    # [E,H,C]
    out = x_l.new_zeros(N_d, H, C).index_add_(0, i, msg)
    out = out.reshape(N_d, H * C)
    if bias is not None:
        out = out + bias
    if return_alpha:
        return out, alpha_pre
    return out


class LinearRef(torch.nn.Module):
    """PyG ``Linear(in, out)``: y = x W^T + b; weight [out, in]."""

    def __init__(self, in_channels: int, out_channels: int, bias: bool = True):
        super().__init__()
        self.weight = torch.nn.Parameter(torch.empty(out_channels, in_channels))
        self.bias = torch.nn.Parameter(torch.empty(out_channels)) if bias else None
        self.reset_parameters()

    def reset_parameters(self):
        # glorot weights / uniform(+-1/sqrt(fan_in)) bias -- irrelevant for parity (weights are copied).
        fan_out, fan_in = self.weight.shape
        a = math.sqrt(6.0 / (fan_in + fan_out))
        with torch.no_grad():
            self.weight.uniform_(-a, a)
            if self.bias is not None:
                b = 1.0 / math.sqrt(fan_in)
                self.bias.uniform_(-b, b)

    def forward(self, x: Tensor) -> Tensor:
        return F.linear(x, self.weight, self.bias)


class GATv2ConvRef(torch.nn.Module):
    """GATv2Conv as instantiated at models/ist_encoder.py:111-131 (Appendix A.1).

    heads=H, concat=True, negative_slope=0.2, add_self_loops=False, bias=True,
    edge_dim=None, share_weights=False, residual=False, aggr='add'.
    """

    def __init__(self, in_channels: Tuple[int, int], out_channels: int, heads: int,
                 negative_slope: float = 0.2, dropout: float = 0.0):
        super().__init__()
        self.heads, self.out_channels = heads, out_channels
        self.negative_slope, self.dropout = negative_slope, dropout
        self.lin_l = LinearRef(in_channels[0], heads * out_channels)
        self.lin_r = LinearRef(in_channels[1], heads * out_channels)
        self.att = torch.nn.Parameter(torch.empty(1, heads, out_channels))
        self.bias = torch.nn.Parameter(torch.zeros(heads * out_channels))
        a = math.sqrt(6.0 / (heads + out_channels))
        with torch.no_grad():
            self.att.uniform_(-a, a)

    def forward(self, x, edge_index: Tensor, keep_mask: Optional[Tensor] = None,
                return_alpha: bool = False):
        H, C = self.heads, self.out_channels
        if isinstance(x, Tensor):
            x_src = x_dst = x
        else:
            x_src, x_dst = x
        x_l = self.lin_l(x_src).view(-1, H, C)
        x_r = self.lin_r(x_dst).view(-1, H, C)
        return gatv2_aggregate(x_l, x_r, edge_index, self.att, self.bias, self.negative_slope,
                               self.dropout, self.training, keep_mask, return_alpha)


class HeteroConvRef(torch.nn.Module):
    """HeteroConv(convs, aggr='sum') (Appendix A.2)."""

    def __init__(self, convs: Dict[EdgeType, torch.nn.Module]):
        super().__init__()
        self.edge_types = list(convs.keys())
        self.convs = torch.nn.ModuleDict({mangle(k): v for k, v in convs.items()})

    def forward(self, x_dict, edge_index_dict, keep_mask_dict=None):
        out: Dict[str, list] = {}
        for et in self.edge_types:
            if et not in edge_index_dict:
                continue
            src, _, dst = et
            conv = self.convs[mangle(et)]
            x = x_dict[src] if src == dst else (x_dict.get(src), x_dict.get(dst))
            km = None if keep_mask_dict is None else keep_mask_dict.get(et)
            out.setdefault(dst, []).append(conv(x, edge_index_dict[et], keep_mask=km))
        return {k: (v[0] if len(v) == 1 else torch.stack(v, 0).sum(0)) for k, v in out.items()}


class HeteroDictLinearRef(torch.nn.Module):
    """HeteroDictLinear(-1, out, types) (Appendix A.3): independent Linear per node type."""

    def __init__(self, in_channels: Dict[str, int], out_channels: int):
        super().__init__()
        self.lins = torch.nn.ModuleDict({k: LinearRef(c, out_channels) for k, c in in_channels.items()})

    def forward(self, x_dict):
        return {k: self.lins[k](x) for k, x in x_dict.items() if k in self.lins}
