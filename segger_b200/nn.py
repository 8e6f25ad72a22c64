"""Drop-in stand-ins for the torch_geometric classes segger composes
(`from torch_geometric.nn import GATv2Conv, Linear, HeteroDictLinear, HeteroConv`,
/root/reference/src/segger/models/ist_encoder.py:1), backed by libsegger_b200 kernels.

Same constructor arguments, parameter names and state-dict keys so that
checkpoints trained with the PyG implementation load unchanged.  Only the configurations segger
uses are implemented; anything else raises NotImplementedError rather than silently deviating.
"""
from __future__ import annotations

import math
from typing import Dict, Iterable, Mapping, Optional, Tuple, Union

import torch
from torch import Tensor
from torch.nn.parameter import UninitializedParameter

from . import ops
from ._lib import ACT_NONE

EdgeType = Tuple[str, str, str]


# ------------------------------------------------------------------------------------------------
# ModuleDict with tuple keys, mangled the way torch_geometric.nn.module_dict.ModuleDict does
# ------------------------------------------------------------------------------------------------
class ModuleDict(torch.nn.ModuleDict):
    """``torch_geometric.nn.module_dict.ModuleDict``: accepts tuple keys, stores them as
    ``'<a___b___c>'``; string keys that collide with attributes or contain '.' are wrapped too."""

    def __init__(self, modules: Optional[Mapping] = None):
        if modules is not None:
            modules = {self.to_internal_key(k): m for k, m in modules.items()}
        super().__init__(modules)

    @classmethod
    def to_internal_key(cls, key) -> str:
        if isinstance(key, tuple):
            assert len(key) > 1
            key = f"<{'___'.join(key)}>"
        assert isinstance(key, str)
        if hasattr(cls, key) or "." in key:
            key = f"<{key}>"
        return key

    @classmethod
    def to_external_key(cls, key: str):
        if key[0] == "<" and key[-1] == ">" and hasattr(cls, key[1:-1]):
            key = key[1:-1]
        if key[0] == "<" and key[-1] == ">" and "___" in key:
            key = tuple(key[1:-1].split("___"))
        return key

    def __getitem__(self, key):
        return super().__getitem__(self.to_internal_key(key))

    def __setitem__(self, key, module):
        return super().__setitem__(self.to_internal_key(key), module)

    def __delitem__(self, key):
        return super().__delitem__(self.to_internal_key(key))

    def __contains__(self, key) -> bool:
        return super().__contains__(self.to_internal_key(key))

    def keys(self):
        return [self.to_external_key(k) for k in super().keys()]

    def items(self):
        return [(self.to_external_key(k), v) for k, v in super().items()]
    # NOTE: __iter__ is inherited from torch.nn.ModuleDict and yields the *internal* string keys --
    # exactly the PyG behaviour SkipGAT.forward trips over.


# ------------------------------------------------------------------------------------------------
# Linear / HeteroDictLinear
# ------------------------------------------------------------------------------------------------
class Linear(torch.nn.Module):
    """``torch_geometric.nn.Linear(in_channels, out_channels, bias=True, weight_initializer=None,
    bias_initializer=None)``; ``in_channels=-1`` is resolved at the first forward."""

    def __init__(self, in_channels: int, out_channels: int, bias: bool = True,
                 weight_initializer: Optional[str] = None, bias_initializer: Optional[str] = None):
        super().__init__()
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.weight_initializer = weight_initializer
        self.bias_initializer = bias_initializer
        if in_channels > 0:
            self.weight = torch.nn.Parameter(torch.empty(out_channels, in_channels))
        else:
            self.weight = UninitializedParameter()
        if bias:
            self.bias = torch.nn.Parameter(torch.empty(out_channels))
        else:
            self.register_parameter("bias", None)
        self.reset_parameters()

    def reset_parameters(self):
        if self.in_channels > 0:
            with torch.no_grad():
                if self.weight_initializer == "glorot":
                    a = math.sqrt(6.0 / (self.weight.size(0) + self.weight.size(1)))
                    self.weight.uniform_(-a, a)
                elif self.weight_initializer == "uniform":
                    b = 1.0 / math.sqrt(self.weight.size(-1))
                    self.weight.uniform_(-b, b)
                elif self.weight_initializer in ("kaiming_uniform", None):
                    torch.nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
                else:
                    raise RuntimeError(f"Linear weight initializer '{self.weight_initializer}' not supported")
        if self.bias is not None and self.in_channels > 0:
            with torch.no_grad():
                if self.bias_initializer == "zeros":
                    self.bias.zero_()
                elif self.bias_initializer is None:
                    b = 1.0 / math.sqrt(self.in_channels)
                    self.bias.uniform_(-b, b)
                else:
                    raise RuntimeError(f"Linear bias initializer '{self.bias_initializer}' not supported")

    def materialize(self, in_channels: int, like: Tensor) -> None:
        if isinstance(self.weight, UninitializedParameter):
            self.in_channels = in_channels
            self.weight.materialize((self.out_channels, in_channels), device=like.device, dtype=torch.float32)
            if self.bias is not None and self.bias.device != like.device:
                self.bias.data = self.bias.data.to(like.device)
            self.reset_parameters()

    def forward(self, x: Tensor) -> Tensor:
        self.materialize(x.size(-1), x)
        return ops.linear(x.to(torch.float32), self.weight, self.bias, ACT_NONE)

    # uninitialised parameters round-trip through state_dict like PyG's / torch's lazy modules
    def _save_to_state_dict(self, destination, prefix, keep_vars):
        if isinstance(self.weight, UninitializedParameter):
            destination[prefix + "weight"] = self.weight
            if self.bias is not None:
                destination[prefix + "bias"] = self.bias if keep_vars else self.bias.detach()
        else:
            super()._save_to_state_dict(destination, prefix, keep_vars)

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys,
                              error_msgs):
        key = prefix + "weight"
        w = state_dict.get(key, None)
        if w is not None and isinstance(w, UninitializedParameter):
            # checkpoint holds a never-materialised layer (e.g. the dead bd-contains-tx conv): stay lazy
            self.in_channels = -1
            if not isinstance(self.weight, UninitializedParameter):
                self.weight = UninitializedParameter()
            mine = self._parameters.pop("weight")
            theirs = state_dict.pop(key)
            try:
                super()._load_from_state_dict(state_dict, prefix, local_metadata, strict, missing_keys,
                                              unexpected_keys, error_msgs)
            finally:
                self._parameters["weight"] = mine
                state_dict[key] = theirs
            return
        if w is not None and isinstance(self.weight, UninitializedParameter):
            self.in_channels = w.size(-1)
            dev = self.bias.device if self.bias is not None else w.device
            self.weight.materialize((self.out_channels, self.in_channels), device=dev, dtype=torch.float32)
        super()._load_from_state_dict(state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys,
                                      error_msgs)

    def __repr__(self) -> str:
        return f"{self.__class__.__name__}({self.in_channels}, {self.out_channels}, bias={self.bias is not None})"


class HeteroDictLinear(torch.nn.Module):
    """``torch_geometric.nn.HeteroDictLinear(in_channels, out_channels, types)``: one independent
    Linear per node type, stored in ``self.lins`` (keys ``lins.tx.weight`` ...)."""

    def __init__(self, in_channels: Union[int, Dict[str, int]], out_channels: int,
                 types: Optional[Iterable[str]] = None, **kwargs):
        super().__init__()
        if isinstance(in_channels, dict):
            self.types = list(in_channels.keys())
            self.in_channels = dict(in_channels)
        else:
            if types is None:
                raise ValueError("HeteroDictLinear: `types` is required when in_channels is an int")
            self.types = list(types)
            self.in_channels = {t: in_channels for t in self.types}
        self.out_channels = out_channels
        self.lins = ModuleDict({t: Linear(c, out_channels, **kwargs) for t, c in self.in_channels.items()})

    def reset_parameters(self):
        for lin in self.lins.values():
            lin.reset_parameters()

    def forward(self, x_dict: Dict[str, Tensor]) -> Dict[str, Tensor]:
        return {k: self.lins[k](x) for k, x in x_dict.items() if k in self.lins}


# ------------------------------------------------------------------------------------------------
# GATv2Conv / HeteroConv
# ------------------------------------------------------------------------------------------------
class GATv2Conv(torch.nn.Module):
    """``torch_geometric.nn.GATv2Conv`` restricted to what segger instantiates
    (models/ist_encoder.py:111-131): bipartite (lazy) in_channels, concat heads, no edge features,
    separate lin_l / lin_r, no residual.  Math: SURVEY.md Appendix A.1.
    """

    def __init__(self, in_channels: Union[int, Tuple[int, int]], out_channels: int, heads: int = 1,
                 concat: bool = True, negative_slope: float = 0.2, dropout: float = 0.0,
                 add_self_loops: bool = True, edge_dim: Optional[int] = None, fill_value="mean",
                 bias: bool = True, share_weights: bool = False, residual: bool = False, **kwargs):
        super().__init__()
        if not concat:
            raise NotImplementedError("segger_b200.GATv2Conv: concat=False is not implemented")
        if edge_dim is not None:
            raise NotImplementedError("segger_b200.GATv2Conv: edge features are not implemented")
        if share_weights or residual:
            raise NotImplementedError("segger_b200.GATv2Conv: share_weights / residual are not implemented")
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.heads = heads
        self.concat = concat
        self.negative_slope = negative_slope
        self.dropout = dropout
        self.add_self_loops = add_self_loops
        self.edge_dim = edge_dim
        self.fill_value = fill_value
        self.residual = residual
        self.share_weights = share_weights
        if isinstance(in_channels, int):
            in_l = in_r = in_channels
        else:
            in_l, in_r = in_channels
        self.lin_l = Linear(in_l, heads * out_channels, bias=bias, weight_initializer="glorot")
        self.lin_r = Linear(in_r, heads * out_channels, bias=bias, weight_initializer="glorot")
        self.att = torch.nn.Parameter(torch.empty(1, heads, out_channels))
        if bias:
            self.bias = torch.nn.Parameter(torch.empty(heads * out_channels))
        else:
            self.register_parameter("bias", None)
        self.reset_parameters()

    def reset_parameters(self):
        self.lin_l.reset_parameters()
        self.lin_r.reset_parameters()
        with torch.no_grad():
            a = math.sqrt(6.0 / (self.att.size(-2) + self.att.size(-1)))
            self.att.uniform_(-a, a)
            if self.bias is not None:
                self.bias.zero_()

    @property
    def initialized(self) -> bool:
        return not (isinstance(self.lin_l.weight, UninitializedParameter)
                    or isinstance(self.lin_r.weight, UninitializedParameter))

    def _edges(self, edge_index: Tensor, n_src: int, n_dst: int) -> Tensor:
        if not self.add_self_loops:
            return edge_index
        # PyG: remove_self_loops then add_self_loops(num_nodes=min(n_src, n_dst)); index plumbing only
        n = min(n_src, n_dst)
        keep = edge_index[0] != edge_index[1]
        loops = torch.arange(n, device=edge_index.device, dtype=edge_index.dtype)
        return torch.cat([edge_index[:, keep], torch.stack([loops, loops])], 1)

    def forward(self, x, edge_index: Tensor, edge_attr=None, return_attention_weights=None):
        if edge_attr is not None:
            raise NotImplementedError("segger_b200.GATv2Conv: edge_attr is not implemented")
        H, C = self.heads, self.out_channels
        if isinstance(x, Tensor):
            x_src = x_dst = x
        else:
            x_src, x_dst = x
            if x_dst is None:
                raise NotImplementedError("segger_b200.GATv2Conv: x_dst=None is not implemented")
        x_l = self.lin_l(x_src)
        x_r = self.lin_r(x_dst)
        ei = self._edges(edge_index, x_src.size(0), x_dst.size(0))
        csr = ops.CSR_CACHE.get(ei, x_src.size(0), x_dst.size(0), transpose=torch.is_grad_enabled())
        ops.validate_csrs(csr)
        training = self.training and self.dropout > 0.0
        seed = ops.new_seed() if training else 0
        out = ops.GATv2AggregateFn.apply(x_l, x_r, self.att, self.bias, csr, H, C, self.negative_slope,
                                         self.dropout, training, seed, False)
        if isinstance(return_attention_weights, bool):
            with torch.no_grad():
                _, _, smax, sden = ops.gatv2_fwd(x_l.detach(), x_r.detach(), self.att.detach(), None, csr, H, C,
                                                 self.negative_slope, 0.0, False, 0, False)
                alpha = ops.gatv2_alpha(x_l.detach(), x_r.detach(), self.att.detach(), csr, H, C,
                                        self.negative_slope, smax, sden)
            return out, (ei, alpha)
        return out

    def __repr__(self) -> str:
        return f"{self.__class__.__name__}({self.in_channels}, {self.out_channels}, heads={self.heads})"


class HeteroConv(torch.nn.Module):
    """``torch_geometric.nn.HeteroConv(convs, aggr='sum')``."""

    def __init__(self, convs: Dict[EdgeType, torch.nn.Module], aggr: Optional[str] = "sum"):
        super().__init__()
        if aggr not in ("sum", "mean", "min", "max", "cat", None):
            raise ValueError(f"HeteroConv: unknown aggr '{aggr}'")
        self.convs = ModuleDict(convs)
        self.aggr = aggr

    def reset_parameters(self):
        for conv in self.convs.values():
            conv.reset_parameters()

    def forward(self, *args_dict, **kwargs_dict):
        out_dict: Dict[str, list] = {}
        for edge_type, conv in self.convs.items():
            src, rel, dst = edge_type
            has_edge_level_arg = False
            args = []
            for value_dict in args_dict:
                if edge_type in value_dict:
                    has_edge_level_arg = True
                    args.append(value_dict[edge_type])
                elif src == dst and src in value_dict:
                    args.append(value_dict[src])
                elif src in value_dict or dst in value_dict:
                    args.append((value_dict.get(src, None), value_dict.get(dst, None)))
            kwargs = {}
            for arg, value_dict in kwargs_dict.items():
                if not arg.endswith("_dict"):
                    raise ValueError(f"Keyword arguments in 'HeteroConv' need to end with '_dict' (got '{arg}')")
                arg = arg[:-5]
                if edge_type in value_dict:
                    has_edge_level_arg = True
                    kwargs[arg] = value_dict[edge_type]
                elif src == dst and src in value_dict:
                    kwargs[arg] = value_dict[src]
                elif src in value_dict or dst in value_dict:
                    kwargs[arg] = (value_dict.get(src, None), value_dict.get(dst, None))
            if not has_edge_level_arg:
                continue
            out = conv(*args, **kwargs)
            out_dict.setdefault(dst, []).append(out)
        for key, value in out_dict.items():
            out_dict[key] = _group(value, self.aggr)
        return out_dict


def _group(xs, aggr):
    if len(xs) == 0:
        return None
    if aggr is None:
        return torch.stack(xs, dim=1)
    if len(xs) == 1:
        return xs[0]
    if aggr == "cat":
        return torch.cat(xs, dim=-1)
    out = torch.stack(xs, dim=0)
    out = getattr(torch, aggr)(out, dim=0)
    return out[0] if isinstance(out, tuple) else out
