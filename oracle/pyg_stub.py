"""Stand-ins with the *API* of the third-party classes segger imports, so that the reference's own
source files can be executed here unmodified (oracle/reference_import.py).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  What is restated here, and why:

  torch_geometric 2.7.0 (pixi.lock:3408; absent from /root/reference, not installable offline)
      ``nn.GATv2Conv``, ``nn.Linear``, ``nn.HeteroDictLinear``, ``nn.HeteroConv``,
      ``nn.module_dict.ModuleDict``  -- constructor signatures, lazy ``-1`` fan-ins, parameter names and
      the argument routing of ``HeteroConv.forward`` follow the published implementation
      (SURVEY.md Appendix A.1-A.3, A.6); the attention math is ``oracle.pyg_ref.gatv2_aggregate``.
  torch_scatter 2.1.2 (pixi.lock:3470)
      ``scatter_max`` -- ``oracle.ist_encoder_ref.scatter_max_ref`` (Appendix A.7).
  lightning 2.6.1
      ``LightningModule`` -- a ``torch.nn.Module`` with the handful of attributes
      ``LitISTEncoder`` touches (``save_hyperparameters``, ``log``, ``trainer``, ``current_epoch``, ``device``).

Only these third-party internals remain "restated" once the reference's own files run on top of them:
everything segger itself wrote (ISTEncoder / SkipGAT / Positional2dEmbedder composition, the losses,
predict_step, kdtree_neighbors / knn_to_edge_index) is then the reference's code, not ours.
"""
from __future__ import annotations

import math
from types import SimpleNamespace
from typing import Dict, Iterable, Mapping, Optional, Tuple, Union

import torch
from torch import Tensor
from torch.nn.parameter import UninitializedParameter

from .pyg_ref import gatv2_aggregate


# ---- torch_geometric.nn.module_dict.ModuleDict -----------------------------------------------------
class ModuleDict(torch.nn.ModuleDict):
    def __init__(self, modules: Optional[Mapping] = None):
        if modules is not None:
            modules = {self.to_internal_key(k): m for k, m in modules.items()}
        super().__init__(modules)

    @classmethod
    def to_internal_key(cls, key) -> str:
        if isinstance(key, tuple):
            key = "<" + "___".join(key) + ">"
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

    def __contains__(self, key) -> bool:
        return super().__contains__(self.to_internal_key(key))

    def keys(self):
        return [self.to_external_key(k) for k in super().keys()]

    def items(self):
        return [(self.to_external_key(k), v) for k, v in super().items()]
    # __iter__ stays torch's: it yields the INTERNAL string keys


# ---- torch_geometric.nn.Linear / HeteroDictLinear ---------------------------------------------------
class Linear(torch.nn.Module):
    def __init__(self, in_channels: int, out_channels: int, bias: bool = True,
                 weight_initializer: Optional[str] = None, bias_initializer: Optional[str] = None):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.weight_initializer, self.bias_initializer = weight_initializer, bias_initializer
        self.weight = (torch.nn.Parameter(torch.empty(out_channels, in_channels)) if in_channels > 0
                       else UninitializedParameter())
        if bias:
            self.bias = torch.nn.Parameter(torch.empty(out_channels))
        else:
            self.register_parameter("bias", None)
        self.reset_parameters()

    def reset_parameters(self):
        if self.in_channels <= 0:
            return
        with torch.no_grad():
            if self.weight_initializer == "glorot":
                a = math.sqrt(6.0 / (self.weight.size(0) + self.weight.size(1)))
                self.weight.uniform_(-a, a)
            else:
                torch.nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
            if self.bias is not None:
                if self.bias_initializer == "zeros":
                    self.bias.zero_()
                else:
                    b = 1.0 / math.sqrt(self.in_channels)
                    self.bias.uniform_(-b, b)

    def forward(self, x: Tensor) -> Tensor:
        if isinstance(self.weight, UninitializedParameter):
            self.in_channels = x.size(-1)
            self.weight.materialize((self.out_channels, self.in_channels))
            self.reset_parameters()
        return torch.nn.functional.linear(x, self.weight, self.bias)

    def _save_to_state_dict(self, destination, prefix, keep_vars):
        # PyG: a never-materialised weight is stored as the UninitializedParameter itself
        if isinstance(self.weight, UninitializedParameter):
            destination[prefix + "weight"] = self.weight
            if self.bias is not None:
                destination[prefix + "bias"] = self.bias if keep_vars else self.bias.detach()
        else:
            super()._save_to_state_dict(destination, prefix, keep_vars)


class HeteroDictLinear(torch.nn.Module):
    def __init__(self, in_channels: Union[int, Dict[str, int]], out_channels: int,
                 types: Optional[Iterable[str]] = None, **kwargs):
        super().__init__()
        if isinstance(in_channels, dict):
            chans = dict(in_channels)
        else:
            chans = {t: in_channels for t in types}
        self.lins = ModuleDict({t: Linear(c, out_channels, **kwargs) for t, c in chans.items()})

    def forward(self, x_dict):
        return {k: self.lins[k](x) for k, x in x_dict.items() if k in self.lins}


# ---- torch_geometric.nn.GATv2Conv -------------------------------------------------------------------
class GATv2Conv(torch.nn.Module):
    """Signature of PyG's GATv2Conv; only the configuration segger instantiates is supported."""

    def __init__(self, in_channels, out_channels: int, heads: int = 1, concat: bool = True,
                 negative_slope: float = 0.2, dropout: float = 0.0, add_self_loops: bool = True,
                 edge_dim=None, fill_value="mean", bias: bool = True, share_weights: bool = False,
                 residual: bool = False, **kwargs):
        super().__init__()
        assert concat and edge_dim is None and not share_weights and not residual
        self.in_channels, self.out_channels, self.heads = in_channels, out_channels, heads
        self.negative_slope, self.dropout, self.add_self_loops = negative_slope, dropout, add_self_loops
        in_l, in_r = (in_channels, in_channels) if isinstance(in_channels, int) else in_channels
        self.lin_l = Linear(in_l, heads * out_channels, bias=bias, weight_initializer="glorot")
        self.lin_r = Linear(in_r, heads * out_channels, bias=bias, weight_initializer="glorot")
        self.att = torch.nn.Parameter(torch.empty(1, heads, out_channels))
        self.bias = torch.nn.Parameter(torch.zeros(heads * out_channels)) if bias else None
        a = math.sqrt(6.0 / (heads + out_channels))
        with torch.no_grad():
            self.att.uniform_(-a, a)
        self.keep_mask: Optional[Tensor] = None     # test hook: injected [E, H] dropout keep mask

    def forward(self, x, edge_index: Tensor, edge_attr=None, return_attention_weights=None):
        assert edge_attr is None
        H, C = self.heads, self.out_channels
        x_src, x_dst = (x, x) if isinstance(x, Tensor) else x
        x_l = self.lin_l(x_src).view(-1, H, C)
        x_r = self.lin_r(x_dst).view(-1, H, C)
        if self.add_self_loops:
            n = min(x_src.size(0), x_dst.size(0))
            keep = edge_index[0] != edge_index[1]
            loops = torch.arange(n, dtype=edge_index.dtype)
            edge_index = torch.cat([edge_index[:, keep], torch.stack([loops, loops])], 1)
        out, alpha = gatv2_aggregate(x_l, x_r, edge_index, self.att, self.bias, self.negative_slope, self.dropout,
                                     self.training, self.keep_mask, return_alpha=True)
        if isinstance(return_attention_weights, bool):
            return out, (edge_index, alpha)
        return out


# ---- torch_geometric.nn.HeteroConv ------------------------------------------------------------------
class HeteroConv(torch.nn.Module):
    def __init__(self, convs, aggr: Optional[str] = "sum"):
        super().__init__()
        self.convs = ModuleDict(convs)
        self.aggr = aggr

    def forward(self, *args_dict, **kwargs_dict):
        out_dict: Dict[str, list] = {}
        for edge_type, conv in self.convs.items():
            src, _, dst = edge_type
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
                arg = arg[:-5]                      # strip `_dict`
                if edge_type in value_dict:
                    has_edge_level_arg = True
                    kwargs[arg] = value_dict[edge_type]
                elif src == dst and src in value_dict:
                    kwargs[arg] = value_dict[src]
                elif src in value_dict or dst in value_dict:
                    kwargs[arg] = (value_dict.get(src, None), value_dict.get(dst, None))
            if not has_edge_level_arg:
                continue
            out_dict.setdefault(dst, []).append(conv(*args, **kwargs))
        res = {}
        for k, v in out_dict.items():
            if len(v) == 1:
                res[k] = v[0]
            else:
                assert self.aggr == "sum"
                res[k] = torch.stack(v, dim=0).sum(dim=0)
        return res


# ---- lightning.LightningModule ----------------------------------------------------------------------
class LightningModule(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.trainer = SimpleNamespace(max_epochs=1, datamodule=None)
        self.current_epoch = 0
        self.logged = {}

    def save_hyperparameters(self, *a, **k):
        pass

    def log(self, name, value, **kwargs):
        self.logged[name] = value

    def setup(self, stage=None):
        return None

    @property
    def device(self):
        for p in self.parameters():
            if not isinstance(p, UninitializedParameter):
                return p.device
        return torch.device("cpu")
