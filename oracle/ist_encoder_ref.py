"""CPU restatement of segger's ISTEncoder / SkipGAT / Positional2dEmbedder.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Follows
/root/reference/src/segger/models/ist_encoder.py line by line:
  sinusoidal_embedding        :22-31
  Positional2dEmbedder        :33-79
  SkipGAT                     :82-211   (dead bd-contains-tx conv kept unmaterialised, Appendix B.1;
                                         attention kwarg is a no-op, Appendix B.2)
  ISTEncoder                  :214-333
State-dict keys follow SURVEY.md Appendix A.6 so weights can be copied 1:1 into
the product modules.  Lazy (-1) fan-ins are resolved at construction time here
(the oracle is told the widths) -- the lazily-initialised behaviour itself is
host logic tested on the product side.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
from torch import Tensor
from torch.nn import functional as F

from .pyg_ref import GATv2ConvRef, HeteroConvRef, HeteroDictLinearRef, LinearRef

TT = ("tx", "neighbors", "tx")
TB = ("tx", "belongs", "bd")
BT = ("bd", "contains", "tx")


def sinusoidal_embedding(x: Tensor, dim: int, max_period: float = 1000) -> Tensor:
    """ist_encoder.py:22-31."""
    half = dim // 2
    freqs = torch.exp(
        -math.log(max_period) * torch.arange(start=0, end=half, dtype=torch.float32) / half
    ).to(device=x.device)
    args = x[:, None].float() * freqs[None]
    embedding = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)
    if dim % 2:
        embedding = torch.cat([embedding, torch.zeros_like(embedding[:, :1])], dim=-1)
    return embedding


class Positional2dEmbedderRef(torch.nn.Module):
    """ist_encoder.py:33-79."""

    def __init__(self, hidden_size: int, frequency_embedding_size: int = 256):
        super().__init__()
        self.dim = hidden_size // 2
        self.mlp = torch.nn.Sequential(
            torch.nn.Linear(frequency_embedding_size, self.dim, bias=True),
            torch.nn.SiLU(),
            torch.nn.Linear(self.dim, self.dim, bias=True),
        )
        self.frequency_embedding_size = frequency_embedding_size

    def forward(self, pos: Tensor, batch: Optional[Tensor] = None) -> Tensor:
        if batch is None:
            pos = pos - pos.min(dim=0).values
            pos = pos / pos.max(dim=0).values
        else:
            nb = int(batch.max()) + 1
            mins = torch.zeros((nb, 2), device=pos.device)      # (device: bench.py --impl torch_cuda runs this on the GPU)
            maxs = torch.zeros((nb, 2), device=pos.device)
            for b in range(nb):
                mask = batch == b
                if mask.any():
                    mins[b] = pos[mask].min(dim=0).values
                    maxs[b] = pos[mask].max(dim=0).values
            pos = (pos - mins[batch]) / (maxs[batch] - mins[batch] + 1e-8)
        shape = pos.shape
        emb = sinusoidal_embedding(pos.flatten(), self.frequency_embedding_size, max_period=10000)
        pos_freq = emb.reshape(shape + (self.frequency_embedding_size,))
        # (.to() is a no-op in fp32; it lets the fp64 copy of the oracle run for conditioning checks)
        return self.mlp(pos_freq.to(self.mlp[0].weight.dtype)).flatten(-2)


class SkipGATRef(torch.nn.Module):
    """ist_encoder.py:82-211.  Returns HeteroConv's x_dict (plain tensors, Appendix B.2)."""

    def __init__(self, in_channels: Dict[str, int], out_channels: int, n_heads: int,
                 dropout: float = 0.2):
        super().__init__()
        ct, cb = in_channels["tx"], in_channels["bd"]
        self.conv = HeteroConvRef({
            TT: GATv2ConvRef((ct, ct), out_channels, n_heads, dropout=dropout),
            TB: GATv2ConvRef((ct, cb), out_channels, n_heads, dropout=dropout),
            # BT is declared by the reference (ist_encoder.py:125-131) but never runs and its lazy
            # parameters are never materialised (Appendix B.1): the oracle omits it.
        })

    def forward(self, x_dict, edge_index_dict, keep_mask_dict=None):
        return self.conv(x_dict, edge_index_dict, keep_mask_dict)


class ISTEncoderRef(torch.nn.Module):
    """ist_encoder.py:214-333."""

    def __init__(self, n_genes: int, bd_in: int, in_channels: int = 16, hidden_channels: int = 32,
                 out_channels: int = 32, n_mid_layers: int = 3, n_heads: int = 3,
                 normalize_embeddings: bool = True, use_positional_embeddings: bool = True,
                 dropout: float = 0.2):
        super().__init__()
        self.normalize_embeddings = normalize_embeddings
        self.use_positional_embeddings = use_positional_embeddings
        self.lin_first = torch.nn.ModuleDict({
            "tx": torch.nn.Embedding(n_genes, in_channels),
            "bd": LinearRef(bd_in, in_channels),
        })
        self.pos_emb = Positional2dEmbedderRef(in_channels)
        w0 = in_channels + (2 * (in_channels // 2) if use_positional_embeddings else 0)
        widths = [w0] + [hidden_channels * n_heads] * (n_mid_layers + 1)
        outs = [hidden_channels] * (n_mid_layers + 1) + [out_channels]
        self.conv_layers = torch.nn.ModuleList(
            SkipGATRef({"tx": w, "bd": w}, o, n_heads, dropout) for w, o in zip(widths, outs)
        )
        last_in = out_channels * n_heads
        self.lin_last = HeteroDictLinearRef({"tx": last_in, "bd": last_in}, out_channels)

    def forward(self, x_dict, edge_index_dict, pos_dict, batch_dict, keep_masks=None,
                return_hidden: bool = False):
        x_dict = {k: self.lin_first[k](x) for k, x in x_dict.items()}
        if self.use_positional_embeddings:
            x_dict = {
                k: torch.cat((x, self.pos_emb(pos_dict[k], batch_dict[k])), -1)
                for k, x in x_dict.items()
            }
        x_dict = {k: F.gelu(x) for k, x in x_dict.items()}
        hidden = [x_dict]
        for li, conv_layer in enumerate(self.conv_layers):
            km = None if keep_masks is None else keep_masks[li]
            x_dict = conv_layer(x_dict, edge_index_dict, km)
            x_dict = {k: F.gelu(x) for k, x in x_dict.items()}
            hidden.append(x_dict)
        x_dict = self.lin_last(x_dict)
        if self.normalize_embeddings:
            x_dict = {k: F.normalize(v, dim=-1) for k, v in x_dict.items()}
        if return_hidden:
            return x_dict, hidden
        return x_dict


def scatter_max_ref(src: Tensor, index: Tensor, dim_size: int):
    """torch_scatter 2.1.2 ``scatter_max`` on CPU (Appendix A.7) -- PARITY UNPINNED.

    Empty segments: out = 0, arg = src.numel(); ties: first occurrence.
    """
    E = src.numel()
    out = src.new_zeros(dim_size)
    arg = torch.full((dim_size,), E, dtype=torch.long)
    if E == 0:
        return out, arg
    index = index.long()
    mx = src.new_full((dim_size,), float("-inf")).scatter_reduce_(0, index, src, "amax", include_self=True)
    has = torch.zeros(dim_size, dtype=torch.bool).index_fill_(0, index, True)
    is_max = src == mx.index_select(0, index)
    eid = torch.where(is_max, torch.arange(E), torch.full((E,), E))
    arg = arg.scatter_reduce_(0, index, eid, "amin", include_self=True)
    out = torch.where(has, mx, out)
    return out, arg


def predict_scores_ref(emb_tx: Tensor, emb_bd: Tensor, edge_index: Tensor, bd_index: Tensor,
                       min_similarity: Optional[float] = None):
    """models/lightning_model.py:275-293 (scoring + arg-max part of predict_step).

    Returns (seg_idx int64 [N_tx], max_sim fp32 [N_tx], max_idx int64 [N_tx]).
    """
    src, dst = edge_index[0].long(), edge_index[1].long()
    sim = torch.cosine_similarity(emb_tx[src], emb_bd[dst])
    max_sim, max_idx = scatter_max_ref(sim, src, emb_tx.shape[0])
    valid = max_idx < dst.shape[0]
    if min_similarity is not None:
        valid &= max_sim >= min_similarity
    dst_idx = bd_index.to(torch.long)
    seg_idx = torch.full_like(max_idx, -1)
    seg_idx[valid] = dst_idx[dst[max_idx[valid]]]
    return seg_idx, max_sim, max_idx
