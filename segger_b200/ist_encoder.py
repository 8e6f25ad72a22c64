"""B200-native drop-in for ``segger.models.ist_encoder``
(/root/reference/src/segger/models/ist_encoder.py): ``Positional2dEmbedder``, ``SkipGAT`` and
``ISTEncoder`` with identical constructor signatures, sub-module names and state-dict keys
(``sinusoidal_embedding`` is folded into ``Positional2dEmbedder.features``, one fused kernel).
All arithmetic runs in libsegger_b200 kernels; the modules below only hold parameters and
schedule launches.
"""
from __future__ import annotations

import logging
import os
from typing import Dict, Optional, Tuple

import torch
from torch import Tensor
from torch.nn import Embedding, Module, ModuleList, Sequential, SiLU
from torch.nn import Linear as NNLinear

from . import ops
from ._lib import ACT_GELU, ACT_NONE, ACT_SILU, require_cuda
from .nn import GATv2Conv, HeteroConv, HeteroDictLinear, Linear, ModuleDict

logger = logging.getLogger(__name__)

TT = ("tx", "neighbors", "tx")
TB = ("tx", "belongs", "bd")
BT = ("bd", "contains", "tx")


def set_num_graphs(batch: Tensor, num_graphs: int) -> Tensor:
    """Tag a ``batch`` vector with its number of graphs (what a PyG ``Batch`` knows as ``num_graphs``) so that no
    device read-back is needed to size the per-tile min/max table.  Batches assembled by ``segger_b200.tiles`` are
    tagged; an untagged vector is measured once (``batch.max()``, one sync, as the reference does at
    ist_encoder.py:67-68) and tagged with the result."""
    batch._sgb_num_graphs = (None if batch.is_inference() else batch._version, int(num_graphs))
    return batch


def _detached(x_dict):
    """The layer inputs kept for `attention_weights`, without their autograd history: holding the graph would keep
    last step's AccumulateGrad nodes (and their stream) alive into the next step, which breaks CUDA-graph capture."""
    return {k: (v.detach() if isinstance(v, Tensor) else v) for k, v in x_dict.items()}


def _known_num_graphs(batch: Optional[Tensor]) -> Optional[int]:
    if batch is None or batch.numel() == 0:
        return 1
    tag = getattr(batch, "_sgb_num_graphs", None)
    if tag is not None and tag[0] == (None if batch.is_inference() else batch._version):
        return tag[1]
    return None


def _num_batches(batch: Optional[Tensor]) -> int:
    n = _known_num_graphs(batch)
    if n is None:
        n = int(batch.max()) + 1
        set_num_graphs(batch, n)
    return n


class Positional2dEmbedder(Module):
    """ist_encoder.py:33-79.  ``forward`` normalises positions per tile, builds the 2 x 256-d sinusoid
    and applies the 2-layer SiLU MLP; returns [N, 2*dim] (x embedding | y embedding)."""

    def __init__(self, hidden_size: int, frequency_embedding_size: int = 256):
        super().__init__()
        self.dim = hidden_size // 2
        self.mlp = Sequential(
            NNLinear(frequency_embedding_size, self.dim, bias=True),
            SiLU(),
            NNLinear(self.dim, self.dim, bias=True),
        )
        self.frequency_embedding_size = frequency_embedding_size
        self._freqs: Dict[torch.device, Tensor] = {}

    def freqs(self, device) -> Tensor:
        f = self._freqs.get(device)
        if f is None:
            f = ops.sinusoid_freqs(self.frequency_embedding_size, 10000, device)
            self._freqs[device] = f
        return f

    def cheb_coef(self, device) -> Optional[Tensor]:
        """[deg, freq] matrix M with sinusoid features = chebyshev basis @ M (None if the layout has a pad column)."""
        if self.frequency_embedding_size % 2:
            return None
        key = ("cheb", device)
        m = self._freqs.get(key)
        if m is None:
            m = ops.cheb_feature_matrix(self.freqs(device))
            self._freqs[key] = m
        return m

    def basis(self, pos: Tensor, batch: Optional[Tensor]) -> Tensor:
        """[2N, deg] Chebyshev basis of the normalised coordinates: the low-rank form of ``features``."""
        return ops.poscheb(pos, batch, _num_batches(batch))

    def features(self, pos: Tensor, batch: Optional[Tensor]) -> Tensor:
        """[2, N, freq] sinusoid features of the normalised coordinates (no parameters involved)."""
        return ops.posfreq(pos, batch, _num_batches(batch), self.frequency_embedding_size, self.freqs(pos.device))

    def forward(self, pos: Tensor, batch: Optional[Tensor] = None) -> Tensor:
        feat = self.features(pos, batch)                   # [2, N, 256]
        N = pos.size(0)
        h = ops.linear(feat.view(2 * N, -1), self.mlp[0].weight, self.mlp[0].bias, ACT_SILU)
        h = ops.linear(h, self.mlp[2].weight, self.mlp[2].bias, ACT_NONE)   # [2N, dim]
        return h.view(2, N, self.dim).permute(1, 0, 2).reshape(N, 2 * self.dim)


class SkipGAT(Module):
    """ist_encoder.py:82-211: HeteroConv over three GATv2Conv (dropout 0.2, aggr='sum').

    As in the reference, only ``tx-neighbors-tx`` and ``tx-belongs-bd`` ever receive edges; the
    ``bd-contains-tx`` conv is declared (its lazy parameters stay unmaterialised, SURVEY Appendix
    B.1) and still runs -- through the generic path -- if a caller does supply such edges.
    ``attention_weights`` is exposed for parity of the interface; unlike the reference (where the
    hook stores junk, Appendix B.2) it is computed on demand from the last forward.
    """

    def __init__(self, in_channels, out_channels: int, n_heads: int, add_self_loops_tx: bool = False) -> None:
        super().__init__()
        self.conv = HeteroConv(
            convs={
                TT: GATv2Conv(in_channels=in_channels, out_channels=out_channels, heads=n_heads,
                              add_self_loops=add_self_loops_tx, dropout=0.2),
                TB: GATv2Conv(in_channels=in_channels, out_channels=out_channels, heads=n_heads,
                              add_self_loops=False, dropout=0.2),
                BT: GATv2Conv(in_channels=in_channels, out_channels=out_channels, heads=n_heads,
                              add_self_loops=False, dropout=0.2),
            },
            aggr="sum",
        )
        self._attn_weights: Dict[Tuple[str, str, str], Tensor] = {}
        self._last = None
        self._last_factor = None

    # -- fused path ----------------------------------------------------------------------------
    def _fusable(self, x_dict, edge_index_dict) -> bool:
        if TT not in edge_index_dict or TB not in edge_index_dict or BT in edge_index_dict:
            return False
        if "tx" not in x_dict or "bd" not in x_dict:
            return False
        if self.conv.convs[TT].add_self_loops:
            return False
        return all(isinstance(x_dict[k], Tensor) and x_dict[k].is_cuda for k in ("tx", "bd"))

    def forward_fused(self, x_dict: Dict[str, Tensor], edge_index_dict, apply_gelu: bool = False,
                      csr: Optional[dict] = None, tx_factor: Optional[Tuple[Tensor, Tensor]] = None) -> Dict[str, Tensor]:
        """``tx_factor = (ids, table)``: the transcript features are cat(table[ids], x_dict['tx']) in factored form
        (first layer of ISTEncoder: table = GELU(gene embedding)); see ops.SkipGATLayerFn."""
        tt, tb = self.conv.convs[TT], self.conv.convs[TB]
        x_tx, x_bd = x_dict["tx"].to(torch.float32), x_dict["bd"].to(torch.float32)
        tx_ids, tx_table = tx_factor if tx_factor is not None else (None, None)
        in_tx = x_tx.size(-1) + (tx_table.size(-1) if tx_table is not None else 0)
        tt.lin_l.materialize(in_tx, x_tx)
        tt.lin_r.materialize(in_tx, x_tx)
        tb.lin_l.materialize(in_tx, x_tx)
        tb.lin_r.materialize(x_bd.size(-1), x_bd)
        need_t = torch.is_grad_enabled()
        if csr is None:
            csr = {}
        N, M = x_tx.size(0), x_bd.size(0)
        csr_tt = csr.get(TT) or ops.CSR_CACHE.get(edge_index_dict[TT], N, N, need_t)
        csr_tb = csr.get(TB) or ops.CSR_CACHE.get(edge_index_dict[TB], N, M, need_t)
        training = self.training and tt.dropout > 0.0
        seed_tt = ops.new_seed() if training else 0
        seed_tb = ops.new_seed() if training else 0
        self._last = (_detached(x_dict), edge_index_dict)
        self._last_factor = None if tx_table is None else (tx_ids, tx_table.detach())
        h_tx, h_bd = ops.SkipGATLayerFn.apply(
            x_tx, x_bd,
            tt.lin_l.weight, tt.lin_l.bias, tt.lin_r.weight, tt.lin_r.bias, tt.att, tt.bias,
            tb.lin_l.weight, tb.lin_l.bias, tb.lin_r.weight, tb.lin_r.bias, tb.att, tb.bias,
            csr_tt, csr_tb, tt.heads, tt.out_channels, tt.negative_slope, tt.dropout, training,
            seed_tt, seed_tb, apply_gelu, torch.is_grad_enabled(), tx_ids, tx_table)
        return {"tx": h_tx, "bd": h_bd}

    def forward(self, x_dict: Dict[str, Tensor], edge_index_dict: Dict[str, Tensor]) -> Dict[str, Tensor]:
        if self._fusable(x_dict, edge_index_dict):
            return self.forward_fused(x_dict, edge_index_dict)
        self._last = (_detached(x_dict), edge_index_dict)
        self._last_factor = None
        # Reference builds {edge: False for edge in self.conv.convs} -- string keys that HeteroConv
        # never matches, i.e. a no-op kwarg (Appendix B.2); we simply do not pass it.
        return self.conv(x_dict, edge_index_dict)

    @property
    def attention_weights(self) -> Dict[Tuple[str, str, str], Tensor]:
        if self._last is None:
            raise AttributeError("Attention weights are empty. Please perform a forward pass.")
        x_dict, edge_index_dict = self._last
        with torch.no_grad():
            was = self.training
            self.eval()
            try:
                conv = self.conv.convs[TT]
                x_tx = x_dict["tx"].detach()
                if getattr(self, "_last_factor", None) is not None:      # factored first layer: rebuild cat(table[ids], x)
                    ids, table = self._last_factor
                    x_tx = torch.cat([table[ids.long()], x_tx], dim=-1)
                _, (_, alpha) = conv(x_tx, edge_index_dict[TT], return_attention_weights=True)
            finally:
                self.train(was)
        self._attn_weights[TT] = alpha
        return self._attn_weights


class ISTEncoder(torch.nn.Module):
    """ist_encoder.py:214-333, same signature and sub-module names.

    forward = input stage (Embedding / Linear + positional MLP + GELU) -> (n_mid_layers + 2) fused
    hetero GATv2 layers with GELU epilogues -> per-type output projection -> L2 normalisation.
    """

    def __init__(self, n_genes: int, in_channels: int = 16, hidden_channels: int = 32, out_channels: int = 32,
                 n_mid_layers: int = 3, n_heads: int = 3, normalize_embeddings: bool = True,
                 use_positional_embeddings: bool = True):
        super().__init__()
        self.normalize_embeddings = normalize_embeddings
        self.use_positional_embeddings = use_positional_embeddings
        self.hparams = locals()
        for k in ["self", "__class__"]:
            self.hparams.pop(k)
        self.lin_first = ModuleDict({
            "tx": Embedding(n_genes, in_channels),
            "bd": Linear(-1, in_channels),
        })
        self.pos_emb = Positional2dEmbedder(in_channels)
        self.conv_layers = ModuleList()
        self.conv_layers.append(SkipGAT((-1, -1), hidden_channels, n_heads))
        for _ in range(n_mid_layers):
            self.conv_layers.append(SkipGAT((-1, -1), hidden_channels, n_heads))
        self.conv_layers.append(SkipGAT((-1, -1), out_channels, n_heads))
        self.lin_last = HeteroDictLinear(-1, out_channels, types=("tx", "bd"))
        logger.debug(f"ISTEncoder: n_genes={n_genes}, in={in_channels}, hidden={hidden_channels}, "
                     f"out={out_channels}, layers={n_mid_layers + 2}")

    def _input_stage(self, k: str, x: Tensor, pos: Optional[Tensor], batch: Optional[Tensor],
                     skip_first: bool = False) -> Tensor:
        first = self.lin_first[k]
        feat = w0 = b0 = w2 = b2 = coef = None
        if self.use_positional_embeddings:
            coef = self.pos_emb.cheb_coef(pos.device)
            feat = self.pos_emb.basis(pos, batch) if coef is not None else self.pos_emb.features(pos, batch)
            w0, b0 = self.pos_emb.mlp[0].weight, self.pos_emb.mlp[0].bias
            w2, b2 = self.pos_emb.mlp[2].weight, self.pos_emb.mlp[2].bias
        if isinstance(first, Embedding):
            return ops.InputStageFn.apply(x, first.weight, None, feat, w0, b0, w2, b2, True, torch.is_grad_enabled(),
                                          coef, skip_first)
        first.materialize(x.size(-1), x)
        return ops.InputStageFn.apply(x, first.weight, first.bias, feat, w0, b0, w2, b2, False,
                                      torch.is_grad_enabled(), coef)

    def forward(self, x_dict: Dict[str, Tensor], edge_index_dict: Dict[str, Tensor], pos_dict: Dict[str, Tensor],
                batch_dict: Dict[str, Tensor]) -> Dict[str, Tensor]:
        for k, x in x_dict.items():
            require_cuda(x)
        with torch.cuda.device(next(iter(x_dict.values())).device):     # launches go to the tensors' device
            return self._forward(x_dict, edge_index_dict, pos_dict, batch_dict)

    def _resolve_meta(self, csrs, batch_dict) -> None:
        """Everything the host must know about a NEW batch, fetched with one device->host copy: the CSR status words
        (out-of-range node ids raise IndexError, as indexing with them would in the reference) and the number of
        tiles in each ``batch`` vector.  Nothing is read when the batch was seen before / was assembled by
        ``segger_b200.tiles`` (tagged) -- the forward then runs without any stream synchronisation."""
        todo = [c for c in csrs if c._src_unique is None] if (ops.VALIDATE and ops._DEFERRED is None) else []
        if ops._DEFERRED is not None:
            ops.validate_csrs(*csrs)
        need = []
        if self.use_positional_embeddings and batch_dict is not None:
            need = [b for b in batch_dict.values() if _known_num_graphs(b) is None]
        if not todo and not need:
            return
        parts = [c.status.to(torch.int64) for c in todo] + [b.max().to(torch.int64).reshape(1) for b in need]
        vals = torch.cat(parts).tolist()
        for b, v in zip(need, vals[2 * len(todo):]):
            set_num_graphs(b, v + 1)
        ops._apply_status(todo, [vals[2 * i:2 * i + 2] for i in range(len(todo))])

    def _build_csrs(self, x_dict, edge_index_dict, mark):
        need_t = torch.is_grad_enabled()
        N, M = x_dict["tx"].size(0), x_dict["bd"].size(0)
        pending = ops.csr_build_overlapped([(edge_index_dict[TT], N, N), (edge_index_dict[TB], N, M)], need_t, after=mark)
        pending.join()
        pending.resolve()                       # status words of a new graph, read from the side stream's copy
        self._resolve_meta(pending.csrs, None)  # (deferred / cached CSRs: the original path)
        return pending

    def _forward(self, x_dict, edge_index_dict, pos_dict, batch_dict):
        fused = (len(self.conv_layers) > 0 and self.conv_layers[0]._fusable(x_dict, edge_index_dict))
        # new graphs are built on a side stream while this stream runs the input stage (ops.csr_build_overlapped):
        # the input stage is enqueued first, the builds wait for the mark recorded here
        mark = ops.csr_overlap_mark(edge_index_dict[TT]) if fused else None
        csr, pending = None, None
        if fused and mark is None:                  # no overlap: build (and validate) first, as a plain caller would
            pending = self._build_csrs(x_dict, edge_index_dict, None)
            csr = {TT: pending.csrs[0], TB: pending.csrs[1]}
        self._resolve_meta([], batch_dict)          # tile counts of the batch vectors (the input stage needs them)
        # The transcript input is cat(GELU(Embedding[gene]), GELU(pos MLP)): its first half takes only n_genes distinct
        # values, so the first layer consumes it in factored form (ids + a [n_genes, in] table) and that half of its
        # projections becomes a table lookup instead of half the GEMM (ops.SkipGATLayerFn).  SEGGER_B200_FACTOR=0 turns
        # the factorisation off (A/B and parity of the two forms).
        first_tx = self.lin_first["tx"] if "tx" in self.lin_first else None
        factor = (fused and self.use_positional_embeddings and isinstance(first_tx, Embedding)
                  and first_tx.weight.size(1) % 4 == 0 and x_dict["tx"].dtype in (torch.int32, torch.int64)
                  and os.environ.get("SEGGER_B200_FACTOR", "1") != "0")
        # Input stage (ist_encoder.py:312-320)
        h_dict = {
            k: self._input_stage(k, x, pos_dict[k] if self.use_positional_embeddings else None,
                                 batch_dict.get(k) if self.use_positional_embeddings and batch_dict is not None else None,
                                 skip_first=(factor and k == "tx"))
            for k, x in x_dict.items()
        }
        tx_factor = None
        if factor:
            tx_factor = (x_dict["tx"].contiguous(), _GeluFn.apply(first_tx.weight))
        if fused and mark is not None:
            pending = self._build_csrs(x_dict, edge_index_dict, mark)
            csr = {TT: pending.csrs[0], TB: pending.csrs[1]}
        # Graph convolutions with GATv2 + GELU (ist_encoder.py:323-325)
        if fused:
            for li, conv_layer in enumerate(self.conv_layers):
                h_dict = conv_layer.forward_fused(h_dict, edge_index_dict, apply_gelu=True, csr=csr,
                                                  tx_factor=tx_factor if li == 0 else None)
        else:
            for conv_layer in self.conv_layers:
                h_dict = conv_layer(h_dict, edge_index_dict)
                h_dict = {k: _GeluFn.apply(v) for k, v in h_dict.items()}
        # Output projection + normalisation (ist_encoder.py:328-332)
        out = {}
        for k, h in h_dict.items():
            if k not in self.lin_last.lins:
                continue
            lin = self.lin_last.lins[k]
            lin.materialize(h.size(-1), h)
            out[k] = ops.OutputStageFn.apply(h, lin.weight, lin.bias, self.normalize_embeddings)
        return out


class _GeluFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        ctx.save_for_backward(x)
        return ops.act_fwd(x, ACT_GELU)

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        return ops.act_bwd(dy.contiguous(), x, ACT_GELU)
