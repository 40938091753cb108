"""nn namespace: MessagePassing, HeteroConv, MLP, LayerNorm, knn, knn_graph ([3P-memory])."""
import math

import torch
import torch.nn.functional as F

from . import conv  # noqa: F401
from .conv import MessagePassing, HeteroConv  # noqa: F401


class Linear(torch.nn.Module):
    """torch_geometric.nn.dense.linear.Linear — deliberately NOT a torch.nn.Linear subclass (the
    reference's orthogonal re-init of the critic, builders/utils_algo_graph.py:195-198, skips it)."""

    def __init__(self, in_channels, out_channels, bias=True):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.weight = torch.nn.Parameter(torch.empty(out_channels, in_channels))
        self.bias = torch.nn.Parameter(torch.empty(out_channels)) if bias else None
        torch.nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        if bias:
            bound = 1 / math.sqrt(in_channels)
            torch.nn.init.uniform_(self.bias, -bound, bound)

    def forward(self, x):
        return F.linear(x, self.weight, self.bias)


class LayerNorm(torch.nn.Module):
    """torch_geometric.nn.norm.LayerNorm, mode='graph': with batch=None the statistics run over
    the ENTIRE input tensor; denominator is (std + eps), not sqrt(var + eps)."""

    def __init__(self, in_channels, eps=1e-5, affine=True, mode="graph"):
        super().__init__()
        self.in_channels, self.eps, self.mode = in_channels, eps, mode
        self.weight = torch.nn.Parameter(torch.ones(in_channels)) if affine else None
        self.bias = torch.nn.Parameter(torch.zeros(in_channels)) if affine else None

    def forward(self, x, batch=None, batch_size=None):
        assert self.mode == "graph" and batch is None
        x = x - x.mean()
        out = x / (x.std(unbiased=False) + self.eps)
        if self.weight is not None and self.bias is not None:
            out = out * self.weight + self.bias
        return out


class MLP(torch.nn.Module):
    def __init__(self, channel_list, *, dropout=0.0, act="relu", act_first=False, norm="batch_norm",
                 plain_last=True, bias=True, **kwargs):
        super().__init__()
        assert act == "relu" and plain_last and dropout == 0.0
        assert norm in (None, "layer_norm")
        self.channel_list = channel_list
        self.act_first = act_first
        self.lins = torch.nn.ModuleList(
            [Linear(a, b, bias=bias) for a, b in zip(channel_list[:-1], channel_list[1:])]
        )
        self.norms = torch.nn.ModuleList()
        for hidden in channel_list[1:-1]:
            self.norms.append(LayerNorm(hidden) if norm is not None else torch.nn.Identity())

    def forward(self, x):
        for lin, norm in zip(self.lins[:-1], self.norms):
            x = lin(x)
            if self.act_first:
                x = F.relu(x)
            x = norm(x)
            if not self.act_first:
                x = F.relu(x)
        return self.lins[-1](x)


def knn(x, y, k, batch_x=None, batch_y=None):
    """torch_cluster.knn: for every y the k nearest x; returns [2, M] = (y index, x index), grouped by y,
    ascending distance.  Squared L2 accumulated as dx*dx + dy*dy + dz*dz in fp32, ties -> lower index."""
    assert batch_x is None and batch_y is None
    if x.numel() == 0 or y.numel() == 0:
        return torch.empty(2, 0, dtype=torch.long, device=x.device)
    d = None
    for c in range(x.size(1)):
        diff = y[:, None, c] - x[None, :, c]
        d = diff * diff if d is None else d + diff * diff
    kk = min(k, x.size(0))
    order = torch.argsort(d, dim=1, stable=True)[:, :kk]
    rows = torch.arange(y.size(0), device=x.device)[:, None].expand(-1, kk)
    return torch.stack([rows.reshape(-1), order.reshape(-1)], dim=0)


def knn_graph(x, k, batch=None, loop=False, flow="source_to_target"):
    assert flow == "source_to_target" and batch is None
    ei = knn(x, x, k if loop else k + 1)
    row, col = ei[1], ei[0]  # row = neighbour (source), col = centre (target)
    if not loop:
        mask = row != col
        row, col = row[mask], col[mask]
    return torch.stack([row, col], dim=0)
