"""torch-geometric-free equivalents of the three PyG modules the critic / transformer baseline use
(`MLP`, `Linear`, `LayerNorm(mode='graph')`), with PyG 2.5.2's parameter names (`lins.N.weight`,
`norms.N.weight`) so reference checkpoints load.  [3P-memory]: PyG is not installed offline; semantics
restated from the pinned version:
  MLP([a, b, c], norm=...): Linear -> norm -> ReLU for hidden layers, plain last Linear;
  LayerNorm(mode='graph', batch=None): statistics over the ENTIRE input tensor, x / (std + eps)."""
import math
from typing import List, Optional

import torch
import torch.nn.functional as F


class Linear(torch.nn.Module):
    # not a torch.nn.Linear subclass on purpose: the reference's critic re-initialisation
    # (builders/utils_algo_graph.py:195-198) skips PyG Linear layers for exactly that reason
    def __init__(self, in_channels: int, out_channels: int, bias: bool = True):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.weight = torch.nn.Parameter(torch.empty(out_channels, in_channels))
        self.bias = torch.nn.Parameter(torch.empty(out_channels)) if bias else None
        self.reset_parameters()

    def reset_parameters(self):
        torch.nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        if self.bias is not None:
            bound = 1 / math.sqrt(self.in_channels)
            torch.nn.init.uniform_(self.bias, -bound, bound)

    def forward(self, x):
        return F.linear(x, self.weight, self.bias)


class GraphLayerNorm(torch.nn.Module):
    """PyG LayerNorm(mode='graph') with batch=None: ONE mean / std over the entire input tensor.

    `groups` > 1 splits the leading axis into that many independent tensors (the time steps the reference
    evaluates in a Python loop, value/gnn_vf_net.py:72-78); `stats_reduce` (data parallelism) maps the local
    per-group (sum, sum of squares, count) to the global mean / biased std with gradients."""

    def __init__(self, in_channels: int, eps: float = 1e-5):
        super().__init__()
        self.in_channels, self.eps = in_channels, eps
        self.weight = torch.nn.Parameter(torch.ones(in_channels))
        self.bias = torch.nn.Parameter(torch.zeros(in_channels))
        self.stats_reduce = None
        # data parallel, fused critic kernels: (all_reduce(double tensor) -> None, world_size); set by DataParallel.attach
        self.fused_all_reduce = None

    def forward(self, x, groups: int = 1):
        shape = x.shape
        xg = x.reshape(groups, -1)
        if self.stats_reduce is None:
            xg = xg - xg.mean(dim=1, keepdim=True)
            out = xg / (xg.std(dim=1, unbiased=False, keepdim=True) + self.eps)
        else:
            mean, std = self.stats_reduce(xg)  # [groups, 1] each
            out = (xg - mean) / (std + self.eps)
        return out.reshape(shape) * self.weight + self.bias


class MLP(torch.nn.Module):
    def __init__(self, channel_list: List[int], norm: Optional[str] = "batch_norm", **kwargs):
        super().__init__()
        if norm not in (None, "layer_norm"):
            raise NotImplementedError("only norm=None | 'layer_norm' is used by the shipped configs")
        self.channel_list = list(channel_list)
        self.lins = torch.nn.ModuleList([Linear(a, b) for a, b in zip(channel_list[:-1], channel_list[1:])])
        self.norms = torch.nn.ModuleList(
            [GraphLayerNorm(h) if norm is not None else torch.nn.Identity() for h in channel_list[1:-1]])

    def forward(self, x, norm_groups: int = 1):
        for lin, norm in zip(self.lins[:-1], self.norms):
            x = lin(x)
            x = F.relu(norm(x, norm_groups) if isinstance(norm, GraphLayerNorm) else norm(x))
        return self.lins[-1](x)
