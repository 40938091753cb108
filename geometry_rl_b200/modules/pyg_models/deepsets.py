"""DeepSets critic body: drop-in for geometry_rl/modules/pyg_models/deepsets.py (class name,
constructor, `one_step`, parameter names).  Dense [B, N, F] MLPs with whole-batch LayerNorm: these are
plain library GEMMs (cuBLAS through torch) plus three reductions — not one of the custom-kernel rows."""
from typing import Dict, List

import torch
from torch import nn

from .pyg_compat import MLP


class DeepSets(nn.Module):
    def __init__(self, input_dim_node, output_dim, hidden_dim=64, norm: List[str] = [None, None], **ignored):
        super().__init__()
        self.input_dim = input_dim_node
        self._device = None
        self.mlp_inner = MLP([input_dim_node, hidden_dim, hidden_dim], norm=norm[0])
        self.mlp_outer = MLP([hidden_dim, hidden_dim, output_dim], norm=norm[1])

    @property
    def device(self):
        if self._device is None:
            self._device = next(self.parameters()).device
        return self._device

    def forward(self, data, input_vector, **kwargs):
        return self.one_step(data, input_vector, **kwargs)

    def one_step(self, graph, u_dict: Dict[str, torch.Tensor], norm_groups: int = 1, **ignored):
        B = len(graph)
        with torch.no_grad():
            x = torch.cat([u_dict[t].reshape(B, -1, u_dict[t].shape[-1]) for t in graph.node_types], dim=1)
        x = self.mlp_inner(x, norm_groups)
        x = x.sum(dim=1)
        return self.mlp_outer(x, norm_groups)
