"""Transformer baseline policy body: drop-in for geometry_rl/modules/pyg_models/transformer_vanilla.py (same class
name, constructor, parameter names).  The 2-layer post-LN encoder over the <= 56 tokens of a graph runs as ONE kernel per
layer and direction (grl_encoder_layer_fwd / _bwd, K5) on batch-major tokens; options the shipped config never selects
(active dropout, concat_global, other widths) take the nn.TransformerEncoder path on torch's kernels."""
from typing import Dict

import torch
from torch import nn

from ... import ops
from .pyg_compat import MLP


class TransformerVanilla(nn.Module):
    def __init__(self, input_dim_node, output_dim, num_layers=2, num_heads=2, hidden_dim=64, dropout=0.1,
                 concat_global=False, **ignored):
        super().__init__()
        self.input_dim = input_dim_node
        self._device = None
        self.concat_global = concat_global
        self.cls_token = nn.Parameter(torch.randn(1, 1, output_dim), requires_grad=True)
        self.embedding = nn.Linear(input_dim_node, hidden_dim)
        self.transformer_encoder_layer = nn.TransformerEncoderLayer(d_model=hidden_dim, nhead=num_heads,
                                                                    dim_feedforward=hidden_dim, dropout=dropout)
        self.transformer_encoder = nn.TransformerEncoder(self.transformer_encoder_layer, num_layers=num_layers,
                                                         enable_nested_tensor=False)
        self.fc_out = MLP([hidden_dim * 2 if concat_global else hidden_dim, output_dim], norm=None)

    @property
    def device(self):
        if self._device is None:
            self._device = next(self.parameters()).device
        return self._device

    def forward(self, data, input_vector, **kwargs):
        return self.one_step(data, input_vector, **kwargs)

    def one_step(self, graph, u_dict: Dict[str, torch.Tensor], **ignored):
        B = len(graph)
        with torch.no_grad():
            x = torch.cat([u_dict[t].reshape(B, -1, u_dict[t].shape[-1]) for t in graph.node_types], dim=1)
        x = self.embedding(x)
        mask = graph.output_mask
        if self.concat_global:
            x = torch.cat((self.cls_token.expand(B, -1, -1), x), dim=1)
            h = self.transformer_encoder(x.permute(1, 0, 2)).permute(1, 0, 2)
            cls = h[:, :1]
            h = h[:, slice(mask.start + 1, mask.stop + 1)]
            h = torch.cat([cls.expand(-1, h.shape[1], -1), h], dim=-1)
        else:
            layers = self.transformer_encoder.layers
            if self.transformer_encoder.norm is None and all(ops.encoder_layer_supported(x, l) for l in layers):
                h = x
                for l in layers:
                    h = ops.encoder_layer(h, l)
                h = h[:, mask]
            else:
                h = self.transformer_encoder(x.permute(1, 0, 2)).permute(1, 0, 2)[:, mask]
        return self.fc_out(h.reshape(-1, h.shape[-1]))
