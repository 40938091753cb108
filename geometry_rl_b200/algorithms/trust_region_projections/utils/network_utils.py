"""utils/network_utils.py:41-73 (`initialize_weights`) for the initialisation types the configs use."""
import torch.nn as nn


def initialize_weights(mod, initialization_type, gain: float = 0.01, scale=1 / 3, init_w=3e-3):
    for p in mod.parameters():
        if initialization_type == "orthogonal":
            if len(p.data.shape) >= 2:
                nn.init.orthogonal_(p.data, gain=gain)
            else:
                p.data.zero_()
        elif initialization_type == "xavier":
            if len(p.data.shape) >= 2:
                nn.init.xavier_uniform_(p.data)
            else:
                p.data.zero_()
        elif initialization_type == "uniform":
            if len(p.data.shape) >= 2:
                p.data.uniform_(-init_w, init_w)
            else:
                p.data.zero_()
        else:
            raise ValueError("Need a valid initialization key")
