"""MessagePassing stand-in ([3P-memory] torch-geometric 2.5.2, flow='source_to_target')."""
import inspect

import torch

from .hetero_conv import HeteroConv  # noqa: F401


class MessagePassing(torch.nn.Module):
    def __init__(self, aggr="add", *, aggr_kwargs=None, flow="source_to_target", node_dim=-2, **kwargs):
        super().__init__()
        assert flow == "source_to_target"
        self.aggr = aggr
        self.node_dim = node_dim
        self.aggr_module = None

    def propagate(self, edge_index, size=None, **kwargs):
        # j = source = edge_index[0]; i = target = edge_index[1]
        msg_params = list(inspect.signature(self.message).parameters)
        coll = {}
        for name in msg_params:
            if name.endswith("_j") or name.endswith("_i"):
                base, which = name[:-2], name[-1]
                data = kwargs[base]
                if isinstance(data, (tuple, list)):
                    data = data[0] if which == "j" else data[1]
                idx = edge_index[0] if which == "j" else edge_index[1]
                coll[name] = data.index_select(self.node_dim, idx)
            else:
                coll[name] = kwargs[name]
        out = self.message(**coll)
        x = kwargs.get("x")
        dim_size = kwargs.get("dim_size")
        if dim_size is None and isinstance(x, (tuple, list)):
            dim_size = x[1].size(self.node_dim)
        aggr_params = list(inspect.signature(self.aggregate).parameters)[1:]
        akw = {}
        for name in aggr_params:
            if name == "edge_index":
                akw[name] = edge_index
            elif name == "dim_size":
                akw[name] = dim_size
            elif name == "index":
                akw[name] = edge_index[1]
            elif name in kwargs:
                akw[name] = kwargs[name]
        out = self.aggregate(out, **akw)
        return self.update(out)

    def message(self, x_j):
        return x_j

    def update(self, inputs):
        return inputs
