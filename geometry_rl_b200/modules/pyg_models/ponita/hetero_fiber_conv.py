"""Heterogeneous wrapper with the reference's name and `convs` parameter layout
(geometry_rl/modules/pyg_models/ponita/hetero_fiber_conv.py:9-66 on top of PyG HeteroConv)."""
from typing import Dict, List, Tuple

import torch
from torch import Tensor

EdgeType = Tuple[str, str, str]


def edge_type_key(edge_type) -> str:
    """Key PyG >= 2.4 uses for tuple-keyed ModuleDicts: '<src___rel___dst>' [3P-memory]."""
    return "<" + "___".join(str.__str__(t) for t in edge_type) + ">"


def key_edge_type(key: str) -> EdgeType:
    return tuple(key[1:-1].split("___"))


class HeteroFiberConv(torch.nn.Module):
    def __init__(self, convs: Dict[EdgeType, torch.nn.Module], aggr: str = "sum"):
        super().__init__()
        if aggr != "sum":
            raise NotImplementedError("HeteroConv default aggr='sum' is the only grouping the reference uses")
        self.convs = torch.nn.ModuleDict({edge_type_key(k): v for k, v in convs.items()})
        self.aggr = aggr

    def forward(self, latent_dict: Dict[str, Tensor], edge_index_dict, edge_attr_dict, fiber_attr_dict,
                edge_set_dict=None) -> Dict[str, Tensor]:
        out_dst: Dict[str, List[Tensor]] = {}
        for key, conv in self.convs.items():
            et = key_edge_type(key)
            src, _, dst = et
            es = edge_set_dict[et]
            if es.n_edges == 0:  # hetero_fiber_conv.py:48-49
                continue
            x = latent_dict[src] if src == dst else (latent_dict[src], latent_dict[dst])
            _, upd = conv(x, edge_index=edge_index_dict[et], edge_attr=edge_attr_dict[et],
                          fiber_attr=fiber_attr_dict[et], edge_set=es)
            out_dst.setdefault(dst, []).append(upd)
        latent_dict = dict(latent_dict)
        for dst, vals in out_dst.items():  # group(values, "sum")
            latent_dict[dst] = vals[0] if len(vals) == 1 else torch.stack(vals, dim=0).sum(dim=0)
        return latent_dict
