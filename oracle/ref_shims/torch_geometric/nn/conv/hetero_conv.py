"""HeteroConv / group stand-ins ([3P-memory] torch-geometric 2.5.2)."""
import torch


def group(xs, aggr):
    if len(xs) == 0:
        return None
    elif aggr is None:
        return torch.stack(xs, dim=1)
    elif len(xs) == 1:
        return xs[0]
    elif aggr == "cat":
        return torch.cat(xs, dim=-1)
    else:
        out = torch.stack(xs, dim=0)
        out = getattr(torch, aggr)(out, dim=0)
        out = out[0] if isinstance(out, tuple) else out
        return out


class _TupleKeyModuleDict(torch.nn.ModuleDict):
    """PyG >= 2.4 `torch_geometric.nn.module_dict.ModuleDict`: tuple keys are stored as
    '<a___b___c>' and handed back as tuples by items()/keys()."""

    @staticmethod
    def to_internal_key(key):
        if isinstance(key, tuple):
            key = "<" + "___".join(str.__str__(k) for k in key) + ">"
        return key.replace(".", "#")

    @staticmethod
    def to_external_key(key):
        key = key.replace("#", ".")
        if key.startswith("<") and key.endswith(">") and "___" in key:
            key = tuple(key[1:-1].split("___"))
        return key

    def __init__(self, modules=None):
        super().__init__()
        if modules is not None:
            for k, v in modules.items():
                self[k] = v

    def __getitem__(self, key):
        return super().__getitem__(self.to_internal_key(key))

    def __setitem__(self, key, module):
        super().__setitem__(self.to_internal_key(key), module)

    def __contains__(self, key):
        return super().__contains__(self.to_internal_key(key))

    def keys(self):
        return [self.to_external_key(k) for k in super().keys()]

    def items(self):
        return [(self.to_external_key(k), v) for k, v in super().items()]


class HeteroConv(torch.nn.Module):
    def __init__(self, convs, aggr="sum"):
        super().__init__()
        self.convs = _TupleKeyModuleDict(convs)
        self.aggr = aggr
