"""Dict-backed Data / HeteroData / Batch with the subset of PyG 2.5.2 behaviour the reference's
pyg_data builders and models touch.  [3P-memory] semantics:
  * node/edge stores keep attributes in insertion order; node_types / edge_types follow insertion;
  * num_nodes of a node store = size(0) of its first tensor attribute;
  * coalesce(): per edge type sort edge_index lexicographically by (row, col) and drop duplicates;
  * Batch.from_data_list: graph-major concatenation, edge indices shifted by per-type node offsets;
  * node_offsets: running offset of node types in store order;
  * to_homogeneous(): nodes concatenated in node_types order, edges in edge_types order.
"""
import copy

import torch

from . import datapipes  # noqa: F401


class _Store:
    def __init__(self, key=None):
        object.__setattr__(self, "_key", key)
        object.__setattr__(self, "_d", {})

    def __setattr__(self, name, value):
        self._d[name] = value

    def __getattr__(self, name):
        d = object.__getattribute__(self, "_d")
        if name in d:
            return d[name]
        if name == "num_nodes":
            for v in d.values():
                if isinstance(v, torch.Tensor):
                    return v.size(0)
            return 0
        raise AttributeError(name)

    def __contains__(self, name):
        return name in self._d

    def keys(self):
        return list(self._d.keys())

    def items(self):
        return self._d.items()

    def clone(self):
        s = _Store(self._key)
        for k, v in self._d.items():
            s._d[k] = v.clone() if isinstance(v, torch.Tensor) else copy.copy(v)
        return s

    def to(self, device):
        for k, v in list(self._d.items()):
            if isinstance(v, torch.Tensor):
                self._d[k] = v.to(device)
        return self


class Data(_Store):
    def __init__(self, **kwargs):
        super().__init__(None)
        for k, v in kwargs.items():
            self._d[k] = v

    def __getattr__(self, name):
        d = object.__getattribute__(self, "_d")
        if name in d:
            return d[name]
        if name in ("edge_attr", "pos", "x", "edge_index"):
            return None
        return super().__getattr__(name)


class HeteroData:
    def __init__(self):
        self.__dict__["_nodes"] = {}
        self.__dict__["_edges"] = {}
        self.__dict__["_glob"] = {}
        self.__dict__["_num_graphs"] = None
        self.__dict__["_node_counts"] = None  # type -> list[int] per graph (Batch only)
        self.__dict__["_edge_counts"] = None

    # ---- store access ---------------------------------------------------------------------
    @staticmethod
    def _norm_key(key):
        if isinstance(key, tuple):
            return tuple(str.__str__(k) if isinstance(k, str) else k for k in key)
        return str.__str__(key) if isinstance(key, str) else key

    def __getitem__(self, key):
        if isinstance(key, int):
            return self.get_example(key)
        key = self._norm_key(key)
        if isinstance(key, tuple):
            if key not in self._edges:
                self._edges[key] = _Store(key)
            return self._edges[key]
        if key not in self._nodes:
            self._nodes[key] = _Store(key)
        return self._nodes[key]

    def __setattr__(self, name, value):
        self._glob[name] = value

    def __getattr__(self, name):
        g = self.__dict__["_glob"]
        if name in g:
            return g[name]
        raise AttributeError(name)

    def __len__(self):
        if self._num_graphs is not None:
            return self._num_graphs
        return len(self._nodes) + len(self._edges)

    @property
    def node_types(self):
        return list(self._nodes.keys())

    @property
    def edge_types(self):
        return list(self._edges.keys())

    @property
    def edge_index_dict(self):
        return {k: s.edge_index for k, s in self._edges.items() if "edge_index" in s}

    @property
    def num_nodes(self):
        return sum(s.num_nodes for s in self._nodes.values())

    @property
    def node_offsets(self):
        out, off = {}, 0
        for k, s in self._nodes.items():
            out[k] = off
            off += s.num_nodes
        return out

    # ---- transforms -------------------------------------------------------------------------
    def coalesce(self):
        for s in self._edges.values():
            ei = s.edge_index
            if ei.numel() == 0:
                continue
            n = int(ei.max()) + 1
            key = ei[0] * n + ei[1]
            key_sorted, perm = torch.sort(key, stable=True)
            keep = torch.ones_like(key_sorted, dtype=torch.bool)
            keep[1:] = key_sorted[1:] != key_sorted[:-1]
            s.edge_index = ei[:, perm][:, keep]
        return self

    def to(self, device):
        for s in list(self._nodes.values()) + list(self._edges.values()):
            s.to(device)
        return self

    def clone(self):
        out = HeteroData()
        out.__dict__["_nodes"] = {k: s.clone() for k, s in self._nodes.items()}
        out.__dict__["_edges"] = {k: s.clone() for k, s in self._edges.items()}
        out.__dict__["_glob"] = dict(self._glob)
        out.__dict__["_num_graphs"] = self._num_graphs
        out.__dict__["_node_counts"] = self._node_counts
        out.__dict__["_edge_counts"] = self._edge_counts
        return out

    def node_type_subgraph(self, node_types):
        keep = [self._norm_key(t) for t in node_types]
        out = HeteroData()
        out.__dict__["_nodes"] = {k: s for k, s in self._nodes.items() if k in keep}
        out.__dict__["_edges"] = {k: s for k, s in self._edges.items() if k[0] in keep and k[2] in keep}
        out.__dict__["_glob"] = dict(self._glob)
        out.__dict__["_num_graphs"] = self._num_graphs
        if self._node_counts is not None:
            out.__dict__["_node_counts"] = {k: v for k, v in self._node_counts.items() if k in keep}
            out.__dict__["_edge_counts"] = {k: v for k, v in self._edge_counts.items() if k in out._edges}
        return out

    def to_homogeneous(self):
        offs = self.node_offsets
        eis = []
        for (src, _, dst), s in self._edges.items():
            off = torch.tensor([[offs[src]], [offs[dst]]], dtype=s.edge_index.dtype, device=s.edge_index.device)
            eis.append(s.edge_index + off)
        return Data(edge_index=torch.cat(eis, dim=-1), num_nodes=self.num_nodes)

    # ---- Batch behaviour ----------------------------------------------------------------------
    def get_example(self, i):
        assert self._num_graphs is not None
        out = HeteroData()
        node_start = {}
        for k, s in self._nodes.items():
            c = self._node_counts[k]
            a = sum(c[:i])
            node_start[k] = a
            st = out[k]
            for name, v in s.items():
                if isinstance(v, torch.Tensor) and v.size(0) == sum(c):
                    st._d[name] = v[a : a + c[i]]
        for k, s in self._edges.items():
            c = self._edge_counts[k]
            a = sum(c[:i])
            st = out[k]
            for name, v in s.items():
                if name == "edge_index":
                    off = torch.tensor([[node_start[k[0]]], [node_start[k[2]]]], dtype=v.dtype, device=v.device)
                    st._d[name] = v[:, a : a + c[i]] - off
                elif isinstance(v, torch.Tensor) and v.size(0) == sum(c):
                    st._d[name] = v[a : a + c[i]]
        return out


class Batch:
    @staticmethod
    def from_data_list(data_list):
        first = data_list[0]
        out = HeteroData()
        node_counts = {k: [d[k].num_nodes for d in data_list] for k in first.node_types}
        edge_counts = {k: [d[k].edge_index.size(1) for d in data_list] for k in first.edge_types}
        for k in first.node_types:
            st = out[k]
            for name in first[k].keys():
                st._d[name] = torch.cat([d[k]._d[name] for d in data_list], dim=0)
        for k in first.edge_types:
            st = out[k]
            src, _, dst = k
            so, do = 0, 0
            eis = []
            for gi, d in enumerate(data_list):
                ei = d[k].edge_index
                off = torch.tensor([[so], [do]], dtype=ei.dtype, device=ei.device)
                eis.append(ei + off)
                so += node_counts[src][gi]
                do += node_counts[dst][gi]
            st._d["edge_index"] = torch.cat(eis, dim=1)
        out.__dict__["_num_graphs"] = len(data_list)
        out.__dict__["_node_counts"] = node_counts
        out.__dict__["_edge_counts"] = edge_counts
        return out
