"""Oracle (test infrastructure): graph topology + per-node feature assembly on CPU.

Restates (paths under geometry_rl/modules/pyg_data/):
  rigid_tasks_data.py:93-150   _preprocess_input (split flat obs by term)          -> split_obs
  rigid_tasks_data.py:257-343  _construct_placeholders (kNN + dense loops + coalesce + batch)
  rope_tasks_data.py:227-300, cloth_tasks_data.py:224-307                          -> build_topology
  rigid_tasks_data.py:152-230, rope:143-201, cloth:144-198  construct_input_vector -> input_vectors
  transforms.py:43-76          HeteroNodeCategorical (one-hot over the PRE-subgraph type list)
  base_data.py:37-43           output_mask
  ../pyg_models/ponita_gcn.py:65-83  homogeneous edge index                          -> OGraph.homogeneous_edge_index
Third-party semantics restated from memory [3P-memory]: torch_cluster knn_graph (k+1 nearest incl.
self, drop self, row0 = neighbour / row1 = centre), PyG coalesce (sort by (row, col), dedup),
Batch.from_data_list (graph-major, per-type offsets), to_homogeneous (node_types / edge_types order).
kNN ties are implementation-defined upstream -> **parity unpinned**; here: squared L2 accumulated
as (dx*dx + dy*dy) + dz*dz in fp32 with separate roundings, ties -> lower index.
"""
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch


# ---------------------------------------------------------------------------------------------
# task specs (node / edge type names and their enum order in the reference)
# ---------------------------------------------------------------------------------------------
@dataclass
class TaskSpec:
    name: str
    all_node_types: List[str]  # NodeType enum order (one-hot width = len)
    particle_type: str  # node type carrying INTERNAL edges and TASK sources
    actuator_type: str
    internal_mode: str  # "knn" | "full"
    edge_types: List[Tuple[str, str, str]]  # insertion order: INTERNAL, AGENT, TASK

    def kept_node_types(self, full_graph_obs: bool) -> List[str]:
        if self.name == "rigid":
            return [t for t in self.all_node_types if t != "target_geometry"]
        if self.name == "rope":
            return list(self.all_node_types)
        keep = [t for t in self.all_node_types if t != "target_hook"]
        return keep if full_graph_obs else [t for t in keep if t != "particles"]


RIGID = TaskSpec("rigid", ["object_geometry", "grippers", "target_geometry"], "object_geometry", "grippers", "knn",
                 [("object_geometry", "internal", "object_geometry"), ("grippers", "agent", "grippers"),
                  ("object_geometry", "task", "grippers")])
ROPE = TaskSpec("rope", ["links", "grippers", "target_geometry"], "links", "grippers", "knn",
                [("links", "internal", "links"), ("grippers", "agent", "grippers"), ("links", "task", "grippers")])
CLOTH = TaskSpec("cloth", ["particles", "grippers", "hole_boundary", "target_hook"], "hole_boundary", "grippers", "full",
                 [("hole_boundary", "internal", "hole_boundary"), ("grippers", "agent", "grippers"),
                  ("hole_boundary", "task", "grippers")])
TASKS = {"rigid": RIGID, "rope": ROPE, "cloth": CLOTH}


@dataclass
class OGraph:
    num_graphs: int
    node_types: List[str]
    nodes_per_graph: Dict[str, int]
    edge_types: List[Tuple[str, str, str]]
    edge_index_dict: Dict[Tuple[str, str, str], torch.Tensor]  # batched, per-type global indices
    edge_counts: Dict[Tuple[str, str, str], torch.Tensor]  # [B] edges per graph
    pos: Dict[str, torch.Tensor] = field(default_factory=dict)
    norm_pos: Dict[str, torch.Tensor] = field(default_factory=dict)
    properties: Dict[str, torch.Tensor] = field(default_factory=dict)
    output_mask_key: Optional[str] = None
    output_mask: slice = slice(None)
    _homo: Optional[torch.Tensor] = None

    def __len__(self):
        return self.num_graphs

    def homogeneous_edge_index(self) -> torch.Tensor:
        """ponita_gcn.py:65-83: per graph, to_homogeneous() (edge types in insertion order, node
        types concatenated in node_types order), shifted by the running node count."""
        if self._homo is None:
            n_tot = sum(self.nodes_per_graph[t] for t in self.node_types)
            offs, o = {}, 0
            for t in self.node_types:
                offs[t] = o
                o += self.nodes_per_graph[t]
            cols, gids = [], []
            for et in self.edge_types:
                src, _, dst = et
                ei = self.edge_index_dict[et]
                g = torch.repeat_interleave(torch.arange(self.num_graphs), self.edge_counts[et])
                ls = ei[0] - g * self.nodes_per_graph[src]
                ld = ei[1] - g * self.nodes_per_graph[dst]
                cols.append(torch.stack([g * n_tot + offs[src] + ls, g * n_tot + offs[dst] + ld]))
                gids.append(g)
            ei = torch.cat(cols, dim=1)
            order = torch.sort(torch.cat(gids), stable=True).indices
            self._homo = ei[:, order]
        return self._homo


# ---------------------------------------------------------------------------------------------
# kNN (torch_cluster.knn_graph k, loop=False) + coalesce
# ---------------------------------------------------------------------------------------------
def sqdist_matrix(p: torch.Tensor) -> torch.Tensor:
    d = None
    for c in range(p.shape[1]):
        diff = p[:, None, c] - p[None, :, c]
        d = diff * diff if d is None else d + diff * diff
    return d


def coalesce(ei: torch.Tensor) -> torch.Tensor:
    if ei.numel() == 0:
        return ei.reshape(2, 0)
    n = int(ei.max()) + 1
    key = ei[0] * n + ei[1]
    key, perm = torch.sort(key, stable=True)
    keep = torch.ones_like(key, dtype=torch.bool)
    keep[1:] = key[1:] != key[:-1]
    return ei[:, perm][:, keep]


def knn_edges(points: torch.Tensor, k: int) -> torch.Tensor:
    """Coalesced [2,E] int64: row0 = neighbour (source), row1 = centre (target)."""
    P = points.shape[0]
    if P == 0:
        return torch.empty(2, 0, dtype=torch.long)
    d = sqdist_matrix(points.float())
    kk = min(k + 1, P)
    order = torch.argsort(d, dim=1, stable=True)[:, :kk]  # includes self at distance 0
    centre = torch.arange(P)[:, None].expand(-1, kk)
    row, col = order.reshape(-1), centre.reshape(-1)
    m = row != col
    return coalesce(torch.stack([row[m], col[m]]))


def full_edges(n: int) -> torch.Tensor:
    idx = [[j, k] for j in range(n) for k in range(n) if j != k]
    return torch.tensor(idx, dtype=torch.long).T.reshape(2, -1)


def bipartite_edges(n_src: int, n_dst: int) -> torch.Tensor:
    idx = [[j, k] for j in range(n_src) for k in range(n_dst)]
    return torch.tensor(idx, dtype=torch.long).T.reshape(2, -1)


# ---------------------------------------------------------------------------------------------
# obs splitting (rigid_tasks_data.py:93-150)
# ---------------------------------------------------------------------------------------------
def split_obs(obs: Dict[str, torch.Tensor], observation_dim, observation_names) -> Dict[str, Dict[str, torch.Tensor]]:
    out = {}
    for group, tensor in obs.items():
        if group.startswith("norm_"):
            dims_key = group[len("norm_"):]
        else:
            dims_key = group
        dims = [d[0] if isinstance(d, (tuple, list)) else d for d in observation_dim[dims_key]]
        names = observation_names[dims_key]
        parts = torch.split(tensor, dims, dim=1)
        out[group] = {}
        for nm, part in zip(names, parts):
            if "vectors" in group:
                part = part.reshape(tensor.shape[0], -1, 3)
            out[group][nm] = part
    return out


# ---------------------------------------------------------------------------------------------
# topology (the *_construct_placeholders restatement)
# ---------------------------------------------------------------------------------------------
def build_topology(task: TaskSpec, position_vectors: Dict[str, torch.Tensor], *, full_graph_obs: bool,
                   output_mask_key: Optional[str], knn_k: int = 3,
                   num_points: Optional[torch.Tensor] = None) -> OGraph:
    B = position_vectors[task.actuator_type].shape[0]
    kept = task.kept_node_types(full_graph_obs)
    npg = {t: position_vectors[t].shape[1] for t in task.all_node_types}
    e_int, e_agent, e_task = task.edge_types
    per_graph = {et: [] for et in task.edge_types}
    A = npg[task.actuator_type]
    agent = full_edges(A) if A > 1 else torch.empty(2, 0, dtype=torch.long)
    for i in range(B):
        pts = position_vectors[task.particle_type][i]
        n_valid = int(num_points[i]) if num_points is not None else pts.shape[0]
        if task.internal_mode == "knn":
            per_graph[e_int].append(knn_edges(pts[:n_valid], knn_k))
        else:
            per_graph[e_int].append(full_edges(pts.shape[0]))
        per_graph[e_agent].append(agent)
        per_graph[e_task].append(bipartite_edges(n_valid, A))

    eid, ecnt = {}, {}
    for et in task.edge_types:
        src, _, dst = et
        if src not in kept or dst not in kept:
            continue
        shifted = []
        for i, ei in enumerate(per_graph[et]):
            off = torch.tensor([[i * npg[src]], [i * npg[dst]]], dtype=torch.long)
            shifted.append(ei + off)
        eid[et] = torch.cat(shifted, dim=1)
        ecnt[et] = torch.tensor([ei.shape[1] for ei in per_graph[et]], dtype=torch.long)

    g = OGraph(num_graphs=B, node_types=kept, nodes_per_graph={t: npg[t] for t in kept},
               edge_types=[et for et in task.edge_types if et in eid], edge_index_dict=eid, edge_counts=ecnt,
               output_mask_key=output_mask_key)
    for t in kept:  # HeteroNodeCategorical: index in the pre-subgraph node type list
        oh = torch.zeros(B * npg[t], len(task.all_node_types))
        oh[:, task.all_node_types.index(t)] = 1
        g.properties[t] = oh
    if output_mask_key is not None:
        start = 0
        for t in kept:
            if t == output_mask_key:
                break
            start += npg[t]
        g.output_mask = slice(start, start + npg[output_mask_key])
    return g


def update_positions(g: OGraph, position_vectors, norm_position_vectors) -> OGraph:
    """*_tasks_data.py `_update_placeholders`: topology re-used, pos / norm_pos refreshed."""
    for t in g.node_types:
        g.pos[t] = position_vectors[t].reshape(-1, 3)
        g.norm_pos[t] = norm_position_vectors[t].reshape(-1, 3)
    return g


# ---------------------------------------------------------------------------------------------
# per-node features (construct_input_vector), training noise off
# ---------------------------------------------------------------------------------------------
def input_vectors(task: TaskSpec, g: OGraph, norm_position_vectors, norm_velocity_vectors, *, dist_as_pos: bool,
                  angular_velocity: bool = True, concat: bool = False):
    scalar_dict, vector_dict, full = {}, {}, {}
    for t in g.node_types:
        pos = g.norm_pos[t]
        zeros = torch.zeros_like(pos)
        if task.name == "rigid":
            if t == task.particle_type:
                target = norm_position_vectors["target_geometry"].reshape(-1, 3)
                corr = pos - target if dist_as_pos else target
            else:
                corr = zeros
            if t in norm_velocity_vectors:
                if t == task.particle_type:
                    n = norm_position_vectors[t].shape[1]
                    vel = norm_velocity_vectors[t].repeat_interleave(n, dim=1).reshape(-1, 3)
                    ang = (norm_velocity_vectors[f"{t}_angular"].repeat_interleave(n, dim=1).reshape(-1, 3)
                           if angular_velocity else torch.zeros_like(vel))
                else:
                    vel = norm_velocity_vectors[t].reshape(-1, 3)
                    ang = (norm_velocity_vectors[f"{t}_angular"].reshape(-1, 3) if angular_velocity
                           else torch.zeros_like(vel))
            else:
                vel, ang = zeros, zeros
            vectors = torch.cat([pos, corr, vel, ang], dim=1)
        elif task.name == "rope":
            if t == task.particle_type:
                target = norm_position_vectors["target_geometry"].reshape(-1, 3)
                corr = pos - target if dist_as_pos else target
            else:
                corr = zeros
            vel = norm_velocity_vectors[t].reshape(-1, 3) if t in norm_velocity_vectors else zeros
            vectors = torch.cat([pos, corr, vel], dim=1)
        else:  # cloth
            if t == "particles":
                init = norm_position_vectors["init_particles"].reshape(-1, 3)
                corr = pos - init if dist_as_pos else init
            elif t == "hole_boundary":
                n = norm_position_vectors["hole_boundary"].shape[1]
                target = torch.repeat_interleave(norm_position_vectors["target_hook"], n, 1).reshape(-1, 3)
                corr = pos - target if dist_as_pos else target
            else:
                corr = zeros
            vel = norm_velocity_vectors[t].reshape(-1, 3) if t in norm_velocity_vectors else zeros
            vectors = torch.cat([pos, corr, vel], dim=1)
        scalar_dict[t] = g.properties[t]
        vector_dict[t] = vectors
        full[t] = torch.cat([g.properties[t], vectors], dim=1)
    return full if concat else (scalar_dict, vector_dict)
