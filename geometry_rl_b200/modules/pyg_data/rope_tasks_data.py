"""Rope tasks (closing / shaping): drop-in for geometry_rl/modules/pyg_data/rope_tasks_data.py."""
import enum
from typing import Tuple

import torch

from .base_data import BaseData


class NodeType(str, enum.Enum):
    LINKS = "links"
    ACTUATOR = "grippers"
    TARGET_GEOMETRY = "target_geometry"


class EdgeLevel(str, enum.Enum):
    INTERNAL = "internal"
    TASK = "task"
    AGENT = "agent"


class EdgeType(Tuple[str, str, str], enum.Enum):
    LINKS_INTERNAL_LINKS = (NodeType.LINKS, EdgeLevel.INTERNAL, NodeType.LINKS)
    ACTUATOR_AGENT_ACTUATOR = (NodeType.ACTUATOR, EdgeLevel.AGENT, NodeType.ACTUATOR)
    LINKS_TASK_ACTUATOR = (NodeType.LINKS, EdgeLevel.TASK, NodeType.ACTUATOR)


class RopeTasksData(BaseData):
    TASK = "rope"
    ALL_NODE_TYPES = ("links", "grippers", "target_geometry")
    PARTICLE_TYPE = "links"
    INTERNAL_MODE = "knn"
    EDGE_TYPES = (("links", "internal", "links"), ("grippers", "agent", "grippers"), ("links", "task", "grippers"))

    def _kept_node_types(self):
        # rope_tasks_data.py:89: all node types, target_geometry stays as isolated nodes
        return list(self.ALL_NODE_TYPES)

    def _vectors(self, data, t, npv, nvv, train):
        # rope_tasks_data.py:160-196
        pos = self._noisy(data[t].norm_pos, train)
        if t == self.PARTICLE_TYPE:
            target = npv["target_geometry"].reshape(-1, 3)
            corr = self._noisy(pos - target if self.dist_as_pos else target, train)
        else:
            corr = torch.zeros_like(pos)
        vel = self._noisy(nvv[t].reshape(-1, 3), train) if t in nvv else torch.zeros_like(pos)
        return torch.cat([pos, corr, vel], dim=1)
