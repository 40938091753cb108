"""Cloth hanging: drop-in for geometry_rl/modules/pyg_data/cloth_tasks_data.py."""
import enum
from typing import Tuple

import torch

from .base_data import BaseData


class NodeType(str, enum.Enum):
    PARTICLES = "particles"
    ACTUATOR = "grippers"
    HOLE_BOUNDARY = "hole_boundary"
    TARGET = "target_hook"


class EdgeLevel(str, enum.Enum):
    INTERNAL = "internal"
    TASK = "task"
    AGENT = "agent"


class EdgeType(Tuple[str, str, str], enum.Enum):
    PARTICLES_INTERNAL_PARTICLES = (NodeType.HOLE_BOUNDARY, EdgeLevel.INTERNAL, NodeType.HOLE_BOUNDARY)
    ACTUATOR_AGENT_ACTUATOR = (NodeType.ACTUATOR, EdgeLevel.AGENT, NodeType.ACTUATOR)
    HOLE_BOUNDARY_TASK_ACTUATOR = (NodeType.HOLE_BOUNDARY, EdgeLevel.TASK, NodeType.ACTUATOR)


class ClothTasksData(BaseData):
    TASK = "cloth"
    ALL_NODE_TYPES = ("particles", "grippers", "hole_boundary", "target_hook")
    PARTICLE_TYPE = "hole_boundary"
    INTERNAL_MODE = "full"  # cloth_tasks_data.py:248-257: fully connected hole-boundary nodes
    EDGE_TYPES = (("hole_boundary", "internal", "hole_boundary"), ("grippers", "agent", "grippers"),
                  ("hole_boundary", "task", "grippers"))

    def _kept_node_types(self):
        # cloth_tasks_data.py:87-91
        keep = [t for t in self.ALL_NODE_TYPES if t != "target_hook"]
        return keep if self.full_graph_obs else [t for t in keep if t != "particles"]

    def _vectors(self, data, t, npv, nvv, train):
        # cloth_tasks_data.py:160-193 (no training noise in the reference's cloth builder)
        pos = data[t].norm_pos
        if t == "particles":
            init = npv["init_particles"].reshape(-1, 3)
            corr = pos - init if self.dist_as_pos else init
        elif t == "hole_boundary":
            n = npv["hole_boundary"].shape[1]
            target = torch.repeat_interleave(npv["target_hook"], n, 1).reshape(-1, 3)
            corr = pos - target if self.dist_as_pos else target
        else:
            corr = torch.zeros_like(pos)
        vel = nvv[t].reshape(-1, 3) if t in nvv else torch.zeros_like(pos)
        return torch.cat([pos, corr, vel], dim=1)
