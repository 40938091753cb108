"""Shared helpers for the parity tests (oracle side)."""
import os

import torch

from geometry_rl_b200.synthetic import CONFIGS, observation_layout, obs_keys
from oracle import graph as og

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    return torch.load(os.path.join(GOLDEN, f"{name}.pt"), weights_only=True)


def rel_err(a, b):
    a, b = a.double(), b.double()
    r = float((a - b).abs().max() / (b.abs().max() + 1e-30))
    return r if r == r else float("inf")


def oracle_graph_from_obs(cfg, obs, *, policy: bool):
    """Oracle topology + features for either the policy graph or the critic graph of `cfg`."""
    dims, names = observation_layout(cfg)
    sel = {k: obs[k] for k in obs_keys(cfg)}
    if policy and cfg.policy_pos_is_norm:
        sel["position_vectors"] = obs["norm_position_vectors"]
        sel["velocity_vectors"] = obs["norm_velocity_vectors"]
    parts = og.split_obs(sel, dims, names)
    task = og.TASKS[cfg.task]
    num_points = parts["infos"]["object_num_points"].long().reshape(-1) if cfg.task == "rigid" else None
    g = og.build_topology(task, parts["position_vectors"], full_graph_obs=not policy,
                          output_mask_key="grippers" if policy else None, num_points=num_points)
    og.update_positions(g, parts["position_vectors"], parts["norm_position_vectors"])
    concat = (not policy) or cfg.model == "transformer"
    feats = og.input_vectors(task, g, parts["norm_position_vectors"], parts["norm_velocity_vectors"],
                             dist_as_pos=policy, angular_velocity=cfg.angular_velocity, concat=concat)
    return g, feats
