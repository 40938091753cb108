"""Synthetic rollouts with the reference's observation layouts (Isaac Sim is not available offline).

The five workloads are the configs BASELINE.json names.  Shapes, bounds and minibatch sizes come from
the reference YAMLs (configs/<name>.yaml, cited per field below); observation group layouts follow
the Orbit observation managers (geometry_rl/orbit/tasks/manipulation/*/config/common_cfg/
observations_cfg.py) as tabulated in SURVEY.md Appendix B.  Object point counts are not recoverable
(meshes are missing blobs); SURVEY §8(d)'s substitute P_g in {8..48}, assigned in contiguous env
blocks and zero-padded to P_max = 48, is used.
"""
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import torch


@dataclass
class PathConfig:
    name: str
    task: str  # "rigid" | "rope" | "cloth"
    model: str  # "hepi" | "empn" | "transformer"
    num_actuators: int
    action_dim: int  # per actuator (configs/algorithm/policy/default.yaml:21 unless overridden)
    output_dim: int
    output_dim_vec: int
    ponita_dim: int
    only_upper_hemisphere: bool
    num_envs: int
    rollout_len: int
    mini_batch_size: int
    mean_bound: float = 0.05
    cov_bound: float = 0.0025
    trust_region_coeff: float = 1.0
    clip_grad_norm: bool = False
    max_grad_norm: float = 1.0
    angular_velocity: bool = True
    post_fc: bool = False
    policy_aux_dim: int = 4  # number of 3-vectors per node for the policy graph
    geometries: Tuple[int, ...] = (8, 12, 16, 20, 24, 32, 40, 48)
    p_max: int = 48
    rope_links: int = 80
    cloth_particles: int = 225
    cloth_hole: int = 10
    gamma: float = 0.99
    gae_lambda: float = 0.95
    critic_coef: float = 0.5
    entropy_coef: float = 0.005
    clip_value: float = 0.2
    lr: float = 3e-4
    ppo_epochs: int = 5
    policy_pos_is_norm: bool = False  # transformer cfg feeds norm_* into the position slots

    @property
    def total_action_dim(self) -> int:
        return self.action_dim * self.num_actuators


CONFIGS: Dict[str, PathConfig] = {
    # configs/rigid_insertion_multi_hepi_trpl_cfg.yaml:44,79,99,109-116,137-146
    "rigid_insertion_multi_hepi_trpl_cfg": PathConfig(
        "rigid_insertion_multi_hepi_trpl_cfg", "rigid", "hepi", 1, 6, 2, 2, 3, True, 1000, 100, 1000),
    # configs/rigid_pushing_multi_empn_trpl_cfg.yaml:98-116; BASELINE.json: 4096 envs x 16 steps
    "rigid_pushing_multi_empn_trpl_cfg": PathConfig(
        "rigid_pushing_multi_empn_trpl_cfg", "rigid", "empn", 1, 3, 1, 1, 2, False, 4096, 16, 4096,
        geometries=(8, 12, 16, 20, 24, 28, 32, 36, 40, 48)),
    # configs/cloth_hanging_multi_hepi_trpl_cfg.yaml:40,74,97-104,125,131-133
    "cloth_hanging_multi_hepi_trpl_cfg": PathConfig(
        "cloth_hanging_multi_hepi_trpl_cfg", "cloth", "hepi", 4, 3, 1, 1, 3, False, 100, 100, 200,
        cov_bound=0.001, trust_region_coeff=4.0, policy_aux_dim=3),
    # configs/rope_shaping_hepi_trpl_cfg.yaml:44,78,101-110,132-140
    "rope_shaping_hepi_trpl_cfg": PathConfig(
        "rope_shaping_hepi_trpl_cfg", "rope", "hepi", 2, 3, 1, 1, 2, False, 200, 200, 200,
        clip_grad_norm=True, policy_aux_dim=3),
    # configs/rigid_insertion_two_agents_multi_transformer_trpl_cfg.yaml:44,79,90-108,133-141
    "rigid_insertion_two_agents_multi_transformer_trpl_cfg": PathConfig(
        "rigid_insertion_two_agents_multi_transformer_trpl_cfg", "rigid", "transformer", 2, 3, 64, 0, 3, False,
        1000, 100, 1000, angular_velocity=False, post_fc=True, policy_aux_dim=4, policy_pos_is_norm=True),
}
CONFIG_ORDER = list(CONFIGS.keys())


def observation_layout(cfg: PathConfig):
    """(observation_dim, observation_names) as the Orbit observation manager reports them
    (`group_obs_term_dim` / `_group_obs_term_names`, builders/utils_algo_graph.py:68-71)."""
    A = cfg.num_actuators
    if cfg.task == "rigid":
        P = cfg.p_max
        e_max = 3 * P
        names = {
            "scalars": ["object_target_distances"],
            "position_vectors": ["grippers", "object_geometry", "target_geometry"],
            "velocity_vectors": (["grippers", "grippers_angular", "object_geometry", "object_geometry_angular"]
                                 if cfg.angular_velocity else ["grippers"]),
            "infos": ["object_num_points", "object_geometry_edges", "object_num_edges"],
        }
        dims = {
            "scalars": [(1,)],
            "position_vectors": [(3 * A,), (3 * P,), (3 * P,)],
            "velocity_vectors": [(3 * A,), (3 * A,), (3,), (3,)] if cfg.angular_velocity else [(3 * A,)],
            "infos": [(1,), (2 * e_max,), (1,)],
        }
    elif cfg.task == "rope":
        L = cfg.rope_links
        names = {
            "scalars": ["rope_target_distances"],
            "position_vectors": ["grippers", "links", "target_geometry"],
            "velocity_vectors": ["grippers", "links"],
        }
        dims = {
            "scalars": [(1,)],
            "position_vectors": [(3 * A,), (3 * L,), (3 * L,)],
            "velocity_vectors": [(3 * A,), (3 * L,)],
        }
    else:
        Np, H = cfg.cloth_particles, cfg.cloth_hole
        names = {
            "scalars": ["hole_target_distances", "cloth_edges_length"],
            "position_vectors": ["grippers", "particles", "init_particles", "hole_boundary", "target_hook"],
            "velocity_vectors": ["grippers", "particles"],
        }
        dims = {
            "scalars": [(H,), (2 * 4 * Np,)],
            "position_vectors": [(3 * A,), (3 * Np,), (3 * Np,), (3 * H,), (3,)],
            "velocity_vectors": [(3 * A,), (3 * Np,)],
        }
    return dims, names


def obs_keys(cfg: PathConfig) -> List[str]:
    keys = ["scalars", "position_vectors", "velocity_vectors", "norm_position_vectors", "norm_velocity_vectors"]
    if cfg.task == "rigid":
        keys.append("infos")
    return keys


def _standardise(x: torch.Tensor) -> torch.Tensor:
    """Stand-in for NDVecNorm (geometry_rl/torchrl/envs/transforms.py:72-171): per-3-vector-component
    standardisation over everything but the last axis, clipped to +-20 (ClipTransform)."""
    flat = x.reshape(-1, 3)
    out = (x - flat.mean(0)) / (flat.std(0) + 1e-2)
    return out.clamp(-20.0, 20.0)


def synthetic_obs(cfg: PathConfig, n: int, generator: torch.Generator, env_ids: Optional[torch.Tensor] = None
                  ) -> Dict[str, torch.Tensor]:
    """`n` frames of flat observation groups, fp32 on CPU.  For rigid tasks frame i belongs to env
    `env_ids[i]` (default i) whose geometry fixes the number of valid object points."""
    dims, names = observation_layout(cfg)
    g = generator

    def randn(*shape, scale=1.0):
        return torch.randn(*shape, generator=g) * scale

    obs: Dict[str, torch.Tensor] = {}
    A = cfg.num_actuators
    if cfg.task == "rigid":
        P = cfg.p_max
        if env_ids is None:
            env_ids = torch.arange(n)
        G = len(cfg.geometries)
        # contiguous env blocks per geometry (orbit/tasks/common/sim_utils.py:21-33)
        block = max(1, -(-cfg.num_envs // G))
        geom = (env_ids % cfg.num_envs) // block
        num_points = torch.tensor(cfg.geometries)[geom.clamp(max=G - 1)]
        valid = (torch.arange(P)[None, :] < num_points[:, None]).float()[..., None]
        obj = randn(n, P, 3, scale=0.05) * valid  # zero padded (orbit/tasks/common/utils.py:193-214)
        tgt = (randn(n, P, 3, scale=0.05) + randn(n, 1, 3, scale=0.2)) * valid
        grip = randn(n, A, 3, scale=0.2)
        pos = torch.cat([grip.reshape(n, -1), obj.reshape(n, -1), tgt.reshape(n, -1)], dim=1)
        if cfg.angular_velocity:
            vel = torch.cat([randn(n, 3 * A), randn(n, 3 * A), randn(n, 3), randn(n, 3)], dim=1)
        else:
            vel = randn(n, 3 * A)
        npos = torch.cat([_standardise(grip).reshape(n, -1), (_standardise(obj) * valid).reshape(n, -1),
                          (_standardise(tgt) * valid).reshape(n, -1)], dim=1)
        nvel = vel.clamp(-20, 20)
        e_max = 3 * P
        infos = torch.cat([num_points.float()[:, None], -torch.ones(n, 2 * e_max), (3 * num_points).float()[:, None]],
                          dim=1)
        obs = {"scalars": randn(n, 1), "position_vectors": pos, "velocity_vectors": vel,
               "norm_position_vectors": npos, "norm_velocity_vectors": nvel, "infos": infos}
        if cfg.policy_pos_is_norm:
            obs["policy_position_vectors"] = npos
    elif cfg.task == "rope":
        L = cfg.rope_links
        t = torch.linspace(0, 1, L)[None, :, None]
        links = torch.cat([t.expand(n, -1, -1) * 0.8, torch.zeros(n, L, 2)], dim=-1) + randn(n, L, 3, scale=0.004)
        links = links + randn(n, 1, 3, scale=0.1)
        tgt = links + randn(n, L, 3, scale=0.05)
        grip = torch.stack([links[:, 0], links[:, -1]], dim=1) + randn(n, A, 3, scale=0.01)
        pos = torch.cat([grip.reshape(n, -1), links.reshape(n, -1), tgt.reshape(n, -1)], dim=1)
        vel = torch.cat([randn(n, 3 * A), randn(n, 3 * L)], dim=1)
        npos = torch.cat([_standardise(grip).reshape(n, -1), _standardise(links).reshape(n, -1),
                          _standardise(tgt).reshape(n, -1)], dim=1)
        obs = {"scalars": torch.zeros(n, 1), "position_vectors": pos, "velocity_vectors": vel,
               "norm_position_vectors": npos, "norm_velocity_vectors": vel.clamp(-20, 20)}
    else:
        Np, H = cfg.cloth_particles, cfg.cloth_hole
        parts = randn(n, Np, 3, scale=0.3)
        init = parts + randn(n, Np, 3, scale=0.05)
        hole = randn(n, H, 3, scale=0.05) + randn(n, 1, 3, scale=0.2)
        hook = randn(n, 1, 3, scale=0.3)
        grip = randn(n, A, 3, scale=0.3)
        pos = torch.cat([grip.reshape(n, -1), parts.reshape(n, -1), init.reshape(n, -1), hole.reshape(n, -1),
                         hook.reshape(n, -1)], dim=1)
        vel = torch.cat([randn(n, 3 * A), randn(n, 3 * Np)], dim=1)
        npos = torch.cat([_standardise(grip).reshape(n, -1), _standardise(parts).reshape(n, -1),
                          _standardise(init).reshape(n, -1), _standardise(hole).reshape(n, -1),
                          _standardise(hook).reshape(n, -1)], dim=1)
        obs = {"scalars": randn(n, H + 2 * 4 * Np), "position_vectors": pos, "velocity_vectors": vel,
               "norm_position_vectors": npos, "norm_velocity_vectors": vel.clamp(-20, 20)}
    for k, v in obs.items():
        obs[k] = v.float().contiguous()
    return obs


def synthetic_rollout(cfg: PathConfig, generator: torch.Generator, num_envs: Optional[int] = None,
                      rollout_len: Optional[int] = None) -> Dict[str, torch.Tensor]:
    """[B_env, T(+1)] rollout tensors for GAE: obs groups are [B_env, T+1, F] (time T = next obs of the
    last step, the `shifted=True` layout of train.py:134-140), reward/done/terminated are [B_env, T]."""
    B = num_envs or cfg.num_envs
    T = rollout_len or cfg.rollout_len
    env_ids = torch.arange(B).repeat_interleave(T + 1)
    obs = synthetic_obs(cfg, B * (T + 1), generator, env_ids=env_ids)
    out = {k: v.reshape(B, T + 1, -1) for k, v in obs.items()}
    out["reward"] = torch.randn(B, T, generator=generator)
    done = torch.zeros(B, T, dtype=torch.bool)
    done[:, -1] = True  # only time-outs end episodes (rigid_tasks/config/common_cfg/terminations_cfg.py:12)
    out["done"] = done
    out["terminated"] = torch.zeros(B, T, dtype=torch.bool)
    return out


LOG_2PI = 1.8378770664093453


def synthetic_minibatch(obs: Dict[str, torch.Tensor], mean: torch.Tensor, var: torch.Tensor, value: torch.Tensor,
                        generator: torch.Generator, drift: float = 0.35) -> Dict[str, torch.Tensor]:
    """Minibatch content `TRPLLoss.forward` consumes, CPU tensors.  The old distribution is
    the current policy output (mean, var) pushed away by up to `drift` old-std units so that both branches of
    the mean and covariance projections are exercised (as after a few Adam steps in a real run); action ~ old
    distribution; advantage / value_target / state_value are random around the critic output."""
    mean, var, value = mean.detach().cpu(), var.detach().cpu(), value.detach().cpu()
    B, k = mean.shape
    g = generator
    s = torch.rand(B, 1, generator=g)
    q_mean = mean + torch.randn(B, k, generator=g) * drift * s * var.sqrt()
    q_var = var * torch.exp(torch.randn(B, k, generator=g) * 0.12 * s)
    action = q_mean + torch.randn(B, k, generator=g) * q_var.sqrt()
    logp = -0.5 * (((action - q_mean) ** 2 / q_var).sum(-1) + k * LOG_2PI + q_var.log().sum(-1))
    batch = {k_: v for k_, v in obs.items()}
    batch.update({
        "action": action, "loc": q_mean, "covariance_matrix": torch.diag_embed(q_var), "sample_log_prob": logp,
        "advantage": torch.randn(B, 1, generator=g),
        "state_value": value + 0.1 * torch.randn(B, 1, generator=g),
        "value_target": value + 0.5 * torch.randn(B, 1, generator=g),
    })
    return {k_: (t.float().contiguous() if torch.is_tensor(t) and t.is_floating_point() else t) for k_, t in batch.items()}
