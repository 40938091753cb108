"""Helpers for the `-m gpu` parity tests: build product modules / graphs on cuda:0."""
import torch

from geometry_rl_b200.synthetic import CONFIGS, observation_layout, obs_keys


def dev():
    return torch.device("cuda:0")


def make_data(cfg, *, policy: bool):
    dims, names = observation_layout(cfg)
    if cfg.task == "rigid":
        from geometry_rl_b200.modules.pyg_data.rigid_tasks_data import RigidTasksData as D
    elif cfg.task == "rope":
        from geometry_rl_b200.modules.pyg_data.rope_tasks_data import RopeTasksData as D
    else:
        from geometry_rl_b200.modules.pyg_data.cloth_tasks_data import ClothTasksData as D
    concat = (not policy) or cfg.model == "transformer"
    return D(observation_dim=dims, observation_names=names, full_graph_obs=not policy, dist_as_pos=policy,
             output_mask_key="grippers" if policy else None, concat_input_vector=concat,
             angular_velocity=cfg.angular_velocity)


def enums(cfg):
    if cfg.task == "rigid":
        from geometry_rl_b200.modules.pyg_data import rigid_tasks_data as m
    elif cfg.task == "rope":
        from geometry_rl_b200.modules.pyg_data import rope_tasks_data as m
    else:
        from geometry_rl_b200.modules.pyg_data import cloth_tasks_data as m
    return m.NodeType, m.EdgeType, m.EdgeLevel


def make_policy_body(cfg):
    NodeType, EdgeType, EdgeLevel = enums(cfg)
    if cfg.model == "hepi":
        from geometry_rl_b200.modules.pyg_models.hepi import HEPi
        from geometry_rl_b200.modules.pyg_models.ponita.conv import FiberBundleConv
        codes = [[1, 0], [0, 1], [0, 1]]
        mp = [[FiberBundleConv(64, 64, 64, groups=64, separable=True, widening_factor=4) if c else None for c in code]
              for code in codes]
        net = HEPi(input_dim_node=len(NodeType) + cfg.policy_aux_dim, input_dim_edge=len(EdgeType) + 4, hidden_dim=64,
                   latent_dim=64, output_dim=cfg.output_dim, output_dim_vec=cfg.output_dim_vec, node_encoder_layers=2,
                   edge_encoder_layers=2, node_decoder_layers=2, node_type_mapping=NodeType, edge_type_mapping=EdgeType,
                   edge_level_mapping=EdgeLevel, message_passing=mp, num_messages=2, device="cuda", num_ori=16, degree=2,
                   ponita_dim=cfg.ponita_dim, only_upper_hemisphere=cfg.only_upper_hemisphere)
    elif cfg.model == "empn":
        from geometry_rl_b200.modules.pyg_models.ponita_gcn import PonitaGCN
        net = PonitaGCN(input_dim_node=len(NodeType) + cfg.policy_aux_dim, output_dim=cfg.output_dim,
                        output_dim_vec=cfg.output_dim_vec, num_layers=2, hidden_dim=64, dropout=0.0, num_ori=16, degree=2,
                        widening_factor=4, attention=False, ponita_dim=cfg.ponita_dim)
    else:
        from geometry_rl_b200.modules.pyg_models.transformer_vanilla import TransformerVanilla
        net = TransformerVanilla(input_dim_node=len(NodeType) + 12, output_dim=64, num_layers=2, num_heads=2,
                                 hidden_dim=64, dropout=0.0, concat_global=False)
    return net.to(dev())


def obs_args(cfg, obs, *, policy: bool):
    out = []
    for k in obs_keys(cfg):
        if policy and cfg.policy_pos_is_norm and k == "position_vectors":
            out.append(obs["norm_position_vectors"].to(dev()))
        elif policy and cfg.policy_pos_is_norm and k == "velocity_vectors":
            out.append(obs["norm_velocity_vectors"].to(dev()))
        else:
            out.append(obs[k].to(dev()))
    return out


def err_report(name, a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    scale = float(b.abs().max()) + 1e-30
    return f"{name}: max|diff|={float((a - b).abs().max()):.3e} scale={scale:.3e} rel={float((a - b).abs().max()) / scale:.3e}"


def rel(a, b):
    """max |a - b| / max |b|; NaN / inf anywhere -> inf, so that `rel(...) >= tol` style checks cannot pass on NaNs."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    r = float((a - b).abs().max()) / (float(b.abs().max()) + 1e-30)
    return r if r == r and bool(torch.isfinite(a).all()) else float("inf")
