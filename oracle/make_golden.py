"""Generate tests/golden/*.pt by executing the UNMODIFIED reference from /root/reference.

TEST INFRASTRUCTURE.  Run in the build container only (the reference checkout does not exist on the
GPU box):  `python -m oracle.make_golden`.  The reference modules are imported from
/root/reference (never copied) against the stand-ins in oracle/ref_shims for the third-party
packages that are not installed.  Each fixture stores inputs, the reference module's state_dict,
outputs and gradients (fp32, small sizes) so that the oracle restatement and the CUDA path can be
checked without the reference being present.
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("GRL_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.join(HERE, "ref_shims"))
sys.path.insert(0, REF)
sys.path.insert(0, ROOT)

from geometry_rl_b200.synthetic import CONFIGS, observation_layout, obs_keys, synthetic_obs  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def _plain(obj):
    """Strip reference-defined Enum members so that fixtures unpickle without the reference present
    (and load under torch.load(weights_only=True))."""
    import enum
    if isinstance(obj, enum.Enum):
        obj = obj.value
    if isinstance(obj, torch.Tensor):
        return obj.detach().clone()
    if isinstance(obj, str):
        return str.__str__(obj) + ""
    if isinstance(obj, dict):
        return {_plain(k): _plain(v) for k, v in obj.items()}
    if isinstance(obj, tuple):
        return tuple(_plain(v) for v in obj)
    if isinstance(obj, list):
        return [_plain(v) for v in obj]
    return obj


def _save(rec, tag):
    torch.save(_plain(rec), os.path.join(OUT, f"{tag}.pt"))


def _sd(module):
    return {k: v.detach().clone() for k, v in module.state_dict().items()}


def _grads(module):
    return {k: (p.grad.detach().clone() if p.grad is not None else None) for k, p in module.named_parameters()}


def _ref_data(cfg, **base_kwargs):
    dims, names = observation_layout(cfg)
    if cfg.task == "rigid":
        from geometry_rl.modules.pyg_data.rigid_tasks_data import RigidTasksData as D, NodeType, EdgeType, EdgeLevel
        base_kwargs.setdefault("angular_velocity", cfg.angular_velocity)
    elif cfg.task == "rope":
        from geometry_rl.modules.pyg_data.rope_tasks_data import RopeTasksData as D, NodeType, EdgeType, EdgeLevel
    else:
        from geometry_rl.modules.pyg_data.cloth_tasks_data import ClothTasksData as D, NodeType, EdgeType, EdgeLevel
    # Python >= 3.11 formats a str-mixin Enum member as "NodeType.X"; the reference (written for the
    # Python 3.10 of Isaac Sim) relies on f"{node_type}_angular" yielding the VALUE
    # (rigid_tasks_data.py:199).  Restore the 3.10 behaviour from outside, source untouched.
    NodeType.__format__ = lambda self, spec: str.__format__(self.value, spec)
    return D(observation_dim=dims, observation_names=names, **base_kwargs), NodeType, EdgeType, EdgeLevel


def _ref_hepi(cfg, NodeType, EdgeType, EdgeLevel):
    from geometry_rl.modules.pyg_models.hepi import HEPi
    from geometry_rl.modules.pyg_models.ponita.conv import FiberBundleConv

    codes = [[1, 0], [0, 1], [0, 1]]  # configs/algorithm/pyg_agent/model/hepi.yaml:17-48
    mp = [[FiberBundleConv(64, 64, 64, groups=64, separable=True, widening_factor=4) if c else None for c in code]
          for code in codes]
    return HEPi(input_dim_node=len(NodeType) + cfg.policy_aux_dim, input_dim_edge=len(EdgeType) + 4, hidden_dim=64,
                latent_dim=64, output_dim=cfg.output_dim, output_dim_vec=cfg.output_dim_vec, node_encoder_layers=2,
                edge_encoder_layers=2, node_decoder_layers=2, node_type_mapping=NodeType, edge_type_mapping=EdgeType,
                edge_level_mapping=EdgeLevel, message_passing=mp, num_messages=2, device="cpu", num_ori=16, degree=2,
                ponita_dim=cfg.ponita_dim, only_upper_hemisphere=cfg.only_upper_hemisphere)


def _obs_list(cfg, obs):
    keys = obs_keys(cfg)
    out = []
    for k in keys:
        if cfg.policy_pos_is_norm and k == "position_vectors":
            out.append(obs["norm_position_vectors"])
        elif cfg.policy_pos_is_norm and k == "velocity_vectors":
            out.append(obs["norm_velocity_vectors"])
        else:
            out.append(obs[k])
    return out


def _graph_record(graph):
    return {
        "node_types": list(graph.node_types),
        "edge_types": [tuple(et) for et in graph.edge_types],
        "edge_index": {"___".join(et): ei.clone() for et, ei in graph.edge_index_dict.items()},
        "pos": {nt: graph[nt].pos.clone() for nt in graph.node_types},
        "norm_pos": {nt: graph[nt].norm_pos.clone() for nt in graph.node_types},
        "properties": {nt: graph[nt].properties.clone() for nt in graph.node_types},
        "output_mask": (graph.output_mask.start, graph.output_mask.stop),
    }


def golden_policy_body(cfg_name, B, seed, tag):
    """HEPi / EMPN body through the reference's own data builder: build_data -> one_step -> loss.backward."""
    cfg = CONFIGS[cfg_name]
    torch.manual_seed(seed)
    gen = torch.Generator().manual_seed(seed)
    obs = synthetic_obs(cfg, B, gen, env_ids=torch.arange(B) * max(1, cfg.num_envs // B))
    data, NodeType, EdgeType, EdgeLevel = _ref_data(cfg, full_graph_obs=False, dist_as_pos=True,
                                                    output_mask_key="grippers", concat_input_vector=False)
    if cfg.model == "hepi":
        net = _ref_hepi(cfg, NodeType, EdgeType, EdgeLevel)
    else:
        from geometry_rl.modules.pyg_models.ponita_gcn import PonitaGCN
        net = PonitaGCN(input_dim_node=len(NodeType) + cfg.policy_aux_dim, output_dim=cfg.output_dim,
                        output_dim_vec=cfg.output_dim_vec, num_layers=2, hidden_dim=64, dropout=0.0, num_ori=16,
                        degree=2, widening_factor=4, attention=False, ponita_dim=cfg.ponita_dim)
    # perturb zero-initialised parameters so that every gradient path is exercised
    with torch.no_grad():
        for n, p in net.named_parameters():
            if n.endswith("bias") and p.abs().max() == 0:
                p.normal_(0, 0.05)
    net.train()
    graph, u = data.build_data(*_obs_list(cfg, obs), train=True)
    net.one_step(graph, u)  # first training-mode forward: callibrate() fires (conv.py:104-105)
    sd = _sd(net)
    graph, u = data.build_data(*_obs_list(cfg, obs), train=True)
    out, hidden = net.one_step(graph, u)
    w_out = torch.randn(out.shape, generator=gen)
    w_hid = torch.randn(hidden.shape, generator=gen)
    loss = (out * w_out).sum() + (hidden * w_hid).sum()
    net.zero_grad()
    loss.backward()
    rec = {
        "config": cfg_name, "B": B, "obs": obs, "state_dict": sd, "graph": _graph_record(graph),
        "scalar_dict": {k: v.clone() for k, v in u[0].items()}, "vector_dict": {k: v.clone() for k, v in u[1].items()},
        "out": out.detach(), "hidden": hidden.detach(), "w_out": w_out, "w_hid": w_hid, "loss": loss.detach(),
        "grads": _grads(net),
    }
    if cfg.model == "empn":
        rec["homo_edge_index"] = net.homogeneous_edge_index(graph).clone()
    _save(rec, tag)
    print(tag, "out", tuple(out.shape), "hidden", tuple(hidden.shape), "loss", float(loss))


def golden_critic(cfg_name, B, seed, tag):
    from geometry_rl.modules.pyg_models.deepsets import DeepSets
    cfg = CONFIGS[cfg_name]
    torch.manual_seed(seed)
    gen = torch.Generator().manual_seed(seed)
    obs = synthetic_obs(cfg, B, gen, env_ids=torch.arange(B) * max(1, cfg.num_envs // B))
    data, NodeType, _, _ = _ref_data(cfg, full_graph_obs=True, dist_as_pos=False, output_mask_key=None,
                                     concat_input_vector=True)
    aux = 12 if cfg.task == "rigid" else 9
    net = DeepSets(input_dim_node=len(NodeType) + aux, output_dim=64, hidden_dim=64, norm=["layer_norm", "layer_norm"])
    with torch.no_grad():
        for n, p in net.named_parameters():
            if "norms" in n:
                p.add_(torch.randn(p.shape, generator=gen) * 0.1)
    cfg_obs = [obs[k] for k in obs_keys(cfg)]
    graph, u = data.build_data(*cfg_obs, train=True)
    out = net.one_step(graph, u)
    w = torch.randn(out.shape, generator=gen)
    (out * w).sum().backward()
    x_tokens = torch.cat([u[nt].reshape(B, -1, u[nt].shape[-1]) for nt in graph.node_types], dim=1)
    _save({"config": cfg_name, "B": B, "obs": obs, "state_dict": _sd(net), "tokens": x_tokens.clone(),
                "out": out.detach(), "w": w, "grads": _grads(net)}, tag)
    print(tag, "out", tuple(out.shape), "tokens", tuple(x_tokens.shape))


def golden_transformer(B, seed, tag):
    from geometry_rl.modules.pyg_models.transformer_vanilla import TransformerVanilla
    cfg = CONFIGS["rigid_insertion_two_agents_multi_transformer_trpl_cfg"]
    torch.manual_seed(seed)
    gen = torch.Generator().manual_seed(seed)
    obs = synthetic_obs(cfg, B, gen, env_ids=torch.arange(B) * max(1, cfg.num_envs // B))
    data, NodeType, _, _ = _ref_data(cfg, full_graph_obs=False, dist_as_pos=True, output_mask_key="grippers",
                                     concat_input_vector=True)
    net = TransformerVanilla(input_dim_node=len(NodeType) + 12, output_dim=64, num_layers=2, num_heads=2, hidden_dim=64,
                             dropout=0.0, concat_global=False)
    net.train()
    graph, u = data.build_data(*_obs_list(cfg, obs), train=True)
    out = net.one_step(graph, u)
    w = torch.randn(out.shape, generator=gen)
    (out * w).sum().backward()
    x_tokens = torch.cat([u[nt].reshape(B, -1, u[nt].shape[-1]) for nt in graph.node_types], dim=1)
    _save({"B": B, "obs": obs, "state_dict": _sd(net), "tokens": x_tokens.clone(), "out": out.detach(), "w": w,
                "output_mask": (graph.output_mask.start, graph.output_mask.stop), "grads": _grads(net)},
          tag)
    print(tag, "out", tuple(out.shape))


class _PolicyStub:
    """The projection layers only call these five helpers of the policy
    (models/policy/gnn_gaussian_policy_diag.py:100-148); bind the reference's own unbound functions."""
    contextual_std = True
    is_diag = True

    def __init__(self):
        from geometry_rl.algorithms.trust_region_projections.models.policy.gnn_gaussian_policy_diag import (
            GNNGaussianPolicyDiag as P)
        for name in ("maha", "log_determinant", "covariance", "precision", "entropy", "log_probability"):
            setattr(self, name, getattr(P, name).__get__(self))


def golden_projection(seed, tag):
    from geometry_rl.algorithms.trust_region_projections.projections.projection_factory import get_projection_layer
    gen = torch.Generator().manual_seed(seed)
    rec = {}
    for k, eps_cov, coeff in ((6, 0.0025, 1.0), (3, 0.0025, 1.0), (12, 0.001, 4.0)):
        B = 64
        q_mean = torch.randn(B, k, generator=gen) * 0.3
        q_v = (0.6 + 0.8 * torch.rand(B, k, generator=gen)) ** 2  # old *variances*
        # a mix of samples inside / outside either bound
        scale = torch.rand(B, 1, generator=gen)
        mean = (q_mean + torch.randn(B, k, generator=gen) * 0.6 * scale).requires_grad_(True)
        v = (q_v * torch.exp(torch.randn(B, k, generator=gen) * 0.15 * scale)).requires_grad_(True)
        action = q_mean + torch.randn(B, k, generator=gen) * q_v.sqrt()
        adv = torch.randn(B, generator=gen)
        for proj_type in ("kl", "w2"):
            proj = get_projection_layer(proj_type=proj_type, mean_bound=0.05, cov_bound=eps_cov,
                                        trust_region_coeff=coeff, scale_prec=True, entropy_schedule=False,
                                        target_entropy=0.0, temperature=0.5, entropy_eq=False, entropy_first=False,
                                        action_dim=k, total_train_steps=1000, cpu=True, dtype=torch.float32)
            policy = _PolicyStub()
            p = (mean, torch.diag_embed(v))
            q = (q_mean, torch.diag_embed(q_v))
            pm, pstd = proj(policy, p, q, 0)
            dist = torch.distributions.MultivariateNormal(pm, pstd)  # objectives/trpl.py:245 (std as covariance)
            logp = dist.log_prob(action)
            ent = dist.entropy()
            tr_loss = proj.get_trust_region_loss(policy, p, (pm, pstd))
            metrics = proj.compute_metrics(policy, p, q, 0)
            total = -(torch.exp(logp - logp.detach()) * adv).mean() - 0.005 * ent.mean() + tr_loss
            g_mean, g_v = torch.autograd.grad(total, (mean, v))
            rec[f"{proj_type}_k{k}"] = {
                "k": k, "eps_mean": 0.05, "eps_cov": eps_cov, "coeff": coeff, "mean": mean.detach().clone(),
                "v": v.detach().clone(), "q_mean": q_mean, "q_v": q_v, "action": action, "adv": adv,
                "proj_mean": pm.detach(), "proj_v": torch.diagonal(pstd, dim1=-2, dim2=-1).detach().clone(),
                "logp": logp.detach(), "entropy": ent.detach(), "tr_loss": tr_loss.detach(),
                "metrics": {m: metrics[m].detach().clone() for m in
                            ("kl", "constraint", "mean_constraint", "mean_constraint_max", "cov_constraint",
                             "cov_constraint_max", "entropy", "entropy_diff")},
                "g_mean": g_mean, "g_v": g_v,
            }
    _save(rec, tag)
    print(tag, list(rec.keys()))


def golden_head(seed, tag):
    """GNNGaussianPolicyDiag.forward (gnn_gaussian_policy_diag.py:26-87) around a stub gnn."""
    from geometry_rl.algorithms.trust_region_projections.models.policy.gnn_gaussian_policy_diag import (
        GNNGaussianPolicyDiag)
    gen = torch.Generator().manual_seed(seed)
    rec = {}
    for post_fc, A, adim in ((False, 1, 6), (False, 4, 3), (True, 2, 3)):
        B = 16
        mean_in = torch.randn(B * A * (adim // 3), 3, generator=gen)
        hidden = torch.randn(B * A, 64, generator=gen)

        class _G(torch.nn.Module):
            device = "cpu"

            def one_step(self, data, iv):
                return hidden if post_fc else (mean_in, hidden)

        class _D:
            def build_data(self, *a, **k):
                return None, None

        torch.manual_seed(seed)
        pol = GNNGaussianPolicyDiag(gnn=_G(), hyper_data=_D(), action_dim=adim * A, num_actuators=A, init="orthogonal",
                                    hidden_sizes=(64, 64), contextual_std=True, init_std=1.0, minimal_std=1e-5,
                                    share_action_dim=True, post_fc=post_fc)
        with torch.no_grad():
            pol._pre_std.weight.normal_(0, 0.3, generator=gen)
            pol._mean.weight.normal_(0, 0.3, generator=gen)
        loc, cov = pol(torch.zeros(B, 1))
        rec[f"post_fc{int(post_fc)}_A{A}"] = {"post_fc": post_fc, "A": A, "action_dim": adim, "B": B, "mean_in": mean_in,
                                              "hidden": hidden, "state_dict": _sd(pol), "loc": loc.detach(),
                                              "cov": cov.detach()}
    _save(rec, tag)
    print(tag, list(rec.keys()))

def golden_policy_wrapper(cfg_name, B, seed, tag):
    """The reference's own CALLER of the policy body: the unmodified `GNNGaussianPolicyDiag` wrapping the unmodified HEPi
    and data builder, called as ProbabilisticActor calls it, `module(*obs) -> (loc, covariance_matrix)`
    (abstract_gnn_gaussian_policy.py:95-109, gnn_gaussian_policy_diag.py:26-87, utils_algo_graph.py:146-158)."""
    from geometry_rl.algorithms.trust_region_projections.models.policy.gnn_gaussian_policy_diag import (
        GNNGaussianPolicyDiag)
    cfg = CONFIGS[cfg_name]
    torch.manual_seed(seed)
    gen = torch.Generator().manual_seed(seed)
    obs = synthetic_obs(cfg, B, gen, env_ids=torch.arange(B) * max(1, cfg.num_envs // B))
    data, NodeType, EdgeType, EdgeLevel = _ref_data(cfg, full_graph_obs=False, dist_as_pos=True,
                                                    output_mask_key="grippers", concat_input_vector=False)
    net = _ref_hepi(cfg, NodeType, EdgeType, EdgeLevel)
    pol = GNNGaussianPolicyDiag(gnn=net, hyper_data=data, action_dim=cfg.total_action_dim, num_actuators=cfg.num_actuators,
                                init="orthogonal", hidden_sizes=(64, 64), contextual_std=True, init_std=1.0,
                                minimal_std=1e-5, share_action_dim=True, post_fc=cfg.post_fc)
    with torch.no_grad():
        for n, p in pol.named_parameters():
            if n.endswith("bias") and p.abs().max() == 0:
                p.normal_(0, 0.05, generator=gen)
        pol._pre_std.weight.normal_(0, 0.3, generator=gen)
    args = _obs_list(cfg, obs)
    pol(*args)  # first training-mode forward: callibrate() fires (conv.py:104-105)
    sd = _sd(pol)
    loc, cov = pol(*args)
    w_loc = torch.randn(loc.shape, generator=gen)
    w_cov = torch.randn(loc.shape, generator=gen)
    loss = (loc * w_loc).sum() + (cov.diagonal(dim1=-2, dim2=-1) * w_cov).sum()
    pol.zero_grad()
    loss.backward()
    _save({"config": cfg_name, "B": B, "obs": obs, "state_dict": sd, "loc": loc.detach(), "cov": cov.detach(),
           "w_loc": w_loc, "w_cov": w_cov, "grads": _grads(pol)}, tag)
    print(tag, "loc", tuple(loc.shape), "cov", tuple(cov.shape))


def golden_value_wrapper(cfg_name, B, T, seed, tag):
    """The reference's own CALLER of the critic body: the unmodified `GNNVFNet` around the unmodified DeepSets and data
    builder, with 2-D observations ([B,F] -> [B,1]) and with 3-D ones ([B,T,F] -> [B,T,1], the per-time-step Python loop of
    value/gnn_vf_net.py:65-86 that GAE(shifted=True) drives)."""
    from geometry_rl.algorithms.trust_region_projections.models.value.gnn_vf_net import GNNVFNet
    from geometry_rl.modules.pyg_models.deepsets import DeepSets
    cfg = CONFIGS[cfg_name]
    torch.manual_seed(seed)
    gen = torch.Generator().manual_seed(seed)
    data, NodeType, _, _ = _ref_data(cfg, full_graph_obs=True, dist_as_pos=False, output_mask_key=None,
                                     concat_input_vector=True)
    aux = 12 if cfg.task == "rigid" else 9
    vf = GNNVFNet(gnn=DeepSets(input_dim_node=len(NodeType) + aux, output_dim=64, hidden_dim=64,
                               norm=["layer_norm", "layer_norm"]), hyper_data=data, init="orthogonal", hidden_sizes=(64, 64))
    with torch.no_grad():
        vf.final.weight.normal_(0, 0.2, generator=gen)
    env_ids = torch.arange(B) * max(1, cfg.num_envs // B)
    frames = [synthetic_obs(cfg, B, gen, env_ids=env_ids) for _ in range(T)]
    obs3 = {k: torch.stack([f[k] for f in frames], dim=1) for k in obs_keys(cfg)}
    v2 = vf(*[obs3[k][:, 0] for k in obs_keys(cfg)])
    v3 = vf(*[obs3[k] for k in obs_keys(cfg)])
    w = torch.randn(v3.shape, generator=gen)
    vf.zero_grad()
    (v3 * w).sum().backward()
    _save({"config": cfg_name, "B": B, "T": T, "obs3": obs3, "state_dict": _sd(vf), "v2": v2.detach(), "v3": v3.detach(),
           "w": w, "grads": _grads(vf)}, tag)
    print(tag, "v2", tuple(v2.shape), "v3", tuple(v3.shape))


def golden_equivariance(tag):
    """Ponita.main() (ponita/ponita.py:372-449) with asserts added externally: four graphs that are
    90-degree rotated copies, S1 grid with 4 orientations -> outputs must be rotated copies."""
    from geometry_rl.modules.pyg_models.ponita.ponita import Ponita
    from geometry_rl.modules.pyg_models.ponita.utils.to_from_sphere import (scalar_to_sphere, vec_to_sphere,
                                                                           sphere_to_scalar, sphere_to_vec)
    torch.manual_seed(0)
    num_ori, bs, nn_, dim, hid = 4, 4, 3, 2, 16
    R = torch.tensor([[0.0, 1.0], [-1.0, 0.0]])
    iv = torch.zeros(bs, nn_, dim)
    iv[0] = torch.tensor([[1.0, 0.0], [0.0, 1.0], [-1.0, 0.0]]) * 2.0
    for i in range(1, 4):
        iv[i] = torch.einsum("ij,bj->bi", R, iv[i - 1])
    isc = torch.ones_like(iv[..., 0]) * 10.0
    pos = iv.clone() * 10.0
    ei = torch.tensor([[0, 1, 3, 4, 6, 7, 9, 10], [2, 2, 5, 5, 8, 8, 11, 11]])
    model = Ponita(2, hid, 1, 4, dim=2, num_ori=num_ori, output_dim_vec=1, task_level="graph").eval()
    s = scalar_to_sphere(isc, model.ori_grid).permute(0, 2, 1).unsqueeze(-1)
    v = vec_to_sphere(iv, model.ori_grid).permute(0, 2, 1).unsqueeze(-1)
    x = torch.cat([s, v], dim=-1).reshape(bs * nn_, num_ori, -1)
    out = model.read_out_layers[-1](model(x, pos.reshape(bs * nn_, dim), edge_index=ei))
    out = out.reshape(bs, nn_, num_ori, -1)
    osc, ovec = torch.split(out, [1, 1], dim=-1)
    osc = sphere_to_scalar(osc)
    ovec = sphere_to_vec(ovec.reshape(-1, num_ori, 1), model.ori_grid).reshape(bs, nn_, 1, dim)
    torch.save({"state_dict": _sd(model), "x": x.detach(), "pos": pos.reshape(bs * nn_, dim), "edge_index": ei,
                "out_scalar": osc.detach(), "out_vec": ovec.detach(), "R": R}, os.path.join(OUT, f"{tag}.pt"))
    print(tag, "max scalar spread", float((osc - osc[:1]).abs().max()))


def main():
    os.makedirs(OUT, exist_ok=True)
    golden_policy_body("rigid_insertion_multi_hepi_trpl_cfg", 6, 11, "hepi_rigid_insertion")
    golden_policy_body("cloth_hanging_multi_hepi_trpl_cfg", 4, 12, "hepi_cloth_hanging")
    golden_policy_body("rope_shaping_hepi_trpl_cfg", 2, 13, "hepi_rope_shaping")
    golden_policy_body("rigid_pushing_multi_empn_trpl_cfg", 5, 14, "empn_rigid_pushing")
    golden_critic("rigid_insertion_multi_hepi_trpl_cfg", 6, 15, "deepsets_rigid")
    golden_critic("rope_shaping_hepi_trpl_cfg", 3, 16, "deepsets_rope")
    golden_transformer(4, 17, "transformer_two_agents")
    golden_projection(18, "projection")
    golden_head(19, "gaussian_head")
    golden_equivariance("ponita_equivariance")
    golden_policy_wrapper("rigid_insertion_multi_hepi_trpl_cfg", 5, 21, "policy_wrapper_rigid_insertion")
    golden_policy_wrapper("cloth_hanging_multi_hepi_trpl_cfg", 3, 22, "policy_wrapper_cloth_hanging")
    golden_value_wrapper("rigid_insertion_multi_hepi_trpl_cfg", 4, 3, 23, "value_wrapper_rigid")


if __name__ == "__main__":
    main()
