"""Oracle (test infrastructure): ONE learner iteration of the reference on the CPU, restated.

Follows examples/torchrl/train.py:134-146,249-316 for one collected batch:
  * advantage phase: GAE(shifted=True) = critic over the T+1 observations of every env, evaluated with the
    reference's Python loop over time (value/gnn_vf_net.py:72-78), then the GAE recurrence (oracle/gae.py);
  * one minibatch of `TRPLLoss.forward` (objectives/trpl.py:275-321): advantage standardisation, policy
    forward (graph features -> HEPi / EMPN / transformer -> Gaussian head), KL / W2 projection, importance
    weights, trust-region loss, entropy bonus, clipped critic loss, ESS and the 8 logged metrics;
  * actor_loss.backward(); critic_loss.backward() (train.py:296-305); optional grad-norm clip; Adam.

Parameters are plain state-dict tensors (the product modules' own `state_dict()` keys, prefixes stripped).
The topology cache mirrors the reference's: ONE placeholder, built from the first batch of a size and re-used for
every later batch of that size, rebuilt when the size changes (rigid_tasks_data.py:254-255; SURVEY 3.4 per-slot quirk).

ITPAL is unavailable -> the KL covariance step uses the restated fp64 dual solve (parity unpinned)."""
from typing import Dict, Mapping, Optional

import torch
import torch.nn.functional as F

from geometry_rl_b200.synthetic import PathConfig, observation_layout, obs_keys
from . import gae as ogae
from . import graph as og
from . import models as om
from . import projection as op


def strip(sd: Mapping[str, torch.Tensor], prefix: str) -> Dict[str, torch.Tensor]:
    return {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}


class OracleAgent:
    def __init__(self, cfg: PathConfig, actor_sd: Mapping[str, torch.Tensor], critic_sd: Mapping[str, torch.Tensor],
                 proj_type: str = "kl", dtype=torch.float32):
        self.cfg, self.proj_type, self.dtype = cfg, proj_type, dtype
        cast = lambda v: v.detach().cpu().to(dtype).clone().requires_grad_(True) if v.is_floating_point() \
            else v.detach().cpu().clone()
        policy = strip(actor_sd, "0.module.") or dict(actor_sd)
        self.actor = {k: cast(v) for k, v in policy.items()}
        vf = strip(critic_sd, "module._network1.") or dict(critic_sd)
        self.critic = {k: cast(v) for k, v in vf.items()}
        self._topo = {True: {}, False: {}}
        self.task = og.TASKS[cfg.task]

    # ---- graph data (pyg_data/*_tasks_data.py) ---------------------------------------------------------
    def _graph(self, obs: Mapping[str, torch.Tensor], *, policy: bool):
        cfg = self.cfg
        dims, names = observation_layout(cfg)
        sel = {k: obs[k].to(self.dtype) for k in obs_keys(cfg)}
        if policy and cfg.policy_pos_is_norm:
            sel["position_vectors"] = sel["norm_position_vectors"]
            sel["velocity_vectors"] = sel["norm_velocity_vectors"]
        parts = og.split_obs(sel, dims, names)
        B = sel["scalars"].shape[0]
        cache = self._topo[policy]
        if B not in cache:
            cache.clear()  # the reference keeps only the latest placeholder (rigid_tasks_data.py:254-255)
            num_points = parts["infos"]["object_num_points"].long().reshape(-1) if cfg.task == "rigid" else None
            cache[B] = og.build_topology(self.task, parts["position_vectors"], full_graph_obs=not policy,
                                         output_mask_key="grippers" if policy else None, num_points=num_points)
        g = og.update_positions(cache[B], parts["position_vectors"], parts["norm_position_vectors"])
        concat = (not policy) or cfg.model == "transformer"
        feats = og.input_vectors(self.task, g, parts["norm_position_vectors"], parts["norm_velocity_vectors"],
                                 dist_as_pos=policy, angular_velocity=cfg.angular_velocity, concat=concat)
        return g, feats

    # ---- networks ---------------------------------------------------------------------------------------
    def policy_forward(self, obs):
        """-> (mean [B,k], variance diagonal [B,k])  (gnn_gaussian_policy_diag.py:26-87)."""
        cfg = self.cfg
        g, feats = self._graph(obs, policy=True)
        body = strip(self.actor, "gnn.")
        B = g.num_graphs
        if cfg.model == "transformer":
            out = om.transformer_forward(body, om.concat_tokens(g, feats), g.output_mask)
        else:
            kw = dict(dim=cfg.ponita_dim, output_dim=cfg.output_dim, output_dim_vec=cfg.output_dim_vec)
            out = (om.hepi_forward if cfg.model == "hepi" else om.empn_forward)(body, g, feats[0], feats[1], **kw)
        return om.gaussian_head(self.actor, out, B, post_fc=cfg.post_fc)

    def critic_forward(self, obs):
        """[B,F] -> [B,1]; [B,T,F] -> [B,T,1] with the reference's per-time-step loop (gnn_vf_net.py:65-86)."""
        first = obs[obs_keys(self.cfg)[0]]
        if first.dim() == 3:
            outs = []
            for t in range(first.shape[1]):
                outs.append(self._critic_one({k: v[:, t] for k, v in obs.items() if torch.is_tensor(v) and v.dim() == 3}))
            c = torch.stack(outs, dim=1)
        else:
            c = self._critic_one(obs)
        return F.linear(c, self.critic["final.weight"], self.critic["final.bias"])

    def _critic_one(self, obs):
        g, feats = self._graph(obs, policy=False)
        return om.deepsets_forward(strip(self.critic, "gnn."), om.concat_tokens(g, feats))

    # ---- advantage phase (train.py:249-252) ----------------------------------------------------------------
    @torch.no_grad()
    def gae(self, rollout: Mapping[str, torch.Tensor]):
        """rollout obs groups [B_env, T+1, F] (shifted layout), reward/done/terminated [B_env, T]."""
        value = self.critic_forward({k: rollout[k] for k in obs_keys(self.cfg)})[..., 0]
        adv, vt = ogae.gae_reverse_loop(rollout["reward"], value, rollout["done"], rollout["terminated"],
                                        self.cfg.gamma, self.cfg.gae_lambda)
        return adv, vt, value

    # ---- TRPLLoss.forward (objectives/trpl.py:275-321) --------------------------------------------------------
    def losses(self, batch: Mapping[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        cfg = self.cfg
        dt = self.dtype
        adv = batch["advantage"].to(dt)
        if adv.numel() > 1:
            adv = (adv - adv.mean()) / adv.std().clamp_min(1e-6)
        mean, var = self.policy_forward(batch)
        q_mean = batch["loc"].to(dt)
        q_var = batch["covariance_matrix"].to(dt)
        q_var = q_var.diagonal(dim1=-2, dim2=-1) if q_var.dim() == 3 else q_var
        if self.proj_type == "kl":
            pm, pv = op.kl_projection(mean, var, q_mean, q_var, cfg.mean_bound, cfg.cov_bound)
        else:
            pm, pv = op.w2_projection(mean, var, q_mean, q_var, cfg.mean_bound, cfg.cov_bound)
        logp = op.mvn_diag_log_prob(batch["action"].to(dt), pm, pv)
        log_weight = (logp - batch["sample_log_prob"].to(dt)).unsqueeze(-1)
        with torch.no_grad():
            lw = log_weight.squeeze(-1)
            ess = (2 * lw.logsumexp(0) - (2 * lw).logsumexp(0)).exp()
        out = {"loss_objective": -(log_weight.exp() * adv).mean()}
        out["loss_trust_region"] = op.trust_region_loss(mean, var, pm, pv, cfg.trust_region_coeff, self.proj_type)
        ent = op.mvn_diag_entropy(pv)
        out["loss_entropy"] = -cfg.entropy_coef * ent.mean()
        # critic (trpl.py:176-229, objectives/utils.py:5-28), l2
        target, old_v = batch["value_target"].to(dt), batch["state_value"].to(dt)
        v = self.critic_forward(batch)
        l = (target - v).pow(2)
        v_clip = old_v + (v - old_v).clamp(-cfg.clip_value, cfg.clip_value)
        out["loss_critic"] = (cfg.critic_coef * torch.max(l, (target - v_clip).pow(2))).mean()
        out["ESS"] = ess.mean() / log_weight.shape[0]
        out.update(op.compute_metrics(mean.detach(), var.detach(), pm.detach(), pv.detach(), self.proj_type))
        out["actor_loss"] = out["loss_objective"] + out["loss_entropy"] + out["loss_trust_region"]
        return out

    def step_grads(self, batch):
        """losses + gradients of actor_loss w.r.t. actor parameters and of loss_critic w.r.t. critic parameters."""
        for p in list(self.actor.values()) + list(self.critic.values()):
            if p.is_floating_point():
                p.grad = None
        out = self.losses(batch)
        out["actor_loss"].backward()
        out["loss_critic"].backward()
        ga = {k: p.grad for k, p in self.actor.items() if p.is_floating_point() and p.requires_grad}
        gc = {k: p.grad for k, p in self.critic.items() if p.is_floating_point() and p.requires_grad}
        return {k: v.detach() for k, v in out.items()}, ga, gc


def make_minibatch(cfg: PathConfig, agent: OracleAgent, obs: Mapping[str, torch.Tensor], generator: torch.Generator,
                   drift: float = 0.35) -> Dict[str, torch.Tensor]:
    """Minibatch whose old distribution is derived from the ORACLE's current policy output (CPU)."""
    from geometry_rl_b200.synthetic import synthetic_minibatch
    with torch.no_grad():
        mean, var = agent.policy_forward(obs)
        v = agent.critic_forward(obs)
    return synthetic_minibatch(obs, mean.float(), var.float(), v.float(), generator, drift)
