"""Advantage module: replaces `torchrl.objectives.value.GAE(gamma, lmbda, value_network, average_gae=False,
shifted=True)` as constructed and called at examples/torchrl/train.py:134-140,249-252.

`adv_module(td[B_env, T])` runs ONE critic call over the T+1 observations of every env (shifted=True: the
rollout's obs plus the last `next` obs) and the reverse-time warp-scan kernel grl_gae_scan, then writes
`advantage`, `value_target`, `state_value` ([B_env, T, 1]) back."""
import torch
from torch import nn

from .... import ops
from .operators import td_get


class GAE(nn.Module):
    def __init__(self, *, gamma: float, lmbda: float, value_network: nn.Module, average_gae: bool = False,
                 shifted: bool = True, **kwargs):
        super().__init__()
        if not shifted:
            raise NotImplementedError("the reference train loop uses shifted=True (train.py:139)")
        self.gamma, self.lmbda = float(gamma), float(lmbda)
        self.average_gae = average_gae
        self.value_network = value_network

    @torch.no_grad()
    def forward(self, td):
        keys = self.value_network.in_keys
        nxt = td_get(td, "next")
        obs = [torch.cat([td_get(td, k), td_get(nxt, k)[:, -1:]], dim=1) for k in keys]  # [B, T+1, F]
        value = self.value_network.module(*obs)  # [B, T+1, 1]
        B, T1 = value.shape[:2]
        reward = td_get(nxt, "reward").reshape(B, T1 - 1)
        done = td_get(nxt, "done").reshape(B, T1 - 1)
        terminated = td_get(nxt, "terminated", done).reshape(B, T1 - 1)
        adv, vt = ops.gae(reward, value.reshape(B, T1), done, terminated, self.gamma, self.lmbda)
        if self.average_gae:
            adv = (adv - adv.mean()) / adv.std().clamp_min(1e-4)
        td["advantage"] = adv.unsqueeze(-1)
        td["value_target"] = vt.unsqueeze(-1)
        td["state_value"] = value[:, :-1]
        nxt["state_value"] = value[:, 1:]
        return td
