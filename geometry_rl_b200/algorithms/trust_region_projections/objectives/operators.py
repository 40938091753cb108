"""TensorDict-free stand-ins for the three torchrl wrappers the train loop builds around the networks
(examples/torchrl/builders/utils_algo_graph.py:146-158,200-203): `ProbabilisticActor(TensorDictModule(policy))`,
`ValueOperator(critic)`.  torchrl / tensordict are not installed offline; these expose exactly the methods
TRPLLoss and GAE call (`get_dist`, `build_dist_from_params`, `get_submodule("0").module`, `__call__(td)`) and
accept any mapping with `.get` (a TensorDict works unchanged)."""
import math
from typing import Mapping, Sequence

import torch
from torch import nn

LOG_2PI = math.log(2.0 * math.pi)


def td_get(td, key, default=None):
    """`td.get(key)` for nested keys given as tuples (TensorDict semantics) on plain nested dicts."""
    if isinstance(key, tuple):
        cur = td
        for k in key:
            if cur is None:
                return default
            cur = cur.get(k, None) if hasattr(cur, "get") else None
        return default if cur is None else cur
    v = td.get(key, None)
    return default if v is None else v


class DiagMultivariateNormal:
    """torch.distributions.MultivariateNormal(loc, covariance_matrix=diag(var)) restricted to what the loss
    reads (objectives/trpl.py:237-246,310): mean, covariance_matrix, log_prob, entropy, sample."""

    def __init__(self, loc: torch.Tensor, covariance_matrix: torch.Tensor = None, var_diag: torch.Tensor = None):
        self.loc = loc
        if var_diag is None:
            var_diag = covariance_matrix.diagonal(dim1=-2, dim2=-1) if covariance_matrix.dim() == loc.dim() + 1 \
                else covariance_matrix
        self.var_diag = var_diag
        self._cov = covariance_matrix if (covariance_matrix is not None and covariance_matrix.dim() == loc.dim() + 1) \
            else None

    @property
    def mean(self):
        return self.loc

    @property
    def covariance_matrix(self):
        if self._cov is None:
            self._cov = torch.diag_embed(self.var_diag)
        return self._cov

    def log_prob(self, x):
        k = x.shape[-1]
        return -0.5 * (((x - self.loc) ** 2 / self.var_diag).sum(-1) + k * LOG_2PI + self.var_diag.log().sum(-1))

    def entropy(self):
        k = self.loc.shape[-1]
        return 0.5 * (k * (1.0 + LOG_2PI) + self.var_diag.log().sum(-1))

    def sample(self, generator=None):
        eps = torch.randn(self.loc.shape, dtype=self.loc.dtype, device=self.loc.device, generator=generator)
        return self.loc + eps * self.var_diag.sqrt()


class _Holder(nn.Module):
    def __init__(self, module):
        super().__init__()
        self.module = module


class PolicyOperator(nn.Module):
    """ProbabilisticActor equivalent: in_keys -> policy -> (loc, covariance_matrix) -> MultivariateNormal."""

    def __init__(self, policy: nn.Module, in_keys: Sequence[str], distribution_class=DiagMultivariateNormal):
        super().__init__()
        self.add_module("0", _Holder(policy))
        self.in_keys = list(in_keys)
        self.out_keys = ["loc", "covariance_matrix", "action", "sample_log_prob"]
        self.distribution_class = distribution_class

    @property
    def policy(self):
        return self.get_submodule("0").module

    def get_dist(self, td: Mapping):
        mean, var = self.policy.forward_diag(*[td_get(td, k) for k in self.in_keys])
        return DiagMultivariateNormal(mean, var_diag=var)

    def build_dist_from_params(self, td: Mapping):
        return DiagMultivariateNormal(td_get(td, "loc"), covariance_matrix=td_get(td, "covariance_matrix"))

    @torch.no_grad()
    def forward(self, td, generator=None):
        dist = self.get_dist(td)
        action = dist.sample(generator)
        td["loc"], td["covariance_matrix"] = dist.mean, dist.covariance_matrix
        td["action"], td["sample_log_prob"] = action, dist.log_prob(action)
        return td


class ValueOperator(nn.Module):
    def __init__(self, module: nn.Module, in_keys: Sequence[str], out_keys=("state_value",)):
        super().__init__()
        self.module = module
        self.in_keys = list(in_keys)
        self.out_keys = list(out_keys)

    def forward(self, td):
        td[self.out_keys[0]] = self.module(*[td_get(td, k) for k in self.in_keys])
        return td
