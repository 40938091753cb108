"""Distances between diagonal Gaussians under the reference's "std := second tuple element" convention
(geometry_rl/algorithms/trust_region_projections/utils/projection_utils.py:34-67,107-149,252-280).

p and q are `(mean [B,k], S)` where S is either the diagonal `[B,k]` or the full `[B,k,k]` diagonal
matrix the reference passes around; only the diagonal is ever read."""
import math
from typing import Tuple

import torch


def diag_of(s: torch.Tensor) -> torch.Tensor:
    return s.diagonal(dim1=-2, dim2=-1) if s.dim() == 3 else s


def gaussian_kl(policy, p, q) -> Tuple[torch.Tensor, torch.Tensor]:
    """projection_utils.py:34-67: (0.5 * maha, 0.5 * (tr(S_q^-1 S_p)^2 - k + logdet_q - logdet_p))."""
    mean, s = p[0], diag_of(p[1])
    mean_o, s_o = q[0], diag_of(q[1])
    k = mean.shape[-1]
    maha_part = 0.5 * ((mean - mean_o) / s_o).pow(2).sum(-1)
    trace_part = (s / s_o).square().sum(-1)
    cov_part = 0.5 * (trace_part - k + 2 * s_o.log().sum(-1) - 2 * s.log().sum(-1))
    return maha_part, cov_part


def gaussian_wasserstein_commutative(policy, p, q, scale_prec=False) -> Tuple[torch.Tensor, torch.Tensor]:
    """projection_utils.py:107-149 for diagonal factors."""
    mean, s = p[0], diag_of(p[1])
    mean_o, s_o = q[0], diag_of(q[1])
    if scale_prec:
        mean_part = ((mean - mean_o) / s_o).pow(2).sum(-1)
        inv = 1.0 / s_o
        cov_part = (1.0 + inv * (s * s) * inv - 2.0 * inv * s).sum(-1)
    else:
        mean_part = ((mean_o - mean) ** 2).sum(-1)
        cov_part = (s_o * s_o + s * s - 2.0 * s_o * s).sum(-1)
    return mean_part, cov_part


def get_entropy_schedule(schedule_type, total_train_steps, dim):
    """projection_utils.py:252-280: f(initial_entropy, target_entropy, temperature, step)."""
    if schedule_type == "linear":
        return lambda ie, te, temp, step: step * (te - ie) / total_train_steps + ie
    if schedule_type == "exp":
        return lambda ie, te, temp, step: dim * te + (ie - dim * te) * temp ** (10 * step / total_train_steps)
    return lambda ie, te, temp, step: te.new([-math.inf])
