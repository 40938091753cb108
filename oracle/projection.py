"""Oracle (test infrastructure): TRPL trust-region projection of a diagonal Gaussian policy.

Restates, on CPU torch tensors, the numerics of
  * projections/base_projection_layer.py:71-100   (mean_projection)
  * projections/base_projection_layer.py:199-273  (__call__ / _projection, entropy bound = -inf)
  * projections/base_projection_layer.py:292-327  (get_trust_region_loss)
  * projections/base_projection_layer.py:332-384  (compute_metrics)
  * projections/kl_projection_layer.py:15-111,162-204 (KL projection, diag branch + ITPAL op)
  * projections/w2_projection_layer.py:15-76      (commuting W2 projection)
  * utils/projection_utils.py:34-67,107-149       (gaussian_kl, gaussian_wasserstein_commutative)
  * models/policy/gnn_gaussian_policy_diag.py:100-137 (maha / log_determinant / covariance / entropy)
(all under geometry_rl/algorithms/trust_region_projections/).

Everything works on the DIAGONAL `v` of the matrix the reference calls "std".  Because the loss
hands `(mean, covariance_matrix)` to the projection (objectives/trpl.py:241) while every routine
treats the second element as a std/Cholesky factor, `v` is numerically the policy *variance*
("std := cov" quirk).  Parity means reproducing that, not fixing it.

ITPAL (`cpp_projection.BatchedDiagCovOnlyProjection`) is not available offline and is unpinned
upstream -> **parity unpinned** for the KL covariance step.  What is restated is its published
algorithm (Otto et al., "Differentiable Trust Region Layers", ICLR 2021, App. B): the projected
precision is the eta-weighted interpolation of old and target precision with the scalar dual eta
at the root of KL(eta) = eps; backward is the implicit-function gradient.
"""
import math
from typing import Tuple

import torch

LOG_2PI = math.log(2.0 * math.pi)


# ----------------------------------------------------------------------------------------------
# Diagonal-Gaussian helpers under the "std := v" convention (gnn_gaussian_policy_diag.py:100-137)
# ----------------------------------------------------------------------------------------------
def maha(mean, mean_other, v):
    return ((mean - mean_other) / v).pow(2).sum(-1)


def log_determinant(v):
    return 2 * v.log().sum(-1)


def entropy(v):
    k = v.shape[-1]
    return 0.5 * (k * math.log(2 * math.e * math.pi) + log_determinant(v))


def gaussian_kl(mean, v, mean_other, v_other) -> Tuple[torch.Tensor, torch.Tensor]:
    """utils/projection_utils.py:34-67 for diagonal factors."""
    k = mean.shape[-1]
    maha_part = 0.5 * maha(mean, mean_other, v_other)
    trace_part = (v / v_other).square().sum(-1)
    cov_part = 0.5 * (trace_part - k + log_determinant(v_other) - log_determinant(v))
    return maha_part, cov_part


def gaussian_w2_commutative(mean, v, mean_other, v_other, scale_prec=True):
    """utils/projection_utils.py:107-149 for diagonal factors (cov = v**2)."""
    if scale_prec:
        mean_part = maha(mean, mean_other, v_other)
        inv = 1.0 / v_other
        cov_part = (1.0 + inv * v * v * inv - 2.0 * inv * v).sum(-1)
    else:
        mean_part = ((mean_other - mean) ** 2).sum(-1)
        cov_part = (v_other * v_other + v * v - 2.0 * v_other * v).sum(-1)
    return mean_part, cov_part


# ----------------------------------------------------------------------------------------------
# Mean projection (base_projection_layer.py:71-100).  Autograd flows through maha -> omega.
# ----------------------------------------------------------------------------------------------
def mean_projection(mean, old_mean, maha_part, eps):
    mask = maha_part > eps
    safe = torch.where(mask, maha_part, torch.ones_like(maha_part))
    omega = torch.where(mask, torch.sqrt(safe / eps) - 1.0, torch.ones_like(maha_part))
    omega = torch.max(-omega, omega)[..., None]
    m = (mean + omega * old_mean) / (1 + omega + 1e-16)
    return torch.where(mask[..., None], m, mean)


# ----------------------------------------------------------------------------------------------
# KL projection of a diagonal covariance: the ITPAL stand-in
# ----------------------------------------------------------------------------------------------
def _kl_cov(c_t, c_old):
    return 0.5 * (c_t / c_old - 1.0 + c_old.log() - c_t.log()).sum(-1)


def _c_tilde(eta, c, c_old):
    eta = eta[..., None]
    return (eta + 1.0) / (eta / c_old + 1.0 / c)


def kl_diag_cov_solve(c: torch.Tensor, c_old: torch.Tensor, eps: float):
    """Solve per sample for eta >= 0 such that KL(N(.,c~(eta)) || N(.,c_old)) = eps (eta = 0 when the
    target already satisfies the bound).  fp64, bracket by doubling then bisection.
    Returns (c_tilde [B,k] fp64, eta [B] fp64)."""
    c64, o64 = c.detach().double(), c_old.detach().double()
    kl0 = _kl_cov(c64, o64)
    active = kl0 > eps
    lo = torch.zeros_like(kl0)
    hi = torch.ones_like(kl0)
    for _ in range(200):  # doubling
        too_small = active & (_kl_cov(_c_tilde(hi, c64, o64), o64) > eps)
        if not bool(too_small.any()):
            break
        lo = torch.where(too_small, hi, lo)
        hi = torch.where(too_small, hi * 2.0, hi)
    for _ in range(200):  # bisection
        mid = 0.5 * (lo + hi)
        above = _kl_cov(_c_tilde(mid, c64, o64), o64) > eps
        lo = torch.where(above, mid, lo)
        hi = torch.where(above, hi, mid)
    eta = torch.where(active, 0.5 * (lo + hi), torch.zeros_like(kl0))
    return _c_tilde(eta, c64, o64), eta


class KLDiagCovProjection(torch.autograd.Function):
    """Stand-in for kl_projection_layer.py:162-204 (KLProjectionGradFunctionDiagCovOnly):
    forward(cov_diag, old_cov_diag, eps) -> projected cov diag; backward -> (d cov, None, None)."""

    @staticmethod
    def forward(ctx, c, c_old, eps):
        eps = float(eps)
        c_t, eta = kl_diag_cov_solve(c, c_old, eps)
        ctx.save_for_backward(c.detach().double(), c_old.detach().double(), eta)
        return c_t.to(c.dtype)

    @staticmethod
    def backward(ctx, g):
        c, o, eta = ctx.saved_tensors
        g64 = g.double()
        e = eta[..., None]
        D = e / o + 1.0 / c
        c_t = (e + 1.0) / D
        dct_dc = (e + 1.0) / (D * D * c * c)  # d c~_i / d c_i at fixed eta
        dct_deta = (1.0 / c - 1.0 / o) / (D * D)  # d c~_i / d eta
        dkl_dct = 0.5 * (1.0 / o - 1.0 / c_t)  # d KL / d c~_i
        num = (g64 * dct_deta).sum(-1, keepdim=True)
        den = (dkl_dct * dct_deta).sum(-1, keepdim=True)
        active = (eta > 0)[..., None]
        den = torch.where(active, den, torch.ones_like(den))
        corr = num * (dkl_dct * dct_dc) / den
        grad = torch.where(active, g64 * dct_dc - corr, g64)
        return grad.to(g.dtype), None, None


def kl_projection(mean, v, old_mean, old_v, eps_mean, eps_cov):
    """kl_projection_layer.py:15-111 (contextual std, diagonal).  Returns (proj_mean, proj_v)."""
    mean_part, _ = gaussian_kl(mean, v, old_mean, old_v)
    proj_mean = mean_projection(mean, old_mean, mean_part, eps_mean)
    cov = v.pow(2)  # policy.covariance(std) = std**2 (gnn_gaussian_policy_diag.py:136-137)
    old_cov = old_v.pow(2)
    proj_cov = KLDiagCovProjection.apply(cov, old_cov, float(eps_cov))
    return proj_mean, proj_cov.sqrt()


def w2_projection(mean, v, old_mean, old_v, eps_mean, eps_cov, scale_prec=True):
    """w2_projection_layer.py:15-68 for diagonal sqrt factors."""
    mean_part, cov_part = gaussian_w2_commutative(mean, v, old_mean, old_v, scale_prec)
    proj_mean = mean_projection(mean, old_mean, mean_part, eps_mean)
    mask = cov_part > eps_cov
    safe = torch.where(mask, cov_part, torch.ones_like(cov_part))
    eta = torch.where(mask, torch.sqrt(safe / eps_cov) - 1.0, torch.ones_like(cov_part))
    eta = torch.max(-eta, eta)[..., None]
    new_v = (v + eta * old_v) / (1.0 + eta + 1e-16)
    return proj_mean, torch.where(mask[..., None], new_v, v)


def trust_region_loss(mean, v, proj_mean, proj_v, coeff, proj_type="kl", scale_prec=True):
    """base_projection_layer.py:292-327: coeff * mean(d_mean(p, sg(p~)) + d_cov(p, sg(p~)))."""
    pm, pv = proj_mean.detach(), proj_v.detach()
    if proj_type == "w2":
        a, b = gaussian_w2_commutative(mean, v, pm, pv, scale_prec)
    else:
        a, b = gaussian_kl(mean, v, pm, pv)
    return (a + b).mean() * coeff


def compute_metrics(mean, v, q_mean, q_v, proj_type="kl", scale_prec=True):
    """base_projection_layer.py:332-384 (aggregate=True; the 8 entries trpl.py:264-271 logs)."""
    with torch.no_grad():
        ent_old, ent = entropy(q_v), entropy(v)
        mean_kl, cov_kl = gaussian_kl(mean, v, q_mean, q_v)
        if proj_type == "w2":
            md, cd = gaussian_w2_commutative(mean, v, q_mean, q_v, scale_prec)
        else:
            md, cd = mean_kl, cov_kl
        return {
            "kl": (mean_kl + cov_kl).mean(),
            "constraint": (md + cd).mean(),
            "mean_constraint": md.mean(),
            "mean_constraint_max": md.max(),
            "cov_constraint": cd.mean(),
            "cov_constraint_max": cd.max(),
            "entropy": ent.mean(),
            "entropy_diff": (ent_old - ent).mean(),
        }


def mvn_diag_log_prob(x, loc, cov_diag):
    """torch.distributions.MultivariateNormal(loc, covariance_matrix=diag(cov_diag)).log_prob(x)
    (objectives/trpl.py:245-246): the projected "std" is passed as the covariance matrix."""
    k = x.shape[-1]
    return -0.5 * (((x - loc) ** 2 / cov_diag).sum(-1) + k * LOG_2PI + cov_diag.log().sum(-1))


def mvn_diag_entropy(cov_diag):
    k = cov_diag.shape[-1]
    return 0.5 * (k * (1.0 + LOG_2PI) + cov_diag.log().sum(-1))
