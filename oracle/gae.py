"""Oracle (test infrastructure): generalized advantage estimation.

Follows the call site examples/torchrl/train.py:134-140,249-252:
`GAE(gamma, lmbda, value_network=critic, average_gae=False, shifted=True)` from torchrl 0.3.1
(third-party, not under /root/reference, pinned by requirements.txt:3 -> **parity unpinned**; its
published recurrence is restated):

    delta_t = r_t + gamma * (1 - terminated_t) * V_{t+1} - V_t
    A_t     = delta_t + gamma * lmbda * (1 - done_t) * A_{t+1},   A_T = 0
    value_target_t = A_t + V_t

with V taken from ONE critic call over the T+1 observations of each env (shifted=True), i.e.
value[:, :-1] / value[:, 1:].  torchrl evaluates the same recurrence with a vectorised
geometric-kernel convolution; the summation order differs, so comparisons are at 1e-5, while the
OUTPUT ORDER [B_env, T] is bit-for-bit the same layout.
"""
import torch


def gae_reverse_loop(reward, value, done, terminated, gamma: float, lmbda: float):
    """reward/done/terminated: [B, T]; value: [B, T+1].  Plain reverse-time loop, fp32, separate
    multiply and add roundings (no fused multiply-add)."""
    B, T = reward.shape
    not_done = (~done).to(reward.dtype)
    not_term = (~terminated).to(reward.dtype)
    adv = torch.zeros_like(reward)
    carry = torch.zeros(B, dtype=reward.dtype)
    for t in range(T - 1, -1, -1):
        delta = reward[:, t] + gamma * not_term[:, t] * value[:, t + 1] - value[:, t]
        carry = delta + (gamma * lmbda) * not_done[:, t] * carry
        adv[:, t] = carry
    return adv, adv + value[:, :-1]


def gae_vectorised(reward, value, done, terminated, gamma: float, lmbda: float):
    """Same recurrence in the closed form torchrl's vectorised path evaluates (discount-matrix
    product), fp64 — an independent check of the loop."""
    B, T = reward.shape
    r, v = reward.double(), value.double()
    nd, nt = (~done).double(), (~terminated).double()
    delta = r + gamma * nt * v[:, 1:] - v[:, :-1]
    c = gamma * lmbda * nd  # coefficient linking t -> t+1
    logc = torch.zeros(B, T + 1, dtype=torch.double)
    adv = torch.zeros_like(delta)
    for t in range(T):
        w = torch.ones(B, dtype=torch.double)
        acc = torch.zeros(B, dtype=torch.double)
        for s in range(t, T):
            acc = acc + w * delta[:, s]
            w = w * c[:, s]
        adv[:, t] = acc
    return adv, adv + v[:, :-1]
