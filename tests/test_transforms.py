"""CPU: NDVecNorm / ReshapeTransform (geometry_rl_b200/torchrl/envs/transforms.py) against a scalar-loop restatement of
torchrl 0.3.1 VecNorm._update driven the way the reference's NDVecNorm._call drives it (transforms.py:135-158).
PARITY UNPINNED: torchrl is not installable offline, so both sides are restatements of the same published arithmetic;
the test pins the vectorised implementation (shared per-component statistics, decay, count = all leading dims,
update-before-standardise, eps clamps) against the naive one."""
import math

import pytest
import torch

from geometry_rl_b200.torchrl.envs import NDVecNorm, ReshapeTransform


def _naive(batches, decay, eps):
    s, q, c = [0.0] * 3, [0.0] * 3, 0.0
    outs = []
    for x in batches:  # [B, n, 3]
        flat = x.reshape(-1, 3).double()
        for k in range(3):
            s[k] = decay * s[k] + float(flat[:, k].sum())
            q[k] = decay * q[k] + float((flat[:, k] ** 2).sum())
        c = decay * c + flat.shape[0]
        o = torch.empty_like(x, dtype=torch.float64)
        for k in range(3):
            mean = s[k] / c
            std = math.sqrt(max(q[k] / c - mean * mean, eps))
            o[..., k] = (x[..., k].double() - mean) / max(std, eps)
        outs.append(o)
    return outs


def test_ndvecnorm_matches_the_naive_running_statistics():
    g = torch.Generator().manual_seed(0)
    batches = [torch.randn(5, 7, 3, generator=g) * torch.tensor([1.0, 10.0, 0.1]) + torch.tensor([0.0, 3.0, -2.0])
               for _ in range(4)]
    reshape = ReshapeTransform(in_keys=["position_vectors"], out_shape=[-1, 3])
    norm = NDVecNorm(in_keys=["position_vectors"], out_keys=["norm_position_vectors"], shapes=[3], decay=0.99999, eps=1e-2)
    ref = _naive(batches, 0.99999, 1e-2)
    for x, r in zip(batches, ref):
        td = {"position_vectors": x.reshape(5, 21)}  # the env emits flat groups
        norm(reshape(td))
        assert td["position_vectors"].shape == (5, 7, 3)
        # fp32 running sums: `ssq / count - mean^2` cancels (component 2: mean -2, std 0.1), as in torchrl's own fp32 buffers
        assert torch.allclose(td["norm_position_vectors"].double(), r, rtol=1e-3, atol=1e-3)
    sd = norm.state_dict()
    assert sd["position_vectors_sum"].shape == (3,) and float(sd["position_vectors_count"]) == pytest.approx(
        sum(35 * 0.99999 ** (3 - i) for i in range(4)), rel=1e-6)
    other = NDVecNorm(in_keys=["position_vectors"], out_keys=["norm_position_vectors"], shapes=[3], decay=0.99999, eps=1e-2)
    other.load_state_dict(sd)
    other.freeze()
    td1, td2 = {"position_vectors": batches[0].clone()}, {"position_vectors": batches[0].clone()}
    other(td1)
    other(td2)
    assert torch.equal(td1["norm_position_vectors"], td2["norm_position_vectors"])  # frozen: statistics do not move


def test_shapes_are_checked():
    with pytest.raises(ValueError):
        NDVecNorm(in_keys=["a", "b"], shapes=[3])
    with pytest.raises(ValueError):
        NDVecNorm(in_keys=["a"], shapes=[3])({"a": torch.zeros(4, 5)})
    with pytest.raises(ValueError):
        ReshapeTransform(in_keys=["a"])
