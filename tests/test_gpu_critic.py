"""GPU: the fused critic kernels (grl_critic_inner_*: first Linear -> whole-tensor LayerNorm -> ReLU -> token sum, forward
and backward, no per-token activation in HBM) against the torch formulation of deepsets.py:34-53 / PyG
MLP + LayerNorm(mode='graph') in fp64: outputs 1e-5, every gradient 1e-5 (fp32 recomputation, fp64 statistics).
The reference fixtures (tests/golden/deepsets_*.pt, value_wrapper_rigid.pt) run through the same kernels in
tests/test_gpu_parity.py / test_gpu_boundary.py."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    r = float((a - b).abs().max()) / (float(b.abs().max()) + 1e-30)
    return r if r == r and bool(torch.isfinite(a).all()) else float("inf")


@pytest.mark.parametrize("B,N,Fd", [(1, 1, 15), (3, 97, 15), (64, 162, 12), (7, 239, 13), (5, 300, 16), (700, 97, 15)])
def test_critic_inner_matches_torch_fp64(B, N, Fd):
    from geometry_rl_b200 import ops
    g = torch.Generator().manual_seed(B * 1000 + N)
    r = lambda *s, scale=1.0: (torch.randn(*s, generator=g) * scale).cuda()
    x = r(B, N, Fd)
    p = dict(w1=r(64, Fd, scale=0.3), b1=r(64, scale=0.2), gamma=1 + r(64, scale=0.2), beta=r(64, scale=0.3))
    w = r(B, 64)
    eps = 1e-5
    leaves = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    ysum = ops.critic_inner(x, leaves["w1"], leaves["b1"], leaves["gamma"], leaves["beta"], eps)
    (ysum * w).sum().backward()

    D = torch.float64
    ref = {k: v.to(D).requires_grad_(True) for k, v in p.items()}
    h = F.linear(x.to(D), ref["w1"], ref["b1"])
    hc = h - h.mean()
    y = torch.relu(hc / (hc.pow(2).mean().sqrt() + eps) * ref["gamma"] + ref["beta"])
    ysum_ref = y.sum(1)
    (ysum_ref * w.to(D)).sum().backward()
    assert _rel(ysum, ysum_ref) < 1e-5, f"ysum rel {_rel(ysum, ysum_ref):.3e}"
    bad = [f"{k}: rel {_rel(leaves[k].grad, ref[k].grad):.3e}" for k in p if _rel(leaves[k].grad, ref[k].grad) >= 1e-5]
    assert not bad, "\n".join(bad)


@pytest.mark.parametrize("name", ["deepsets_rigid", "deepsets_rope"])
def test_fused_and_torch_critic_bodies_agree_on_the_reference_fixture(name):
    """DeepSets with the fused per-token half vs its torch formulation (`fused_inner = False`) on the fixture inputs of the
    unmodified reference: both within 1e-5 (outputs and gradients) of the reference's own results."""
    from geometry_rl_b200.modules.pyg_models.deepsets import DeepSets
    from tests.helpers import load_golden
    rec = load_golden(name)
    tokens = rec["tokens"].cuda()
    for fused in (True, False):
        net = DeepSets(input_dim_node=tokens.shape[-1], output_dim=64, hidden_dim=64, norm=["layer_norm", "layer_norm"]).cuda()
        net.load_state_dict(rec["state_dict"], strict=True)
        net.fused_inner = fused

        class _G:  # one node type holding all tokens
            node_types = ["all"]

            def __len__(self):
                return tokens.shape[0]

        out = net.one_step(_G(), {"all": tokens.reshape(-1, tokens.shape[-1])})
        (out * rec["w"].cuda()).sum().backward()
        assert _rel(out, rec["out"]) < 1e-5, f"fused={fused} out rel {_rel(out, rec['out']):.3e}"
        params = dict(net.named_parameters())
        bad = [f"fused={fused} {k}: rel {_rel(params[k].grad, gr):.3e}" for k, gr in rec["grads"].items()
               if gr is not None and _rel(params[k].grad, gr) >= 1e-5]
        assert not bad, "\n".join(bad)
