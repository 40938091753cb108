"""GPU: K5, the fused transformer encoder layer (grl_encoder_layer_fwd / _bwd) behind TransformerVanilla
(geometry_rl/modules/pyg_models/transformer_vanilla.py:30-36,76-92).

1. the layer through the C ABI against torch's own nn.TransformerEncoderLayer evaluated in fp64 on the same parameters
   and inputs: output, input gradient and all twelve parameter gradients within 1e-5 of each tensor's max magnitude
   (north_star's fp32 bound); single token, ragged sizes, the shipped S = 50, the maximum S = 56, more graphs than CTAs;
2. the whole policy body: kernel path == library path (the reference's own call) on outputs and every parameter gradient;
3. more tokens than the kernel holds are refused, not silently routed elsewhere."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(a, b):
    a, b = a.double(), b.double()
    r = float((a - b).abs().max() / (b.abs().max() + 1e-30))
    return r if r == r else float("inf")


@pytest.mark.parametrize("B,S", [(1, 1), (3, 7), (5, 50), (300, 50), (2, 56), (149, 33)])
def test_encoder_layer_matches_torch_fp64(B, S):
    from geometry_rl_b200 import ops
    torch.manual_seed(10 * B + S)
    layer = torch.nn.TransformerEncoderLayer(d_model=64, nhead=2, dim_feedforward=64, dropout=0.0).cuda()
    with torch.no_grad():  # non-trivial affine parameters and biases
        for p in layer.parameters():
            if p.dim() == 1:
                p.add_(0.2 * torch.randn_like(p))
    ref = copy.deepcopy(layer).double()
    x = torch.randn(B, S, 64, device="cuda")
    w = torch.randn(B, S, 64, device="cuda")
    assert ops.encoder_layer_supported(x, layer)

    xk = x.clone().requires_grad_(True)
    out = ops.encoder_layer(xk, layer)
    (out * w).sum().backward()

    xr = x.double().requires_grad_(True)
    out_ref = ref(xr.transpose(0, 1)).transpose(0, 1)
    (out_ref * w.double()).sum().backward()

    bad = []
    if _rel(out, out_ref) >= 1e-5:
        bad.append(f"out {_rel(out, out_ref):.2e}")
    if _rel(xk.grad, xr.grad) >= 1e-5:
        bad.append(f"grad_x {_rel(xk.grad, xr.grad):.2e}")
    for (n, p), (_, q) in zip(layer.named_parameters(), ref.named_parameters()):
        assert p.grad is not None, n
        if _rel(p.grad, q.grad) >= 1e-5:
            bad.append(f"grad {n} {_rel(p.grad, q.grad):.2e}")
    assert not bad, "\n".join(bad)
    assert _rel(out, out_ref) > 0, "bit-identical to fp64: the kernel did not run"


def test_encoder_backward_is_deterministic():
    from geometry_rl_b200 import ops
    torch.manual_seed(0)
    layer = torch.nn.TransformerEncoderLayer(d_model=64, nhead=2, dim_feedforward=64, dropout=0.0).cuda()
    x = torch.randn(400, 50, 64, device="cuda")
    grads = []
    for _ in range(2):
        layer.zero_grad()
        xk = x.clone().requires_grad_(True)
        ops.encoder_layer(xk, layer).square().sum().backward()
        grads.append([xk.grad.clone()] + [p.grad.clone() for p in layer.parameters()])
    assert all(torch.equal(a, b) for a, b in zip(*grads))


def test_transformer_body_kernel_path_equals_library_path(monkeypatch):
    from geometry_rl_b200 import ops
    from geometry_rl_b200.synthetic import CONFIGS, synthetic_obs
    from tests import gpu_helpers as G
    cfg = CONFIGS["rigid_insertion_two_agents_multi_transformer_trpl_cfg"]
    B = 16
    gen = torch.Generator().manual_seed(3)
    obs = synthetic_obs(cfg, B, gen, env_ids=torch.arange(B) * max(1, cfg.num_envs // B))
    net = G.make_policy_body(cfg)
    net.train()
    data = G.make_data(cfg, policy=True)
    graph, u = data.build_data(*G.obs_args(cfg, obs, policy=True), train=True)

    def run():
        net.zero_grad()
        out = net.one_step(graph, u)
        out.square().sum().backward()
        return out.detach().clone(), {n: p.grad.clone() for n, p in net.named_parameters() if p.grad is not None}

    from geometry_rl_b200 import _lib as L
    c0 = L.launch_count
    out_k, g_k = run()
    assert L.launch_count - c0 >= 4, "the encoder kernels did not launch"
    monkeypatch.setattr(ops, "encoder_layer_supported", lambda x, layer: False)
    out_l, g_l = run()
    assert _rel(out_k, out_l) < 1e-5
    assert set(g_k) == set(g_l)
    for n in g_l:
        assert _rel(g_k[n], g_l[n]) < 2e-5, (n, _rel(g_k[n], g_l[n]))


def test_too_many_tokens_are_refused():
    from geometry_rl_b200 import ops
    layer = torch.nn.TransformerEncoderLayer(d_model=64, nhead=2, dim_feedforward=64, dropout=0.0).cuda()
    x = torch.randn(2, ops.ENCODER_MAX_TOKENS + 1, 64, device="cuda")
    assert not ops.encoder_layer_supported(x, layer)
    with pytest.raises(RuntimeError):
        ops.encoder_layer(x, layer)
