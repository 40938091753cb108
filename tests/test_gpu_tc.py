"""GPU: tcgen05 building blocks (geometry_rl_b200/csrc/grl_tc.cuh) — one-CTA bf16 GEMM self-test against a
torch fp32 matmul of the bf16-rounded operands (exact products, fp32 accumulation: 1e-5 relative)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("N,K", [(64, 16), (64, 64), (256, 64), (64, 256)])
def test_tcgen05_gemm_selftest(N, K):
    from geometry_rl_b200 import _lib as L
    g = torch.Generator().manual_seed(N * 1000 + K)
    A = torch.randn(128, K, generator=g)
    B = torch.randn(N, K, generator=g)
    D = torch.full((128, N), float("nan"), device="cuda")
    Ad, Bd = A.cuda(), B.cuda()  # keep the device copies alive across the launch
    L.call("grl_tc_selftest_gemm", L.ptr(Ad), L.ptr(Bd), L.ptr(D), N, K)
    torch.cuda.synchronize()
    ref = A.bfloat16().float() @ B.bfloat16().float().t()
    err = float((D.cpu() - ref).abs().max()) / float(ref.abs().max())
    assert err < 1e-5, f"tcgen05 GEMM N={N} K={K}: rel err {err}"


def _conv_inputs(B, n_per, deg, seed):
    """Random homogeneous graph batch + layer parameters for one FiberConvFn call."""
    from geometry_rl_b200 import ops
    g = torch.Generator().manual_seed(seed)
    N = B * n_per
    src = torch.randint(0, n_per, (B, n_per * deg), generator=g)
    dst = torch.arange(n_per).repeat_interleave(deg)[None].expand(B, -1)
    base = (torch.arange(B) * n_per)[:, None]
    coo = torch.stack([(src + base).reshape(-1), (dst + base).reshape(-1)]).cuda()
    edge_ptr = (torch.arange(B + 1) * n_per * deg).cuda()
    es = ops.build_edge_set(coo, edge_ptr, B, n_per, n_per)
    r = lambda *s, scale=1.0: (torch.randn(*s, generator=g) * scale).cuda()
    p = dict(x=r(N, 16, 64), basis=r(es.n_edges, 16, 64, scale=0.5), fk=r(16, 16, 64, scale=0.3), wk=r(64, 64, scale=0.12),
             bias=r(64, scale=0.1), ln_g=1 + r(64, scale=0.1), ln_b=r(64, scale=0.1), w1=r(256, 64, scale=0.12),
             b1=r(256, scale=0.1), w2=r(64, 256, scale=0.06), b2=r(64, scale=0.1))
    return es, p


@pytest.mark.parametrize("B,n_per", [(3, 7), (40, 49)])
def test_fiber_conv_bf16_tensor_core_path_within_1e_2(B, n_per):
    """bf16 MLP path (tcgen05) vs the strict fp32 FFMA path of the same operator: outputs within 1e-2
    (north_star's bound for the bf16 path), relative to the tensor's max magnitude."""
    from geometry_rl_b200 import ops
    es, p = _conv_inputs(B, n_per, 3, 11)
    outs = {}
    for mode in ("fp32", "bf16"):
        ops.set_precision(mode)
        try:
            outs[mode] = ops.fiber_conv(p["x"], None, p["basis"], p["fk"], p["wk"], p["bias"], p["ln_g"], p["ln_b"], p["w1"],
                                        p["b1"], p["w2"], p["b2"], es)
        finally:
            ops.set_precision("fp32")
    torch.cuda.synchronize()
    delta = outs["bf16"] - p["x"]  # the residual x_dst is exact in both paths: compare the update itself
    ref = outs["fp32"] - p["x"]
    err = float((delta - ref).abs().max()) / float(ref.abs().max())
    assert err < 1e-2, f"bf16 node path rel err {err}"
    assert err > 0, "bf16 path returned the fp32 result bit-for-bit: tensor-core kernel not exercised"
