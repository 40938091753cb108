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


@pytest.mark.parametrize("K", [64, 256])
def test_tcgen05_gemm_with_the_a_operand_in_tensor_memory(K):
    """TS mode: A rows written into TMEM by their owning threads (tcgen05.st, packed fp16 pairs, lane = row), B from shared
    memory.  Pins the operand layout: element (m, k) = half (k & 1) of column k / 2 of lane m."""
    from geometry_rl_b200 import _lib as L
    g = torch.Generator().manual_seed(77 + K)
    A = torch.randn(128, K, generator=g)
    B = torch.randn(64, K, generator=g)
    D = torch.full((128, 64), float("nan"), device="cuda")
    Ad, Bd = A.cuda(), B.cuda()
    L.call("grl_tc_selftest_gemm_ts", L.ptr(Ad), L.ptr(Bd), L.ptr(D), K)
    torch.cuda.synchronize()
    ref = A.half().float() @ B.half().float().t()
    err = float((D.cpu() - ref).abs().max()) / float(ref.abs().max())
    assert err < 1e-5, f"TS-mode GEMM K={K}: rel err {err}"


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


def _idesc(M, N, a_mn=0, b_mn=0, a_bf16=1, b_bf16=1):
    return (1 << 4) | (a_bf16 << 7) | (b_bf16 << 10) | (a_mn << 15) | (b_mn << 16) | ((N >> 3) << 17) | ((M >> 4) << 24)


def _chunked(X, dtype=torch.bfloat16):
    """[rows][cols] -> the library's operand image [cols/8][rows][8] (bf16 or fp16)."""
    r, c = X.shape
    return X.to(dtype).reshape(r, c // 8, 8).permute(1, 0, 2).contiguous().cuda()


def _debug_mma(a_img, b_img, N, n_ksteps, a_desc, b_desc, idesc):
    from geometry_rl_b200 import _lib as L
    D = torch.full((128, N), float("nan"), device="cuda")
    L.call("grl_tc_debug_mma", a_img.data_ptr(), a_img.numel() * 2, b_img.data_ptr(), b_img.numel() * 2, L.ptr(D), N,
           n_ksteps, a_desc[0], a_desc[1], a_desc[2], b_desc[0], b_desc[1], b_desc[2], idesc, 0)
    torch.cuda.synchronize()
    return D.cpu()


def test_tcgen05_transposed_operands_share_the_same_image():
    """The [chunk][row][8] image of a [rows][cols] tile is BOTH a K-major operand (M/N = rows, K = cols:
    LBO = rows*16, SBO = 128, 2*LBO per K step) and an MN-major operand (K = rows, M/N = cols: SBO = rows*16,
    LBO = 128, 256 B per K step).  The backward kernels rely on this to form X^T Y weight gradients and the
    W^T products without any transposed copies."""
    g = torch.Generator().manual_seed(3)
    rows = 128
    X = torch.randn(rows, 128, generator=g)   # A^T: [k rows][m]
    Y = torch.randn(rows, 64, generator=g)    # B^T: [k rows][n]
    Xb, Yb = X.bfloat16().float(), Y.bfloat16().float()
    mn = (128, rows * 16, 256)                # (LBO, SBO, advance per K step) of the MN-major view
    # (1) weight-gradient form: D[m][n] = sum_rows X[row][m] Y[row][n]
    D = _debug_mma(_chunked(X), _chunked(Y), 64, rows // 16, mn, mn, _idesc(128, 64, 1, 1))
    ref = Xb.t() @ Yb
    assert float((D - ref).abs().max() / ref.abs().max()) < 1e-5
    # (2) activation x W form with W stored [k'][c] K-major-for-forward, read MN-major for the backward:
    #     D[m][c] = sum_k' A[m][k'] W[k'][c],  A K-major (rows = m, K = k'), W image = chunked([k' rows][c cols])
    A = torch.randn(128, 128, generator=g)
    W = torch.randn(128, 64, generator=g)
    kmaj = (128 * 16, 128, 2 * 128 * 16)
    D = _debug_mma(_chunked(A), _chunked(W), 64, 128 // 16, kmaj, mn, _idesc(128, 64, 0, 1))
    ref = A.bfloat16().float() @ W.bfloat16().float()
    assert float((D - ref).abs().max() / ref.abs().max()) < 1e-5


@pytest.mark.parametrize("a_bf16,b_bf16", [(0, 0), (1, 1)])
@pytest.mark.parametrize("N,K", [(64, 64), (80, 80), (256, 80)])
def test_tcgen05_kind_f16_operand_formats(a_bf16, b_bf16, N, K):
    """kind::f16 with fp16 operands (the forward node kernel stages LayerNorm / GELU outputs as fp16; gradients stay
    bf16) and N = 80 / K = 80 shapes (bias folded into the contraction as an extra K chunk, bias gradient as an extra
    N chunk).  A and B must share one format: measured on B200, fp16 x bf16 raises an illegal-instruction error."""
    g = torch.Generator().manual_seed(17 + N)
    A = torch.randn(128, K, generator=g)
    B = torch.randn(N, K, generator=g)
    ta = torch.bfloat16 if a_bf16 else torch.float16
    tb = torch.bfloat16 if b_bf16 else torch.float16
    D = _debug_mma(_chunked(A, ta), _chunked(B, tb), N, K // 16, (128 * 16, 128, 2 * 128 * 16), (N * 16, 128, 2 * N * 16),
                   _idesc(128, N, 0, 0, a_bf16, b_bf16))
    ref = A.to(ta).float() @ B.to(tb).float().t()
    assert float((D - ref).abs().max() / ref.abs().max()) < 1e-5


@pytest.mark.parametrize("B,n_per", [(3, 7), (40, 49)])
def test_fiber_conv_bf16_backward_within_1e_2(B, n_per):
    """Every gradient of the fused convolution (inputs, edge basis, all 11 parameter tensors) from the bf16
    tensor-core path vs the strict fp32 path: within 1e-2 of each tensor's max magnitude."""
    from geometry_rl_b200 import ops
    es, p = _conv_inputs(B, n_per, 3, 23)
    g = torch.Generator().manual_seed(5)
    w = torch.randn(p["x"].shape, generator=g).cuda()
    names = ["x", "basis", "fk", "wk", "bias", "ln_g", "ln_b", "w1", "b1", "w2", "b2"]
    grads = {}
    for mode in ("fp32", "bf16"):
        leaves = {k: p[k].clone().requires_grad_(True) for k in names}
        ops.set_precision(mode)
        try:
            out = ops.fiber_conv(leaves["x"], None, leaves["basis"], leaves["fk"], leaves["wk"], leaves["bias"], leaves["ln_g"],
                                 leaves["ln_b"], leaves["w1"], leaves["b1"], leaves["w2"], leaves["b2"], es)
            (out * w).sum().backward()
        finally:
            ops.set_precision("fp32")
        grads[mode] = {k: leaves[k].grad.clone() for k in names}
    torch.cuda.synchronize()
    bad = []
    for k in names:
        a, b = grads["bf16"][k], grads["fp32"][k]
        ref = b - w if k == "x" else b  # the identity part of d out / d x is exact in both paths
        got = a - w if k == "x" else a
        err = float((got - ref).abs().max()) / float(ref.abs().max())
        if not err < 1e-2:
            bad.append(f"{k}: rel err {err:.3e}")
    assert not bad, "\n".join(bad)


@pytest.mark.parametrize("name", ["hepi_rigid_insertion", "hepi_cloth_hanging", "hepi_rope_shaping", "empn_rigid_pushing"])
def test_policy_body_bf16_path_matches_reference_fixture_within_1e_2(name):
    """The whole policy body on the bf16 tensor-core path (bf16 edge basis, tcgen05 contractions) against the
    fixtures recorded from the UNMODIFIED reference: outputs and every parameter gradient within 1e-2 of the
    tensor's max magnitude (north_star: "bf16 MLP path within 1e-2")."""
    from geometry_rl_b200 import ops
    from geometry_rl_b200.synthetic import CONFIGS
    from tests import gpu_helpers as G
    from tests.helpers import load_golden
    rec = load_golden(name)
    cfg = CONFIGS[rec["config"]]
    net = G.make_policy_body(cfg)
    net.load_state_dict(rec["state_dict"], strict=True)
    net.train()
    data = G.make_data(cfg, policy=True)
    graph, u = data.build_data(*G.obs_args(cfg, rec["obs"], policy=True), train=True)
    ops.set_precision("bf16")
    try:
        out, hidden = net.one_step(graph, u)
        loss = (out * rec["w_out"].cuda()).sum() + (hidden * rec["w_hid"].cuda()).sum()
        loss.backward()
    finally:
        ops.set_precision("fp32")
    assert G.rel(out, rec["out"]) < 1e-2, G.err_report("out", out, rec["out"])
    assert G.rel(hidden, rec["hidden"]) < 1e-2, G.err_report("hidden", hidden, rec["hidden"])
    assert G.rel(out, rec["out"]) > 1e-7, "bit-identical to fp32: the bf16 kernels did not run"
    params = dict(net.named_parameters())
    bad = []
    for k, gref in rec["grads"].items():
        if gref is None:
            continue
        if G.rel(params[k].grad, gref) >= 1e-2:
            bad.append(G.err_report(k, params[k].grad, gref))
    assert not bad, "\n".join(bad)


@pytest.mark.parametrize("n", [1, 3, 4, 1023, 4096 * 1024 + 5])
def test_absmax_is_exact(n):
    """grl_absmax returns the bit pattern of max |x| (NaNs ignored): it sets the power-of-two scale of the fp16
    gradient operands, so it must be exact and order-independent."""
    from geometry_rl_b200 import _lib as L
    g = torch.Generator().manual_seed(n)
    x = (torch.randn(n, generator=g) * 10 ** torch.randint(-6, 3, (n,), generator=g).float()).cuda()
    if n > 3:
        x[1] = float("nan")
    out = torch.zeros(1, dtype=torch.int32, device="cuda")
    L.call("grl_absmax", L.ptr(x), n, L.ptr(out))
    torch.cuda.synchronize()
    ref = torch.nan_to_num(x, nan=0.0).abs().max()
    assert out.view(torch.float32).item() == ref.item()


@pytest.mark.parametrize("n_p,n", [(1, 1), (3, 31), (7, 33), (8, 64), (148, 50000), (296, 4096), (301, 5)])
def test_reduce_partials_fixed_order_sum(n_p, n):
    """grl_reduce_partials: sum over the partial slots (any count, outputs not a multiple of 32), accumulate mode, and
    bit-identical results on repetition (fixed summation order)."""
    from geometry_rl_b200 import _lib as L
    g = torch.Generator().manual_seed(n_p * 1000 + n)
    part = torch.randn(n_p, n, generator=g).cuda()
    out = torch.full((n,), float("nan"), device="cuda")
    L.call("grl_reduce_partials", L.ptr(part), n_p, n, L.ptr(out), 0)
    ref = part.double().sum(0)
    assert float((out.double() - ref).abs().max()) <= 1e-6 * max(1.0, float(ref.abs().max())) * max(1, n_p) ** 0.5
    out2 = torch.empty_like(out)
    L.call("grl_reduce_partials", L.ptr(part), n_p, n, L.ptr(out2), 0)
    assert torch.equal(out, out2)
    L.call("grl_reduce_partials", L.ptr(part), n_p, n, L.ptr(out2), 1)
    assert float((out2.double() - 2 * ref).abs().max()) <= 2e-6 * max(1.0, float(ref.abs().max())) * max(1, n_p) ** 0.5


@pytest.mark.parametrize("scale", [1e-7, 1.0, 3e4])
def test_fiber_conv_backward_is_invariant_to_the_gradient_scale(scale):
    """The tensor-core backward stages gradients as fp16 after multiplying them by a power of two taken from
    grl_absmax(grad_out).  Gradients of the loss `scale * L` must therefore equal `scale` times those of L up to fp32
    rounding, whether the incoming gradients are O(1e-7) (mean losses over large minibatches) or O(1e4)."""
    from geometry_rl_b200 import ops
    es, p = _conv_inputs(6, 21, 3, 31)
    g = torch.Generator().manual_seed(9)
    w = torch.randn(p["x"].shape, generator=g).cuda()
    names = ["x", "basis", "fk", "wk", "bias", "ln_g", "ln_b", "w1", "b1", "w2", "b2"]
    grads = {}
    ops.set_precision("bf16")
    try:
        for sc in (1.0, scale):
            leaves = {k: p[k].clone().requires_grad_(True) for k in names}
            out = ops.fiber_conv(leaves["x"], None, leaves["basis"], leaves["fk"], leaves["wk"], leaves["bias"], leaves["ln_g"],
                                 leaves["ln_b"], leaves["w1"], leaves["b1"], leaves["w2"], leaves["b2"], es)
            (out * (w * sc)).sum().backward()
            grads[sc] = {k: leaves[k].grad.clone() for k in names}
    finally:
        ops.set_precision("fp32")
    torch.cuda.synchronize()
    pow2 = 2.0 ** round(torch.log2(torch.tensor(scale)).item())
    for k in ["fk", "bias", "ln_g", "ln_b", "w1", "b1", "w2", "b2"]:  # produced by the node kernels
        a, b = grads[scale][k] / scale, grads[1.0][k]
        # exact when the scale is a power of two (the fp16 payload is then bit-identical), tiny otherwise
        tol = 1e-6 if scale == pow2 else 2e-3
        err = float((a - b).abs().max()) / float(b.abs().max())
        assert err <= tol, f"{k}: gradient not scale-covariant (rel {err:.2e} at scale {scale})"
