"""CPU, world_size 2, gloo: the data-parallel host logic (geometry_rl_b200/parallel.py).  A minibatch sharded
over two ranks must give the single-process statistics, critic loss and gradients.  The CUDA
kernels are not involved: the DeepSets critic and the reduction helpers are torch code."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _critic(seed=0):
    from geometry_rl_b200.modules.pyg_models.deepsets import DeepSets
    torch.manual_seed(seed)
    net = DeepSets(input_dim_node=15, output_dim=64, hidden_dim=64, norm=["layer_norm", "layer_norm"])
    final = torch.nn.Linear(64, 1)
    return net, final


class _G:
    node_types = ["a"]

    def __init__(self, B):
        self.B = B

    def __len__(self):
        return self.B


def _critic_loss(net, final, x, target, old_v, mean_fn, groups=1):
    v = final(net.one_step(_G(x.shape[0]), {"a": x.reshape(-1, x.shape[-1])}, norm_groups=groups))
    l = (target - v).pow(2)
    v_clip = old_v + (v - old_v).clamp(-0.2, 0.2)
    return mean_fn(0.5 * torch.max(l, (target - v_clip).pow(2)))


# ---- torch mirror of the staged reduction kernels of csrc/grl_trpl.cu (trpl_loss_stats / _sums / _finalize) ---------------
# terms[b] = (log_w, tr_mean, tr_cov, kl, H_dist, H_p, H_proj, advantage), fp64
def _loss_terms(B, g):
    t = torch.randn(B, 8, generator=g, dtype=torch.float64)
    t[:, 1:4] = t[:, 1:4].abs() * 0.01
    t[:, 7] = t[:, 7] * 3 + 1
    return t


def _stage1(t, stats):
    stats[0], stats[1], stats[2], stats[3] = t[:, 7].sum(), (t[:, 7] ** 2).sum(), float(t.shape[0]), t[:, 0].max()


def _stage2(t, stats, sums):
    n = stats[2]
    loc = stats[0] / n
    var = ((stats[1] - n * loc * loc) / (n - 1)).clamp_min(0)
    inv = 1.0 / var.sqrt().clamp_min(1e-6)
    e = (t[:, 0] - stats[3]).exp()
    sums[0] = (t[:, 0].exp() * (t[:, 7] - loc) * inv).sum()
    sums[1], sums[2] = (t[:, 1] + t[:, 2]).sum(), t[:, 4].sum()
    sums[3], sums[4], sums[5], sums[6], sums[7] = e.sum(), (e * e).sum(), t[:, 1].sum(), t[:, 2].sum(), t[:, 3].sum()
    sums[8], sums[9] = t[:, 5].sum(), (t[:, 6] - t[:, 5]).sum()
    sums[10], sums[11] = t[:, 1].max(), t[:, 2].max()
    stats[4], stats[5] = loc, inv


def _stage3(stats, G, entropy_coef, tr_coeff):
    n = stats[2]
    return torch.stack([-G[0] / n, G[1] / n * tr_coeff, -entropy_coef * G[2] / n, G[2] / n, G[3] * G[3] / G[4] / n, G[7] / n,
                        (G[5] + G[6]) / n, G[5] / n, G[10], G[6] / n, G[11], G[8] / n, G[9] / n])


def _worker(rank, world, port, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from geometry_rl_b200.modules.pyg_models.pyg_compat import GraphLayerNorm
        from geometry_rl_b200.parallel import DataParallel
        dp = DataParallel(side_group=True)  # second communicator for the critic branch (two-stream update)
        assert dp.side is not None and dp.side.world_size == world and dp.side.side is None
        g = torch.Generator().manual_seed(5)
        B, N = 8, 7
        x = torch.randn(B, N, 15, generator=g)
        target, old_v = torch.randn(B, 1, generator=g), torch.randn(B, 1, generator=g)
        adv = torch.randn(B, 1, generator=g) * 3 + 1
        lw = torch.randn(B, generator=g)
        sl = slice(rank * B // world, (rank + 1) * B // world)

        # statistics
        m, s = dp.mean_std_unbiased(adv[sl])
        lse = dp.logsumexp(lw[sl])
        # critic: sharded loss with global LayerNorm statistics + summed gradients
        net, final = _critic()
        class _FakeLoss:  # DataParallel.attach wires the critic's graph-LayerNorm statistics to the SIDE communicator
            critic_network = net
        dp.attach(_FakeLoss)
        assert _FakeLoss.dp is dp
        hooks = [mod.stats_reduce for mod in net.modules() if isinstance(mod, GraphLayerNorm)]
        assert hooks and all(h.__self__ is dp.side for h in hooks)
        loss = _critic_loss(net, final, x[sl], target[sl], old_v[sl], lambda t: t.sum() / (t.numel() * world))
        loss.backward()
        params = list(net.parameters()) + list(final.parameters())
        n = dp.allreduce_grads(params)
        loss_g = loss.detach().clone()
        dist.all_reduce(loss_g)
        # per-sample metrics
        agg = dp.aggregate_metrics({"kl": lw[sl].abs()})
        # staged reduction of the fused TRPL loss (grl_trpl_loss_fwd stages 1-3 with ops.TrplLossFn's two all-gathers)
        terms = _loss_terms(B, torch.Generator().manual_seed(9))
        stats, sums = torch.zeros(8, dtype=torch.float64), torch.zeros(16, dtype=torch.float64)
        def make_global(buf, lo, hi, n_sum):  # ONE all-gather per stage + grl_dp_combine's rank-order sum / max (torch mirror)
            g_ = dp.all_gather_small(buf[lo:hi])
            assert g_.shape == (world, hi - lo)
            acc = g_[0].clone()
            for r in range(1, world):
                acc[:n_sum] = acc[:n_sum] + g_[r, :n_sum]
                acc[n_sum:] = torch.maximum(acc[n_sum:], g_[r, n_sum:])
            buf[lo:hi] = acc
        _stage1(terms[sl], stats)
        make_global(stats, 0, 4, 3)
        _stage2(terms[sl], stats, sums)
        make_global(sums, 3, 12, 7)
        scal = _stage3(stats, sums, 0.01, 4.0)
        share = scal[:3].clone()
        dist.all_reduce(share)  # the three losses are per-rank shares: their sum is the global loss
        if rank == 0:
            torch.save({"mean": m, "std": s, "lse": lse, "loss": loss_g, "grads": [p.grad.clone() for p in params],
                        "n": n, "kl": agg["kl"], "kl_max": agg["kl_max"], "scalars": scal, "loss_sum": share}, out_path)
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_sharded_step_equals_single_process(tmp_path):
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    out_path = str(tmp_path / "rank0.pt")
    procs = [ctx.Process(target=_worker, args=(r, world, port, out_path)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=240)
        assert p.exitcode == 0
    res = torch.load(out_path, weights_only=True)

    g = torch.Generator().manual_seed(5)
    B, N = 8, 7
    x = torch.randn(B, N, 15, generator=g)
    target, old_v = torch.randn(B, 1, generator=g), torch.randn(B, 1, generator=g)
    adv = torch.randn(B, 1, generator=g) * 3 + 1
    lw = torch.randn(B, generator=g)
    assert torch.allclose(res["mean"], adv.mean(), atol=1e-6)
    assert torch.allclose(res["std"], adv.std(), atol=1e-5)
    assert torch.allclose(res["lse"], lw.logsumexp(0), atol=1e-6)
    assert torch.allclose(res["kl"], lw.abs().mean(), atol=1e-6) and torch.allclose(res["kl_max"], lw.abs().max())
    # staged fused-loss reduction: metrics are global, the three losses are shares that sum to the single-process values
    terms = _loss_terms(B, torch.Generator().manual_seed(9))
    stats, sums = torch.zeros(8, dtype=torch.float64), torch.zeros(16, dtype=torch.float64)
    _stage1(terms, stats)
    _stage2(terms, stats, sums)
    ref = _stage3(stats, sums, 0.01, 4.0)
    assert torch.allclose(res["scalars"][4:], ref[4:], rtol=1e-12, atol=1e-14)
    assert torch.allclose(res["loss_sum"], ref[:3], rtol=1e-12, atol=1e-14)
    assert not torch.allclose(res["scalars"][:3], ref[:3], rtol=1e-3)  # a rank's own share is not the global loss
    a_n = (terms[:, 7] - terms[:, 7].mean()) / terms[:, 7].std()  # and the mirror itself matches the plain formulas
    assert torch.allclose(ref[0], -(terms[:, 0].exp() * a_n).mean(), rtol=1e-10)
    assert torch.allclose(ref[4], (2 * terms[:, 0].logsumexp(0) - (2 * terms[:, 0]).logsumexp(0)).exp() / B, rtol=1e-10)
    net, final = _critic()
    loss = _critic_loss(net, final, x, target, old_v, lambda t: t.mean())
    loss.backward()
    params = list(net.parameters()) + list(final.parameters())
    assert res["n"] == sum(p.numel() for p in params)
    assert torch.allclose(res["loss"], loss.detach(), rtol=1e-5, atol=1e-7)
    for a, p in zip(res["grads"], params):
        assert torch.allclose(a, p.grad, rtol=2e-4, atol=1e-7), float((a - p.grad).abs().max())


def test_grouped_graph_layer_norm_equals_per_step_loop():
    """GNNVFNet batches the T+1 critic calls; GraphLayerNorm(groups=T+1) keeps the loop's per-step statistics."""
    net, final = _critic(1)
    g = torch.Generator().manual_seed(2)
    T, B, N = 4, 3, 5
    x = torch.randn(T, B, N, 15, generator=g) * torch.arange(1, T + 1).view(T, 1, 1, 1)
    loop = torch.stack([net.one_step(_G(B), {"a": x[t].reshape(-1, 15)}) for t in range(T)])
    batched = net.one_step(_G(T * B), {"a": x.reshape(-1, 15)}, norm_groups=T).reshape(T, B, -1)
    assert torch.allclose(loop, batched, rtol=1e-5, atol=1e-6)
    assert not torch.allclose(loop, net.one_step(_G(T * B), {"a": x.reshape(-1, 15)}).reshape(T, B, -1), atol=1e-3)
