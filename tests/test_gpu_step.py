"""GPU parity of the whole update step (`-m gpu`): graph features -> policy -> head -> TRPL projection ->
TRPLLoss -> critic -> both backward passes, and of the advantage phase (batched-over-time critic + GAE
kernel), against the CPU oracle (oracle/step.py) on the same seeded synthetic inputs.
Tolerances: losses and gradients 1e-5 (north_star's fp32 bound), gradients relative to each tensor's max magnitude.
Measured (tools/fp32_parity_report.py -> profiles/r02_fp32_parity.json): the CUDA path is at most 7.0e-6 from the fp32
oracle and 5.2e-6 from an fp64 evaluation of it over all five configs — closer to fp64 than the CPU fp32 oracle itself
(1.1e-5 worst, tests/test_fp32_noise_floor.py).  W2 covariance gradients 1e-4: the reference's own fp32 result is 4e-5 away
from an fp64 evaluation there (tests/test_oracle_golden.py::test_w2_grad_noise_floor)."""
import pytest
import torch

from geometry_rl_b200.synthetic import CONFIGS, synthetic_obs, synthetic_rollout
from tests.helpers import load_golden
from tests import gpu_helpers as G

pytestmark = pytest.mark.gpu

SIZES = {"rigid_insertion_multi_hepi_trpl_cfg": 48, "rigid_pushing_multi_empn_trpl_cfg": 40,
         "cloth_hanging_multi_hepi_trpl_cfg": 24, "rope_shaping_hepi_trpl_cfg": 6,
         "rigid_insertion_two_agents_multi_transformer_trpl_cfg": 16}


def _setup(cfg_name, proj_type="kl", seed=0):
    from geometry_rl_b200 import learner
    from geometry_rl_b200.tensors import to_device
    from oracle.step import OracleAgent, make_minibatch
    cfg = CONFIGS[cfg_name]
    B = SIZES[cfg_name]
    actor, critic, projection, loss_module, adv_module = learner.build_agent(cfg, G.dev(), proj_type=proj_type, seed=seed)
    gen = torch.Generator().manual_seed(99 + seed)
    obs = synthetic_obs(cfg, B, gen, env_ids=torch.arange(B) * max(1, cfg.num_envs // B))
    with torch.no_grad():  # one-time calibration (train.py:72-74) before the weights are mirrored
        actor.get_dist(to_device(obs, G.dev()))
        for n, p in actor.named_parameters():  # zero-initialised biases would hide bias-gradient bugs
            if n.endswith("bias") and float(p.abs().max()) == 0:
                p.normal_(0, 0.05)
    oracle = OracleAgent(cfg, actor.state_dict(), critic.state_dict(), proj_type=proj_type)
    mb = make_minibatch(cfg, oracle, obs, gen)
    return cfg, actor, critic, loss_module, adv_module, oracle, mb, gen


@pytest.mark.parametrize("cfg_name", list(SIZES.keys()))
@pytest.mark.parametrize("proj_type", ["kl", "w2"])
def test_update_step_matches_oracle(cfg_name, proj_type):
    if proj_type == "w2" and cfg_name not in ("rigid_insertion_multi_hepi_trpl_cfg", "cloth_hanging_multi_hepi_trpl_cfg"):
        pytest.skip("W2 covered on two configs")
    from geometry_rl_b200 import learner
    from geometry_rl_b200.tensors import to_device
    cfg, actor, critic, loss_module, _, oracle, mb, _ = _setup(cfg_name, proj_type)
    ref, ga, gc = oracle.step_grads(mb)
    lrn = learner.Learner(cfg, actor, critic, loss_module)
    out = lrn.compute_losses(to_device(mb, G.dev()))
    out["actor_loss"].backward()
    out["loss_critic"].backward()
    bad = []
    for k in ("loss_objective", "loss_trust_region", "loss_entropy", "loss_critic", "ESS", "kl", "constraint",
              "mean_constraint", "mean_constraint_max", "cov_constraint", "cov_constraint_max", "entropy", "entropy_diff"):
        a, b = float(out[k]), float(ref[k])
        if not abs(a - b) <= 1e-5 * abs(b) + 2e-7:
            bad.append(f"{k}: {a} vs {b}")
    gtol = 1e-4 if proj_type == "w2" else 1e-5
    pol = dict(actor.get_submodule("0").module.named_parameters())
    n_checked = 0
    for k, g in ga.items():
        if k not in pol or g is None or float(g.abs().max()) == 0.0:  # buffers (ori_grid) are not parameters
            continue
        if pol[k].grad is None:
            bad.append(f"{k}: missing grad")
        elif G.rel(pol[k].grad, g) >= gtol:
            bad.append(G.err_report(k, pol[k].grad, g))
        n_checked += 1
    vf = dict(critic.module._network1.named_parameters())
    for k, g in gc.items():
        if G.rel(vf[k].grad, g) >= 1e-5:
            bad.append(G.err_report("critic " + k, vf[k].grad, g))
        n_checked += 1
    assert n_checked >= 30
    assert not bad, "\n".join(bad)


@pytest.mark.parametrize("proj_type", ["kl", "w2"])
@pytest.mark.parametrize("B,k", [(1, 3), (37, 6), (4096, 12), (1500, 3)])
def test_fused_trpl_loss_matches_torch_formulation(proj_type, B, k):
    """ops.trpl_loss (grl_trpl_fwd + grl_trpl_loss_fwd / _bwd + grl_trpl_bwd) against the same quantities written
    with torch ops in fp64 on top of ops.trpl_project (objectives/trpl.py:231-321, base_projection_layer.py:292-384):
    13 scalars at 1e-6, gradients w.r.t. (mean, v) of an arbitrary combination of the three losses at 2e-6."""
    import math
    from geometry_rl_b200 import _lib, ops
    gen = torch.Generator().manual_seed(B * 100 + k)
    q_mean = torch.randn(B, k, generator=gen) * 0.3
    q_v = (0.5 + torch.rand(B, k, generator=gen)) ** 2
    s = torch.rand(B, 1, generator=gen)
    mean = (q_mean + torch.randn(B, k, generator=gen) * 0.5 * s).cuda().requires_grad_(True)
    v = (q_v * torch.exp(torch.randn(B, k, generator=gen) * 0.2 * s)).cuda().requires_grad_(True)
    q_mean, q_v = q_mean.cuda(), q_v.cuda()
    action = (q_mean + torch.randn(B, k, generator=gen).cuda() * q_v.sqrt())
    prev_lp = (-0.5 * (((action - q_mean) ** 2 / q_v).sum(-1) + k * math.log(2 * math.pi) + q_v.log().sum(-1)))
    adv = torch.randn(B, 1, generator=gen).cuda() * 3 + 1
    em, ec, c_h, c_tr = 0.05, 0.0025, 0.01, 4.0
    w = torch.tensor([0.7, 1.3, -0.4]).cuda()  # upstream gradients of the three losses

    l_obj, l_tr, l_ent, sc = ops.trpl_loss(mean, v, q_mean, q_v, action, prev_lp, adv, em, ec, proj_type, c_h, c_tr, True)
    (w[0] * l_obj + w[1] * l_tr + w[2] * l_ent).backward()
    g_mean, g_v = mean.grad.clone(), v.grad.clone()
    mean.grad = v.grad = None

    pm, pv = ops.trpl_project(mean, v, q_mean, q_v, em, ec, proj_type)
    D = torch.float64
    m64, v64, pm64, pv64 = mean.to(D), v.to(D), pm.to(D), pv.to(D)
    a_n = adv.to(D).reshape(-1)
    if B > 1:
        a_n = (a_n - a_n.mean()) / a_n.std().clamp_min(1e-6)
    logp = -0.5 * (((action.to(D) - pm64) ** 2 / pv64).sum(-1) + k * math.log(2 * math.pi) + pv64.log().sum(-1))
    lw = logp - prev_lp.to(D)
    ent_dist = 0.5 * (k * (1 + math.log(2 * math.pi)) + pv64.log().sum(-1))
    pmd, pvd = pm64.detach(), pv64.detach()

    def kl(m_, s_, mo_, so_):
        return 0.5 * ((m_ - mo_) / so_).pow(2).sum(-1), 0.5 * ((s_ / so_).square().sum(-1) - k + 2 * so_.log().sum(-1)
                                                                - 2 * s_.log().sum(-1))

    def w2(m_, s_, mo_, so_):
        return ((m_ - mo_) / so_).pow(2).sum(-1), (1.0 + s_ * s_ / (so_ * so_) - 2.0 * s_ / so_).sum(-1)

    trm, trc = (kl if proj_type == "kl" else w2)(m64, v64, pmd, pvd)
    klm, klc = kl(m64, v64, pmd, pvd)
    ref = {"loss_objective": -(lw.exp() * a_n).mean(), "loss_trust_region": (trm + trc).mean() * c_tr,
           "loss_entropy": -c_h * ent_dist.mean(), "dist_entropy": ent_dist.mean(),
           "ESS": (2 * lw.logsumexp(0) - (2 * lw).logsumexp(0)).exp() / B, "kl": (klm + klc).mean(),
           "constraint": (trm + trc).mean(), "mean_constraint": trm.mean(), "mean_constraint_max": trm.max(),
           "cov_constraint": trc.mean(), "cov_constraint_max": trc.max(),
           "entropy": (0.5 * (k * math.log(2 * math.e * math.pi) + 2 * v64.log().sum(-1))).mean(),
           "entropy_diff": (pv64.log().sum(-1) - v64.log().sum(-1)).mean()}
    (w[0] * ref["loss_objective"] + w[1] * ref["loss_trust_region"] + w[2] * ref["loss_entropy"]).backward()
    bad = []
    for key, ix in _lib.LOSS_SCALAR_INDEX.items():
        a, b = float(sc[ix]), float(ref[key])
        if not abs(a - b) <= 1e-6 * abs(b) + 1e-7:
            bad.append(f"{key}: {a} vs {b}")
    for key, t in (("loss_objective", l_obj), ("loss_trust_region", l_tr), ("loss_entropy", l_ent)):
        assert float(t) == float(sc[_lib.LOSS_SCALAR_INDEX[key]])
    if G.rel(g_mean, mean.grad) >= 2e-6:
        bad.append(G.err_report("grad mean", g_mean, mean.grad))
    if G.rel(g_v, v.grad) >= 2e-6:
        bad.append(G.err_report("grad v", g_v, v.grad))
    assert not bad, "\n".join(bad)


def test_fused_and_unfused_loss_modules_agree():
    """TRPLLoss.fused switches between ops.trpl_loss and the torch formulation: same 12 outputs, same actor gradients."""
    from geometry_rl_b200 import learner
    from geometry_rl_b200.tensors import to_device
    cfg, actor, critic, loss_module, _, oracle, mb, _ = _setup("rigid_pushing_multi_empn_trpl_cfg")
    lrn = learner.Learner(cfg, actor, critic, loss_module)
    res = {}
    for fused in (True, False):
        loss_module.fused = fused
        actor.zero_grad(set_to_none=True)
        out = lrn.compute_losses(to_device(mb, G.dev()))
        out["actor_loss"].backward()
        res[fused] = ({k: float(t) for k, t in out.items()},
                      {k: p.grad.detach().clone() for k, p in actor.named_parameters() if p.grad is not None})
    (o1, g1), (o0, g0) = res[True], res[False]
    assert set(o1) == set(o0)
    bad = [f"{k}: {o1[k]} vs {o0[k]}" for k in o0 if not abs(o1[k] - o0[k]) <= 1e-5 * abs(o0[k]) + 2e-7]
    assert set(g1) == set(g0)
    bad += [G.err_report(k, g1[k], g0[k]) for k in g0 if G.rel(g1[k], g0[k]) >= 1e-5]
    assert not bad, "\n".join(bad)


def test_gaussian_head_matches_reference_fixture():
    from geometry_rl_b200.algorithms.trust_region_projections.models.policy.gnn_gaussian_policy_diag import (
        GNNGaussianPolicyDiag)
    rec = load_golden("gaussian_head")
    for key, r in rec.items():
        class _G(torch.nn.Module):
            device = "cuda"

            def one_step(self, data, iv):
                h = r["hidden"].cuda()
                return h if r["post_fc"] else (r["mean_in"].cuda(), h)

        class _D:
            def build_data(self, *a, **k):
                return None, None

        pol = GNNGaussianPolicyDiag(gnn=_G(), hyper_data=_D(), action_dim=r["action_dim"] * r["A"], num_actuators=r["A"],
                                    init="orthogonal", hidden_sizes=(64, 64), contextual_std=True, init_std=1.0,
                                    minimal_std=1e-5, share_action_dim=True, post_fc=r["post_fc"]).cuda()
        missing = pol.load_state_dict({k: v for k, v in r["state_dict"].items()}, strict=True)
        loc, cov = pol(torch.zeros(r["B"], 1, device="cuda"))
        assert cov.shape == r["cov"].shape
        assert G.rel(loc, r["loc"]) < 1e-6, key
        assert G.rel(cov, r["cov"]) < 1e-6, key


@pytest.mark.parametrize("cfg_name", ["rigid_insertion_multi_hepi_trpl_cfg", "rope_shaping_hepi_trpl_cfg",
                                      "cloth_hanging_multi_hepi_trpl_cfg"])
def test_advantage_phase_matches_oracle(cfg_name):
    """adv_module(td[B_env,T]): ONE batched critic call over T+1 steps (per-step LayerNorm statistics kept) +
    grl_gae_scan == the reference's per-step critic loop + reverse GAE loop."""
    cfg, actor, critic, loss_module, adv_module, oracle, _, gen = _setup(cfg_name)
    Benv, T = 5, 11
    roll = synthetic_rollout(cfg, gen, num_envs=Benv, rollout_len=T)
    a_ref, vt_ref, v_ref = oracle.gae(roll)
    keys = critic.in_keys
    td = {k: roll[k][:, :-1].cuda() for k in keys}
    td["next"] = {k: roll[k][:, 1:].cuda() for k in keys}
    td["next"].update({"reward": roll["reward"].unsqueeze(-1).cuda(), "done": roll["done"].unsqueeze(-1).cuda(),
                       "terminated": roll["terminated"].unsqueeze(-1).cuda()})
    adv_module(td)
    assert td["advantage"].shape == (Benv, T, 1) and td["value_target"].shape == (Benv, T, 1)
    assert G.rel(td["state_value"][..., 0], v_ref[:, :-1]) < 1e-5, G.err_report("value", td["state_value"][..., 0], v_ref[:, :-1])
    assert G.rel(td["advantage"][..., 0], a_ref) < 1e-5, G.err_report("adv", td["advantage"][..., 0], a_ref)
    assert G.rel(td["value_target"][..., 0], vt_ref) < 1e-5


def test_learner_update_changes_parameters_and_is_deterministic():
    """Two learners from the same seed fed the same minibatch end with bit-identical parameters
    (deterministic segmented sums and fixed-order partial reductions: no atomics anywhere)."""
    from geometry_rl_b200 import learner
    from geometry_rl_b200.tensors import to_device
    res = []
    for _ in range(2):
        cfg, actor, critic, loss_module, _, _, mb, _ = _setup("rigid_insertion_multi_hepi_trpl_cfg", seed=3)
        lrn = learner.Learner(cfg, actor, critic, loss_module)
        before = torch.cat([p.detach().reshape(-1).clone() for p in actor.parameters()])
        for _ in range(2):
            lrn.update(to_device(mb, G.dev()))
        after = torch.cat([p.detach().reshape(-1) for p in actor.parameters()])
        assert float((after - before).abs().max()) > 0
        res.append(after.clone())
    assert torch.equal(res[0], res[1])


def test_graph_replay_and_prefetched_inputs_match_eager_updates():
    """Learner.capture / update_graphed and the prefetching input pipeline (prefetch / update_prefetched: pinned host
    minibatch -> staging on a copy stream -> the graph's inputs) replay exactly the eager update: bit-identical
    parameters after three updates on two alternating minibatches."""
    from geometry_rl_b200 import learner
    from geometry_rl_b200.tensors import to_device
    finals = {}
    for mode in ("eager", "graph", "prefetch"):
        cfg, actor, critic, loss_module, _, _, mb, _ = _setup("rigid_pushing_multi_empn_trpl_cfg", seed=4)
        mb2 = dict(mb)  # second minibatch: same slots (the cached per-slot topology belongs to the slot's geometry),
        mb2["advantage"] = -mb["advantage"]  # other advantages
        host = [{k: v.pin_memory() for k, v in b.items() if torch.is_tensor(v)} for b in (mb, mb2)]
        dev = [to_device(b, G.dev()) for b in (mb, mb2)]
        lrn = learner.Learner(cfg, actor, critic, loss_module)
        if mode == "eager":
            for i in range(3):  # capture() undoes its warm-up updates: the first replay is the first update
                lrn.update(dev[i % 2])
        else:
            lrn.capture(dev[0], warmup=3)
            if mode == "graph":
                for i in range(3):
                    lrn.update_graphed(dev[i % 2])
            else:
                lrn.prefetch(host[0])
                for i in range(3):
                    lrn.update_prefetched(host[(i + 1) % 2] if i < 2 else None)
        torch.cuda.synchronize()
        finals[mode] = torch.cat([p.detach().reshape(-1).clone() for p in list(actor.parameters()) + list(critic.parameters())])
    assert torch.equal(finals["graph"], finals["prefetch"])
    assert G.rel(finals["graph"], finals["eager"]) < 1e-6, G.err_report("graph vs eager", finals["graph"], finals["eager"])

