"""GPU parity of the BENCHED path: the whole update step on the 16-bit tensor-core path (`ops.set_precision("bf16")`, fused
edge kernels, tcgen05 node kernels) against the CPU oracle (oracle/step.py, fp32 restatement of the reference step).

  * all four message-passing configs, eager: the 13 loss / metric scalars and every actor gradient (after the projection
    and the loss backward) within north_star's 16-bit bound, 1e-2 of each tensor's max magnitude; the critic branch is
    not a 16-bit path (fp32 library GEMMs, TF32 off in tests and in bench.py) and stays at the fp32 bound, 1e-5;
  * the headline workload at B = 1024 through `Learner.capture` / `update_graphed` (CUDA-graph replay, critic branch on
    its second stream): gradients read back from the learner's flat bucket;
  * `Learner.fit` over a `DeviceRolloutBuffer` == the same permutation fed by hand (bit-identical parameters).

Scalars: the trust-region metrics and entropies are smooth functions of (mean, cov), so they inherit the outputs' 1e-2
relative bound; `loss_objective = -mean(w * A_hat)` is a difference of O(1) terms (A_hat is standardised), so its
absolute error is bounded by 1e-2 of mean|w * A_hat| ~ 1, i.e. 1e-2 absolute, not by 1e-2 of its own (small) value."""
import pytest
import torch

from geometry_rl_b200.synthetic import CONFIGS, synthetic_obs
from tests import gpu_helpers as G

pytestmark = pytest.mark.gpu

SCALARS = ("loss_objective", "loss_trust_region", "loss_entropy", "loss_critic", "ESS", "kl", "constraint",
           "mean_constraint", "mean_constraint_max", "cov_constraint", "cov_constraint_max", "entropy", "entropy_diff")
MP_CONFIGS = {"rigid_insertion_multi_hepi_trpl_cfg": 48, "rigid_pushing_multi_empn_trpl_cfg": 40,
              "cloth_hanging_multi_hepi_trpl_cfg": 24, "rope_shaping_hepi_trpl_cfg": 6}


@pytest.fixture(autouse=True)
def _sixteen_bit_mode():
    from geometry_rl_b200 import ops
    ops.set_precision("bf16")
    yield
    ops.set_precision("fp32")


def _setup(cfg_name, B, seed=0):
    from geometry_rl_b200 import learner
    from geometry_rl_b200.tensors import to_device
    cfg = CONFIGS[cfg_name]
    actor, critic, projection, loss_module, adv_module = learner.build_agent(cfg, G.dev(), seed=seed)
    gen = torch.Generator().manual_seed(199 + seed)
    obs = synthetic_obs(cfg, B, gen, env_ids=torch.arange(B) * max(1, cfg.num_envs // B) % cfg.num_envs)
    with torch.no_grad():  # one-time calibration (train.py:72-74) before the weights are mirrored
        actor.get_dist(to_device(obs, G.dev()))
        for n, p in actor.named_parameters():
            if n.endswith("bias") and float(p.abs().max()) == 0:
                p.normal_(0, 0.05)
    return cfg, actor, critic, loss_module, obs, gen


def _compare(out, ref, grads_actor, ga, grads_critic, gc, report):
    bad = []
    for k in SCALARS:
        a, b = float(out[k]), float(ref[k])
        tol = 1e-2 if k == "loss_objective" else 1e-2 * abs(b) + 1e-5
        report.append(f"{k}: {a:.6g} vs {b:.6g}")
        if not abs(a - b) <= tol:
            bad.append(f"{k}: {a} vs {b}")
    n = 0
    for k, g in ga.items():
        if k not in grads_actor or g is None or float(g.abs().max()) == 0.0:
            continue
        if grads_actor[k] is None:
            bad.append(f"{k}: missing grad")
        elif G.rel(grads_actor[k], g) >= 1e-2:
            bad.append(G.err_report(k, grads_actor[k], g))
        n += 1
    for k, g in gc.items():
        if G.rel(grads_critic[k], g) >= 1e-5:
            bad.append(G.err_report("critic " + k, grads_critic[k], g))
        n += 1
    assert n >= 30
    return bad


@pytest.mark.parametrize("cfg_name", list(MP_CONFIGS))
def test_16bit_update_step_matches_oracle(cfg_name):
    from geometry_rl_b200 import learner, _lib
    from geometry_rl_b200.tensors import to_device
    from oracle.step import OracleAgent, make_minibatch
    cfg, actor, critic, loss_module, obs, gen = _setup(cfg_name, MP_CONFIGS[cfg_name])
    oracle = OracleAgent(cfg, actor.state_dict(), critic.state_dict())
    mb = make_minibatch(cfg, oracle, obs, gen)
    ref, ga, gc = oracle.step_grads(mb)
    lrn = learner.Learner(cfg, actor, critic, loss_module)
    c0 = _lib.launch_count
    out = lrn.compute_losses(to_device(mb, G.dev()))
    out["actor_loss"].backward()
    out["loss_critic"].backward()
    assert _lib.launch_count > c0
    pol = {k: p.grad for k, p in actor.get_submodule("0").module.named_parameters()}
    vf = {k: p.grad for k, p in critic.module._network1.named_parameters()}
    report = []
    bad = _compare(out, ref, pol, ga, vf, gc, report)
    assert not bad, "\n".join(bad + ["--"] + report)
    # the 16-bit kernels really ran: results differ from the fp32 oracle beyond fp32 noise
    assert abs(float(out["kl"]) - float(ref["kl"])) > 1e-7 * abs(float(ref["kl"]))


def test_graphed_two_stream_16bit_update_at_B1024_matches_oracle():
    """The path bench.py times: CUDA-graph replay of the whole update (fused Adam included), critic branch on its own
    stream, 16-bit kernels, at a minibatch of 1024 HEPi graphs (CPU oracle: a few seconds)."""
    from geometry_rl_b200 import learner
    from geometry_rl_b200.tensors import to_device
    from oracle.step import OracleAgent, make_minibatch
    cfg_name, B = "rigid_insertion_multi_hepi_trpl_cfg", 1024
    cfg, actor, critic, loss_module, obs, gen = _setup(cfg_name, B, seed=2)
    oracle0 = OracleAgent(cfg, actor.state_dict(), critic.state_dict())
    mb = make_minibatch(cfg, oracle0, obs, gen)
    lrn = learner.Learner(cfg, actor, critic, loss_module)
    assert lrn._critic_stream is not None
    dev_mb = to_device(mb, G.dev())
    before = {k: v.detach().clone() for k, v in actor.state_dict().items()}
    lrn.capture(dev_mb, warmup=3)  # restores parameters and optimiser state after its warm-up updates
    for k, v in actor.state_dict().items():
        assert torch.equal(v, before[k]), f"capture() changed {k}"
    oracle = OracleAgent(cfg, actor.state_dict(), critic.state_dict())
    ref, ga, gc = oracle.step_grads(mb)
    out = lrn.update_graphed(dev_mb)
    torch.cuda.synchronize()
    flat = lrn.flat_grads()
    pol = {k[len("actor.0.module."):]: v for k, v in flat.items() if k.startswith("actor.0.module.")}
    vf = {k[len("critic.module._network1."):]: v for k, v in flat.items() if k.startswith("critic.module._network1.")}
    report = []
    bad = _compare(out, ref, pol, ga, vf, gc, report)
    assert not bad, "\n".join(bad + ["--"] + report)
    moved = max(float((v - before[k]).abs().max()) for k, v in actor.state_dict().items() if v.is_floating_point())
    assert moved > 0, "the replayed update did not change the parameters"


def test_learner_fit_equals_feeding_the_same_permutation_by_hand():
    """Learner.fit(buffer, epochs) (train.py:255-316 over the device-resident rollout buffer: gather each minibatch into
    the captured update's static inputs, replay) vs the same index chunks fed through update_graphed by hand."""
    from geometry_rl_b200 import learner
    from geometry_rl_b200.rollout import DeviceRolloutBuffer
    from geometry_rl_b200.tensors import to_device
    from oracle.step import OracleAgent, make_minibatch
    cfg_name, n_frames, mb_size = "rigid_pushing_multi_empn_trpl_cfg", 96, 32
    finals = []
    for mode in ("fit", "manual"):
        cfg, actor, critic, loss_module, obs, gen = _setup(cfg_name, n_frames, seed=5)
        oracle = OracleAgent(cfg, actor.state_dict(), critic.state_dict())
        frames = to_device(make_minibatch(cfg, oracle, obs, gen), G.dev())
        frames = {k: v for k, v in frames.items() if torch.is_tensor(v)}
        buf = DeviceRolloutBuffer(n_frames, mb_size, G.dev(), drop_last=True, generator=torch.Generator().manual_seed(7))
        buf.extend(frames)
        lrn = learner.Learner(cfg, actor, critic, loss_module)
        if mode == "fit":
            res = lrn.fit(buf, epochs=2)
            assert len(res) == 2 * (n_frames // mb_size)
        else:
            chunks = [idx for _ in range(2) for idx in buf.sample_indices()]
            lrn.capture({k: v[:mb_size] for k, v in frames.items()})
            for idx in chunks:
                lrn.update_graphed({k: v.index_select(0, idx) for k, v in frames.items()})
        torch.cuda.synchronize()
        assert lrn.num_network_updates == 2 * (n_frames // mb_size)
        finals.append(torch.cat([p.detach().reshape(-1).clone() for p in list(actor.parameters()) + list(critic.parameters())]))
    assert torch.equal(finals[0], finals[1])
