"""Checker behind __graft_entry__.smoke(): one small HEPi update step + GAE on cuda:0, compared with the CPU oracle.
The oracle is the checker here, never the thing run: every product tensor below comes out of libgrl_b200 / torch CUDA.
Lives at the repo root (not inside geometry_rl_b200/) because it imports oracle/, which the product package must not."""
import torch


def run_smoke(cfg_name: str = "rigid_insertion_multi_hepi_trpl_cfg", B: int = 32, verbose: bool = True):
    from geometry_rl_b200 import learner
    from geometry_rl_b200.synthetic import CONFIGS, synthetic_obs, synthetic_rollout
    from geometry_rl_b200.tensors import to_device
    from oracle.step import OracleAgent, make_minibatch

    dev = torch.device("cuda:0")
    cfg = CONFIGS[cfg_name]
    actor, critic, projection, loss_module, adv_module = learner.build_agent(cfg, dev, seed=0)
    gen = torch.Generator().manual_seed(1234)
    obs = synthetic_obs(cfg, B, gen, env_ids=torch.arange(B) * max(1, cfg.num_envs // B))
    # warm-up forward in training mode = the one-time calibration of train.py:72-74
    with torch.no_grad():
        actor.get_dist(to_device(obs, dev))
    oracle = OracleAgent(cfg, actor.state_dict(), critic.state_dict())
    mb = make_minibatch(cfg, oracle, obs, gen)
    ref, ga, gc = oracle.step_grads(mb)

    lrn = learner.Learner(cfg, actor, critic, loss_module)
    out = lrn.compute_losses(to_device(mb, dev))
    out["actor_loss"].backward()
    out["loss_critic"].backward()
    worst = 0.0
    for k in ("loss_objective", "loss_trust_region", "loss_entropy", "loss_critic", "ESS", "kl"):
        err = abs(float(out[k]) - float(ref[k])) / (abs(float(ref[k])) + 1e-6)
        worst = max(worst, err)
        assert err < 1e-4, f"smoke: {k} {float(out[k])} vs oracle {float(ref[k])}"
    pol = dict(actor.get_submodule("0").module.named_parameters())
    for k, g in ga.items():
        if k not in pol or g is None or float(g.abs().max()) == 0.0:  # buffers (ori_grid) are not parameters
            continue
        err = float((pol[k].grad.cpu() - g).abs().max()) / float(g.abs().max())
        worst = max(worst, err)
        assert err < 1e-4, f"smoke: grad {k} rel err {err}"
    lrn.actor_optim.step()
    lrn.critic_optim.step()

    # the same step on the BENCHED path: 16-bit tensor-core kernels (fused edge kernels + tcgen05 node kernels),
    # gradients within north_star's 1e-2 of the CPU oracle at the same (updated) parameters
    from geometry_rl_b200 import ops
    lrn.actor_optim.zero_grad()
    lrn.critic_optim.zero_grad()
    oracle16 = OracleAgent(cfg, actor.state_dict(), critic.state_dict())
    ref16, ga16, _ = oracle16.step_grads(mb)
    ops.set_precision("bf16")
    try:
        out16 = lrn.compute_losses(to_device(mb, dev))
        out16["actor_loss"].backward()
    finally:
        ops.set_precision("fp32")
    worst16 = 0.0
    for k in ("loss_trust_region", "loss_entropy", "kl"):
        err = abs(float(out16[k]) - float(ref16[k])) / (abs(float(ref16[k])) + 1e-6)
        worst16 = max(worst16, err)
        assert err < 1e-2, f"smoke (16-bit): {k} {float(out16[k])} vs oracle {float(ref16[k])}"
    for k, g in ga16.items():
        if k not in pol or g is None or float(g.abs().max()) == 0.0:
            continue
        err = float((pol[k].grad.cpu() - g).abs().max()) / float(g.abs().max())
        worst16 = max(worst16, err)
        assert err < 1e-2, f"smoke (16-bit): grad {k} rel err {err}"
    assert worst16 > 1e-7, "smoke (16-bit): identical to fp32, the tensor-core kernels did not run"

    # GAE on a tiny rollout: batched-over-time critic + warp-scan kernel vs the per-step loop + reverse loop
    roll = synthetic_rollout(cfg, gen, num_envs=6, rollout_len=9)
    oracle2 = OracleAgent(cfg, actor.state_dict(), critic.state_dict())
    a_ref, vt_ref, _ = oracle2.gae(roll)
    keys = critic.in_keys
    td = {k: roll[k][:, :-1].to(dev) for k in keys}
    td["next"] = {k: roll[k][:, 1:].to(dev) for k in keys}
    td["next"].update({"reward": roll["reward"].unsqueeze(-1).to(dev), "done": roll["done"].unsqueeze(-1).to(dev),
                       "terminated": roll["terminated"].unsqueeze(-1).to(dev)})
    adv_module(td)
    err = float((td["advantage"][..., 0].cpu() - a_ref).abs().max()) / float(a_ref.abs().max())
    assert err < 1e-4, f"smoke: GAE rel err {err}"
    torch.cuda.synchronize()
    if verbose:
        print(f"smoke ok: {cfg_name} B={B} worst rel err fp32 path {max(worst, err):.2e}, 16-bit path {worst16:.2e}")
