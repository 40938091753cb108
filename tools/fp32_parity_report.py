"""Measured fp32-path parity (run on the GPU box): worst per-tensor gradient deviation (max|a-b| / max|b| and relative L2) of the
strict fp32 CUDA path vs the CPU oracle evaluated in fp32 AND in fp64, for the whole update step of every config.
usage: python tools/fp32_parity_report.py > gpurun_out/fp32_parity.json"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from geometry_rl_b200 import learner  # noqa: E402
from geometry_rl_b200.synthetic import CONFIGS, synthetic_obs  # noqa: E402
from geometry_rl_b200.tensors import to_device  # noqa: E402
from oracle.step import OracleAgent, make_minibatch  # noqa: E402

SIZES = {"rigid_insertion_multi_hepi_trpl_cfg": 48, "rigid_pushing_multi_empn_trpl_cfg": 40,
         "cloth_hanging_multi_hepi_trpl_cfg": 24, "rope_shaping_hepi_trpl_cfg": 6,
         "rigid_insertion_two_agents_multi_transformer_trpl_cfg": 16}
dev = torch.device("cuda:0")
report = {}
for name, B in SIZES.items():
    cfg = CONFIGS[name]
    actor, critic, projection, loss_module, adv = learner.build_agent(cfg, dev, seed=0)
    gen = torch.Generator().manual_seed(99)
    obs = synthetic_obs(cfg, B, gen, env_ids=torch.arange(B) * max(1, cfg.num_envs // B))
    with torch.no_grad():
        actor.get_dist(to_device(obs, dev))
        for n, p in actor.named_parameters():
            if n.endswith("bias") and float(p.abs().max()) == 0:
                p.normal_(0, 0.05)
    o32 = OracleAgent(cfg, actor.state_dict(), critic.state_dict())
    o64 = OracleAgent(cfg, actor.state_dict(), critic.state_dict(), dtype=torch.float64)
    mb = make_minibatch(cfg, o32, obs, gen)
    _, ga32, gc32 = o32.step_grads(mb)
    _, ga64, gc64 = o64.step_grads(mb)
    lrn = learner.Learner(cfg, actor, critic, loss_module)
    out = lrn.compute_losses(to_device(mb, dev))
    out["actor_loss"].backward()
    out["loss_critic"].backward()
    pol = dict(actor.get_submodule("0").module.named_parameters())
    vf = dict(critic.module._network1.named_parameters())
    rec = {"gpu_vs_fp32": [0, ""], "gpu_vs_fp64": [0, ""], "fp32_vs_fp64": [0, ""], "gpu_vs_fp32_l2": [0, ""]}

    def upd(key, a, b, k):
        a, b = a.detach().double().cpu(), b.detach().double().cpu()
        r = float((a - b).abs().max()) / (float(b.abs().max()) + 1e-30)
        if r > rec[key][0]:
            rec[key] = [r, k]

    for src32, src64, params, tag in ((ga32, ga64, pol, ""), (gc32, gc64, vf, "critic ")):
        for k, g64 in src64.items():
            if k not in params or g64 is None or float(g64.abs().max()) == 0 or params[k].grad is None:
                continue
            upd("gpu_vs_fp32", params[k].grad, src32[k], tag + k)
            upd("gpu_vs_fp64", params[k].grad, g64, tag + k)
            upd("fp32_vs_fp64", src32[k], g64, tag + k)
            a, b = params[k].grad.detach().double().cpu(), src32[k].double()
            r2 = float((a - b).norm() / (b.norm() + 1e-30))
            if r2 > rec["gpu_vs_fp32_l2"][0]:
                rec["gpu_vs_fp32_l2"] = [r2, tag + k]
    report[name] = rec
print(json.dumps(report, indent=1))
