"""torch.profiler view of one eager update step: which ATen ops (not libgrl kernels) cost device time.
usage: python tools/torch_glue_profile.py [config]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

from geometry_rl_b200 import learner, ops  # noqa: E402
from geometry_rl_b200.tensors import to_device  # noqa: E402
from geometry_rl_b200.synthetic import CONFIGS, synthetic_minibatch, synthetic_obs  # noqa: E402

cfg = CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "rigid_pushing_multi_empn_trpl_cfg"]
dev = torch.device("cuda")
torch.backends.cuda.matmul.allow_tf32 = True
ops.set_precision("bf16")
actor, critic, projection, loss_module, adv = learner.build_agent(cfg, dev, seed=0)
lrn = learner.Learner(cfg, actor, critic, loss_module, overlap_critic=False)
B = cfg.mini_batch_size
gen = torch.Generator().manual_seed(1)
obs = synthetic_obs(cfg, B, gen, env_ids=torch.arange(B) % cfg.num_envs)
with torch.no_grad():
    d = actor.get_dist(to_device(obs, dev))
    v = critic.module(*[obs[k].to(dev) for k in critic.in_keys])
mb = to_device(synthetic_minibatch(obs, d.mean, d.var_diag, v, gen), dev)
for _ in range(3):
    lrn.update(mb)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True) as prof:
    lrn.update(mb)
    torch.cuda.synchronize()
print(prof.key_averages(group_by_input_shape=True).table(sort_by="self_cuda_time_total", row_limit=45, max_name_column_width=60,
                                                          max_shapes_column_width=70))
