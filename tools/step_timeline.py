"""Device timeline of ONE graph-replayed update step (torch.profiler / CUPTI kernel records): which kernels sit on
the critical path between the library's own kernels.  Writes gpurun_out/timeline_<config>.json (compact list of
[name, stream, start_us, dur_us]) for analysis off the GPU box, and prints the per-stream totals.
usage: python tools/step_timeline.py [config] [precision] [minibatch]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

from geometry_rl_b200 import learner, ops  # noqa: E402
from geometry_rl_b200.synthetic import CONFIGS, synthetic_minibatch, synthetic_obs  # noqa: E402
from geometry_rl_b200.tensors import to_device  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "rigid_pushing_multi_empn_trpl_cfg"
precision = sys.argv[2] if len(sys.argv) > 2 else "bf16"
cfg = CONFIGS[name]
dev = torch.device("cuda")
torch.backends.cuda.matmul.allow_tf32 = False
ops.set_precision(precision)
actor, critic, projection, loss_module, adv = learner.build_agent(cfg, dev, seed=0)
lrn = learner.Learner(cfg, actor, critic, loss_module)
B = int(sys.argv[3]) if len(sys.argv) > 3 else cfg.mini_batch_size
gen = torch.Generator().manual_seed(1)
obs = synthetic_obs(cfg, B, gen, env_ids=torch.arange(B) % cfg.num_envs)
with torch.no_grad():
    d = actor.get_dist(to_device(obs, dev))
    v = critic.module(*[obs[k].to(dev) for k in critic.in_keys])
mb = to_device(synthetic_minibatch(obs, d.mean, d.var_diag, v, gen), dev)
lrn.capture(mb, warmup=3)
for _ in range(3):
    lrn.update_graphed(mb)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    lrn.update_graphed(mb)
    torch.cuda.synchronize()
ev = []
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA:
        ev.append([e.name, int(getattr(e, "device_index", 0)), float(e.time_range.start), float(e.time_range.end - e.time_range.start)])
trace = os.path.join(ROOT, "gpurun_out", f"trace_{name}.json")
os.makedirs(os.path.dirname(trace), exist_ok=True)
prof.export_chrome_trace(trace)
# compact list with stream ids from the chrome trace (prof.events() does not expose the stream)
t = json.load(open(trace))
rows = []
for e in t.get("traceEvents", []):
    if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and "ts" in e:
        rows.append([e["name"], e.get("args", {}).get("stream"), e["ts"], e["dur"]])
rows.sort(key=lambda r: r[2])
t0 = rows[0][2] if rows else 0
for r in rows:
    r[2] = round(r[2] - t0, 3)
out = os.path.join(ROOT, "gpurun_out", f"timeline_{name}_{precision}.json")
json.dump(rows, open(out, "w"))
os.remove(trace)
span = max(r[2] + r[3] for r in rows) if rows else 0
print(f"{len(rows)} device activities, span {span / 1e3:.3f} ms")
by_stream = {}
for r in rows:
    by_stream.setdefault(r[1], 0.0)
    by_stream[r[1]] += r[3]
print({k: round(v / 1e3, 3) for k, v in by_stream.items()})
