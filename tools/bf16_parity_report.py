"""Print the relative error (max|diff| / max|ref|) of the bf16 tensor-core path against the reference fixtures for
every output and parameter gradient of the four message-passing configs (the numbers behind
tests/test_gpu_tc.py::test_policy_body_bf16_path_matches_reference_fixture_within_1e_2)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from geometry_rl_b200 import ops  # noqa: E402
from geometry_rl_b200.synthetic import CONFIGS  # noqa: E402
from tests import gpu_helpers as G  # noqa: E402
from tests.helpers import load_golden  # noqa: E402

names = sys.argv[1:] or ["hepi_rigid_insertion", "hepi_cloth_hanging", "hepi_rope_shaping", "empn_rigid_pushing"]
for name in names:
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
    torch.cuda.synchronize()
    errs = {"out": G.rel(out, rec["out"]), "hidden": G.rel(hidden, rec["hidden"])}
    params = dict(net.named_parameters())
    for k, gref in rec["grads"].items():
        if gref is not None:
            errs["grad " + k] = G.rel(params[k].grad, gref)
    worst = sorted(errs.items(), key=lambda kv: -kv[1])
    print(f"{name}: out {errs['out']:.2e} hidden {errs['hidden']:.2e} | worst grads: " +
          ", ".join(f"{k.replace('grad ', '')} {v:.2e}" for k, v in worst[:6]))
