"""GPU: the drop-in boundary at the level of the reference's own CALLERS.

The fixtures were produced by the UNMODIFIED reference wrappers driving the unmodified reference bodies
(oracle/make_golden.py::golden_policy_wrapper / golden_value_wrapper):
  * `GNNGaussianPolicyDiag(gnn=HEPi, hyper_data=*TasksData)(*obs) -> (loc, covariance_matrix)` — what ProbabilisticActor
    calls (abstract_gnn_gaussian_policy.py:95-109, gnn_gaussian_policy_diag.py:26-87, utils_algo_graph.py:146-158);
  * `GNNVFNet(gnn=DeepSets, hyper_data=*TasksData)(*obs)` for 2-D and 3-D observations — what ValueOperator / GAE call
    (value/gnn_vf_net.py:50-102, utils_algo_graph.py:200-203).
Here the repo's wrappers + bodies + data builders load those state dicts with strict=True and must reproduce outputs
(1e-5) and every parameter gradient (1e-5) from the same flat observation tensors."""
import pytest
import torch

from geometry_rl_b200.synthetic import CONFIGS
from tests import gpu_helpers as G
from tests.helpers import load_golden

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["policy_wrapper_rigid_insertion", "policy_wrapper_cloth_hanging"])
def test_policy_wrapper_matches_the_reference_caller(name):
    from geometry_rl_b200.algorithms.trust_region_projections.models.policy.gnn_gaussian_policy_diag import (
        GNNGaussianPolicyDiag)
    rec = load_golden(name)
    cfg = CONFIGS[rec["config"]]
    pol = GNNGaussianPolicyDiag(gnn=G.make_policy_body(cfg), hyper_data=G.make_data(cfg, policy=True),
                                action_dim=cfg.total_action_dim, num_actuators=cfg.num_actuators, init="orthogonal",
                                hidden_sizes=(64, 64), contextual_std=True, init_std=1.0, minimal_std=1e-5,
                                share_action_dim=True, post_fc=cfg.post_fc).to(G.dev())
    pol.load_state_dict(rec["state_dict"], strict=True)
    loc, cov = pol(*G.obs_args(cfg, rec["obs"], policy=True))
    assert loc.shape == rec["loc"].shape and cov.shape == rec["cov"].shape
    assert G.rel(loc, rec["loc"]) < 1e-5, G.err_report("loc", loc, rec["loc"])
    assert G.rel(cov, rec["cov"]) < 1e-5, G.err_report("cov", cov, rec["cov"])
    ((loc * rec["w_loc"].cuda()).sum() + (cov.diagonal(dim1=-2, dim2=-1) * rec["w_cov"].cuda()).sum()).backward()
    params = dict(pol.named_parameters())
    bad = []
    for k, g in rec["grads"].items():
        if g is None:
            assert params[k].grad is None or float(params[k].grad.abs().max()) == 0.0, k
        elif G.rel(params[k].grad, g) >= 1e-5:
            bad.append(G.err_report(k, params[k].grad, g))
    assert not bad, "\n".join(bad)


def test_value_wrapper_matches_the_reference_caller():
    from geometry_rl_b200.algorithms.trust_region_projections.models.value.gnn_vf_net import GNNVFNet
    from geometry_rl_b200.modules.pyg_models.deepsets import DeepSets
    from geometry_rl_b200.synthetic import obs_keys
    rec = load_golden("value_wrapper_rigid")
    cfg = CONFIGS[rec["config"]]
    vf = GNNVFNet(gnn=DeepSets(input_dim_node=3 + 12, output_dim=64, hidden_dim=64, norm=["layer_norm", "layer_norm"]),
                  hyper_data=G.make_data(cfg, policy=False), init="orthogonal", hidden_sizes=(64, 64)).to(G.dev())
    vf.load_state_dict(rec["state_dict"], strict=True)
    obs3 = [rec["obs3"][k].cuda() for k in obs_keys(cfg)]
    v2 = vf(*[o[:, 0] for o in obs3])
    v3 = vf(*obs3)  # ONE batched call over the T steps; the reference loops over time
    assert v2.shape == rec["v2"].shape and v3.shape == rec["v3"].shape
    assert G.rel(v2, rec["v2"]) < 1e-5, G.err_report("v2", v2, rec["v2"])
    assert G.rel(v3, rec["v3"]) < 1e-5, G.err_report("v3", v3, rec["v3"])
    (v3 * rec["w"].cuda()).sum().backward()
    params = dict(vf.named_parameters())
    bad = [G.err_report(k, params[k].grad, g) for k, g in rec["grads"].items()
           if g is not None and G.rel(params[k].grad, g) >= 1e-5]
    assert not bad, "\n".join(bad)
