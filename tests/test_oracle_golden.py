"""CPU: the oracle restatement against fixtures produced by the UNMODIFIED reference
(oracle/make_golden.py).  Tolerance: fp32, 1e-5 relative to the tensor's max magnitude (north_star)."""
import pytest
import torch

from geometry_rl_b200.synthetic import CONFIGS
from oracle import models as om
from oracle import projection as op
from tests.helpers import load_golden, oracle_graph_from_obs, rel_err

TOL = 1e-5


@pytest.mark.parametrize("name", ["hepi_rigid_insertion", "hepi_cloth_hanging", "hepi_rope_shaping",
                                  "empn_rigid_pushing"])
def test_policy_body_matches_reference(name):
    rec = load_golden(name)
    cfg = CONFIGS[rec["config"]]
    g, (sc, vec) = oracle_graph_from_obs(cfg, rec["obs"], policy=True)
    # topology: bit-exact coalesced edge lists, same type order
    assert g.node_types == rec["graph"]["node_types"]
    assert [tuple(e) for e in g.edge_types] == [tuple(e) for e in rec["graph"]["edge_types"]]
    for et in g.edge_types:
        assert torch.equal(g.edge_index_dict[et], rec["graph"]["edge_index"]["___".join(et)])
    assert (g.output_mask.start, g.output_mask.stop) == rec["graph"]["output_mask"]
    for nt in g.node_types:
        assert torch.equal(sc[nt], rec["scalar_dict"][nt])
        assert torch.equal(vec[nt], rec["vector_dict"][nt])
        assert torch.equal(g.pos[nt], rec["graph"]["pos"][nt])

    sd = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in rec["state_dict"].items()}
    kw = dict(dim=cfg.ponita_dim, output_dim=cfg.output_dim, output_dim_vec=cfg.output_dim_vec)
    if cfg.model == "hepi":
        out, hidden = om.hepi_forward(sd, g, sc, vec, **kw)
    else:
        assert torch.equal(g.homogeneous_edge_index(), rec["homo_edge_index"])
        out, hidden = om.empn_forward(sd, g, sc, vec, **kw)
    assert rel_err(out, rec["out"]) < TOL
    assert rel_err(hidden, rec["hidden"]) < TOL
    loss = (out * rec["w_out"]).sum() + (hidden * rec["w_hid"]).sum()
    loss.backward()
    checked = 0
    for k, gref in rec["grads"].items():
        if gref is None:
            assert sd[k].grad is None or float(sd[k].grad.abs().max()) == 0.0, k
            continue
        assert rel_err(sd[k].grad, gref) < 5 * TOL, k
        checked += 1
    assert checked >= 20


@pytest.mark.parametrize("name", ["deepsets_rigid", "deepsets_rope"])
def test_deepsets_matches_reference(name):
    rec = load_golden(name)
    cfg = CONFIGS[rec["config"]]
    g, feats = oracle_graph_from_obs(cfg, rec["obs"], policy=False)
    tokens = om.concat_tokens(g, feats)
    assert torch.equal(tokens, rec["tokens"])
    sd = {k: v.clone().requires_grad_(True) for k, v in rec["state_dict"].items()}
    out = om.deepsets_forward(sd, tokens)
    assert rel_err(out, rec["out"]) < TOL
    (out * rec["w"]).sum().backward()
    for k, gref in rec["grads"].items():
        assert rel_err(sd[k].grad, gref) < 5 * TOL, k


def test_transformer_matches_reference():
    rec = load_golden("transformer_two_agents")
    cfg = CONFIGS["rigid_insertion_two_agents_multi_transformer_trpl_cfg"]
    g, feats = oracle_graph_from_obs(cfg, rec["obs"], policy=True)
    tokens = om.concat_tokens(g, feats)
    assert torch.equal(tokens, rec["tokens"])
    assert (g.output_mask.start, g.output_mask.stop) == rec["output_mask"]
    sd = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in rec["state_dict"].items()}
    out = om.transformer_forward(sd, tokens, g.output_mask)
    assert rel_err(out, rec["out"]) < TOL
    (out * rec["w"]).sum().backward()
    for k, gref in rec["grads"].items():
        if gref is None:
            continue
        assert rel_err(sd[k].grad, gref) < 5 * TOL, k


def test_gaussian_head_matches_reference():
    rec = load_golden("gaussian_head")
    for key, r in rec.items():
        gnn_out = r["hidden"] if r["post_fc"] else (r["mean_in"], r["hidden"])
        loc, var = om.gaussian_head(r["state_dict"], gnn_out, r["B"], post_fc=r["post_fc"])
        assert rel_err(loc, r["loc"]) < TOL, key
        assert rel_err(torch.diag_embed(var), r["cov"]) < TOL, key


@pytest.mark.parametrize("key", ["kl_k6", "w2_k6", "kl_k3", "w2_k3", "kl_k12", "w2_k12"])
def test_projection_matches_reference(key):
    r = load_golden("projection")[key]
    ptype = key.split("_")[0]
    mean = r["mean"].clone().requires_grad_(True)
    v = r["v"].clone().requires_grad_(True)
    if ptype == "kl":
        pm, pv = op.kl_projection(mean, v, r["q_mean"], r["q_v"], r["eps_mean"], r["eps_cov"])
    else:
        pm, pv = op.w2_projection(mean, v, r["q_mean"], r["q_v"], r["eps_mean"], r["eps_cov"])
    assert rel_err(pm, r["proj_mean"]) < TOL
    assert rel_err(pv, r["proj_v"]) < TOL
    logp = op.mvn_diag_log_prob(r["action"], pm, pv)
    ent = op.mvn_diag_entropy(pv)
    assert rel_err(logp, r["logp"]) < TOL
    assert rel_err(ent, r["entropy"]) < TOL
    trl = op.trust_region_loss(mean, v, pm, pv, r["coeff"], ptype)
    assert rel_err(trl, r["tr_loss"]) < TOL
    met = op.compute_metrics(mean, v, r["q_mean"], r["q_v"], ptype)
    for m, val in r["metrics"].items():
        assert rel_err(met[m], val) < 2 * TOL or abs(float(met[m] - val)) < 1e-6, m
    total = -(torch.exp(logp - logp.detach()) * r["adv"]).mean() - 0.005 * ent.mean() + trl
    g_mean, g_v = torch.autograd.grad(total, (mean, v))
    assert rel_err(g_mean, r["g_mean"]) < 5 * TOL
    assert rel_err(g_v, r["g_v"]) < 5 * TOL
    # the fixture must exercise both branches
    active = (op.gaussian_kl(r["mean"], r["v"], r["q_mean"], r["q_v"])[0] > r["eps_mean"])
    assert bool(active.any()) and bool((~active).any())


def test_w2_grad_noise_floor():
    """The reference's fp32 W2 covariance gradient is only accurate to a few 1e-5 (cancellation in
    trace(I + c - 2 S_q^-1 S)): an fp64 evaluation of the same formulas deviates from the fp32 fixture by
    2e-5 .. 5e-5 relative.  This is the noise floor behind the 1e-4 W2 gradient tolerance of the GPU tests."""
    worst = 0.0
    for key in ("w2_k6", "w2_k3", "w2_k12"):
        r = load_golden("projection")[key]
        d = torch.float64
        mean, v = r["mean"].to(d).requires_grad_(True), r["v"].to(d).requires_grad_(True)
        pm, pv = op.w2_projection(mean, v, r["q_mean"].to(d), r["q_v"].to(d), r["eps_mean"], r["eps_cov"])
        logp, ent = op.mvn_diag_log_prob(r["action"].to(d), pm, pv), op.mvn_diag_entropy(pv)
        trl = op.trust_region_loss(mean, v, pm, pv, r["coeff"], "w2")
        total = -(torch.exp(logp - logp.detach()) * r["adv"].to(d)).mean() - 0.005 * ent.mean() + trl
        _, g_v = torch.autograd.grad(total, (mean, v))
        worst = max(worst, rel_err(g_v.detach(), r["g_v"]))
    assert 1e-5 < worst < 1e-4, worst


def test_kl_projection_kkt_and_gradcheck():
    torch.manual_seed(0)
    B, k, eps = 32, 6, 0.0025
    o = (0.5 + torch.rand(B, k, dtype=torch.double)) ** 2
    c = o * torch.exp(torch.randn(B, k, dtype=torch.double) * 0.3)
    c_t, eta = op.kl_diag_cov_solve(c, o, eps)
    kl = 0.5 * (c_t / o - 1 + o.log() - c_t.log()).sum(-1)
    kl0 = 0.5 * (c / o - 1 + o.log() - c.log()).sum(-1)
    assert bool(((kl0 <= eps) == (eta == 0)).all())
    assert float((kl[eta > 0] - eps).abs().max()) < 1e-12
    assert torch.equal(c_t[eta == 0], c[eta == 0])
    cg = c.clone().requires_grad_(True)
    assert torch.autograd.gradcheck(lambda x: op.KLDiagCovProjection.apply(x, o, eps), (cg,), eps=1e-7, atol=1e-6)


def test_equivariance_kat():
    """Ponita.main() rotated-copies input (ponita.py:391-408): outputs of graph i are the 90-degree
    rotations of graph 0 (S1 grid with 4 orientations) — recorded reference outputs satisfy it."""
    r = load_golden("ponita_equivariance")
    osc, ovec, R = r["out_scalar"], r["out_vec"], r["R"]
    assert float((osc - osc[:1]).abs().max()) < 1e-6
    v = ovec[0]
    for i in range(1, 4):
        v = torch.einsum("ij,ncj->nci", R, v)
        assert float((ovec[i] - v).abs().max()) < 1e-6
