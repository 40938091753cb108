"""Oracle (test infrastructure): functional CPU restatement of the reference networks.

Parameters are passed as a flat `state_dict`-style mapping with the *reference's own key names*
(e.g. `processor.1.convs.<object_geometry___task___grippers>.kernel.weight`), so the same weights can
be fed to the unmodified reference (golden generation), to this oracle, and to the CUDA modules.

Reference lines restated (paths under geometry_rl/modules/pyg_models/):
  ponita/ponita.py:53-97      orientation grids           -> ori_grid
  ponita/ponita.py:233-244    PolynomialFeatures          -> poly_features
  hepi.py:109-123             compute_invariants          -> invariants
  hepi.py:76-89               basis_fn / fiber_basis_fn   -> basis_mlp
  ponita/utils/to_from_sphere.py:4-17                     -> lift / readout
  ponita/conv.py:71-149       FiberBundleConv             -> fiber_bundle_conv
  ponita/hetero_fiber_conv.py:33-64 + hepi.py:125-190     -> hepi_forward
  ponita/ponita.py:149-185,219-230,349-369, ponita_gcn.py:88-146 -> empn_forward
  deepsets.py:34-53 (+ PyG MLP / LayerNorm(mode='graph') [3P-memory]) -> deepsets_forward
  transformer_vanilla.py:50-92                            -> transformer_forward
  algorithms/.../policy/gnn_gaussian_policy_diag.py:26-87 -> gaussian_head
"""
import math
from typing import Dict, List, Mapping, Sequence, Tuple

import torch
import torch.nn.functional as F


# ---------------------------------------------------------------------------------------------
# grids, features
# ---------------------------------------------------------------------------------------------
def ori_grid(dim: int, n: int, only_upper_hemisphere: bool = False) -> torch.Tensor:
    """ponita/ponita.py:53-97.  S1: n equispaced angles; S2: Fibonacci lattice (offset 0.5)."""
    if dim == 2:
        ang = torch.linspace(0, 2 * torch.pi - (2 * torch.pi / n), n)
        return torch.stack((torch.cos(ang), torch.sin(ang)), dim=1)
    i = torch.arange(n)
    theta = (math.pi * i * (1 + math.sqrt(5))) % (2 * math.pi)
    scale = 1 if only_upper_hemisphere else 2
    phi = torch.acos(1 - scale * (i + 0.5) / (n - 1 + 2 * 0.5))
    return torch.stack((torch.cos(theta) * torch.sin(phi), torch.sin(theta) * torch.sin(phi), torch.cos(phi)), dim=-1)


def poly_features(x: torch.Tensor, degree: int = 2) -> torch.Tensor:
    """[x, x(x)x, (x(x)x)(x)x] flattened: 2+4+8 = 14 for the spatial pair, 1+1+1 = 3 for the fibre."""
    feats = [x]
    for _ in range(degree):
        feats.append((feats[-1][..., :, None] * x[..., None, :]).flatten(-2, -1))
    return torch.cat(feats, -1)


def invariants(grid: torch.Tensor, pos_src: torch.Tensor, pos_dst: torch.Tensor):
    rel = (pos_src - pos_dst)[:, None, :]
    i1 = (rel * grid[None]).sum(-1, keepdim=True)
    i2 = (rel - i1 * grid[None]).norm(dim=-1, keepdim=True)
    i3 = (grid[None, :, :] * grid[:, None, :]).sum(-1, keepdim=True)
    return torch.cat([i1, i2], -1), i3


def basis_mlp(x: torch.Tensor, sd: Mapping[str, torch.Tensor], prefix: str, degree: int = 2) -> torch.Tensor:
    h = poly_features(x, degree)
    h = F.gelu(F.linear(h, sd[f"{prefix}.1.weight"], sd[f"{prefix}.1.bias"]))
    return F.gelu(F.linear(h, sd[f"{prefix}.3.weight"], sd[f"{prefix}.3.bias"]))


def lift(scalars: torch.Tensor, vectors: torch.Tensor, grid: torch.Tensor) -> torch.Tensor:
    """scalar_to_sphere / vec_to_sphere + concat: [N,S], [N,3V] -> [N,O,S+V]."""
    dim = grid.shape[-1]
    s = scalars[:, None, :].expand(-1, grid.shape[0], -1)
    v = vectors.view(scalars.shape[0], -1, 3)[..., :dim]
    v = torch.einsum("bcd,nd->bnc", v, grid)
    return torch.cat([s, v], dim=-1)


# ---------------------------------------------------------------------------------------------
# one FiberBundleConv (HEPi flavour, conv.py:71-114) — also the body of the EMPN layer
# ---------------------------------------------------------------------------------------------
def fiber_bundle_conv(x_src, x_dst, edge_index, kernel_basis, fiber_basis, sd, prefix,
                      fiber_transposed=False, mlp_names=("node_mlp.0", "node_mlp.1", "node_mlp.3"),
                      intermediates: dict = None):
    kern = F.linear(kernel_basis, sd[f"{prefix}.kernel.weight"])
    msg = kern * x_src[edge_index[0]]
    x1 = torch.zeros_like(x_dst).index_add_(0, edge_index[1], msg)
    fk = F.linear(fiber_basis, sd[f"{prefix}.fiber_kernel.weight"])
    if fiber_transposed:  # ponita.py:166  "boc,poc->bpc"
        x2 = torch.einsum("boc,poc->bpc", x1, fk) / fk.shape[-2]
    else:  # conv.py:90  "boc,opc->bpc"
        x2 = torch.einsum("boc,opc->bpc", x1, fk) / fk.shape[-2]
    x2 = x2 + sd[f"{prefix}.bias"]
    if intermediates is not None:
        intermediates.update(kern=kern, x1=x1, x2=x2, fk=fk)
    return x2


def convnext_update(x_in, x2, sd, ln, lin1, lin2):
    y = F.layer_norm(x2, (x2.shape[-1],), sd[f"{ln}.weight"], sd[f"{ln}.bias"], 1e-5)
    h = F.gelu(F.linear(y, sd[f"{lin1}.weight"], sd[f"{lin1}.bias"]))
    return x_in + F.linear(h, sd[f"{lin2}.weight"], sd[f"{lin2}.bias"])


def _edge_key(et: Tuple[str, str, str]) -> str:
    return "<" + "___".join(et) + ">"


# ---------------------------------------------------------------------------------------------
# HEPi (hepi.py:125-190)
# ---------------------------------------------------------------------------------------------
def hepi_forward(sd: Mapping[str, torch.Tensor], graph, scalar_dict, vector_dict, *, dim: int,
                 output_dim: int, output_dim_vec: int, num_messages: int = 2):
    grid = sd["ori_grid"]
    num_ori = grid.shape[0]
    latent = {}
    for nt in graph.node_types:
        latent[nt] = F.linear(lift(scalar_dict[nt], vector_dict[nt], grid), sd["node_encoder.weight"])

    kb, fb = {}, {}
    for et in graph.edge_types:
        src, _, dst = et
        ei = graph.edge_index_dict[et]
        ps = graph.pos[src][ei[0]][..., :dim]
        pd = graph.pos[dst][ei[1]][..., :dim]
        sp, oi = invariants(grid, ps, pd)
        kb[et] = basis_mlp(sp, sd, "basis_fn")
        fb[et] = basis_mlp(oi, sd, "fiber_basis_fn")

    for i in range(num_messages):
        outs: Dict[str, List[torch.Tensor]] = {}
        # ModuleDict order == insertion order == EdgeLevel order (hepi.py:93-104)
        conv_keys = [k for k in sd.keys() if k.startswith(f"processor.{i}.convs.") and k.endswith(".kernel.weight")]
        for key in conv_keys:
            ek = key[len(f"processor.{i}.convs."):-len(".kernel.weight")]
            et = tuple(ek[1:-1].split("___"))
            ei = graph.edge_index_dict[et]
            if ei.numel() == 0:
                continue  # hetero_fiber_conv.py:48-49
            src, _, dst = et
            prefix = f"processor.{i}.convs.{ek}"
            x2 = fiber_bundle_conv(latent[src], latent[dst], ei, kb[et], fb[et], sd, prefix)
            upd = convnext_update(latent[dst], x2, sd, f"{prefix}.node_mlp.0", f"{prefix}.node_mlp.1",
                                  f"{prefix}.node_mlp.3")
            outs.setdefault(dst, []).append(upd)
        for dst, vals in outs.items():  # group(..., "sum")
            latent[dst] = vals[0] if len(vals) == 1 else torch.stack(vals, 0).sum(0)

    lat = latent[graph.output_mask_key]
    out = F.linear(lat, sd["decoder.weight"], sd["decoder.bias"])
    out_scalar, out_vec = out.split([output_dim, output_dim_vec], dim=-1)
    hidden = lat.mean(dim=-2)
    out_scalar = out_scalar.mean(dim=-2)
    out_vec = torch.einsum("boc,od->bcd", out_vec, grid) / num_ori
    res = out_vec * out_scalar.unsqueeze(-1)
    if dim == 2:
        res = torch.cat([res, torch.zeros_like(res[..., :1])], dim=-1)
    return res.reshape(-1, res.shape[-1]), hidden.reshape(-1, hidden.shape[-1])


# ---------------------------------------------------------------------------------------------
# EMPN = PonitaGCN -> Ponita (ponita_gcn.py:88-146, ponita.py:349-369)
# ---------------------------------------------------------------------------------------------
def empn_forward(sd: Mapping[str, torch.Tensor], graph, scalar_dict, vector_dict, *, dim: int,
                 output_dim: int, output_dim_vec: int, num_layers: int = 2):
    grid = sd["ponita.ori_grid"]
    num_ori = grid.shape[0]
    B = graph.num_graphs
    xs, ps = [], []
    for nt in graph.node_types:
        x = lift(scalar_dict[nt], vector_dict[nt], grid)
        xs.append(x.reshape(B, -1, *x.shape[1:]))
        ps.append(graph.pos[nt].reshape(B, -1, 3)[..., :dim])
    x = torch.cat(xs, dim=1)
    pos = torch.cat(ps, dim=1)
    n_per_graph = x.shape[1]
    x = x.reshape(-1, *x.shape[2:])
    pos = pos.reshape(-1, dim)
    ei = graph.homogeneous_edge_index()

    sp, oi = invariants(grid, pos[ei[0]], pos[ei[1]])
    kb = basis_mlp(sp, sd, "ponita.basis_fn")
    fb = basis_mlp(oi, sd, "ponita.fiber_basis_fn")
    x = F.linear(x, sd["ponita.x_embedder.weight"])
    for l in range(num_layers):
        p = f"ponita.interaction_layers.{l}"
        x2 = fiber_bundle_conv(x, x, ei, kb, fb, sd, f"{p}.conv", fiber_transposed=True)
        x = convnext_update(x, x2, sd, f"{p}.norm", f"{p}.linear_1", f"{p}.linear_2")

    hidden = x.reshape(B, n_per_graph, num_ori, -1)
    out = F.linear(hidden, sd["linear.weight"], sd["linear.bias"])
    out_scalar, out_vec = out.split([output_dim, output_dim_vec], dim=-1)
    hidden = hidden.mean(dim=-2)
    out_scalar = out_scalar.mean(dim=-2)
    out_vec = torch.einsum("bnoc,od->bncd", out_vec, grid) / num_ori
    m = graph.output_mask
    hidden, out_scalar, out_vec = hidden[:, m], out_scalar[:, m], out_vec[:, m]
    res = out_vec * out_scalar.unsqueeze(-1)
    if dim == 2:
        res = torch.cat([res, torch.zeros_like(res[..., :1])], dim=-1)
    return res.reshape(-1, res.shape[-1]), hidden.reshape(-1, hidden.shape[-1])


# ---------------------------------------------------------------------------------------------
# DeepSets critic body (deepsets.py:34-53) with PyG MLP / LayerNorm(mode='graph') [3P-memory]
# ---------------------------------------------------------------------------------------------
def _pyg_graph_layer_norm(x, w, b, eps=1e-5):
    x = x - x.mean()
    out = x / (x.std(unbiased=False) + eps)
    return out * w + b


def _pyg_mlp(x, sd, prefix, n_lins):
    for i in range(n_lins - 1):
        x = F.linear(x, sd[f"{prefix}.lins.{i}.weight"], sd[f"{prefix}.lins.{i}.bias"])
        if f"{prefix}.norms.{i}.weight" in sd:
            x = _pyg_graph_layer_norm(x, sd[f"{prefix}.norms.{i}.weight"], sd[f"{prefix}.norms.{i}.bias"])
        x = F.relu(x)
    i = n_lins - 1
    return F.linear(x, sd[f"{prefix}.lins.{i}.weight"], sd[f"{prefix}.lins.{i}.bias"])


def deepsets_forward(sd, x):
    """x: [B, N, F] (all node types concatenated per graph).  Returns [B, out]."""
    h = _pyg_mlp(x, sd, "mlp_inner", 2)
    return _pyg_mlp(h.sum(dim=1), sd, "mlp_outer", 2)


def concat_tokens(graph, input_vector_dict):
    B = graph.num_graphs
    return torch.cat([input_vector_dict[nt].reshape(B, -1, input_vector_dict[nt].shape[-1])
                      for nt in graph.node_types], dim=1)


# ---------------------------------------------------------------------------------------------
# Transformer baseline (transformer_vanilla.py:50-92, concat_global=False):
# nn.TransformerEncoder, post-LN, ReLU FF, dropout 0, restated explicitly
# ---------------------------------------------------------------------------------------------
def transformer_forward(sd, x, output_mask: slice, num_layers=2, num_heads=2):
    h = F.linear(x, sd["embedding.weight"], sd["embedding.bias"])  # [B,S,D]
    B, S, D = h.shape
    hd = D // num_heads
    for l in range(num_layers):
        p = f"transformer_encoder.layers.{l}"
        qkv = F.linear(h, sd[f"{p}.self_attn.in_proj_weight"], sd[f"{p}.self_attn.in_proj_bias"])
        q, k, v = qkv.split(D, dim=-1)
        q = q.view(B, S, num_heads, hd).transpose(1, 2)
        k = k.view(B, S, num_heads, hd).transpose(1, 2)
        v = v.view(B, S, num_heads, hd).transpose(1, 2)
        att = torch.softmax((q @ k.transpose(-1, -2)) / math.sqrt(hd), dim=-1)
        a = (att @ v).transpose(1, 2).reshape(B, S, D)
        a = F.linear(a, sd[f"{p}.self_attn.out_proj.weight"], sd[f"{p}.self_attn.out_proj.bias"])
        h = F.layer_norm(h + a, (D,), sd[f"{p}.norm1.weight"], sd[f"{p}.norm1.bias"], 1e-5)
        f = F.linear(F.relu(F.linear(h, sd[f"{p}.linear1.weight"], sd[f"{p}.linear1.bias"])),
                     sd[f"{p}.linear2.weight"], sd[f"{p}.linear2.bias"])
        h = F.layer_norm(h + f, (D,), sd[f"{p}.norm2.weight"], sd[f"{p}.norm2.bias"], 1e-5)
    h = h[:, output_mask].reshape(-1, D)
    return F.linear(h, sd["fc_out.lins.0.weight"], sd["fc_out.lins.0.bias"])


# ---------------------------------------------------------------------------------------------
# Gaussian head (gnn_gaussian_policy_diag.py:26-87; contextual std)
# ---------------------------------------------------------------------------------------------
def gaussian_head(sd, gnn_out, batch_size: int, *, post_fc: bool, init_std=1.0, minimal_std=1e-5):
    """Returns (loc [B,k], var_diag [B,k]); the reference returns diag_embed(std)**2 == diag(var)."""
    shift = torch.log(torch.exp(torch.tensor(init_std) - torch.tensor(minimal_std)) - 1.0)
    if post_fc:
        hidden = gnn_out
        mean = F.linear(hidden, sd["_mean.weight"], sd["_mean.bias"])
    else:
        mean, hidden = gnn_out
    std = F.softplus(F.linear(hidden, sd["_pre_std.weight"], sd["_pre_std.bias"]) + shift) + minimal_std
    std = std.reshape(batch_size, -1)
    return mean.reshape(batch_size, -1), std ** 2
