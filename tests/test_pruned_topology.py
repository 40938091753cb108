"""CPU: index bookkeeping of the pruned EMPN topology (GraphBatch.homogeneous_pruned / ops.build_sub_edge_set).

The derivation is plain torch index arithmetic, so it is checked here without a GPU: a dense two-layer message
passing over the padded homogeneous graph, read at the output nodes, must equal layer 1 over the live nodes followed
by layer 2 over the "live -> output" sub edge set, and the src-sorted view handed to the backward kernels must
address exactly the sub-edges through the parent's edge order."""
import torch

from geometry_rl_b200 import ops
from geometry_rl_b200.modules.pyg_data.graph import GraphBatch


def _edge_set_cpu(coo, n_nodes):
    """Python restatement of ops.build_edge_set's conventions (dst-sorted, ties in COO order; src-sorted view of
    the edge-ordered list, ties in edge order)."""
    order = torch.sort(coo[1], stable=True).indices
    src, dst = coo[0][order], coo[1][order]

    def rowptr(keys):
        rp = torch.zeros(n_nodes + 1, dtype=torch.int64)
        rp[1:] = torch.cumsum(torch.bincount(keys, minlength=n_nodes), 0)
        return rp.to(torch.int32)

    so = torch.sort(src, stable=True).indices
    return ops.EdgeSet(n_nodes, n_nodes, coo.shape[1], coo, None, rowptr(dst), src.to(torch.int32), dst.to(torch.int32),
                       order.to(torch.int32), rowptr(src), so.to(torch.int32))


def _edge_set_bipartite_cpu(coo, n_src, n_dst):
    order = torch.sort(coo[1], stable=True).indices
    src, dst = coo[0][order], coo[1][order]

    def rowptr(keys, n):
        rp = torch.zeros(n + 1, dtype=torch.int64)
        rp[1:] = torch.cumsum(torch.bincount(keys, minlength=n), 0)
        return rp.to(torch.int32)

    so = torch.sort(src, stable=True).indices
    return ops.EdgeSet(n_src, n_dst, coo.shape[1], coo, None, rowptr(dst, n_dst), src.to(torch.int32), dst.to(torch.int32),
                       order.to(torch.int32), rowptr(src, n_src), so.to(torch.int32))


def test_hetero_pruned_equals_dense_at_the_output_type():
    """HEPi schedule on a toy graph: step 0 object->object, step 1 object->gripper; a third node type is isolated."""
    gen = torch.Generator().manual_seed(3)
    B, P, A, T = 6, 8, 2, 5
    internal, task = [], []
    for g in range(B):
        nv = int(torch.randint(0, P + 1, (1,), generator=gen))
        for i in range(nv):
            for j in torch.randperm(nv, generator=gen)[:3].tolist():
                if j != i:
                    internal.append((g * P + j, g * P + i))
            for a in range(A):
                task.append((g * P + i, g * A + a))
    graph = GraphBatch(B, {"obj": P, "grip": A, "target": T}, torch.device("cpu"))
    graph.output_mask_key = "grip"
    e_int, e_task = ("obj", "internal", "obj"), ("obj", "task", "grip")
    graph.edge_types = [e_int, e_task]
    graph.edge_sets[e_int] = _edge_set_bipartite_cpu(torch.tensor(internal).t().contiguous(), B * P, B * P)
    graph.edge_sets[e_task] = _edge_set_bipartite_cpu(torch.tensor(task).t().contiguous(), B * P, B * A)
    pr = graph.hetero_pruned()
    assert set(pr.live_ids) == {"obj", "grip"} and len(pr.live_ids["grip"]) == B * A
    assert len(pr.live_ids["obj"]) < B * P
    x_obj = torch.randn(B * P, 3, generator=gen, dtype=torch.float64)
    x_grip = torch.randn(B * A, 3, generator=gen, dtype=torch.float64)
    w_int = torch.randn(len(internal), generator=gen, dtype=torch.float64)
    w_task = torch.randn(len(task), generator=gen, dtype=torch.float64)
    es_i, es_t = graph.edge_sets[e_int], graph.edge_sets[e_task]
    h = _layer(x_obj, x_obj, es_i.edge_src, es_i.edge_dst, w_int)
    dense = _layer(h, x_grip, es_t.edge_src, es_t.edge_dst, w_task)
    ci, ct = pr.edge_sets[e_int], pr.edge_sets[e_task]
    xo = x_obj[pr.live_ids["obj"]]
    h1 = _layer(xo, xo, ci.edge_src, ci.edge_dst, w_int)
    pruned = _layer(h1, x_grip[pr.live_ids["grip"]], ct.edge_src, ct.edge_dst, w_task)
    assert torch.equal(pruned, dense)
    for es in (ci, ct):  # CSR rows consistent with the renumbered edge lists
        assert torch.equal((es.rowptr_dst[1:] - es.rowptr_dst[:-1]).long(), torch.bincount(es.edge_dst.long(), minlength=es.n_dst))
        assert torch.equal((es.rowptr_src[1:] - es.rowptr_src[:-1]).long(), torch.bincount(es.edge_src.long(), minlength=es.n_src))
        ss = es.edge_src[es.src_eid.long()]
        assert bool((ss[1:] >= ss[:-1]).all()) and bool((es.edge_dst[1:] >= es.edge_dst[:-1]).all())


def _random_batch(B, P, A, gen):
    n_tot = P + A
    coo = []
    for g in range(B):
        nv = int(torch.randint(0, P + 1, (1,), generator=gen))
        base = g * n_tot
        for i in range(nv):  # "internal": up to 3 random neighbours -> i
            for j in torch.randperm(nv, generator=gen)[:3].tolist():
                if j != i:
                    coo.append((base + j, base + i))
            for a in range(A):  # "task": every valid point -> every actuator
                coo.append((base + i, base + P + a))
        for a in range(A):  # "agent": actuators pairwise
            for b in range(A):
                if a != b:
                    coo.append((base + a + P, base + b + P))
    coo = torch.tensor(coo, dtype=torch.int64).t().contiguous()
    graph = GraphBatch(B, {"object_geometry": P, "grippers": A}, torch.device("cpu"))
    graph.output_mask = slice(P, P + A)
    graph._homo_cache["homo"] = _edge_set_cpu(coo, B * n_tot)
    return graph


def _layer(x_src, x_dst, edge_src, edge_dst, w):
    """x_dst + sum over in-edges of w_e * x_src[src_e] (stand-in for the fibre-bundle convolution)."""
    out = x_dst.clone()
    out.index_add_(0, edge_dst.long(), w[:, None] * x_src[edge_src.long()])
    return out


def test_pruned_equals_dense_at_the_output_nodes():
    gen = torch.Generator().manual_seed(0)
    for B, P, A in [(5, 7, 1), (4, 6, 2), (3, 1, 1)]:
        graph = _random_batch(B, P, A, gen)
        es = graph.homogeneous()
        pr = graph.homogeneous_pruned()
        n_tot = P + A
        x0 = torch.randn(B * n_tot, 3, generator=gen, dtype=torch.float64)
        w = torch.randn(es.n_edges, generator=gen, dtype=torch.float64)  # per-edge weight in EDGE order
        # dense: two layers over all nodes, read at the output nodes
        h = _layer(x0, x0, es.edge_src, es.edge_dst, w)
        h = _layer(h, h, es.edge_src, es.edge_dst, w)
        dense = h.reshape(B, n_tot, 3)[:, graph.output_mask].reshape(-1, 3)
        # pruned: layer 1 over the live nodes, layer 2 over the sub edge set (same edge order -> same per-edge weights)
        c, sub = pr.es, pr.sub
        assert c.n_edges == es.n_edges and c.n_src == len(pr.live_ids) == c.n_dst
        xl = x0[pr.live_ids]
        h1 = _layer(xl, xl, c.edge_src, c.edge_dst, w)
        out = _layer(h1, h1[sub.out_ids], sub.edge_src, sub.edge_dst, w[sub.eids])
        assert torch.equal(out, dense)
        # every dropped node is isolated and not an output node
        dead = torch.ones(B * n_tot, dtype=torch.bool)
        dead[pr.live_ids] = False
        touched = torch.zeros(B * n_tot, dtype=torch.bool)
        touched[es.edge_src.long()] = True
        touched[es.edge_dst.long()] = True
        assert not bool((dead & touched).any())
        assert sub.n_dst == B * A


def test_sub_edge_set_csr_views():
    gen = torch.Generator().manual_seed(1)
    graph = _random_batch(6, 9, 2, gen)
    pr = graph.homogeneous_pruned()
    c, sub = pr.es, pr.sub
    # compact CSR rows are consistent with the compact edge lists
    for rp, keys in [(c.rowptr_dst, c.edge_dst)]:
        deg = torch.bincount(keys.long(), minlength=c.n_dst)
        assert torch.equal((rp[1:] - rp[:-1]).long(), deg)
    assert bool((c.edge_dst[1:] >= c.edge_dst[:-1]).all())
    src_sorted = c.edge_src[c.src_eid.long()]
    assert bool((src_sorted[1:] >= src_sorted[:-1]).all())
    assert torch.equal((c.rowptr_src[1:] - c.rowptr_src[:-1]).long(), torch.bincount(c.edge_src.long(), minlength=c.n_src))
    # forward view: dst-sorted, ranks 0..n_dst-1, rowptr matches, edges are the parent's edges into out_ids
    assert bool((sub.edge_dst[1:] >= sub.edge_dst[:-1]).all())
    assert torch.equal((sub.rowptr_dst[1:] - sub.rowptr_dst[:-1]).long(), torch.bincount(sub.edge_dst.long(), minlength=sub.n_dst))
    assert torch.equal(sub.out_ids[sub.edge_dst.long()], c.edge_dst[sub.eids].long())
    assert torch.equal(sub.edge_src, c.edge_src[sub.eids])
    into_out = torch.isin(c.edge_dst.long(), sub.out_ids)
    assert torch.equal(into_out.nonzero().squeeze(1), sub.eids)
    # backward view: walking rowptr_src / src_eid_parent visits each sub-edge exactly once, grouped by source node in
    # edge order, and reaches (src, dst rank) through the PARENT's arrays
    seen = []
    for n in range(sub.n_src):
        lo, hi = int(sub.rowptr_src[n]), int(sub.rowptr_src[n + 1])
        e = sub.src_eid_parent[lo:hi].long()
        assert bool((c.edge_src[e] == n).all())
        assert bool((e[1:] > e[:-1]).all())
        assert torch.equal(sub.out_ids[sub.edge_dst_parent[e].long()], c.edge_dst[e].long())
        seen.append(e)
    assert torch.equal(torch.sort(torch.cat(seen)).values, sub.eids)
    assert int(sub.rowptr_src[-1]) == sub.n_edges == int(sub.rowptr_dst[-1])
