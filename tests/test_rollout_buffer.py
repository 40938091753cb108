"""CPU: device-resident rollout buffer + sampler without replacement (geometry_rl_b200/rollout.py).
The buffer is plain torch indexing, device-agnostic; it is exercised here on the CPU device."""
import pytest
import torch

from geometry_rl_b200.rollout import DeviceRolloutBuffer, flatten_rollout, shard_by_env


def _rollout(B, T, seed=0):
    g = torch.Generator().manual_seed(seed)
    ids = torch.arange(B * T).reshape(B, T)
    return {"frame_id": ids, "obs": torch.randn(B, T, 5, generator=g), "advantage": torch.randn(B, T, 1, generator=g),
            "done": torch.zeros(B, T, dtype=torch.bool)}


@pytest.mark.parametrize("n,mb,drop_last", [(24, 6, False), (25, 6, False), (25, 6, True), (5, 8, False), (5, 8, True)])
def test_one_pass_visits_every_frame_once(n, mb, drop_last):
    flat = flatten_rollout(_rollout(n, 1))
    buf = DeviceRolloutBuffer(n, mb, "cpu", drop_last=drop_last, generator=torch.Generator().manual_seed(1))
    buf.extend(flat)
    assert len(buf) == n
    batches = list(buf)
    sizes = [b["frame_id"].shape[0] for b in batches]
    full, rem = divmod(n, mb)
    assert sizes == [mb] * full + ([rem] if rem and not drop_last else [])
    assert len(batches) == buf.num_batches()
    seen = torch.cat([b["frame_id"] for b in batches]) if batches else torch.empty(0, dtype=torch.long)
    assert seen.unique().numel() == seen.numel()  # without replacement
    if not drop_last:
        assert torch.equal(seen.sort().values, torch.arange(n))
    for b in batches:  # rows stay together across entries
        assert torch.equal(b["obs"], flat["obs"][b["frame_id"]])
        assert torch.equal(b["advantage"], flat["advantage"][b["frame_id"]])
        assert b["done"].dtype == torch.bool


def test_each_pass_reshuffles_and_is_seed_reproducible():
    flat = flatten_rollout(_rollout(8, 4))
    orders = []
    for seed in (3, 3, 4):
        buf = DeviceRolloutBuffer(32, 8, "cpu", generator=torch.Generator().manual_seed(seed))
        buf.extend(flat)
        p1 = torch.cat([b["frame_id"] for b in buf])
        p2 = torch.cat([b["frame_id"] for b in buf])
        assert not torch.equal(p1, p2)  # a fresh permutation per pass
        orders.append(p1)
    assert torch.equal(orders[0], orders[1]) and not torch.equal(orders[0], orders[2])


def test_extend_overwrites_the_previous_collection_like_a_ring():
    buf = DeviceRolloutBuffer(12, 4, "cpu")
    a, b = flatten_rollout(_rollout(3, 4, seed=1)), flatten_rollout(_rollout(3, 4, seed=2))
    b["frame_id"] = b["frame_id"] + 100
    buf.extend(a)
    buf.extend(b)  # same size as the capacity: the new collection replaces the old one (train.py:255 every iteration)
    assert len(buf) == 12
    seen = torch.cat([x["frame_id"] for x in buf]).sort().values
    assert torch.equal(seen, torch.arange(100, 112))
    half = {k: v[:6] for k, v in a.items()}
    buf.extend(half)  # partial write wraps in place
    seen = torch.cat([x["frame_id"] for x in buf]).sort().values
    assert torch.equal(seen, torch.cat([torch.arange(0, 6), torch.arange(106, 112)]))
    with pytest.raises(ValueError):
        buf.extend({k: torch.cat([v, v]) for k, v in a.items()})
    with pytest.raises(KeyError):
        buf.extend({"frame_id": a["frame_id"]})


def test_env_shards_partition_the_global_pass():
    """Data parallel: rank r stores the frames of its env block and draws batch / world per step; the union of the
    ranks' chunks over a pass is every frame exactly once."""
    B, T, world, mb = 8, 3, 4, 8
    roll = _rollout(B, T)
    seen = []
    for r in range(world):
        shard = shard_by_env(roll, r, world)
        assert shard["frame_id"].shape[0] == B // world
        buf = DeviceRolloutBuffer(B * T // world, mb // world, "cpu", generator=torch.Generator().manual_seed(10 + r))
        buf.extend(flatten_rollout(shard))
        ids = torch.cat([b["frame_id"] for b in buf])
        envs = ids // T
        assert bool(((envs >= r * B // world) & (envs < (r + 1) * B // world)).all())  # only its own environments
        assert buf.num_batches() == B * T // mb
        seen.append(ids)
    assert torch.equal(torch.cat(seen).sort().values, torch.arange(B * T))
    with pytest.raises(ValueError):
        shard_by_env(roll, 0, 3)


def test_gather_into_static_inputs():
    flat = flatten_rollout(_rollout(4, 4))
    buf = DeviceRolloutBuffer(16, 4, "cpu", generator=torch.Generator().manual_seed(0))
    buf.extend(flat)
    static = {"obs": torch.empty(4, 5), "advantage": torch.empty(4, 1)}
    for idx in buf.sample_indices():
        buf.gather_into(idx, static)
        assert torch.equal(static["obs"], flat["obs"][idx]) and torch.equal(static["advantage"], flat["advantage"][idx])


def test_minibatch_epochs_loop():
    from geometry_rl_b200.rollout import run_minibatch_epochs
    flat = flatten_rollout(_rollout(5, 2))  # 10 frames
    buf = DeviceRolloutBuffer(10, 4, "cpu", generator=torch.Generator().manual_seed(2))
    buf.extend(flat)
    seen = []
    out = run_minibatch_epochs(lambda mb: seen.append(mb["frame_id"].clone()) or len(seen), buf, epochs=3)
    assert out == list(range(1, 10))  # 3 passes x (2 full chunks + 1 tail of 2)
    for e in range(3):
        ids = torch.cat(seen[3 * e:3 * e + 3])
        assert torch.equal(ids.sort().values, torch.arange(10))
    # static inputs: full chunks are gathered in place (same tensors every call), the tail is a fresh dict
    static = {"frame_id": torch.empty(4, dtype=torch.long), "obs": torch.empty(4, 5)}
    calls = []
    run_minibatch_epochs(lambda mb: calls.append((mb is static, mb["frame_id"].clone(), mb["obs"].clone())), buf, 1, static)
    assert [c[0] for c in calls] == [True, True, False]
    for is_static, ids, obs in calls:
        assert torch.equal(obs, flat["obs"][ids])
    assert torch.equal(torch.cat([c[1] for c in calls]).sort().values, torch.arange(10))
