"""Device-resident rollout buffer + minibatch sampler: the step either side of the update.

The reference stores every collected batch on the CPU and moves each minibatch to the GPU one by one
(examples/torchrl/train.py:120 `storing_device="cpu"`, :126-131 `TensorDictReplayBuffer(LazyTensorStorage(
frames_per_batch), SamplerWithoutReplacement(), batch_size=mini_batch_size)`, :255 `extend`, :259-261
`for batch in data_buffer: batch = batch.to(device)`).  Here the flattened `[frames_per_batch]` rollout stays in HBM
and a minibatch is one `index_select` per entry with a device-side permutation.

Sampler semantics (torchrl 0.3.1 `SamplerWithoutReplacement`, `drop_last=False`; [3P-memory], torchrl is not
installed offline): a fresh random permutation of the stored frames per pass; consecutive `batch_size` chunks of it;
the last chunk of a pass may be shorter; iterating the buffer yields exactly one pass.

Data parallelism (north_star: "minibatches shard by environment"): every rank stores only the frames of ITS
environments (contiguous env block, `shard_by_env`) and draws `batch_size / world_size` of them per step, so a global
minibatch is the union of the ranks' chunks and no sample is visited twice in a pass.

Fixed-shape option: `drop_last=True` skips the short tail chunk, which keeps every minibatch the shape a captured CUDA
graph (`Learner.capture`) was recorded with."""
from typing import Dict, Iterator, Mapping, Optional

import torch


def flatten_rollout(td: Mapping[str, torch.Tensor], batch_dims: int = 2) -> Dict[str, torch.Tensor]:
    """`data.reshape(-1)` of train.py:252 for a plain mapping: merge the leading [B_env, T] dims of every tensor."""
    out = {}
    for k, v in td.items():
        if torch.is_tensor(v):
            out[k] = v.reshape(-1, *v.shape[batch_dims:])
    return out


def shard_by_env(td: Mapping[str, torch.Tensor], rank: int, world_size: int) -> Dict[str, torch.Tensor]:
    """Rows [rank * B_env / world, (rank + 1) * B_env / world) of every [B_env, ...] tensor (equal shards required)."""
    out = {}
    for k, v in td.items():
        if torch.is_tensor(v):
            b = v.shape[0]
            if b % world_size:
                raise ValueError(f"{k}: {b} environments do not split evenly over {world_size} ranks")
            n = b // world_size
            out[k] = v[rank * n:(rank + 1) * n]
    return out


class DeviceRolloutBuffer:
    def __init__(self, capacity: int, batch_size: int, device, *, drop_last: bool = False,
                 generator: Optional[torch.Generator] = None):
        if capacity <= 0 or batch_size <= 0:
            raise ValueError("capacity and batch_size must be positive")
        self.capacity, self.batch_size, self.device, self.drop_last = int(capacity), int(batch_size), torch.device(device), drop_last
        self.generator = generator
        self._storage: Dict[str, torch.Tensor] = {}
        self._len = 0
        self._cursor = 0  # next write position (ring, like LazyTensorStorage + round-robin writer)

    def __len__(self) -> int:
        return self._len

    # ---- writer (train.py:255 data_buffer.extend(data_reshape)) ---------------------------------------------
    def extend(self, flat: Mapping[str, torch.Tensor]) -> None:
        n = None
        for k, v in flat.items():
            if not torch.is_tensor(v):
                continue
            if n is None:
                n = v.shape[0]
            elif v.shape[0] != n:
                raise ValueError(f"{k}: leading dim {v.shape[0]} != {n}")
        if n is None or n == 0:
            return
        if n > self.capacity:
            raise ValueError(f"{n} frames do not fit a buffer of {self.capacity}")
        if not self._storage:  # lazy allocation from the first batch (LazyTensorStorage)
            for k, v in flat.items():
                if torch.is_tensor(v):
                    self._storage[k] = torch.empty((self.capacity,) + tuple(v.shape[1:]), dtype=v.dtype, device=self.device)
        elif set(self._storage) != {k for k, v in flat.items() if torch.is_tensor(v)}:
            raise KeyError("entries differ from the ones the storage was allocated with")
        first = min(n, self.capacity - self._cursor)
        for k, dst in self._storage.items():
            src = flat[k]
            dst[self._cursor:self._cursor + first].copy_(src[:first], non_blocking=True)
            if first < n:  # wrap around
                dst[:n - first].copy_(src[first:], non_blocking=True)
        self._cursor = (self._cursor + n) % self.capacity
        self._len = min(self.capacity, self._len + n)

    # ---- sampler (train.py:259 `for k, batch in enumerate(data_buffer)`) -------------------------------------
    def permutation(self) -> torch.Tensor:
        g = self.generator
        if g is not None and g.device != self.device:
            return torch.randperm(self._len, generator=g, device=g.device).to(self.device)
        return torch.randperm(self._len, generator=g, device=self.device)

    def num_batches(self) -> int:
        full, rem = divmod(self._len, self.batch_size)
        return full + (1 if rem and not self.drop_last else 0)

    def __iter__(self) -> Iterator[Dict[str, torch.Tensor]]:
        if self._len == 0:
            return
        perm = self.permutation()
        for i in range(self.num_batches()):
            idx = perm[i * self.batch_size:(i + 1) * self.batch_size]
            yield {k: v.index_select(0, idx) for k, v in self._storage.items()}

    def sample_indices(self):
        """The index chunks of one pass (for callers that gather into pre-allocated graph inputs themselves)."""
        perm = self.permutation()
        return [perm[i * self.batch_size:(i + 1) * self.batch_size] for i in range(self.num_batches())]

    def gather_into(self, idx: torch.Tensor, out: Dict[str, torch.Tensor]) -> None:
        """out[k][:] = storage[k][idx] for every entry `out` holds (static inputs of a captured update)."""
        for k, dst in out.items():
            torch.index_select(self._storage[k], 0, idx, out=dst)


def run_minibatch_epochs(update_fn, buffer: DeviceRolloutBuffer, epochs: int, static_inputs: Optional[Dict[str, torch.Tensor]] = None):
    """train.py:258-316: `epochs` passes over the buffer, one `update_fn(minibatch)` per chunk of the pass's permutation.

    With `static_inputs` (the input tensors of a captured update, `Learner._static`) every full-size chunk is gathered
    straight into them and `update_fn` is called with that same dict — no per-minibatch allocation, fixed addresses for
    the CUDA graph; a short tail chunk (only with `drop_last=False`) cannot use the fixed-shape inputs and is handed to
    `update_fn` as a freshly gathered dict.  Returns the list of `update_fn` results."""
    results = []
    for _ in range(epochs):
        for idx in buffer.sample_indices():
            if static_inputs is not None and idx.numel() == buffer.batch_size:
                buffer.gather_into(idx, static_inputs)
                results.append(update_fn(static_inputs))
            else:
                results.append(update_fn({k: v.index_select(0, idx) for k, v in buffer._storage.items()}))
    return results
