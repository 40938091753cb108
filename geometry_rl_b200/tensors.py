"""Small tensor helpers shared by the learner-side callers (bench.py, tests)."""
import torch


def to_device(batch, device):
    return {k: (v.to(device) if torch.is_tensor(v) else v) for k, v in batch.items()}
