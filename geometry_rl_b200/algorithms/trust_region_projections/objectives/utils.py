"""objectives/utils.py:5-28 with torchrl's `distance_loss(..., "l2")` [3P-memory] written out."""
import torch


def distance_loss(v1: torch.Tensor, v2: torch.Tensor, loss_function: str) -> torch.Tensor:
    if loss_function == "l2":
        return (v1 - v2).pow(2)
    if loss_function == "l1":
        return (v1 - v2).abs()
    if loss_function == "smooth_l1":
        return torch.nn.functional.smooth_l1_loss(v1, v2, reduction="none")
    raise NotImplementedError(f"Unknown loss {loss_function}")


def _clip_value_loss(old_state_value, state_value, clip_value, target_return, loss_value, loss_critic_type):
    state_value_clipped = old_state_value + (state_value - old_state_value).clamp(-clip_value, clip_value)
    loss_value_clipped = distance_loss(target_return, state_value_clipped, loss_function=loss_critic_type)
    loss_value = torch.max(loss_value, loss_value_clipped)
    with torch.no_grad():
        clip_fraction = (state_value / old_state_value).clamp(1 - clip_value, 1 + clip_value).abs()
    return loss_value, clip_fraction
