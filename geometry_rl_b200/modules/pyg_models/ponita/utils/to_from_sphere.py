"""Sphere lift / readout helpers with the reference's names
(geometry_rl/modules/pyg_models/ponita/utils/to_from_sphere.py:4-17).  Only used on tiny tensors
(readout of the actuator nodes, calibration); the per-node lift itself is fused into grl_embed_fwd."""
import torch


def vec_to_sphere(vec, ori_grid):
    return torch.einsum("bcd,nd->bnc", vec, ori_grid)


def scalar_to_sphere(scalar, ori_grid):
    return scalar.unsqueeze(-2).expand(*scalar.shape[:-1], ori_grid.shape[-2], scalar.shape[-1])


def sphere_to_vec(spherical_signal, ori_grid):
    return torch.einsum("bnc,nd->bcd", spherical_signal, ori_grid) / ori_grid.shape[-2]


def sphere_to_scalar(spherical_signal):
    return spherical_signal.mean(dim=-2)
