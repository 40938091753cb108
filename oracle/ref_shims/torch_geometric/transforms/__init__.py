"""Compose / BaseTransform / Cartesian / Distance — only what pyg_data/transforms.py touches."""
import torch


class BaseTransform:
    def __call__(self, data):
        return self.forward(data)

    def forward(self, data):
        raise NotImplementedError


class Compose(BaseTransform):
    def __init__(self, transforms):
        self.transforms = transforms

    def forward(self, data):
        for t in self.transforms:
            data = t(data)
        return data


class Cartesian(BaseTransform):
    def __init__(self, norm=True, max_value=None, cat=True):
        self.norm, self.max, self.cat = norm, max_value, cat

    def forward(self, data):
        (row, col), pos, pseudo = data.edge_index, data.pos, data.edge_attr
        cart = pos[row] - pos[col]
        cart = cart.view(-1, 1) if cart.dim() == 1 else cart
        if self.norm and cart.numel() > 0:
            max_value = cart.abs().max() if self.max is None else self.max
            cart = cart / (2 * max_value) + 0.5
        if pseudo is not None and self.cat:
            pseudo = pseudo.view(-1, 1) if pseudo.dim() == 1 else pseudo
            data.edge_attr = torch.cat([pseudo, cart.type_as(pseudo)], dim=-1)
        else:
            data.edge_attr = cart
        return data


class Distance(BaseTransform):
    def __init__(self, norm=True, max_value=None, cat=True):
        self.norm, self.max, self.cat = norm, max_value, cat

    def forward(self, data):
        (row, col), pos, pseudo = data.edge_index, data.pos, data.edge_attr
        dist = torch.norm(pos[col] - pos[row], p=2, dim=-1).view(-1, 1)
        if self.norm and dist.numel() > 0:
            dist = dist / (dist.max() if self.max is None else self.max)
        if pseudo is not None and self.cat:
            pseudo = pseudo.view(-1, 1) if pseudo.dim() == 1 else pseudo
            data.edge_attr = torch.cat([pseudo, dist.type_as(pseudo)], dim=-1)
        else:
            data.edge_attr = dist
        return data
