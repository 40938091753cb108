"""Commuting W2 projection: drop-in for projections/w2_projection_layer.py:14-76."""
from ..utils.projection_utils import gaussian_wasserstein_commutative
from .base_projection_layer import BaseProjectionLayer


class WassersteinProjectionLayer(BaseProjectionLayer):
    KERNEL_TYPE = "w2"

    def trust_region_value(self, policy, p, q):
        return gaussian_wasserstein_commutative(policy, p, q, scale_prec=self.scale_prec)
