"""KL projection: drop-in for projections/kl_projection_layer.py:14-111 (diagonal, contextual std).
The ITPAL op `cpp_projection.BatchedDiagCovOnlyProjection` (:162-204) is replaced by the in-register
fp64 dual solve + implicit gradient of grl_trpl_fwd / grl_trpl_bwd."""
from .base_projection_layer import BaseProjectionLayer


class KLProjectionLayer(BaseProjectionLayer):
    KERNEL_TYPE = "kl"
