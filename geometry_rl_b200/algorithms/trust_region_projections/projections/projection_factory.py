"""projections/projection_factory.py:9-48 for the projection types the hot path reaches."""
from .base_projection_layer import BaseProjectionLayer
from .kl_projection_layer import KLProjectionLayer
from .w2_projection_layer import WassersteinProjectionLayer


def get_projection_layer(proj_type: str = "", **kwargs) -> BaseProjectionLayer:
    if (not proj_type or proj_type.isspace()
            or proj_type.lower() in ["ppo", "sac", "td3", "mpo", "vlearn", "vtrace", "awr", "entropy"]):
        return BaseProjectionLayer(proj_type, **kwargs)
    if proj_type.lower() == "w2":
        return WassersteinProjectionLayer(proj_type, **kwargs)
    if proj_type.lower() == "kl":
        return KLProjectionLayer(proj_type, **kwargs)
    if proj_type.lower() in ("w2_non_com", "frob", "papi"):
        raise NotImplementedError(f"projection '{proj_type}' is outside the hot path (policy_type gnn_diag + proj_type "
                                  "kl|w2 are the only combinations the shipped configs reach; SURVEY 2 row 9)")
    raise ValueError(f"Invalid projection type {proj_type}.")
