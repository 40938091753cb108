from .to_from_sphere import scalar_to_sphere, sphere_to_scalar, sphere_to_vec, vec_to_sphere  # noqa: F401
