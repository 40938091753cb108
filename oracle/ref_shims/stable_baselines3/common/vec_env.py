"""TEST-ONLY empty stand-in."""


class VecEnvWrapper:  # noqa: D401
    pass


class VecNormalize:  # noqa: D401
    pass
