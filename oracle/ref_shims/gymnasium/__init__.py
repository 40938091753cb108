"""TEST-ONLY empty stand-in (network_utils.py imports it at module level; the hot path never uses it)."""


class Env:  # noqa: D401
    pass


class Wrapper:  # noqa: D401
    pass


class _Spaces:
    class Box:  # noqa: D401
        pass


spaces = _Spaces()
