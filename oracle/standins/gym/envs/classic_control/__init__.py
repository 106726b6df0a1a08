from . import rendering  # noqa: F401
