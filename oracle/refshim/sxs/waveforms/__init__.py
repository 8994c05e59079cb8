from . import alignment  # noqa: F401
