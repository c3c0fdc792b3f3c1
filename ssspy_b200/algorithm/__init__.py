from .projection_back import projection_back  # noqa: F401
