from .flooring import EPS, add_flooring, identity, max_flooring  # noqa: F401
