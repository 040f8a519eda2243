from ._impl import interp2D_fast, interp2D_fast_layers  # noqa: F401
