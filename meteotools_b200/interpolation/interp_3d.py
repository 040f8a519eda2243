from ._impl import interp3D_fast, interp3D_fast_layers  # noqa: F401
