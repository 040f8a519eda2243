from ._impl import interp1D_fast, interp1D_fast_layers, interp1D_nonequal, interp1D_nonequal_layers  # noqa: F401
