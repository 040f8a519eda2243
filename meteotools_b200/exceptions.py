"""Exception classes of the reference (meteotools/exceptions.py:2-27), same names."""


class DimensionError(Exception):
    """Array rank / shape is wrong."""


class UnitError(Exception):
    """Unsupported unit."""


class InputError(Exception):
    """Insufficient or inconsistent inputs."""


class LengthError(Exception):
    """Wrong number of elements."""


class BackendError(RuntimeError):
    """The CUDA backend is missing, failed to load, or a device call failed (no CPU fallback)."""
