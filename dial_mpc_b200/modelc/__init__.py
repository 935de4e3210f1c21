"""Model compiler: MJCF subset -> flat CompiledModel (see mjcf.py)."""
from .mjcf import CompiledModel, compile_mjcf, mass_matrix_fp64  # noqa: F401
