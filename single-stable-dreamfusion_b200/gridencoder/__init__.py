"""Drop-in for the reference's ``gridencoder`` package (gridencoder/grid.py)."""
from .grid import GridEncoder, grid_encode  # noqa: F401
