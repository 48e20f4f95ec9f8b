"""`src.grid_functions` — the module name the reference's unmodified scripts import; re-exports plenoxels_b200.grid_functions."""
from plenoxels_b200.grid_functions import *  # noqa: F401,F403
