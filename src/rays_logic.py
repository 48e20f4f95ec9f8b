"""`src.rays_logic` — the module name the reference's unmodified scripts import; re-exports plenoxels_b200.rays_logic."""
from plenoxels_b200.rays_logic import *  # noqa: F401,F403
