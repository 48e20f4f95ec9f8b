"""`src.ray_sampling` — the module name the reference's unmodified scripts import; re-exports plenoxels_b200.ray_sampling."""
from plenoxels_b200.ray_sampling import *  # noqa: F401,F403
