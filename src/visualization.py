"""`src.visualization` — the module name the reference's unmodified scripts import; re-exports plenoxels_b200.visualization."""
from plenoxels_b200.visualization import *  # noqa: F401,F403
