"""`src.utils` — the module name the reference's unmodified scripts import; re-exports plenoxels_b200.utils."""
from plenoxels_b200.utils import *  # noqa: F401,F403
