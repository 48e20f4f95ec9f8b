"""`src.data_processing` — the module name the reference's unmodified scripts import; re-exports plenoxels_b200.data_processing."""
from plenoxels_b200.data_processing import *  # noqa: F401,F403
