"""Name-compatible shim: the reference's scripts do `from src.grid_functions import ...` (scripts/train.py:10-12).
With this repo's root on PYTHONPATH those imports resolve to the B200 implementation in `plenoxels_b200/`."""
