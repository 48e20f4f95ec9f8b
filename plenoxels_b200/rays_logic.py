"""`src.rays_logic` — imported by the reference's scripts/compare_inference_to_image.py:10 but missing from the reference
tree; provided so that script (and scripts/main.py) import unchanged."""
from .ray_sampling import compute_alpha_weighted_pixels  # noqa: F401
