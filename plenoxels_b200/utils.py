"""Drop-in for the reference's `src/utils.py` (one unused helper)."""
import torch


def tensor_linspace(start, end, steps=10, device="cuda"):
    """Vectorised linspace: out[..., i] interpolates start -> end — src/utils.py:3-32."""
    assert start.size() == end.size()
    w_end = torch.linspace(0, 1, steps=steps, device=device).to(start)
    w_start = torch.linspace(1, 0, steps=steps, device=device).to(start)
    shape = (1,) * start.dim() + (steps,)
    return w_start.view(shape) * start.contiguous().unsqueeze(-1) + w_end.view(shape) * end.contiguous().unsqueeze(-1)
