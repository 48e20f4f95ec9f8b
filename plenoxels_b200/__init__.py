"""plenoxels_b200 — B200-native (sm_100a) voxel-grid volume renderer and training step of DanJbk/Plenoxels.

Layout (only what the hot path needs):
  csrc/            CUDA kernels + the extern "C" boundary (include/plenoxel_abi.h) -> libplenoxel_b200.so
  _lib.py          ctypes binding (no CPU fallback: a missing library or a CPU tensor raises PlxError)
  ops.py           torch.autograd wrappers over the C ABI (render_rays, render_train, adam_step, ...)
  trainer.py       one-call training step (VoxelTrainer) and the peer-memory multi-GPU step (PeerVoxelTrainer)
  lazy.py          storage-less tensor handles that route the reference's unfused call sequence to the fused march
  grid_functions / ray_sampling / rays_logic / visualization / data_processing / utils
                   the reference's module and function names (re-exported as `src.*` for its unmodified scripts)
  synth.py         seeded synthetic scenes for tests and benchmarks
"""
__version__ = "0.1.0"
