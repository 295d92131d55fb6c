"""lennard-jones-cuda_b200 — B200-native Lennard-Jones MD step (sm_100a CUDA behind a C ABI).

The directory name is not a valid Python identifier; load it with `ljpkg.load()` from the repo
root (tests, bench.py and __graft_entry__.py all do) which registers it as
`lennard_jones_cuda_b200`.

  .ljmd       ctypes binding of include/ljmd.h (the product path; raises without the CUDA library)
  .snapshots  seeded synthetic initial configurations
  .dist       one-process-per-GPU plumbing over torch.distributed (unique-id exchange, barriers, max-over-ranks)
"""
from . import snapshots  # noqa: F401
from . import ljmd  # noqa: F401
from . import dist  # noqa: F401

__all__ = ["ljmd", "snapshots", "dist"]
